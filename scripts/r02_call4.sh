#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 300 python scripts/tc_trace.py ) > gpurun_out/r02d_trace_fwd.log 2>&1
( timeout 300 python scripts/tc_trace_bwd.py ) > gpurun_out/r02d_trace_bwd.log 2>&1
PROF_T=256 PROF_STEPS=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/r02d_launches_T256.csv python scripts/profile_step.py > gpurun_out/r02d_ncu.log 2>&1
tail -n 30 gpurun_out/r02d_trace_fwd.log; tail -n 25 gpurun_out/r02d_trace_bwd.log; tail -n 3 gpurun_out/r02d_ncu.log; wc -l gpurun_out/r02d_launches_T256.csv
