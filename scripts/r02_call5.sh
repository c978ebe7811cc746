#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 600 python scripts/fwd_variants.py ) > gpurun_out/r02e_fwd_variants.log 2>&1
( timeout 300 python scripts/tc_trace.py ) > gpurun_out/r02e_trace_fwd.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x -k "tc_engine_vs_oracle or full_size_tc or gather" ) > gpurun_out/r02e_pytest.log 2>&1
tail -n 8 gpurun_out/r02e_fwd_variants.log; tail -n 24 gpurun_out/r02e_trace_fwd.log; tail -n 6 gpurun_out/r02e_pytest.log
