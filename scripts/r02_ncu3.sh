#!/bin/bash
# final ncu pass of round 2: launch list of one eager train step (one stream) + --set full captures of the GEMM variants after the
# epilogue / conv-operand changes, the conv operand producers and the loss kernels.  1 GPU.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
export ZEGGS_LANES=0
PROF_T=256 PROF_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02q_launches_T256.csv python scripts/profile_step.py > gpurun_out/r02q_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
PROF_T=256 PROF_STEPS=1 timeout 900 $NCU -k regex:^tc_gemm_kernel --launch-skip 8 -c 30 -f -o gpurun_out/r02q_ncu_tcgemm python scripts/profile_step.py > gpurun_out/r02q_ncu_tcgemm.log 2>&1
PROF_T=256 PROF_STEPS=1 timeout 600 $NCU -k regex:^im2col_split -c 6 -f -o gpurun_out/r02q_ncu_im2col python scripts/profile_step.py > gpurun_out/r02q_ncu_im2col.log 2>&1
ls -la gpurun_out/r02q* | tail -8; tail -n 2 gpurun_out/r02q_ncu_tcgemm.log gpurun_out/r02q_ncu_im2col.log
