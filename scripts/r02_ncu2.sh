#!/bin/bash
# second ncu pass of round 2 (after the loss rewrite, the GEMM tile variants and the lanes): launch list of one eager train step and
# --set full captures of the BPTT kernel, the new GEMM variants and the warp-per-frame loss kernels.  1 GPU.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
export ZEGGS_LANES=0          # one stream: the launch list then reads in program order
PROF_T=256 PROF_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_launches_T256.csv python scripts/profile_step.py > gpurun_out/r02k_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
PROF_T=256 PROF_STEPS=1 timeout 600 $NCU -k regex:^decoder_bwd_tc_kernel -c 1 -f -o gpurun_out/r02k_ncu_bwd python scripts/profile_step.py > gpurun_out/r02k_ncu_bwd.log 2>&1
PROF_T=256 PROF_STEPS=1 timeout 900 $NCU -k regex:^tc_gemm_kernel --launch-skip 14 -c 22 -f -o gpurun_out/r02k_ncu_tcgemm python scripts/profile_step.py > gpurun_out/r02k_ncu_tcgemm.log 2>&1
PROF_T=256 PROF_STEPS=1 timeout 600 $NCU -k regex:^loss_ -c 5 -f -o gpurun_out/r02k_ncu_loss python scripts/profile_step.py > gpurun_out/r02k_ncu_loss.log 2>&1
ls -la gpurun_out/r02k* | tail -8; tail -n 2 gpurun_out/r02k_ncu_bwd.log gpurun_out/r02k_ncu_tcgemm.log gpurun_out/r02k_ncu_loss.log
