#!/bin/bash
# ncu --set full captures of the kernels the verdict asked evidence for (1 GPU; each capture replays its kernels ~40x)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
PROF_T=256 PROF_STEPS=1 timeout 900 $NCU -k regex:decoder_.*_tc_kernel -c 2 -f -o gpurun_out/r02_ncu_decoder python scripts/profile_step.py > gpurun_out/r02_ncu_decoder.log 2>&1
PROF_T=256 PROF_STEPS=1 timeout 900 $NCU -k regex:tc_gemm_kernel -c 12 -f -o gpurun_out/r02_ncu_tcgemm python scripts/profile_step.py > gpurun_out/r02_ncu_tcgemm.log 2>&1
PROF_T=256 PROF_STEPS=1 timeout 900 $NCU -k regex:loss_ -c 6 -f -o gpurun_out/r02_ncu_loss python scripts/profile_step.py > gpurun_out/r02_ncu_loss.log 2>&1
timeout 600 $NCU -k regex:mel_kernel -c 2 -f -o gpurun_out/r02_ncu_mel python scripts/mel_prof.py > gpurun_out/r02_ncu_mel.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8; tail -n 2 gpurun_out/r02_ncu_decoder.log gpurun_out/r02_ncu_tcgemm.log gpurun_out/r02_ncu_loss.log gpurun_out/r02_ncu_mel.log
