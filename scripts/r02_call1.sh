#!/bin/bash
# round 2, GPU call 1: graph sanity bench, full GPU test suite, full bench line, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 --graph 0 ) > gpurun_out/r02_bench_quick_eager.json 2> gpurun_out/r02_bench_quick_eager.err
( time timeout 2400 python -m pytest tests -m gpu -q -s ) > gpurun_out/r02_pytest_gpu.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -n 30 gpurun_out/r02_pytest_gpu.log
cat gpurun_out/r02_bench_quick.json gpurun_out/r02_bench_quick_eager.json gpurun_out/r02_bench_full.json gpurun_out/r02_bench_ref.json
tail -n 5 gpurun_out/r02_bench_quick.err gpurun_out/r02_bench_full.err
