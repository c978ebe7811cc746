#!/bin/bash
# 2 GPUs: NCCL data-parallel equivalence test + the bench line at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -s ) 2>&1 | tail -8
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --alt 0 ) > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02h_bench_n2.json") if l.startswith("{")][-1]
print(d["n_gpus"], d["ms_per_step"], d["value"], d["kernel_ms_per_step"], d["e2e"]["value"], d["config"]["cuda_graph"], d.get("other_configs",{}).get("batch_inference_64x60s"))
PY
tail -3 gpurun_out/r02h_bench_n2.err
