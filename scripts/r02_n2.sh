#!/bin/bash
# 2 GPUs: NCCL data-parallel equivalence test (2-GPU gradients == 1-GPU gradients on the concatenated batch) + bench at N=2 and N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -s ) > gpurun_out/r02o_pytest_dp.log 2>&1
Q="--steps 20 --warmup 3 --alt 0 --extras 0 --no-cpu"
timeout 300 python bench.py --gpus 1 $Q > gpurun_out/r02o_n1.json 2> gpurun_out/r02o_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $Q > gpurun_out/r02o_n2.json 2> gpurun_out/r02o_n2.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r02o_pytest_dp.log | tail
python - <<PY
import json
base=None
for n in (1,2):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02o_n{n}.json") if l.startswith("{")][-1]
        if n==1: base=d["value"]
        print(n, d["ms_per_step"], d["value"], "eff", round(d["value"]/(n*base),4) if base else None, d["kernel_ms_per_step"].get("allreduce"), "enq", d["host_enqueue_ms_per_step"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(n, "FAILED", e)
PY
tail -n 3 gpurun_out/r02o_n2.err
