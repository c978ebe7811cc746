#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --alt 0 --extras 0 --no-cpu"
timeout 300 python bench.py --gpus 1 $Q > gpurun_out/r02p_n1.json 2> gpurun_out/r02p_n1.err
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N $Q > gpurun_out/r02p_n$N.json 2> gpurun_out/r02p_n$N.err
done
python - <<PY
import json
base=None
for n in (1,4,8):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r02p_n{n}.json") if l.startswith("{")][-1]
        if n==1: base=d["value"]
        print(n, d["ms_per_step"], d["value"], "eff", round(d["value"]/(n*base),4) if base else None, d["kernel_ms_per_step"].get("allreduce"), "enq", d["host_enqueue_ms_per_step"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(n, "FAILED", e)
PY
tail -n 3 gpurun_out/r02p_n8.err
