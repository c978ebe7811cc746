#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 400 python bench.py --steps 300 --warmup 5 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02s_soak.json 2> gpurun_out/r02s_soak.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02s_soak.json") if l.startswith("{")][-1]
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["steps"], d["clocks"], d["loss"])
PY
nvidia-smi --query-gpu=memory.used,clocks.sm,power.draw --format=csv
tail -3 gpurun_out/r02s_soak.err
