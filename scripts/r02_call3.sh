#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s -k "loss or pose_to_bvh or generate_gesture or train_step" ) > gpurun_out/r02c_pytest.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r02c_pytest.log | tail -20
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02c_bench.json") if l.startswith("{")][-1]
print(d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"]["value"])
PY
tail -3 gpurun_out/r02c_bench.err
