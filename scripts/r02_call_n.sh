#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "gemm or encoder or train_step or lanes or full_size or label_style or graph_replayed" ) > gpurun_out/r02r_pytest.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r02r_pytest.log | tail -20
python - <<PY
import json
for f in ("r02r_bench.json",):
    try:
        d=[json.loads(l) for l in open("gpurun_out/"+f) if l.startswith("{")][-1]
        print(f, d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"]["value"], d["value"], d["config"].get("cuda_graph"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/r02r_bench.err
