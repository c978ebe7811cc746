#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 300 python scripts/tc_trace_bwd.py ) 2>&1 | tail -18
( timeout 900 python -m pytest tests -m gpu -q -x -k "tc or golden or curve or graph" ) 2>&1 | tail -4
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02g_bench.json") if l.startswith("{")][-1]
print(d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/r02g_bench.err
