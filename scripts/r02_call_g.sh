#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
( ZEGGS_LANE_PRIORITY=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02g_bench_noprio.json 2> gpurun_out/r02g_bench_noprio.err
python - <<PY
import json
for f in ("r02g_bench.json","r02g_bench_noprio.json"):
    try:
        d=[json.loads(l) for l in open("gpurun_out/"+f) if l.startswith("{")][-1]
        print(f, d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"]["value"], d["value"], d["config"].get("cuda_graph"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/r02g_bench.err
