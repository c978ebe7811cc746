#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x -k "encoder or train_step or graph_replayed or label_style or sgemm" ) > gpurun_out/r02e2_pytest.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02e2_bench.json 2> gpurun_out/r02e2_bench.err
( ZEGGS_BENCH_LOSS_LAG=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --alt 0 --extras 0 ) > gpurun_out/r02e2_bench_lag0.json 2> gpurun_out/r02e2_bench_lag0.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r02e2_pytest.log | tail -20
python - <<PY
import json
for f in ("r02e2_bench.json","r02e2_bench_lag0.json"):
    d=[json.loads(l) for l in open("gpurun_out/"+f) if l.startswith("{")][-1]
    print(f, d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"]["value"], d["value"])
PY
tail -3 gpurun_out/r02e2_bench.err
