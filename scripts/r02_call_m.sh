#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02n_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r02n_pytest.log | tail -20
python - <<PY
import json
for f in ("r02n_bench.json",):
    try:
        d=[json.loads(l) for l in open("gpurun_out/"+f) if l.startswith("{")][-1]
        print(f, d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"]["value"], d["value"], d["config"].get("cuda_graph"))
        print(json.dumps(d.get("alt_config"))[:300])
        oc=d.get("other_configs",{})
        for k,v in oc.items(): print(k, json.dumps(v)[:400])
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/r02n_bench.err
