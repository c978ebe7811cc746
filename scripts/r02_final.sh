#!/bin/bash
# final validation of the round: smoke, the full GPU suite (with the shipped v1 weights present), the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02z_smoke.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -s -rs ) > gpurun_out/r02z_pytest.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
tail -3 gpurun_out/r02z_smoke.log
grep -E "passed|failed|FAILED|Error|SKIP" gpurun_out/r02z_pytest.log | tail -12
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02z_bench.json") if l.startswith("{")][-1]
print(d["ms_per_step"], d["value"], d["e2e"], d["steps"], d["warmup"], d["gpu_launches"], d["clocks"], d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value"))
print(d["kernel_ms_per_step"])
PY
tail -3 gpurun_out/r02z_bench.err
