#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
Q="--steps 10 --warmup 3 --no-cpu --alt 0 --extras 0"
for v in "g1p1:--graph 1 --pin 1" "g1p0:--graph 1 --pin 0" "g0p1:--graph 0 --pin 1" "g0p0:--graph 0 --pin 0"; do
  n=${v%%:*}; f=${v#*:}
  ( timeout 300 python bench.py $Q $f ) > gpurun_out/r02b_$n.json 2> gpurun_out/r02b_$n.err
done
( time timeout 1800 python -m pytest tests -m gpu -q -s -k "graph or golden or curve or strict or loudness" ) > gpurun_out/r02b_pytest.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r02b_bench_full.json 2> gpurun_out/r02b_bench_full.err
tail -n 15 gpurun_out/r02b_pytest.log
for n in g1p1 g1p0 g0p1 g0p0; do python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r02b_$n.json") if l.startswith("{")][-1]
    print("$n", d["ms_per_step"], "enq", d["host_enqueue_ms_per_step"], "e2e", d["e2e"]["value"], "graph", d["config"]["cuda_graph"], d["kernel_ms_per_step"], "launches", d["gpu_launches"])
except Exception as e:
    print("$n", "FAILED", e)
PY
tail -n 3 gpurun_out/r02b_$n.err; done
cat gpurun_out/r02b_bench_full.json; tail -n 5 gpurun_out/r02b_bench_full.err
