#!/bin/bash
OUT=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -6 $OUT/r2d_tests.log
for wl in c1 c2d; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2d_bench_$wl.json 2> $OUT/r2d_bench_$wl.err
  echo "$wl rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["plan_first_band"], d["roofline"]["phases_ms_rank0"])
except Exception as e:
    print("$wl no line", e)
PY
done
