#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/r2c_tests.log 2>&1
echo "tests rc=$?"; tail -4 $OUT/r2c_tests.log
timeout 1200 python bench.py --workload c4 --steps 3 --warmup 3 > $OUT/r2c_bench_c4.json 2> $OUT/r2c_bench_c4.err
echo "c4 rc=$?"; tail -3 $OUT/r2c_bench_c4.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench_c4.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"], d["config"]["plan_first_band"], d["config"]["plane_stacks_rank0"], d["roofline"]["phases_ms_rank0"])
except Exception as e:
    print("c4 no line", e)
PY
