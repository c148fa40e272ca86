#!/bin/bash
OUT=gpurun_out
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
timeout 600 python bench.py --workload c3 --steps 50 --warmup 3 > $OUT/r2c_bench_c3.json 2> $OUT/r2c_bench_c3.err; echo "c3 rc=$?"; tail -2 $OUT/r2c_bench_c3.err; cut -c1-600 $OUT/r2c_bench_c3.json
bash tools/profile_bands.sh r2c > $OUT/r2c_profile.log 2>&1; tail -3 $OUT/r2c_profile.log
timeout 900 python bench.py > $OUT/r2c_bench_default.json 2> $OUT/r2c_bench_default.err
echo "bench rc=$?"; tail -2 $OUT/r2c_bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phases_ms_rank0"], d["cpu_baseline"], d["f64_default_path"])
PY
