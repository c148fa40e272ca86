#!/bin/bash
# end-of-round evidence at HEAD: GPU tests, smoke, default bench line, reference arm, c1 / c2d / c4 lines, launch list
OUT=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2d_gputests.log 2>&1
echo "tests rc=$?"; tail -3 $OUT/r2d_gputests.log
timeout 300 python __graft_entry__.py smoke > $OUT/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2d_smoke.log
timeout 900 python bench.py > $OUT/r2d_bench_default.json 2> $OUT/r2d_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2d_bench_reference.json 2> $OUT/r2d_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2d_bench_c1.json 2> $OUT/r2d_bench_c1.err; echo "c1 rc=$?"
timeout 600 python bench.py --workload c2d --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2d_bench_c2d.json 2> $OUT/r2d_bench_c2d.err; echo "c2d rc=$?"
timeout 1200 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2d_bench_c4.json 2> $OUT/r2d_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("default", "reference", "c1", "c2d", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/r2d_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("f64_default_path") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), d.get("scaling"))
    except Exception as e:
        print(f, "no line", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2d_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f64 > $OUT/r2d_ncu_launches.log 2>&1; echo "launches rc=$?"
