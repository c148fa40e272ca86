#!/bin/bash
# first GPU run of the TMA-fed column kernels: engine + parity tests, then A/B bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_tests_cols2.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2b_tests_cols2.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f64 > gpurun_out/r2b_bench_cols2.json 2> gpurun_out/r2b_bench_cols2.err
echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench_cols2.err
PFBG_COLS=old PFBG_ROWS=old timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f64 > gpurun_out/r2b_bench_colsold.json 2> gpurun_out/r2b_bench_colsold.err
echo "bench old rc=$?"
python - <<'PY'
import json
for n in ("cols2", "colsold"):
    try:
        d = json.loads(open(f"gpurun_out/r2b_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phases_ms_rank0"])
    except Exception as e:
        print(n, "no line", e)
PY
