#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_wgridder.py tests/test_gpu_fft.py tests/test_gpu_cols2.py tests/test_gpu_split.py tests/test_gpu_batch.py tests/test_gpu_coverage.py -x -q -m gpu 2>&1 | tail -3
for b in 1 7; do python tools/prof_band.py $b 3 c2 2>&1 | tail -2; PFBG_ROWS_ODD=pair python tools/prof_band.py $b 3 c2 2>&1 | tail -2; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f64 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['ms_per_band'])"
PFBG_ROWS_ODD=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f64 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['ms_per_band'])"
