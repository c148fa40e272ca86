#!/bin/bash
OUT=gpurun_out
for r in 0 1 2 3; do
  echo "== PFBG_R8=$r band 0 / 7"
  PFBG_R8=$r python tools/prof_band.py 0 3 2>&1 | tail -1
  PFBG_R8=$r python tools/prof_band.py 7 3 2>&1 | tail -1
done
PFBG_R8=3 timeout 600 python -m pytest tests/test_gpu_cols2.py tests/test_gpu_fft.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -3
