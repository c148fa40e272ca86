#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_wgridder.py tests/test_gpu_dropins.py tests/test_gpu_coverage.py -x -q -m gpu > $OUT/r2d_mma2_tests.log 2>&1
echo "tests rc=$?"; tail -8 $OUT/r2d_mma2_tests.log
timeout 300 python tools/prof_band.py 0 3 c2d 2>&1 | tail -2
timeout 300 python tools/prof_band.py 0 3 c1 2>&1 | tail -2
PFBG_MMA_NT8=0 timeout 300 python tools/prof_band.py 0 3 c1 2>&1 | tail -1
