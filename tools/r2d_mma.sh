#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_wgridder.py -x -q -m gpu > $OUT/r2d_mma_tests.log 2>&1
echo "tests rc=$?"; tail -15 $OUT/r2d_mma_tests.log
timeout 300 python tools/prof_band.py 0 3 c2d 2>&1 | tail -2
PFBG_WIDE_MMA=0 timeout 300 python tools/prof_band.py 0 3 c2d 2>&1 | tail -1
