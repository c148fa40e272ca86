#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cols2.py -x -q -m gpu > $OUT/r2b_tests_cols2.log 2>&1
echo "tests rc=$?"; tail -3 $OUT/r2b_tests_cols2.log
for d in 0 16; do
  echo "== debug $d band 0 / 7"
  PFBG_COLS2_DEBUG=$d python tools/prof_band.py 0 3 2>&1 | tail -1
  PFBG_COLS2_DEBUG=$d python tools/prof_band.py 7 3 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cols2" -s 4 -c 2 -f -o $OUT/r2b_cols2_d0 python tools/prof_band.py 0 2 > $OUT/r2b_ncu_cols2.log 2>&1
ncu -i $OUT/r2b_cols2_d0.ncu-rep --page raw --csv > $OUT/r2b_cols2_raw_d0.csv 2>/dev/null
python tools/ncu_summary.py $OUT/r2b_cols2_raw_d0.csv
