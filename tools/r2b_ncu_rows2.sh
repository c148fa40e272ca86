#!/bin/bash
OUT=gpurun_out
python tools/prof_band.py 0 2 > $OUT/r2b_prof_plain_b0.log 2>&1 || { tail -5 $OUT/r2b_prof_plain_b0.log; exit 1; }
tail -1 $OUT/r2b_prof_plain_b0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rows2" -s 4 -c 2 -f -o $OUT/r2b_rows2 python tools/prof_band.py 0 2 > $OUT/r2b_ncu_rows2.log 2>&1
ncu -i $OUT/r2b_rows2.ncu-rep --page raw --csv > $OUT/r2b_rows2_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/r2b_rows2_raw.csv
