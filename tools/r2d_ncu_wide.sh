#!/bin/bash
# ncu --set full of the fp64 wide run kernels (C2 geometry, eps 1e-7, band 0).  usage (GPU box): bash tools/r2d_ncu_wide.sh TAG
TAG=${1:-r2d}; OUT=gpurun_out
python tools/prof_band.py 0 2 c2d > $OUT/${TAG}_plain_c2d.log 2>&1 || { echo plain failed; tail -5 $OUT/${TAG}_plain_c2d.log; exit 1; }
tail -2 $OUT/${TAG}_plain_c2d.log
ncu --set full --clock-control none --import-source on -k regex:"runs_wide" -s 4 -c 2 \
    -f -o $OUT/${TAG}_wide python tools/prof_band.py 0 2 c2d > $OUT/${TAG}_ncu_wide.log 2>&1
ncu -i $OUT/${TAG}_wide.ncu-rep --page raw --csv > $OUT/${TAG}_wide_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_wide.ncu-rep --page source --csv > $OUT/${TAG}_wide_source.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_wide_raw.csv > $OUT/${TAG}_wide_summary.md
cat $OUT/${TAG}_wide_summary.md | head -60
