#!/bin/bash
# tests of the DMMA run kernels, then ncu --set full of both (C2 geometry, eps 1e-7, band 0).  usage: bash tools/r2d_ncu_mma.sh TAG
TAG=${1:-r2d_mma}; OUT=gpurun_out
true

python tools/prof_band.py 0 2 c2d > $OUT/${TAG}_plain_c2d.log 2>&1 || { echo plain failed; tail -5 $OUT/${TAG}_plain_c2d.log; exit 1; }
tail -2 $OUT/${TAG}_plain_c2d.log
ncu --set full --clock-control none --import-source on -k regex:"runs_mma" -s 4 -c 2 \
    -f -o $OUT/${TAG} python tools/prof_band.py 0 2 c2d > $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_raw.csv > $OUT/${TAG}_summary.md
cat $OUT/${TAG}_summary.md | head -40
