#!/bin/bash
OUT=gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2d_c2d_launches.csv python tools/prof_band.py 0 1 c2d > $OUT/r2d_c2d_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2d_c2d_launches.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size"); ib=h.index("Block Size")
for r in rows[1:]:
    if "cub" in r[ik] or "k_bin" in r[ik]: continue
    print(r[ik][:60], r[ig], r[ib], r[iv])
PY
