#!/bin/bash
# kernel-by-kernel times of one C4 band (10240^2, 62.5 M vis, fp32): which transform pass costs what
OUT=gpurun_out
true

timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2d_c4_launches.csv python tools/prof_band.py 0 1 c4 > $OUT/r2d_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2d_c4_launches.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size") if "Grid Size" in h else None; ib=h.index("Block Size") if "Block Size" in h else None
for r in rows[1:]:
    print(r[ik][:70], r[ig] if ig else "", r[ib] if ib else "", r[iv])
PY
