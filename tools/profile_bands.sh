#!/bin/bash
# ncu evidence of one build: plain run first (must exit 0), then `--set full` captures of the six hot kernels of a
# Hessian apply for bands 0 and 7 of C2, and the launch list of the bench command.  Outputs under gpurun_out/.
# usage (GPU box): bash tools/profile_bands.sh TAG
TAG=${1:-r2}; OUT=gpurun_out
for b in 0 7; do
  python tools/prof_band.py $b 2 > $OUT/prof_plain_b$b.log 2>&1 || { echo "plain run of band $b failed"; tail -5 $OUT/prof_plain_b$b.log; exit 1; }
  tail -1 $OUT/prof_plain_b$b.log
  ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_cols|k_grid_runs|k_degrid_runs" -s 12 -c 6 \
      -f -o $OUT/${TAG}_b$b python tools/prof_band.py $b 2 > $OUT/ncu_full_b$b.log 2>&1
  ncu -i $OUT/${TAG}_b$b.ncu-rep --page raw --csv > $OUT/${TAG}_raw_b$b.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/${TAG}_raw_b$b.csv > $OUT/${TAG}_summary_b$b.md
  # gpurun brings back at most 64 MiB: keep the report of band 0 only (~50 MB each)
  [ $b != 0 ] && rm -f $OUT/${TAG}_b$b.ncu-rep
done
python tools/ncu_traffic.py $OUT/traffic.json c2 "ncu --set full --clock-control none, tools/prof_band.py, C2 bands 0 and 7 (fp32, eps 1e-5), build $TAG" \
   band0=$OUT/${TAG}_raw_b0.csv band7=$OUT/${TAG}_raw_b7.csv
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f64 > $OUT/bench_for_launches.json 2> $OUT/bench_for_launches.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f64 > $OUT/ncu_launches.log 2>&1
ls -la $OUT | tail -20
