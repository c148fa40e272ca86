#!/bin/bash
# bench.py exactly the way the driver launches it on N GPUs (no extra environment): gpurun --gpus N -- bash tools/r2d_n.sh N
N=${1:-2}
name=gpurun_out/r2d_bench_c2_n${N}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
  bench.py --gpus $N --steps 20 --warmup 3 > $name.json 2> $name.err
echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM_THREADS" $name.err | tail -5
python - "$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c = d["config"]
    print({k: d[k] for k in ("value", "ms_per_step", "scaling", "n_gpus", "gpu_launches")})
    print({k: c[k] for k in ("ms_per_band", "ms_per_rank", "band_owner", "plane_offloads", "modelled_ms_per_rank")})
    print("e2e", d["e2e"]["value"], "clocks", d["clocks"])
except Exception as e:
    print("no line:", e)
PY
