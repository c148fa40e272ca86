#!/bin/bash
# the three multi-band workloads the way the driver launches bench.py on N GPUs
N=${1:-8}; OUT=gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
run() {  # name, bench args...
  name=$OUT/r2c_bench_$1_n$N; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N "$@" > $name.json 2> $name.err
  echo "rc=$? $name"; grep -v "^\*\*\*\|OMP_NUM_THREADS" $name.err | tail -3
  python - "$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c = d["config"]
    print({k: d[k] for k in ("value", "unit", "ms_per_step", "scaling", "n_gpus", "gpu_launches")}, "e2e", d["e2e"]["value"])
    print({k: c.get(k) for k in ("ms_per_rank", "band_owner", "plane_offloads", "plane_stacks_rank0", "collective")})
except Exception as e:
    print("no line:", e)
PY
}
run c2 --steps 20 --warmup 3
run c3 --workload c3 --steps 50 --warmup 3
run c4 --workload c4 --steps 5 --warmup 3
