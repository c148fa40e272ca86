#!/bin/bash
# Development driver of the band-split path on an N-GPU box (gpurun --gpus N -- bash tools/strong_dev.sh N BANDS):
# the two-process IPC parity test, then the bench with and without plane offload.
N=${1:-2}; BANDS=${2:-7,0}; OUT=gpurun_out; TAG=${3:-dev}
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 400 python -m pytest tests/test_gpu_split.py -x -q 2>&1 | tail -15 > $OUT/r2_split_ipc_$TAG.log
cat $OUT/r2_split_ipc_$TAG.log
for mode in "" "--no-offload"; do
  name=$OUT/r2_bench_n${N}_${TAG}${mode// /}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 20 --warmup 3 --bands $BANDS --no-cpu-baseline $mode > $name.json 2> $name.err
  echo "rc=$? $name"; tail -3 $name.err
  python - "$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c = d["config"]
    print({k: d[k] for k in ("value", "ms_per_step", "scaling", "n_gpus")})
    print({k: c[k] for k in ("ms_per_band", "ms_per_rank", "band_owner", "plane_offloads", "modelled_ms_per_rank", "helper_phases_ms")})
    print("e2e", d["e2e"]["value"])
except Exception as e:
    print("no line:", e)
PY
done
