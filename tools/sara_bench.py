"""BASELINE config 3 ("pfb sara: 8-band L21 wavelet deconv, 4096^2, pcg + forward-backward iterations, NCCL L21
all-reduce"): time the device-resident primal-dual iteration (Psi^T, fused l21 dual update, Psi, PSF-convolution
Hessian gradient, primal step, convergence norm) with the 8 bands sharded over the ranks.

  python tools/sara_bench.py [--iters 50] [--nx 4096]                                  # one GPU, 8 bands
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29522 \
      tools/sara_bench.py                                                               # bands sharded over N GPUs

Prints one JSON line on rank 0: ms per primal-dual iteration (max over ranks), and the split measured on rank 0.
Synthetic Gaussian PSFs and a point-source sky; the numbers are rates, not science."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfb_imaging_b200 import dist  # noqa: E402
from pfb_imaging_b200.plan import good_size  # noqa: E402
from pfb_imaging_b200.psf import HessPSF, PsfGradient  # noqa: E402
from pfb_imaging_b200.sara import L21, PrimalDual, PsiNocopyt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--nband", type=int, default=8)
    ap.add_argument("--nlevel", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init()
    rank = dist.rank()
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    nx = ny = args.nx
    nxp = nyp = good_size(int(1.4 * nx))
    bands = dist.local_bands(args.nband)
    nb = len(bands)
    bases = ["self", "db1", "db2", "db3"]
    rng = np.random.default_rng(100 + rank)
    # Gaussian PSFs, one width per band; abspsf = |rfft2| (operators/hessian.py:300-312)
    xx = np.fft.fftfreq(nxp)[:, None] * nxp
    yy = np.fft.rfftfreq(nyp)[None, :] * nyp
    abspsf = np.stack([np.exp(-2.0 * (np.pi * (1.5 + 0.2 * b)) ** 2 * ((xx / nxp) ** 2 + (yy / nyp) ** 2)) for b in bands])
    truth = np.zeros((nb, nx, ny))
    truth[:, rng.integers(0, nx, 200), rng.integers(0, ny, 200)] = np.exp(rng.standard_normal(200))
    hess = HessPSF(nx, ny, abspsf, beam=None, eta=1e-3)
    dirty = hess.dot(truth) + 1e-3 * rng.standard_normal(truth.shape)
    psi = PsiNocopyt(nb, nx, ny, bases, args.nlevel, 1, device=dev)
    reg = L21(psi, bases, nu=len(bases))
    hooks = dict(reduce_tensor=dist.allreduce_sum, reduce_scalars=dist.allreduce_sum) if world > 1 else {}
    pd = PrimalDual(tol=0.0, maxit=3, verbosity=0, positivity=1, **hooks)
    pd.setup(reg, 1.0 + 1e-3)
    pd.set_grad(PsfGradient(hess, dirty))
    x0 = np.zeros_like(dirty)
    pd.solve(x0, 1e-4)  # warm-up (allocations, first launches)
    # steady-state cost per iteration: difference of a long and a short solve (the upload of the weights / the
    # download of the model at the two ends of solve() are per-call, not per-iteration, costs)
    times = {}
    for n in (5, 5 + args.iters):
        pd.maxit = n
        pd.reset()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        x = pd.solve(x0, 1e-4)
        torch.cuda.synchronize()
        times[n] = time.perf_counter() - t0
    ms = (times[5 + args.iters] - times[5]) / args.iters * 1e3
    per_call_ms = (times[5] - 5 * ms * 1e-3) * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", dev))
    if world > 1:
        dist.allreduce_max(t)
    # split on rank 0 (device tensors, CUDA events)
    split = {}
    if rank == 0:
        d = torch.device("cuda", dev)
        x_t = torch.from_numpy(np.ascontiguousarray(x)).to(d)
        v_t = torch.empty(psi.coeff_shape, dtype=torch.float64, device=d)
        o_t = torch.empty_like(x_t)
        grad = PsfGradient(hess, dirty)

        def timed(fn, n=5):
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        split["psi_dot_ms"] = timed(lambda: psi.dot_dev(x_t, v_t))
        split["psi_hdot_ms"] = timed(lambda: psi.hdot_dev(v_t, o_t))
        split["psf_hessian_grad_ms"] = timed(lambda: grad.device_apply(x_t, o_t))
        print(json.dumps({"workload": f"config 3: {args.nband} bands x {nx}^2, bases {','.join(bases)}, {args.nlevel} levels, "
                                      f"PSF-convolution Hessian {nxp}^2, fp64, positivity 1",
                          "n_gpus": world, "bands_per_rank": nb, "iters": args.iters,
                          "ms_per_pd_iteration": float(t.item()), "per_solve_call_overhead_ms_rank0": per_call_ms, "rank0_split_ms": split,
                          "model_flux": float(np.abs(x).sum())}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
