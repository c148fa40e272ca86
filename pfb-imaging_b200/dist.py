"""Band sharding and the cross-band reductions of the deconvolution loops.

Imaging bands are independent for gridding / degridding / the Hessian
(/root/reference/src/pfb_imaging/operators/band_worker.py:1-18), so band b lives on rank
``b % world`` (one process per GPU, like one ``_BandWorkerImpl`` actor per band) and the data
path needs no collective.  The reference does its cross-band maths on the driver in numpy after
``ray.get``; here they are ``torch.distributed`` all-reduces (NCCL over NVLink on GPUs, gloo in
the CPU tests):

  band sum of the L21 prox          prox/prox_21m.py:123-135      -> allreduce_sum
  pcg / power-method / PD scalars   opt/pcg.py:35-85, opt/power_method.py:60-66,
                                    opt/primal_dual.py:429        -> vdot_allreduce / allreduce_sum
  positivity mode 2                 prox/positivity.py:22-32      -> allreduce_min
  MFS residual / PSF / wsum sums    core/grid.py:439-445, core/sara.py:154-158 -> allreduce_sum
"""

from __future__ import annotations

import os

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def is_initialized() -> bool:
    try:
        d = _dist()
        return d.is_available() and d.is_initialized()
    except Exception:
        return False


def world_size() -> int:
    return _dist().get_world_size() if is_initialized() else 1


def rank() -> int:
    return _dist().get_rank() if is_initialized() else 0


def init(backend: str | None = None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    import torch

    d = _dist()
    if d.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        d.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        d.init_process_group(backend)


def band_owner(band: int, world: int | None = None) -> int:
    return band % (world_size() if world is None else world)


def local_bands(nband: int, r: int | None = None, world: int | None = None) -> list[int]:
    r = rank() if r is None else r
    world = world_size() if world is None else world
    return [b for b in range(nband) if b % world == r]


def _reduce(arr, op_name):
    """All-reduce a numpy array (any float dtype) or a torch tensor in place; returns it."""
    if world_size() == 1:
        return arr
    import torch

    d = _dist()
    op = {"sum": d.ReduceOp.SUM, "min": d.ReduceOp.MIN, "max": d.ReduceOp.MAX}[op_name]
    if isinstance(arr, torch.Tensor):
        d.all_reduce(arr, op=op)
        return arr
    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a)
    if d.get_backend() == "nccl":
        tg = t.cuda(non_blocking=False)
        d.all_reduce(tg, op=op)
        t.copy_(tg)
    else:
        d.all_reduce(t, op=op)
    if a is not arr:
        arr[...] = a
    return arr


def allreduce_sum(arr):
    return _reduce(arr, "sum")


def allreduce_min(arr):
    return _reduce(arr, "min")


def allreduce_max(arr):
    return _reduce(arr, "max")


def vdot_allreduce(*pairs):
    """Sum over ranks of the local dot products <a,b> for every (a, b) pair: one message of
    len(pairs) doubles (the 2-3 scalars of pcg / power method, opt/pcg.py:35-41)."""
    loc = np.array([float(np.vdot(a, b).real) for a, b in pairs], dtype=np.float64)
    return allreduce_sum(loc)


def l21_band_sum(v_local):
    """Band-axis sum of the dual variable (prox/prox_21m.py:123-135).

    `v_local` has shape (nband_local, nbasis, ny, nx); returns the (nbasis, ny, nx) sum over ALL
    bands of all ranks."""
    s = np.ascontiguousarray(v_local.sum(axis=0))
    return allreduce_sum(s)


class BandShardedHessian:
    """(nband, nx, ny) LinearOperator over band-sharded ``BandHessian`` operators: each rank applies
    its own bands; ``dot`` returns the local bands only (no gather — callers reduce what they need),
    ``dot_full`` assembles the full cube on every rank with one all-reduce (for drop-in use)."""

    def __init__(self, nband, make_band):
        self.nband = nband
        self.bands = local_bands(nband)
        self.ops = {b: make_band(b) for b in self.bands}

    def dot(self, x_local):
        return np.stack([self.ops[b].dot(x_local[i]) for i, b in enumerate(self.bands)]) if self.bands else x_local

    def dot_full(self, x):
        out = np.zeros_like(x)
        for b in self.bands:
            out[b] = self.ops[b].dot(x[b])
        return allreduce_sum(out)

    def close(self):
        for o in self.ops.values():
            o.close()
