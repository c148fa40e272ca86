"""Band-sharded primal-dual == single-rank primal-dual (NCCL all-reduces of the l21 band sum and the norms).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sara_sharded_check.py

Every rank solves the full 4-band problem locally (reference) and its own shard with the cross-band hooks;
the shard must equal the corresponding bands of the full solution."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfb_imaging_b200 import dist  # noqa: E402
from pfb_imaging_b200.sara import L21, PrimalDual, PsiNocopyt  # noqa: E402


def main():
    dist.init()
    r, world = dist.rank(), dist.world_size()
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    nband, nx, ny, nlevel = 4, 256, 192, 2
    bases = ["self", "db1", "db2", "db3"]
    rng = np.random.default_rng(7)
    y = 0.1 * rng.standard_normal((nband, nx, ny))
    y[:, 100:120, 50:90] += np.linspace(1.0, 0.4, nband)[:, None, None]
    w = rng.uniform(0.5, 1.5, (len(bases),))

    def solve(bands, hooks, positivity):
        psi = PsiNocopyt(len(bands), nx, ny, bases, nlevel, 1, device=dev)
        reg = L21(psi, bases, nu=len(bases))
        reg.l1weight = reg.l1weight * w[:, None, None]
        yb = y[bands]
        pd = PrimalDual(tol=1e-14, maxit=30, verbosity=0, positivity=positivity, **hooks)
        pd.setup(reg, 1.0)
        pd.set_grad(lambda xx: xx - yb)
        return pd.solve(np.zeros_like(yb), 0.03)

    worst = 0.0
    for positivity in (0, 1, 2):
        full = solve(list(range(nband)), {}, positivity)
        mine = dist.local_bands(nband)
        shard = solve(mine, dict(reduce_tensor=dist.allreduce_sum, reduce_scalars=dist.allreduce_sum), positivity)
        worst = max(worst, float(np.abs(shard - full[mine]).max()))
    t = torch.tensor([worst], device=torch.device("cuda", dev), dtype=torch.float64)
    dist.allreduce_max(t)
    if r == 0:
        print(f"sara sharded check: world={world}, max |sharded - single| = {float(t.item()):.3e}")
        assert float(t.item()) < 1e-11
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
