import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import bench
from pfb_imaging_b200 import wgridder as W
prec = sys.argv[1] if len(sys.argv) > 1 else "single"
nxp = int(sys.argv[2]) if len(sys.argv) > 2 else 5760
cfg = dict(bench.WORKLOADS["c2"], precision=prec)
d, cell, x = bench.make_inputs(cfg, 0)
eps = 1e-5 if prec == "single" else 1e-7
gp = W.plan_for(d["uvw"], d["freq"], npix_x=nxp, npix_y=nxp, pixsize_x=cell, pixsize_y=cell, epsilon=eps, flip_v=True,
                divide_by_n=False, precision=prec, mask=d["mask"], sigma_min=1.1, sigma_max=3.0, device=0)
gp.bind_weights(d["wgt"])
print({k: gp.info()[k] for k in ("nu", "nv", "nplanes", "W", "sigma", "total_bytes")})
dev = torch.device("cuda", 0)
rdt = torch.float32 if prec == "single" else torch.float64
x_d = torch.zeros((nxp, nxp), dtype=rdt, device=dev); x_d[nxp // 2, nxp // 2] = 1
out_d = torch.empty_like(x_d)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2): gp.hessian_dev(x_d.data_ptr(), None, 1.0, 0.0, out_d.data_ptr(), s)
torch.cuda.synchronize()
gp.set_profiling(True)
rec = []
for _ in range(3):
    gp.hessian_dev(x_d.data_ptr(), None, 1.0, 0.0, out_d.data_ptr(), s); torch.cuda.synchronize(); rec.append(gp.timings())
print(prec, nxp, "phases ms", np.round(np.median(np.array(rec), axis=0), 2), "total", round(float(np.median(np.array(rec), axis=0).sum()), 2))
