"""One C2 band, device-resident Hessian applies (for ncu / quick timing).  usage: prof_band.py [band] [napply] [workload]"""
import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import bench
from pfb_imaging_b200 import wgridder as W
band = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nap = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"
cfg = bench.WORKLOADS[wl]
d, cell, x = bench.make_inputs(cfg, band)
gp = W.plan_for(d["uvw"], d["freq"], npix_x=cfg["nx"], npix_y=cfg["nx"], pixsize_x=cell, pixsize_y=cell,
                epsilon=cfg["epsilon"], flip_v=True, divide_by_n=False, precision=cfg["precision"],
                mask=d["mask"], sigma_min=1.1, sigma_max=3.0, device=0)
gp.bind_weights(d["wgt"])
dev = torch.device("cuda", 0)
x_d = torch.from_numpy(x).to(dev); out_d = torch.empty_like(x_d)
stream = torch.cuda.current_stream().cuda_stream
wsum = float(d["wgt"].sum(dtype=np.float64))
print(gp.info())
for _ in range(nap):
    gp.hessian_dev(x_d.data_ptr(), None, wsum, 0.0, out_d.data_ptr(), stream)
torch.cuda.synchronize()
gp.set_profiling(True)
rec = []
for _ in range(nap):
    gp.hessian_dev(x_d.data_ptr(), None, wsum, 0.0, out_d.data_ptr(), stream); torch.cuda.synchronize(); rec.append(gp.timings())
print("phases ms", np.round(np.median(np.array(rec), axis=0), 3), "total", round(float(np.median(np.array(rec), axis=0).sum()), 3))
