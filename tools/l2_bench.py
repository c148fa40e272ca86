"""l2 re-weighting kernels (pfbg_l2_reweight) on device-resident arrays of one C2 band: ms per call and GB/s of the
algorithmic traffic (pass 1: 2p + 1 B per sample, pass 2: 2p + 2p B; p = 4).  usage: python tools/l2_bench.py [nvis]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pfb_imaging_b200 import _lib

nvis = int(sys.argv[1]) if len(sys.argv) > 1 else 24998400
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rv = torch.randn((1, nvis, 2), device=dev, dtype=torch.float32, generator=g)
wgt = torch.rand((1, nvis), device=dev, dtype=torch.float32, generator=g) + 0.5
mask = (torch.rand((nvis,), device=dev, generator=g) > 0.05).to(torch.uint8)
lib = _lib.load()
ovar = (C.c_double * 1)()
applied = C.c_int32(0)
stream = torch.cuda.current_stream().cuda_stream


def call():
    _lib.check(lib.pfbg_l2_reweight(_lib.PFBG_F32, 0, rv.data_ptr(), None, mask.data_ptr(), wgt.data_ptr(), nvis, 1, 5.0,
                                    C.cast(ovar, C.c_void_p), C.cast(C.byref(applied), C.c_void_p), _lib.DEVICE_PTRS,
                                    stream))


for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
ref = float(((rv[0, :, 0] ** 2 + rv[0, :, 1] ** 2).double() * mask).sum() / mask.sum())
print(f"nvis {nvis}: {ms:.3f} ms per call (two kernels + one 136-byte read-back), {nvis * 25 / ms / 1e6:.0f} GB/s algorithmic; "
      f"ovar {ovar[0]:.9f} vs torch fp64 {ref:.9f}, applied {applied.value}")
