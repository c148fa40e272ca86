"""CPU: world_size-2 gloo test of the band sharding / cross-band reductions (host logic of the
multi-GPU path; the gridder itself needs no collective)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, os.environ["PFBG_ROOT"])
    from pfb_imaging_b200 import dist
    dist.init("gloo")
    r, w = dist.rank(), dist.world_size()
    assert w == 2
    nband = 5
    mine = dist.local_bands(nband)
    assert mine == [b for b in range(nband) if b % 2 == r]
    assert all(dist.band_owner(b) == b % 2 for b in range(nband))
    rng = np.random.default_rng(0)            # same stream on both ranks
    cube = rng.standard_normal((nband, 3, 8, 6))
    # L21 band sum (prox_21m.py:123-135)
    s = dist.l21_band_sum(cube[mine])
    assert np.allclose(s, cube.sum(axis=0), atol=1e-12)
    # pcg scalars (pcg.py:35-41): r.y and p.Ap over the band-sharded cube
    a, b = rng.standard_normal((2, nband, 8, 6))
    got = dist.vdot_allreduce((a[mine], b[mine]), (a[mine], a[mine]))
    assert np.allclose(got, [np.vdot(a, b), np.vdot(a, a)])
    # positivity mode 2 (positivity.py:22-32): min over bands
    img = rng.standard_normal((nband, 8, 6))
    m = dist.allreduce_min(img[mine].min(axis=0))
    assert np.array_equal(m, img.min(axis=0))
    # float32 and non-contiguous inputs
    f = np.asfortranarray(rng.standard_normal((4, 4)).astype(np.float32))
    g = dist.allreduce_sum(f.copy(order="F"))
    assert np.allclose(g, 2 * f, atol=1e-6)
    # band-sharded operator protocol with a stand-in band operator
    class Op:
        def __init__(self, b): self.b = b
        def dot(self, x): return (self.b + 1.0) * x
        def close(self): pass
    H = dist.BandShardedHessian(nband, Op)
    x = rng.standard_normal((nband, 8, 6))
    full = H.dot_full(x)
    assert np.allclose(full, (np.arange(nband) + 1.0)[:, None, None] * x)
    loc = H.dot(x[mine])
    assert np.allclose(loc, full[mine])
    print("RANK", r, "OK")
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_band_sharding_and_reductions_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PFBG_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RANK {r} OK" in o


def test_single_process_is_identity():
    from pfb_imaging_b200 import dist

    a = np.arange(6.0).reshape(2, 3)
    assert dist.world_size() == 1 and dist.rank() == 0
    assert dist.allreduce_sum(a) is a
    assert dist.local_bands(4) == [0, 1, 2, 3]
    assert np.allclose(dist.vdot_allreduce((a, a)), [np.vdot(a, a)])
