"""TEST INFRASTRUCTURE ONLY — fast CPU restatement: oracle/wgridder_np.py with the
per-visibility loops in C/OpenMP (oracle/cwgridder.c) and the plane FFTs on scipy.fft
(pocketfft, threaded).  This is the ``cpu_baseline`` / ``--impl reference`` arm of bench.py
("port": ducc0 itself cannot be installed here) and a checker for mid-size tests."""

import concurrent.futures as cf
import ctypes as C
import os
import threading
import subprocess
import time

import numpy as np
import scipy.fft as sfft

from . import wgridder_np as wg

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libcwgridder.so")
        src = os.path.join(HERE, "cwgridder.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(so)
        _lib.cw_num_threads.restype = C.c_int
    return _lib


def nthreads():
    return lib().cw_num_threads()


def set_threads(n=None):
    """Use `n` OpenMP threads (default: every core this process may run on), whatever OMP_NUM_THREADS says."""
    lib().cw_set_threads(int(n or _ncores()))
    return nthreads()


def _ncores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _fft_workers(P):
    return max(1, _ncores() // max(1, min(P, _ncores())))


def _run_planes(fn, P):
    """Planes are independent: run them on a thread pool (numpy/scipy release the GIL)."""
    nt = min(P, _ncores())
    if nt <= 1:
        for p in range(P):
            fn(p)
        return
    with cf.ThreadPoolExecutor(max_workers=nt) as ex:
        list(ex.map(fn, range(P)))


def _crop(f, plan):
    """Image pixels sit in the four corners of the (nu, nv) transform (i' = i - nx/2 mod nu)."""
    hx, hy = plan.nx // 2, plan.ny // 2
    out = np.empty((plan.nx, plan.ny), dtype=f.dtype)
    out[:hx, :hy] = f[-hx:, -hy:]
    out[:hx, hy:] = f[-hx:, :hy]
    out[hx:, :hy] = f[:hx, -hy:]
    out[hx:, hy:] = f[:hx, :hy]
    return out


def _pad(val, plan):
    hx, hy = plan.nx // 2, plan.ny // 2
    pl = np.zeros((plan.nu, plan.nv), dtype=val.dtype)
    pl[-hx:, -hy:] = val[:hx, :hy]
    pl[-hx:, :hy] = val[:hx, hy:]
    pl[:hx, -hy:] = val[hx:, :hy]
    pl[:hx, :hy] = val[hx:, hy:]
    return pl


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _bin(plan, uvw, freq, mask):
    b = wg.bin_indices(plan, uvw, freq, mask)
    for k in ("gu", "gv", "gw", "iu0", "iv0", "ip0", "key", "order"):
        b[k] = np.ascontiguousarray(b[k])
    return b


def vis2dirty_c(plan, uvw, freq, vis, wgt=None, mask=None, b=None, timings=None):
    t0 = time.perf_counter()
    b = _bin(plan, uvw, freq, mask) if b is None else b
    W, P, nu, nv = plan.W, plan.nplanes, plan.nu, plan.nv
    a = np.asarray(vis).astype(np.complex128).ravel()[b["idx"]]
    if wgt is not None:
        a = a * np.asarray(wgt, dtype=np.float64).ravel()[b["idx"]]
    a = np.where(b["conj"], np.conj(a), a)  # Hermitian fold (bin_indices)
    a = np.ascontiguousarray(a * np.exp(2j * np.pi * wg._vis_phase(plan, b)))
    t1 = time.perf_counter()
    grid = np.zeros((P, nu, nv), dtype=np.complex128)
    lib().cw_grid(C.c_int64(a.size), _p(b["order"]), _p(b["gu"]), _p(b["gv"]), _p(b["gw"]), _p(b["iu0"]), _p(b["iv0"]),
                  _p(b["ip0"]), _p(b["key"]), _p(a), W, C.c_double(plan.beta), nu, nv, P, int(plan.do_wgridding), _p(grid))
    t2 = time.perf_counter()
    corr, nu_, ipx, ipy = wg._image_factors(plan)
    img = np.zeros((plan.nx, plan.ny))
    lock = threading.Lock()

    def one(p):
        f = sfft.ifft2(grid[p], workers=_fft_workers(P), norm="forward")
        f = _crop(f, plan)
        if plan.do_wgridding:
            ph = (2.0 * np.pi * (plan.w0 + p * plan.dw)) * nu_
            r = f.real * np.cos(ph) + f.imag * np.sin(ph)  # Re(f e^{-i ph})
        else:
            r = f.real
        with lock:
            np.add(img, r, out=img)

    _run_planes(one, P)
    img *= corr
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(prep=t1 - t0, vis=t2 - t1, planes=t3 - t2)
    return img


def dirty2vis_c(plan, uvw, freq, dirty, mask=None, b=None, timings=None):
    t0 = time.perf_counter()
    b = _bin(plan, uvw, freq, mask) if b is None else b
    W, P, nu, nv = plan.W, plan.nplanes, plan.nu, plan.nv
    t1 = time.perf_counter()
    corr, nu_, ipx, ipy = wg._image_factors(plan)
    x = np.asarray(dirty, dtype=np.float64) * corr
    grid = np.empty((P, nu, nv), dtype=np.complex128)

    def one(p):
        if plan.do_wgridding:
            ph = (2.0 * np.pi * (plan.w0 + p * plan.dw)) * nu_
            val = np.empty(x.shape, dtype=np.complex128)
            np.multiply(x, np.cos(ph), out=val.real)
            np.multiply(x, np.sin(ph), out=val.imag)
        else:
            val = x.astype(np.complex128)
        grid[p] = sfft.fft2(_pad(val, plan), workers=_fft_workers(P), overwrite_x=True)

    _run_planes(one, P)
    t2 = time.perf_counter()
    n = b["idx"].size
    out = np.zeros(n, dtype=np.complex128)
    lib().cw_degrid(C.c_int64(n), _p(b["order"]), _p(b["gu"]), _p(b["gv"]), _p(b["gw"]), _p(b["iu0"]), _p(b["iv0"]),
                    _p(b["ip0"]), W, C.c_double(plan.beta), nu, nv, P, int(plan.do_wgridding), _p(grid), _p(out))
    out *= np.exp(-2j * np.pi * wg._vis_phase(plan, b))
    np.conjugate(out, out=out, where=b["conj"])
    t3 = time.perf_counter()
    nrow, nchan = np.asarray(uvw).shape[0], np.asarray(freq).size
    vis = np.zeros(nrow * nchan, dtype=np.complex128)
    vis[b["idx"]] = out
    if timings is not None:
        timings.update(prep=t1 - t0, planes=t2 - t1, vis=t3 - t2)
    return vis.reshape(nrow, nchan)
