"""Generate tests/golden/solvers.npz by EXECUTING the reference's own solver functions (build container only;
/root/reference does not exist on the GPU box).

  python tests/golden/make_golden_solvers.py

The function bodies are cut out of the reference files' ASTs at generation time and executed with numpy / numba
(their modules import ray, dask, numexpr and ducc0, which are absent here); nothing is copied into this repo:
  opt/pcg.py            _nb_fused_alpha_update, _nb_fused_beta_update, _nb_norm_diff, pcg_numba (:23-199), pcg (:202-314)
  opt/power_method.py   _nb_vdot_pair, _nb_normalize, power_method_numba (:14-92), power_method
  operators/hessian.py  HessianTree.dot (:439-522) with ducc0.fft.r2c / c2r replaced by numpy.fft (same maths)
Recorded: the iterates after a FIXED number of iterations (maxit = minit, tol = 0) of an SPD operator, so that the
comparison pins the recurrence itself rather than a converged solution.
"""
import ast
import os
import sys
import time as _time

import numpy as np

R = "/root/reference/src/pfb_imaging"
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def extract(path, names, ns):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = ast.get_source_segment(src, node)
            # decorators are part of the segment only for FunctionDef via decorator_list lines above: rebuild
            deco = "".join("@" + ast.get_source_segment(src, d) + "\n" for d in getattr(node, "decorator_list", []))
            exec(compile(deco + code, path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing


def spd(n, seed):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.geomspace(1.0, 40.0, n)
    return (q * lam) @ q.T


def main():
    import numba
    from numba import njit, prange

    ns = dict(np=np, njit=njit, prange=prange, time=_time.time,
              _FAST_JIT={"nogil": True, "cache": False, "parallel": True, "fastmath": True})

    def norm_diff(x, xp):
        return np.linalg.norm(x - xp) / np.linalg.norm(x)

    ns["norm_diff"] = norm_diff

    class _Numexpr:  # numexpr is absent: the three element-wise expressions of pcg are evaluated by numpy
        @staticmethod
        def evaluate(expr, out=None, local_dict=None, casting=None):
            out[...] = eval(expr, {"__builtins__": {}}, dict(local_dict))
            return out

    ns["ne"] = _Numexpr
    extract(R + "/opt/pcg.py", ["_nb_fused_alpha_update", "_nb_fused_beta_update", "_nb_norm_diff", "pcg_numba", "pcg"], ns)
    out = {}
    n1, n2 = 12, 16
    a = spd(n1 * n2, 5)
    aop = lambda v: (a @ v.ravel()).reshape(v.shape)  # noqa: E731
    rng = np.random.default_rng(11)
    b = rng.standard_normal((n1, n2))
    x0 = 0.1 * rng.standard_normal((n1, n2))
    out.update(pcg_a=a, pcg_b=b, pcg_x0=x0)
    for k in (1, 2, 5, 12):
        out[f"pcg_numba_k{k}"] = ns["pcg_numba"](aop, b, x0=x0.copy(), tol=0.0, maxit=k, minit=k, verbosity=0)
    try:
        d = np.diag(a).reshape(n1, n2)
        for k in (3, 9):
            out[f"pcg_k{k}"] = ns["pcg"](aop, b, x0=x0.copy(), precond=lambda v: v / d, tol=0.0, maxit=k, minit=k,
                                         verbosity=0, backtrack=False)
    except Exception as e:  # the python-loop variant leans on numexpr
        print("pcg (python loop) not executed:", e)
    xs, rs = ns["pcg_numba"](aop, b, x0=None, tol=1e-9, maxit=400, minit=1, verbosity=0, return_resid=True)
    out.update(pcg_numba_conv=xs, pcg_numba_conv_resid=rs)

    class _Log:
        def info(self, *a, **k):
            pass

    import scipy.linalg

    ns2 = dict(np=np, njit=njit, prange=prange, time=_time.time, norm=scipy.linalg.norm, log=_Log(),
               _FAST_JIT=ns["_FAST_JIT"])
    extract(R + "/opt/power_method.py", ["_nb_vdot_pair", "_nb_normalize", "power_method_numba", "power_method"], ns2)
    b0 = rng.standard_normal((n1, n2))
    out["pm_b0"] = b0
    for k in (1, 4, 25):
        beta, bv = ns2["power_method_numba"](aop, (n1, n2), b0=b0.copy(), tol=0.0, maxit=k, verbosity=0)
        out[f"pm_numba_beta_k{k}"], out[f"pm_numba_b_k{k}"] = beta, bv
    beta, bv = ns2["power_method"](aop, (n1, n2), b0=b0.copy(), tol=1e-10, maxit=3000, verbosity=0)
    out["pm_beta_conv"], out["pm_b_conv"] = beta, bv

    # HessianTree.dot with numpy.fft standing in for ducc0.fft (r2c forward inorm=0, c2r inorm=2 = numpy's irfft2)
    def r2c(x, axes, nthreads, forward, inorm, out):
        out[...] = np.fft.rfft2(x, axes=axes)
        return out

    def c2r(x, axes, forward, out, lastsize, inorm, nthreads, allow_overwriting_input):
        out[...] = np.fft.irfft2(x, s=(x.shape[0], lastsize), axes=axes)
        return out

    ns3 = dict(np=np, r2c=r2c, c2r=c2r, empty_noncritical=lambda shape, dtype: np.empty(shape, dtype=dtype))
    extract(R + "/operators/hessian.py", ["HessianTree"], ns3)
    nx, ny, nxp, nyp, ncorr = 40, 36, 56, 60, 2
    parts = []
    beams = [rng.uniform(0.6, 1.0, (ncorr, nx, ny)), None, None]
    beams[1] = beams[0]  # two partitions share a beam, the third has its own
    beams[2] = rng.uniform(0.5, 1.0, (ncorr, nx, ny))
    for p in range(3):
        psf = np.zeros((ncorr, nxp, nyp))
        psf[:, nxp // 2 - 4: nxp // 2 + 5, nyp // 2 - 4: nyp // 2 + 5] = rng.standard_normal((ncorr, 9, 9)) * 0.1
        psf[:, nxp // 2, nyp // 2] = 3.0 + p
        psfhat = np.abs(np.fft.rfft2(np.fft.ifftshift(psf, axes=(1, 2)), axes=(1, 2)))
        parts.append(dict(psfhat=psfhat, beam=beams[p], wsum=np.array([3.0 + p, 2.0 + p])))
    ht = ns3["HessianTree"](parts, nx, ny, nxp, nyp, eta=0.3, nthreads=1)
    x = rng.standard_normal((ncorr, nx, ny))
    out["ht_x"] = x
    out["ht_dot"] = ht.dot(x)
    out["ht_dot_wsum"] = ns3["HessianTree"](parts, nx, ny, nxp, nyp, eta=0.1, nthreads=1, wsum=17.0).dot(x)
    for p in range(3):
        out[f"ht_psfhat{p}"], out[f"ht_beam{p}"], out[f"ht_wsum{p}"] = parts[p]["psfhat"], parts[p]["beam"], parts[p]["wsum"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "solvers.npz"), **out)
    print("wrote tests/golden/solvers.npz:", sorted(out))


if __name__ == "__main__":
    main()
