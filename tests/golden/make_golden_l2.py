"""Golden vectors for the l2 (Student-t) re-weighting step of image_data_products
(/root/reference/src/pfb_imaging/operators/gridder.py:509-532), made by EXECUTING the reference's own
statements in this container: the `if l2_reweight_dof:` block is cut out of the function's AST at
generation time and run in a scratch namespace (nothing is copied into the repo).

  l2_reweight.npz  <- cases (ncorr=1): complex128 / complex64 residuals, with and without prior
                      weights `wgtp`, 7 % flagged samples, dof = 2 and 5; the all-zero residual case
                      (weights become None)

Usage:  python tests/golden/make_golden_l2.py
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/src/pfb_imaging/operators/gridder.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def l2_block():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "image_data_products")
    blk = [n for n in fn.body if isinstance(n, ast.If) and isinstance(n.test, ast.Name) and n.test.id == "l2_reweight_dof"]
    blk = [n for n in blk if any(isinstance(m, ast.Assign) and getattr(m.targets[0], "id", "") == "ressq" for m in n.body)]
    assert len(blk) == 1
    return compile(ast.Module(body=blk, type_ignores=[]), REF, "exec")


def run(code, residual_vis, wgt, mask, dof, wgtp=None):
    ns = {"np": np, "residual_vis": residual_vis, "wgt": wgt.copy(), "mask": mask, "l2_reweight_dof": dof}
    if wgtp is None:
        ns["dsp"] = None
    else:  # the reference reads the prior weights from a dataset: stand-in with the same attribute path
        ns["dsp"] = "x"
        ns["xds_from_list"] = lambda lst, drop_all_but=None: [types.SimpleNamespace(WEIGHT=types.SimpleNamespace(values=wgtp))]
    exec(code, ns)
    return ns["wgt"]


def main():
    code = l2_block()
    rng = np.random.default_rng(20261018)
    nrow, nchan = 301, 5
    out = {}
    k = 0
    for cdt, rdt in ((np.complex128, np.float64), (np.complex64, np.float32)):
        for use_p in (False, True):
            for dof in (2.0, 5.0):
                rv = (rng.standard_normal((1, nrow, nchan)) + 1j * rng.standard_normal((1, nrow, nchan))).astype(cdt)
                rv[0, ::17] *= 8.0  # outliers the re-weighting is there to suppress
                wgt = rng.uniform(0.5, 1.5, (1, nrow, nchan)).astype(rdt)
                mask = (rng.uniform(size=(nrow, nchan)) > 0.07).astype(np.uint8)
                wgtp = rng.uniform(0.5, 1.5, (1, nrow, nchan)).astype(rdt) if use_p else None
                res = run(code, rv, wgt, mask, dof, wgtp)
                out[f"rv_{k}"], out[f"wgt_{k}"], out[f"mask_{k}"], out[f"dof_{k}"] = rv, wgt, mask, dof
                out[f"wgtp_{k}"] = wgtp if use_p else np.zeros(0)
                out[f"out_{k}"] = res
                k += 1
    out["ncase"] = k
    rv = np.zeros((1, 8, 3), np.complex128)
    res = run(code, rv, np.ones((1, 8, 3)), np.ones((8, 3), np.uint8), 2.0)
    assert res is None
    out["zero_is_none"] = True
    np.savez_compressed(os.path.join(HERE, "l2_reweight.npz"), **out)
    print("wrote l2_reweight.npz", k, "cases")


if __name__ == "__main__":
    main()
