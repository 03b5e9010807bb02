"""OPTIONAL fp32 storage mode (north_star: "1e-5 in an optional fp32 mode"; the reference ships a float32 FISTA,
algorithms.py:199-201): D is kept in float32 in HBM, every product and sum stays fp64.  The mode is therefore EXACTLY
the fp64 algorithm on the design matrix rounded to float32 — tested as such (1e-9 against the fp64 oracle fed the
rounded matrix) — and within 1e-5 of the fp64 run on the unrounded one."""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rbl_oracle as O  # noqa: E402


def _rel(a, b):
    return np.linalg.norm(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n,d", [(1000, 200), (777, 201), (6000, 1000), (300, 41), (50, 7), (3000, 1500), (257, 5000),
                                 (100003, 64)])
def test_fp32_storage_kernels(n, d):
    """matvec, fused pass, Gram matrix, active-row gather, transposed sparse dual pass on float32 rows"""
    import ctypes

    from rbl_b200 import _cabi
    from rbl_b200.engine import AdmmEngine

    rng = np.random.default_rng(n * 13 + d)
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    e = AdmmEngine(X, y, "binary_cross_entropy", np.ones(n) / n, storage="fp32")
    assert e.D.dtype == torch.float32 and e.ld % 4 == 0
    D = (-y[:, None] * X).astype(np.float32).astype(np.float64)          # what the mode stores
    np.testing.assert_array_equal(e.D[:, :d].cpu().numpy().astype(np.float64), D)
    assert float(e.D[:, d:].abs().sum()) == 0.0
    x, b = rng.normal(size=d), rng.normal(size=n)
    xd, bd = e.vec(x), e.vec(b)
    ref = D @ x
    scale = np.abs(D) @ np.abs(x)
    assert np.max(np.abs(e.matvec(xd).cpu().numpy() - ref) / scale) < 1e-14      # fp64 arithmetic on the fp32 rows
    _cabi.check(e.lib.rbl_fused_pass(e.h, e.D.data_ptr(), xd.data_ptr(), bd.data_ptr(), e.r.data_ptr(),
                                     e.red.data_ptr(), e._stream()))
    rref = b - ref
    red = e.red.cpu().numpy()
    assert np.max(np.abs(red[:d] - D.T @ rref) / (np.abs(D).T @ np.abs(rref))) < 1e-13
    assert abs(red[d] - rref @ rref) < 1e-13 * (rref @ rref)
    if d <= 4096 and n >= 2 * d:
        G = e.gram().cpu().numpy()[:, :d]
        Gref = D.T @ D
        assert np.max(np.abs(G - Gref)) < 1e-12 * np.max(np.abs(Gref))
    e.close()


@pytest.mark.parametrize("w_mode", ["gram", "stream"])
@pytest.mark.parametrize("tag,n,d,wf,args,loss,B,kw", [
    ("C2 twin", 6000, 100, "superquantile", [0.8], "binary_cross_entropy", None, dict(l1_reg=0.01)),
    ("C2 twin l2", 6000, 101, "superquantile", [0.8], "binary_cross_entropy", None, dict(l2_reg=0.01)),
    ("C3 twin", 6000, 50, "ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01)),
    ("C4 twin hinge", 2400, 201, "aorr", [0.2, 0.8], "hinge", None, dict(l2_reg=1e-4)),
])
def test_fp32_mode_is_the_fp64_algorithm_on_the_rounded_matrix(tag, n, d, wf, args, loss, B, kw, w_mode, monkeypatch):
    from src.optim.algorithms import ADMMmethod, Optimizer

    monkeypatch.setenv("RBL_W_MODE", w_mode)
    rng = np.random.default_rng(len(tag) * 77 + n)
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:6] = rng.normal(size=6)
    y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
    X32 = X.astype(np.float32).astype(np.float64)
    s = ADMMmethod(X, y, wf, loss, B=B, args=args, max_iter=30, tol=1e-9, _storage="fp32", **kw)
    assert s.engine.storage == "fp32" and s.engine.D.dtype == torch.float32
    o = O.OracleADMM(X32, y, wf, loss, B=B, args=args, max_iter=30, tol=1e-9, **kw)
    for i in range(30):
        o.w, o.z, o.lam, o.rho = s.w.reshape(-1).copy(), s.z.reshape(-1).copy(), s.lagrangian.reshape(-1).copy(), s.rho
        with contextlib.redirect_stdout(io.StringIO()):
            Optimizer.main_loop(s, i, 0.0, False)
        o.step()
        assert _rel(s.z, o.z) < 1e-9 and _rel(s.w, o.w) < 1e-9, (tag, i, _rel(s.w, o.w), _rel(s.z, o.z))
    s.engine.close()
    # against the fp64 run on the unrounded matrix: objective within 1e-5 (north_star's fp32 tolerance)
    a = ADMMmethod(X, y, wf, loss, B=B, args=args, max_iter=60, tol=1e-9, **kw)
    b = ADMMmethod(X, y, wf, loss, B=B, args=args, max_iter=60, tol=1e-9, _storage="fp32", **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        wa, wb = a.main_loop(verbose=False), b.main_loop(verbose=False)
        oa = a.objective.get_arrogate_loss(torch.from_numpy(wa).double())
        ob_on_exact = a.objective.get_arrogate_loss(torch.from_numpy(wb).double())   # fp32-mode w, fp64 objective
    assert abs(oa - ob_on_exact) <= 1e-5 * abs(oa), (tag, oa, ob_on_exact)
    assert _rel(wb, wa) < 1e-4, (tag, _rel(wb, wa))
    a.engine.close()
    b.engine.close()


def test_fista_mirror_honours_dtype(golden_dir):
    """FISTA(..., dtype=torch.float32) — the reference's default (fast_lasso.py:22): float32 storage, float32
    result; within the reference's own float32 scatter of its float32 output, and within 1e-6 of the float64 call."""
    import os

    from src.util.fast_lasso import FISTA

    g = np.load(os.path.join(golden_dir, "fista.npz"))
    d = np.load(os.path.join(golden_dir, "data_300x40.npz"))
    D = -d["y"] * d["X"]
    for lam in (0.5, 20.0, 300.0):
        w64 = FISTA(g["w0"], D, g["b"], lam, np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000, dtype=torch.float64)
        w32 = FISTA(g["w0"], D, g["b"], lam, np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000)
        assert w32.dtype == np.float32 and w64.dtype == np.float64
        assert np.linalg.norm(w32 - w64) <= 1e-5 * max(np.linalg.norm(w64), 1e-3), lam
        ref32 = g[f"ref_f32_{lam}"]
        assert np.linalg.norm(w32 - ref32) <= 2e-2 * np.linalg.norm(ref32) + 1e-6
    with pytest.raises(ValueError):
        FISTA(g["w0"], D, g["b"], 1.0, np.float32(17), np.float32(2.5), dtype=torch.float16)
