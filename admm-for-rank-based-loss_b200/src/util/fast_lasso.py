"""Drop-in for src/util/fast_lasso.py: FISTA(beta, X, y, lam, L, eta, tol, max_iter, dtype) on the B200.

Same argument meaning as the reference (:22); the backtracking loop (:40-67) runs as a device-resident
state machine with one fused pass over X per line-search trial (librbl_b200: rbl_fista_*).
`dtype=torch.float64`: everything in float64.  `dtype=torch.float32` (the reference's default, :22): the
OPTIONAL fp32 mode — X is kept in float32 in HBM (half the bytes per pass) while every product, sum and the
FISTA state stay float64, and the result is returned as float32 like the reference's (:69).  That is tighter
than the reference's all-float32 arithmetic (chaotic at its own tolerance, overflowing L_cur: SURVEY.md §8a
vi-vii), by design.  The global torch default dtype is NOT mutated (:23-26 side effect not reproduced)."""
import numpy as np
import torch

from rbl_b200.engine import AdmmEngine


def soft_thr(x, alpha):  # reference :15-19 (host helper kept for API parity)
    return torch.maximum(torch.abs(x) - alpha, torch.zeros_like(x)) * torch.sign(x)


def FISTA(beta, X, y, lam, L, eta, tol=1e-4, max_iter=5000, dtype=torch.float32, return_info=False):
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    # D = -(-1) * X = X exactly
    if dtype not in (torch.float32, torch.float64):
        raise ValueError(f"dtype must be torch.float32 or torch.float64 (got {dtype})")
    f32 = dtype == torch.float32
    eng = AdmmEngine(X, -np.ones(n), "binary_cross_entropy", np.ones(n) / n, storage="fp32" if f32 else "fp64")
    try:
        b = eng.vec(np.asarray(y, dtype=np.float64))
        w0 = eng.vec(np.asarray(beta, dtype=np.float64))
        w, info = eng.fista(w0, b, lam, L=L, eta=eta, tol=tol, max_iter=max_iter)
        out = w.cpu().numpy().astype(np.float32) if f32 else w.cpu().numpy()
    finally:
        eng.close()
    return (out, info) if return_info else out
