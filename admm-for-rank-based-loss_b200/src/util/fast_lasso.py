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


class FistaSession:
    """Extension for a caller that loops: X is uploaded once (and, in Gram mode, G = X^T X built once), then any number
    of FISTA solves run on it — `FistaSession(X, dtype)(beta, y, lam, L, eta, ...)` returns what `FISTA(beta, X, y, lam,
    L, eta, ...)` returns, bit for bit.  (The reference's own loop calls FISTA(w, D, b, ...) on the same D every ADMM
    iteration, algorithms.py:199-201.)  Use as a context manager or call close()."""

    def __init__(self, X, dtype=torch.float32):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError(f"dtype must be torch.float32 or torch.float64 (got {dtype})")
        X = np.asarray(X, dtype=np.float64)
        n = X.shape[0]
        self.f32 = dtype == torch.float32
        # D = -(-1) * X = X exactly
        self.engine = AdmmEngine(X, -np.ones(n), "binary_cross_entropy", np.ones(n) / n,
                                 storage="fp32" if self.f32 else "fp64")

    def __call__(self, beta, y, lam, L, eta, tol=1e-4, max_iter=5000, return_info=False):
        eng = self.engine
        if eng is None:
            raise RuntimeError("FistaSession is closed")
        b = eng.vec(np.asarray(y, dtype=np.float64))
        w0 = eng.vec(np.asarray(beta, dtype=np.float64))
        w, info = eng.fista(w0, b, lam, L=L, eta=eta, tol=tol, max_iter=max_iter)
        out = w.cpu().numpy().astype(np.float32) if self.f32 else w.cpu().numpy()
        return (out, info) if return_info else out

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def FISTA(beta, X, y, lam, L, eta, tol=1e-4, max_iter=5000, dtype=torch.float32, return_info=False):
    with FistaSession(X, dtype) as session:
        return session(beta, y, lam, L, eta, tol=tol, max_iter=max_iter, return_info=return_info)
