"""Drop-in for src/util/w_LBFGS.py (l2 branch, w_flag == 2): scipy L-BFGS-B on the host, the objective
and gradient  f = rho/2 ||D w - b||^2 + reg/2 ||w||^2,  g = rho D^T (D w - b) + reg w  come from ONE
fused pass over D on the B200 (rbl_fused_pass) instead of D@w plus DTD@w - D.T@b (:31-45)."""
import numpy as np

from rbl_b200.engine import AdmmEngine


def w_solver(w_flag, w0, z, lagrangian, rho, DTD, D, reg, t=None):
    if w_flag != 2:
        if w_flag == 1:
            raise NotImplementedError("smoothed-l1 w-step (smoothADMMmethod) is not part of this path yet")
        raise ValueError("w_flag can only be 1 or 2.")
    D = np.asarray(D, dtype=np.float64)
    n, d = D.shape
    eng = AdmmEngine(D, -np.ones(n), "binary_cross_entropy", np.ones(n) / n)
    try:
        eng.set_state(w=np.asarray(w0).reshape(-1))
        eng.b.copy_(eng.vec(np.asarray(z).reshape(-1) + np.asarray(lagrangian).reshape(-1) / rho))
        eng.w_step_lbfgs(rho, reg)
        w = eng.w.cpu().numpy().reshape(-1, 1)
    finally:
        eng.close()
    return w
