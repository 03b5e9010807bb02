"""Drop-in for src/util/w_LBFGS.py (l2 branch w_flag == 2, smoothed-l1 branch w_flag == 1): scipy L-BFGS-B on the host, the objective
and gradient  f = rho/2 ||D w - b||^2 + reg/2 ||w||^2,  g = rho D^T (D w - b) + reg w  come from ONE
fused pass over D on the B200 (rbl_fused_pass) instead of D@w plus DTD@w - D.T@b (:31-45)."""
import numpy as np

from rbl_b200.engine import AdmmEngine


def w_solver(w_flag, w0, z, lagrangian, rho, DTD, D, reg, t=None):
    if w_flag not in (1, 2):
        raise ValueError("w_flag can only be 1 or 2.")
    if w_flag == 1 and t is None:
        raise ValueError("The smoothess parameter is not given.")
    D = np.asarray(D, dtype=np.float64)
    n, d = D.shape
    eng = AdmmEngine(D, -np.ones(n), "binary_cross_entropy", np.ones(n) / n)
    try:
        eng.set_state(w=np.asarray(w0).reshape(-1))
        eng.b.copy_(eng.vec(np.asarray(z).reshape(-1) + np.asarray(lagrangian).reshape(-1) / rho))
        if w_flag == 2:
            eng.w_step_lbfgs(rho, reg)
        else:
            def huber(w):  # wl1_fun_smooth / wl1_fun_smooth_gradient (:11-28)
                small = np.abs(w) <= t
                R = 0.25 * reg * float(np.sum(np.square(w[small]))) / t + 0.5 * reg * float(np.sum(np.abs(w[~small]) - 0.5 * t))
                return R, np.where(small, 0.5 * reg * w / t, 0.5 * reg * np.sign(w))
            eng.w_step_lbfgs(rho, reg, reg_fg=huber)
        w = eng.w.cpu().numpy().reshape(-1, 1)
    finally:
        eng.close()
    return w
