"""Drop-in for src/util/w_LBFGS.py (l2 branch w_flag == 2, smoothed-l1 branch w_flag == 1): L-BFGS-B with scipy's
control flow (:48-62) — the library's own (rbl_lbfgs_gram) over f/g evaluations on G = D^T D when the engine runs
in Gram mode, scipy's over one fused pass over D per evaluation (rbl_fused_pass) otherwise — for
f = rho/2 ||D w - b||^2 + R(w),  g = rho D^T (D w - b) + R'(w)  (:11-45)."""
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
            eng.w_step_lbfgs(rho, reg, huber_t=float(t))  # wl1_fun_smooth / wl1_fun_smooth_gradient (:11-28)
        w = eng.w.cpu().numpy().reshape(-1, 1)
    finally:
        eng.close()
    return w
