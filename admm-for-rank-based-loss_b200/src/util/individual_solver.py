"""Drop-in for src/util/individual_solver.py: individual_solver(loss, sigma_array, rho, m_array) —
element-wise prox  argmin_z sigma*loss(z) + rho/2 (z-m)^2  on the B200 (rbl_prox_elementwise):
bracketed Newton to machine precision for BCE (reference: damped vector Newton with a global stop,
:90-109), closed form for hinge (reference: early-exit bisection, :15-42)."""
import numpy as np

from rbl_b200 import _cabi
from rbl_b200.engine import AdmmEngine, LOSS_IDS


def individual_solver(loss, sigma_array, rho, m_array):
    if loss == "multinomial_cross_entropy":
        return None  # `pass` in the reference (:124-125)
    if loss not in LOSS_IDS:
        raise ValueError(
            f"Unrecognized loss '{loss}'! Options: ['binary_cross_entropy', 'multinomial_cross_entropy','hinge']")
    sigma = np.ascontiguousarray(sigma_array, dtype=np.float64).reshape(-1)
    m = np.ascontiguousarray(m_array, dtype=np.float64).reshape(-1)
    n = m.size
    eng = AdmmEngine(np.zeros((n, 2)), np.ones(n), loss, np.ones(n) / n)
    try:
        sd, md = eng.vec(sigma), eng.vec(m)
        _cabi.check(eng.lib.rbl_prox_elementwise(eng.h, LOSS_IDS[loss], sd.data_ptr(), md.data_ptr(), n, float(rho),
                                                 eng.z_sorted.data_ptr(), eng._stream()))
        out = eng.z_sorted.cpu().numpy()
    finally:
        eng.close()
    return out.reshape(np.asarray(m_array).shape)
