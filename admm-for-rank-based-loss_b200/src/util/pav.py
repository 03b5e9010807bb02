"""Drop-in for src/util/pav.py: PAV_solver(sigma, m_sorted, rho, loss, B, sigma2).get_opt(maxiter).

The isotonic prox runs on the B200 (element prox + tree-merge PAV, csrc/pav_core.h).  `get_opt`
returns ONE array like the reference's (pav.py:178) — note the reference's own caller unpacks two
values (algorithms.py:97,101) and therefore crashes as shipped; our algorithms.py does not."""
import time

import numpy as np

from rbl_b200.engine import AdmmEngine, LOSS_IDS
from rbl_b200 import _cabi


def _device_pav(sigma, m_sorted, rho, loss, clip=None):
    sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    m = np.ascontiguousarray(m_sorted, dtype=np.float64).reshape(-1)
    n = m.size
    eng = AdmmEngine(np.zeros((n, 2)), np.ones(n), loss, sigma, clip=clip)
    try:
        eng.m_sorted.copy_(eng.vec(m))
        _cabi.check(eng.lib.rbl_pav_prox(eng.h, LOSS_IDS[loss], eng.m_sorted.data_ptr(), float(rho),
                                         eng.z_sorted.data_ptr(), eng._stream()))
        z = eng.z_sorted.cpu().numpy()
    finally:
        eng.close()
    if clip is not None:
        z = np.maximum(z, clip)
    return z


class PAV_solver(object):
    def __init__(self, sigma_array, m_array, rho, loss="binary_cross_entropy", B=None, sigma_array2=None):
        if loss not in LOSS_IDS:
            raise ValueError(
                f"Unrecognized loss '{loss}'! Options: ['binary_cross_entropy', 'multinomial_cross_entropy','hinge']")
        if B is not None and loss != "binary_cross_entropy":
            raise ValueError("ehrm only can be with binary_cross_entropy.")
        self.time = time.time()
        self.rho, self.loss, self.B = rho, loss, B
        self.n = sigma_array.shape[0]
        self._sigma = sigma_array if B is None else sigma_array2
        self._m = m_array

    def get_opt(self, maxiter=10000):
        res = _device_pav(self._sigma, self._m, self.rho, self.loss, clip=self.B)
        self.time = time.time() - self.time
        return res
