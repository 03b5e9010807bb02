"""Drop-in for src/util/PAV_cpt.py: PAV_solver_CPT(sigma1, sigma2, B, m_sorted, rho).get_opt().

The reference's EHRM z-step forms two candidates, min(prox_{sigma1}(m), B) and max(prox_{sigma2}(m), B), and picks
between them by comparing two SCALARS (PAV_cpt.py:222-226), i.e. all or nothing: candidate 1 when
func_value(sigma1, cand1) <= func_value(sigma2, cand2), else candidate 2 — and pools the winner (:229-288).  Here
the two sums are formed on the B200 (rbl_ehrm_candidate_sums), the winner's isotonic prox runs there too, and the
clip at B is applied to it."""
import numpy as np

from rbl_b200 import _cabi
from rbl_b200.engine import AdmmEngine


class PAV_solver_CPT(object):
    def __init__(self, sigma_array1, sigma_array2, B, m_array, rho, multi=False):
        self.rho, self.B = rho, B
        self._s1 = np.ascontiguousarray(sigma_array1, dtype=np.float64).reshape(-1)
        self._s2 = np.ascontiguousarray(sigma_array2, dtype=np.float64).reshape(-1)
        self._m = np.ascontiguousarray(m_array, dtype=np.float64).reshape(-1)
        self.selected = None   # 1 or 2 after get_opt()
        self.fvals = None      # (fval1, fval2)

    def get_opt(self):
        n = self._m.size
        eng = AdmmEngine(np.zeros((n, 2)), np.ones(n), "binary_cross_entropy", None,
                         ehrm=(self._s1, self._s2, float(self.B)))
        try:
            eng.m_sorted.copy_(eng.vec(self._m))
            eng._ehrm_select(self.rho)
            _cabi.check(eng.lib.rbl_pav_prox(eng.h, 0, eng.m_sorted.data_ptr(), float(self.rho),
                                             eng.z_sorted.data_ptr(), eng._stream()))
            z = eng.z_sorted.cpu().numpy()
            self.selected = 1 if eng.clip_mode == 2 else 2
            self.fvals = eng.ehrm_stats["last"]
        finally:
            eng.close()
        return np.minimum(z, self.B) if self.selected == 1 else np.maximum(z, self.B)
