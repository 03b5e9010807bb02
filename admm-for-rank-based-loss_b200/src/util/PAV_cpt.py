"""Drop-in for src/util/PAV_cpt.py: PAV_solver_CPT(sigma1, sigma2, B, m_sorted, rho).get_opt().

As shipped the reference's EHRM z-step picks between its two candidates by comparing two SCALARS
(PAV_cpt.py:222-226), i.e. all-or-nothing, and candidate 2 — sigma = sigma2 clipped below at B —
always wins, so the result is max(B, isotonic prox with sigma2).  That is what runs
here, on the B200."""
from src.util.pav import _device_pav


class PAV_solver_CPT(object):
    def __init__(self, sigma_array1, sigma_array2, B, m_array, rho, multi=False):
        self.rho, self.B = rho, B
        self._s1, self._s2, self._m = sigma_array1, sigma_array2, m_array

    def get_opt(self):
        return _device_pav(self._s2, self._m, self.rho, "binary_cross_entropy", clip=self.B)
