"""Drop-in for the reference's src/optim/objective.py: same names, arguments and error behaviour;
the rank-weighted objective is evaluated on the B200 (one pass over D, key-only radix sort of the
margins, fused weighted-loss reduction) through librbl_b200.so."""
import numpy as np
import torch

from rbl_b200 import spectra as _sp
from rbl_b200.engine import DeviceProblem

_LOSSES = ("binary_cross_entropy", "multinomial_cross_entropy", "hinge")


def get_loss(name, n_class=None):  # reference :26-37 — returns the loss NAME here; the kernels take ids
    if name not in _LOSSES:
        raise ValueError(
            f"Unrecognized loss '{name}'! Options: ['binary_cross_entropy', 'multinomial_cross_entropy', 'hinge']"
        )
    if name == "multinomial_cross_entropy":
        # a stub in the reference's solver as well (individual_solver.py:124-125: `pass`)
        raise NotImplementedError("multinomial_cross_entropy has no ADMM z-step in the reference either")
    return name


def _as_torch(w):
    return torch.from_numpy(np.ascontiguousarray(w)) if isinstance(w, np.ndarray) else w


def get_weights(name, args=None):
    """objective.py:166-187 — returns callables n -> torch.float64 tensor, like the reference"""
    res = _sp.get_weights(name, args)
    if isinstance(res, tuple):
        return tuple((lambda n, f=f: torch.from_numpy(f(n))) for f in res)
    return lambda n: torch.from_numpy(res(n))


get_erm_weights = lambda n: torch.from_numpy(_sp.get_erm_weights(n))  # noqa: E731
get_extremile_weights = lambda n, r: torch.from_numpy(_sp.get_extremile_weights(n, r))  # noqa: E731
get_superquantile_weights = lambda n, q: torch.from_numpy(_sp.get_superquantile_weights(n, q))  # noqa: E731
get_esrm_weights = lambda n, rho: torch.from_numpy(_sp.get_esrm_weights(n, rho))  # noqa: E731
get_aorr_weights = lambda n, a, b: torch.from_numpy(_sp.get_aorr_weights(n, a, b))  # noqa: E731
get_aorr_dc_weights = lambda n, k, m: torch.from_numpy(_sp.get_aorr_dc_weights(n, k, m))  # noqa: E731
get_cpt_weights_a = lambda n: torch.from_numpy(_sp.get_cpt_weights_a(n))  # noqa: E731
get_cpt_weights_b = lambda n: torch.from_numpy(_sp.get_cpt_weights_b(n))  # noqa: E731


class rankbasedObjective:
    """reference :39-94.  X, y are torch (or numpy) arrays on the host; D = -y*X is built on the
    device once and kept there (`self.problem`)."""

    def __init__(self, X, y, weight_function="erm", loss="binary_cross_entropy", l2_reg=None, l1_reg=None, B=None,
                 n_class=None, args=None, _problem=None, _n=None):
        self.n, self.d = X.shape
        if _n is not None:
            self.n = _n  # global row count of a row-sharded problem
        wf = get_weights(weight_function, args)
        if isinstance(wf, tuple):
            self.weight_function, self.weight_function2 = wf
            self.alphas = self.weight_function(self.n).reshape(-1, 1)
            self.betas = self.weight_function2(self.n).reshape(-1, 1)
        else:
            self.weight_function = wf
            self.alphas = self.weight_function(self.n).reshape(-1, 1)
            self.betas = self.alphas
        self.loss_name = get_loss(loss, n_class=n_class)
        self.loss = self.loss_name
        if B is not None:
            if loss != "binary_cross_entropy":
                raise ValueError("erhm only can be with the binary_cross_entropy.")
            self.B = torch.tensor(B)
            if self.B > 0:
                self.lossB = self.B + torch.log(1 + torch.exp(-self.B))
            else:
                self.lossB = torch.log(1 + torch.exp(self.B))
        else:
            self.B = None
            self.lossB = None
        self.n_class = n_class
        self.l2_reg = l2_reg
        self.l1_reg = l1_reg
        self.problem = _problem if _problem is not None else DeviceProblem(X, y)
        self._alphas_dev = self.problem.vec(self.alphas)

    def get_arrogate_loss(self, w, include_reg=True):
        """reference :71-87.  The reference's EHRM branch weights both sides of lossB with `alphas`
        (:76), so the value is sum(alphas * sorted losses) for every weight function."""
        wd = self.problem.vec(_as_torch(w))
        risk, w2, w1 = self.problem.objective_terms(wd, self._alphas_dev, self.loss_name)
        if self.l2_reg and include_reg:
            risk += 0.5 * self.l2_reg * w2
        if self.l1_reg and include_reg:
            risk += 0.5 * self.l1_reg * w1
        return risk
