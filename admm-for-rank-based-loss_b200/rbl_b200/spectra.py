"""Rank weights (spectra) in ascending-rank order — src/optim/objective.py:97-187.

Built on the host in float64 with the reference's formulas (they are O(n) setup, not hot path);
the CPT weights are vectorised instead of the reference's per-element Python loops (:153-164)."""
import math

import numpy as np


def get_erm_weights(n):  # objective.py:97-98
    return np.ones(n, dtype=np.float64) / n


def get_extremile_weights(n, r):  # :101-105
    i = np.arange(n, dtype=np.float64)
    return ((i + 1) ** r - i ** r) / (n ** r)


def get_superquantile_weights(n, q):  # :108-117
    w = np.zeros(n, dtype=np.float64)
    idx = math.floor(n * q)
    frac = 1 - (n - idx - 1) / (n * (1 - q))
    if frac > 1e-12:
        w[idx] = frac
        w[(idx + 1):] = 1 / (n * (1 - q))
    else:
        w[idx:] = 1 / (n - idx)
    return w


def get_esrm_weights(n, rho):  # :120-123
    i = np.arange(n, dtype=np.float64)
    upper = np.exp(rho * ((i + 1) / n))
    lower = np.exp(rho * (i / n))
    return math.exp(-rho) * (upper - lower) / (1 - math.exp(-rho))


def get_aorr_weights(n, qlow, qup):  # :126-136
    w = np.zeros(n, dtype=np.float64)
    lo = math.floor(n * qlow)
    up = math.floor(n * qup)
    frac = 1 - (up - lo - 1) / (n * (qup - qlow))
    if frac > 1e-12:
        w[lo] = frac
        w[(lo + 1):up] = 1 / (n * (qup - qlow))
    else:
        w[lo:up] = 1 / (up - lo)
    return w


def get_aorr_dc_weights(n, k, m):  # :139-145
    if k <= m:
        raise ValueError("need args[0] > args[1]!")
    w = np.zeros(n, dtype=np.float64)
    w[m + 1:k] = 1 / (k - m)
    w[k + 1] = 1 - (k - m - 1) / (k - m)
    return w


def distort(p, gamma):  # :148-150
    return p ** gamma / ((p ** gamma + (1 - p) ** gamma) ** (1 / gamma))


def _cpt(n, which):
    """the reference loops over Python floats (:153-164): the differences distort((i+1)/n) - distort(i/n) cancel
    ~log10(n) digits, so the last bit of every pow matters — a vectorised numpy pow is 3e-9 off in sigma at n = 4M.
    librbl_b200 makes the same scalar libm calls in the same order (host code, rbl_cpt_weights): bit-identical."""
    import ctypes

    from . import _cabi

    out = np.empty(int(n), dtype=np.float64)
    _cabi.check(_cabi.load().rbl_cpt_weights(int(n), which, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    return out


def get_cpt_weights_a(n):  # :153-157
    return _cpt(n, 0)


def get_cpt_weights_b(n):  # :160-164
    return _cpt(n, 1)


def get_weights(name, args=None):  # :166-187
    if name == "erm":
        return get_erm_weights
    elif name == "ehrm":
        return get_cpt_weights_a, get_cpt_weights_b
    elif args is None:
        raise ValueError("args for framework is None!")
    else:
        if name == "extremile":
            return lambda n: get_extremile_weights(n, args[0])
        elif name == "superquantile":
            return lambda n: get_superquantile_weights(n, args[0])
        elif name == "esrm":
            return lambda n: get_esrm_weights(n, args[0])
        elif name == "aorr":
            return lambda n: get_aorr_weights(n, args[0], args[1])
        elif name == "aorr_dc":
            return lambda n: get_aorr_dc_weights(n, args[0], args[1])
        else:
            raise ValueError(
                f"Unrecognized framework '{name}'! Options: ['erm','extremile','superquantile','esrm','aorr','aorr_dc','ehrm']"
            )
