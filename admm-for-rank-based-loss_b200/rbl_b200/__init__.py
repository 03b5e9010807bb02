"""rbl_b200 — B200-native ADMM hot path for rank-based losses (SRM / EHRM / AoRR).

Host layer over the C ABI of librbl_b200.so (include/rbl_b200.h).  The drop-in entry point the
reference's drivers import is `src.optim.algorithms.ADMMmethod` in this directory."""
from . import _cabi  # noqa: F401
from ._cabi import RblError  # noqa: F401
