"""Data ingest on the B200 — the step before the hot path.

The reference prepares its synthetic data on the host: `make_classification` -> `preprocessing.scale`
(src/util/load_data.py:105-115) -> `train_test_split` (run_SRM.py:26).  The generator is scikit-learn's host RNG
stream and stays on the host; the two O(n d) array passes after it run here on the device, so the matrix crosses
PCIe once and never returns: `standardize_` (`rbl_standardize_columns`) and `split_rows` / `train_test_split_device`
(`rbl_gather_rows`, with the index permutation of scikit-learn's own ShuffleSplit so the split is the same rows).
There is no CPU path: without the library / a B200 these raise."""
import ctypes

import numpy as np
import torch

from . import _cabi
from .engine import _require_cuda


def to_device_padded(X, device=None):
    """host (or device) n x d array -> device float64 tensor n x ld, ld even, padding column zero; returns (Xp, d)"""
    device = _require_cuda(device)
    Xt = X if torch.is_tensor(X) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64))
    n, d = int(Xt.shape[0]), int(Xt.shape[1])
    ld = d + (d & 1)
    with torch.cuda.device(device):
        if ld == d:
            return Xt.to(device=device, dtype=torch.float64).contiguous(), d
        Xp = torch.zeros((n, ld), dtype=torch.float64, device=device)
        Xp[:, :d] = Xt.to(device=device, dtype=torch.float64)
        return Xp, d


def standardize_(Xp, d=None):
    """In place: every column of the device tensor Xp (n x ld, ld even, contiguous) becomes (x - mean) / std, as
    `sklearn.preprocessing.scale` computes it (population std; columns with std < 10 eps keep scale 1).
    Returns (mean, scale) as device tensors of d entries."""
    lib = _cabi.load()
    if not (torch.is_tensor(Xp) and Xp.is_cuda and Xp.dtype == torch.float64 and Xp.dim() == 2 and Xp.is_contiguous()):
        raise ValueError("standardize_ needs a contiguous float64 CUDA tensor n x ld")
    n, ld = int(Xp.shape[0]), int(Xp.shape[1])
    if ld & 1:
        raise ValueError("leading dimension must be even (see to_device_padded)")
    d = ld if d is None else int(d)
    dev = Xp.device
    with torch.cuda.device(dev):
        nb = ctypes.c_int64(0)
        _cabi.check(lib.rbl_standardize_scratch_bytes(dev.index or 0, ld, ctypes.byref(nb)))
        scratch = torch.empty((nb.value + 7) // 8, dtype=torch.float64, device=dev)
        mean = torch.empty(ld, dtype=torch.float64, device=dev)
        scale = torch.empty(ld, dtype=torch.float64, device=dev)
        _cabi.check(lib.rbl_standardize_columns(dev.index or 0, Xp.data_ptr(), n, d, ld, mean.data_ptr(),
                                                scale.data_ptr(), scratch.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    return mean[:d], scale[:d]


def split_rows(Xp, idx, d=None):
    """out[i, :] = Xp[idx[i], :] on the device (idx: host or device integer array); same leading dimension."""
    lib = _cabi.load()
    n, ld = int(Xp.shape[0]), int(Xp.shape[1])
    d = ld if d is None else int(d)
    dev = Xp.device
    with torch.cuda.device(dev):
        it = idx if torch.is_tensor(idx) else torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64))
        it = it.to(device=dev, dtype=torch.int64).contiguous()
        if it.numel() and (int(it.min()) < 0 or int(it.max()) >= n):
            raise IndexError("row index out of range")
        out = torch.empty((it.numel(), ld), dtype=torch.float64, device=dev)
        if it.numel():
            _cabi.check(lib.rbl_gather_rows(dev.index or 0, Xp.data_ptr(), ld, it.data_ptr(), it.numel(), d,
                                            out.data_ptr(), ld, torch.cuda.current_stream().cuda_stream))
    return out


def train_test_split_device(Xp, *host_arrays, test_size=None, train_size=None, random_state=None, d=None):
    """`sklearn.model_selection.train_test_split(X, *arrays, test_size=..., random_state=...)` with X on the device:
    the SAME row permutation (scikit-learn's ShuffleSplit on the host produces the indices), the rows of X gathered
    on the device, the small host arrays (labels, groups) indexed on the host.  Returns
    [X_train, X_test, a_train, a_test, ...] like the original."""
    from sklearn.model_selection import ShuffleSplit

    n = int(Xp.shape[0])
    cv = ShuffleSplit(n_splits=1, test_size=0.25 if test_size is None and train_size is None else test_size,
                      train_size=train_size, random_state=random_state)
    train, test = next(cv.split(np.zeros(n)))
    out = [split_rows(Xp, train, d), split_rows(Xp, test, d)]
    for a in host_arrays:
        a = np.asarray(a)
        out += [a[train], a[test]]
    return out
