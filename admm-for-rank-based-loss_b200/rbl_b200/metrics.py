"""Test-set metrics on the B200 — the step after the hot path.

`DeviceTestSet` keeps X_test / y_test / group resident in HBM; every evaluation is ONE pass over X_test
(`rbl_test_metrics`, include/rbl_b200.h) and a 128-byte read-back of the 16 numbers from which
`calculate_accuracy` (reference src/util/calculate_acc.py:3-19) and `calculate_statistics`
(src/util/fair_metric.py:3-40) are formed on the host.  There is no CPU path: without the library / a B200 the
constructor raises."""
import numpy as np
import torch

from . import _cabi
from .engine import _require_cuda

LOSS_IDS = {"binary_cross_entropy": 0, "hinge": 1}


def _loss_id(loss):
    if loss not in LOSS_IDS:  # calculate_acc.py:17-18
        raise ValueError(f"loss '{loss}' is not supported! Options: ['binary_cross_entropy','hinge']")
    return LOSS_IDS[loss]


class DeviceTestSet:
    """X (n x d, numpy or torch, host or device), y (n or n x 1, labels -1 / +1), group (n, {0, 1}; optional).
    sharded=True (one process per GPU, torch.distributed initialised): X, y, group are THIS rank's rows of the
    test set; all 16 numbers are sums over rows, so one all-reduce of 128 bytes gives every rank the metrics of the
    whole set."""

    def __init__(self, X, y, group=None, device=None, sharded=False, dist_group=None):
        self.lib = _cabi.load()
        self.device = _require_cuda(device)
        self.sharded, self.dist_group = bool(sharded), dist_group
        with torch.cuda.device(self.device):
            Xt = X if torch.is_tensor(X) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64))
            if Xt.dim() != 2:
                raise ValueError("X must be n x d")
            self.n, self.d = int(Xt.shape[0]), int(Xt.shape[1])
            self.X = Xt.to(device=self.device, dtype=torch.float64).contiguous()
            if self.d & 1:  # even leading dimension: every row 16-byte aligned for the kernel's double2 loads
                Xp = torch.zeros((self.n, self.d + 1), dtype=torch.float64, device=self.device)
                Xp[:, :self.d] = self.X
                self.X = Xp
            yt = y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y).reshape(-1).astype(np.float64))
            self.y = yt.reshape(-1).to(device=self.device, dtype=torch.float64).contiguous()
            if self.y.numel() != self.n:
                raise ValueError(f"y has {self.y.numel()} entries for {self.n} rows")
            self.group = None
            if group is not None:
                gt = group if torch.is_tensor(group) else torch.from_numpy(np.ascontiguousarray(group).reshape(-1))
                self.group = gt.reshape(-1).to(device=self.device, dtype=torch.int32).contiguous()
                if self.group.numel() != self.n:
                    raise ValueError(f"group has {self.group.numel()} entries for {self.n} rows")
            import ctypes
            nb = ctypes.c_int64(0)
            _cabi.check(self.lib.rbl_metrics_scratch_bytes(self.device.index or 0, ctypes.byref(nb)))
            self._scratch = torch.zeros((nb.value + 7) // 8, dtype=torch.float64, device=self.device)
            self._out = torch.zeros(16, dtype=torch.float64, device=self.device)
            self._out_host = torch.zeros(16, dtype=torch.float64).pin_memory()
            self._w = torch.empty(self.d, dtype=torch.float64, device=self.device)

    def launch(self, w, threshold=0.5, loss="binary_cross_entropy"):
        """Enqueue one evaluation on the current stream; returns the device tensor of the 16 numbers."""
        lid = _loss_id(loss)
        with torch.cuda.device(self.device):
            if torch.is_tensor(w) and w.is_cuda:
                wd = w.reshape(-1).to(device=self.device, dtype=torch.float64).contiguous()
            else:
                wh = torch.from_numpy(np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1)))
                self._w.copy_(wh, non_blocking=False)
                wd = self._w
            if wd.numel() != self.d:
                raise ValueError(f"w has {wd.numel()} entries for {self.d} columns")
            _cabi.check(self.lib.rbl_test_metrics(
                self.device.index or 0, self.X.data_ptr(), self.n, self.d, self.X.stride(0), wd.data_ptr(),
                self.y.data_ptr(), self.group.data_ptr() if self.group is not None else None, lid, float(threshold),
                self._out.data_ptr(), self._scratch.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return self._out

    def counts(self, w, threshold=0.5, loss="binary_cross_entropy"):
        """-> numpy[16], layout of rbl_test_metrics (include/rbl_b200.h)."""
        out = self.launch(w, threshold, loss)
        if self.sharded:
            torch.distributed.all_reduce(out, group=self.dist_group)  # every entry is a sum over rows
        self._out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._out_host.numpy().copy()

    def accuracy(self, w, threshold=0.5, loss="binary_cross_entropy"):
        c = self.counts(w, threshold, loss)
        return float(c[0] / c[1])  # np.mean(binary_preds == y_test), calculate_acc.py:11,16

    def statistics(self, w, threshold=0.5):
        if self.group is None:
            raise ValueError("calculate_statistics needs the group labels of the test set")
        return statistics_from_counts(self.counts(w, threshold, "binary_cross_entropy"))


def statistics_from_counts(c):
    """(SPD, DI, EOD, AOD, TI, FNRD) from the 16 numbers — fair_metric.py:13-40; group 0 is the reference's G1,
    group 1 its G2.  Empty classes divide by zero into nan / inf exactly as numpy does there (no exception)."""
    c = np.asarray(c, dtype=np.float64)
    n = c[1]
    g = c[2:14].reshape(2, 6)
    with np.errstate(divide="ignore", invalid="ignore"):
        size, pp, tp, fn, tn, fp = (g[:, k] for k in range(6))
        share = pp / size                                   # :13-14
        spd = share[1] - share[0]                           # :15
        di = np.inf if share[0] == 0 else share[1] / share[0]   # :24-27
        tpr, fpr, fnr = tp / (tp + fn), fp / (fp + tn), fn / (tp + fn)   # :28-33
        eod = tpr[1] - tpr[0]
        aod = 0.5 * (fpr[1] - fpr[0] + eod)
        mu = c[14] / n                                      # :36
        ti = c[15] / (n * mu) - np.log(mu)                  # :37-38: mean((b/mu) log(b/mu)), expanded
        fnrd = fnr[1] - fnr[0]
    return float(spd), float(di), float(eod), float(aod), float(ti), float(fnrd)
