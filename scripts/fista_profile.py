"""Dev tool: per-step cost of the FISTA state machine vs the bare pass (B200 box)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from rbl_b200.engine import AdmmEngine, _pow_table

def ev():
    return torch.cuda.Event(enable_timing=True)

n, d = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
X = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
e = AdmmEngine(X, -torch.ones(n, dtype=torch.float64, device=dev), "binary_cross_entropy", np.ones(n) / n)
del X
lib, h, s = e.lib, e.h, e._stream
b = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
w0 = torch.zeros(d, dtype=torch.float64, device=dev)
tab = _pow_table(np.float32(2.5))
_cabi.check(lib.rbl_fista_config(h, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
K = 60
def run_steps():
    _cabi.check(lib.rbl_fista_begin(h, w0.data_ptr(), 0.5, 0, 17.0, 0.0, 100000, s()))
    _cabi.check(lib.rbl_fista_steps(h, e.D.data_ptr(), b.data_ptr(), K, s()))
def run_pass():
    for _ in range(K):
        _cabi.check(lib.rbl_fused_pass(h, e.D.data_ptr(), w0.data_ptr(), b.data_ptr(), e.r.data_ptr(), e.red.data_ptr(), s()))
def run_matvec():
    for _ in range(K):
        e.matvec(w0, e.Dw)
for name, fn in [("fista step (pass+reduce+update+combine)", run_steps), ("fused pass + reduce + copy", run_pass), ("matvec pass", run_matvec)]:
    fn(); torch.cuda.synchronize()
    a, c = ev(), ev()
    a.record(); fn(); c.record(); torch.cuda.synchronize()
    t = a.elapsed_time(c) / K * 1e3
    gb = (n * d * 8) / (t * 1e-6) / 1e9
    print(f"n={n} d={d}: {name}: {t:.1f} us/step  ({gb:.0f} GB/s of D)")
hi = (ctypes.c_int32 * 8)(); hd = (ctypes.c_double * 4)()
_cabi.check(lib.rbl_fista_poll(h, s(), hi, hd)); print("state", list(hi), list(hd))
