"""Dev tool: hinted (splitter) sort vs LSD sort timings at n keys, B200 box."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from rbl_b200.engine import AdmmEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
e = AdmmEngine(np.zeros((n, 2)), np.ones(n), "binary_cross_entropy", np.ones(n) / n)
rng = np.random.default_rng(0)
base = rng.normal(size=n)
m0, m1 = e.vec(base), e.vec(base + 1e-4 * rng.normal(size=n))
s = e._stream
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
lsd = lambda: _cabi.check(e.lib.rbl_sort_margins(e.h, m0.data_ptr(), e.m_sorted.data_ptr(), e.perm.data_ptr(), s()))
t_lsd = timeit(lsd)
lsd()
flip = [0]
def near():
    flip[0] ^= 1
    _cabi.check(e.lib.rbl_sort_margins_near(e.h, (m1 if flip[0] else m0).data_ptr(), e.perm.data_ptr(),
                                            e.m_sorted.data_ptr(), e.perm.data_ptr(), s()))
t_near = timeit(near)
st = (ctypes.c_int32 * 4)()
_cabi.check(e.lib.rbl_sort_stats(e.h, s(), st))
print(f"n={n}: LSD {t_lsd:.1f} us, hinted {t_near:.1f} us (route {st[1]}, largest bucket {st[2]}, buckets {st[0]})")
