"""Dev tool: z-step component timings at the states a real solve passes through (B200 box)."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from src.optim.algorithms import ADMMmethod, Optimizer

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

n, d = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(17)
X = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
ws = torch.zeros(d, dtype=torch.float64, device=dev); ws[:10] = torch.randn(10, generator=g, dtype=torch.float64, device=dev)
y = torch.sign(X @ ws + 0.1 * torch.randn(n, generator=g, dtype=torch.float64, device=dev)); y[y == 0] = 1
s = ADMMmethod(X.cpu().numpy(), y.cpu().numpy().reshape(-1, 1), "superquantile", "binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=1000, tol=1e-6)
del X
e = s.engine; lib, h, st = e.lib, e.h, e._stream
marks = {0, 1, 3, 10, 20, 33, 60, 100, 150, 200} if len(sys.argv) < 4 else {int(sys.argv[3])}
for it in range(max(marks) + 1):
    if it in marks:
        rho = float(s.rho)
        t_sort = timeit(lambda: _cabi.check(lib.rbl_sort_margins(h, e.m_glob.data_ptr(), e.m_sorted.data_ptr(), e.perm.data_ptr(), st())))
        t_pav = timeit(lambda: _cabi.check(lib.rbl_pav_prox(h, 0, e.m_sorted.data_ptr(), rho, e.z_sorted.data_ptr(), st())))
        z = e.z_sorted.cpu().numpy(); dz = np.flatnonzero(np.diff(z) != 0)
        runs = np.diff(np.concatenate([[-1], dz, [n - 1]]))
        print(f"it {it:3d} rho {rho:.3e}: sort {t_sort:7.1f} us  pav {t_pav:7.1f} us  blocks {len(runs)} largest {runs.max()} "
              f"fista passes last {e.fista_stats['last_passes']}", flush=True)
    with contextlib.redirect_stdout(io.StringIO()):
        if Optimizer.main_loop(s, it, 0.0, False): break
