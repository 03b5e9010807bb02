"""Dev tool: wall time of the first ADMM iterations of the c2 solve one by one (eager iterations, the transposed copy,
the graph capture), B200 box.  The headline solve spends ~35 of its ~147 ms in iterations 0..4."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
import bench as B
from rbl_b200 import engine as E
from src.optim.algorithms import ADMMmethod

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda", 0)
quiet = contextlib.redirect_stdout(io.StringIO())
rng = np.random.default_rng(5)
Xw = rng.standard_normal((4096, d)); yw = np.sign(Xw[:, 0] + 0.1 * rng.standard_normal(4096)).reshape(-1, 1)
kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.01, args=[0.8])
with quiet:
    w = ADMMmethod(Xw, yw, max_iter=100, tol=1e-6, **kw); w.advance(0, 8)
w.engine.close()
Xh, yh = B.gen_rows_device(torch, dev, 0, n, n, d, pin=True)
marks = []
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(self, *a, **k)
        torch.cuda.synchronize(); marks.append((name, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(cls, name, g)
for nm in ("_build_transpose", "_capture_iteration", "z_step", "w_step_fista", "_dual_launch"):
    wrap(E.AdmmEngine, nm)
torch.cuda.synchronize(); t0 = time.perf_counter()
s = ADMMmethod(Xh.numpy(), yh.numpy().reshape(-1, 1), max_iter=100000, tol=1e-6, **kw)
torch.cuda.synchronize(); print("constructor %.1f ms" % ((time.perf_counter() - t0) * 1e3), s.engine.build_times)
i = 0
with quiet:
    for _ in range(9):
        marks.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        i, done = s.advance(i, 1)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        sys.stderr.write("iteration %d: %.2f ms  %s  fista %s nnz %s graph %s\n" % (
            i - 1, dt, " ".join("%s=%.2f" % m for m in marks), s.engine.fista_stats.get("last_passes"),
            s.engine.dual_stats.get("nnz_last"), s.engine._graph is not None))
