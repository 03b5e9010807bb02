"""Dev tool: GPU time of one replayed iteration graph (back-to-back replays, no host sync in between) vs the
per-iteration time of the native loop (one launch + one sync per iteration), B200 box."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from src.optim.algorithms import ADMMmethod

n, d = 1_000_000, 1000
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(17)
X = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
ws = torch.zeros(d, dtype=torch.float64, device=dev); ws[:10] = torch.randn(10, generator=g, dtype=torch.float64, device=dev)
y = torch.sign(X @ ws + 0.1 * torch.randn(n, generator=g, dtype=torch.float64, device=dev)); y[y == 0] = 1
s = ADMMmethod(X.cpu().numpy(), y.cpu().numpy().reshape(-1, 1), "superquantile", "binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=1000, tol=1e-6)
del X
e = s.engine
with contextlib.redirect_stdout(io.StringIO()):
    s.advance(0, 40)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
with contextlib.redirect_stdout(io.StringIO()):
    s.advance(40, 30)
b.record(); torch.cuda.synchronize()
t_loop = a.elapsed_time(b) / 30 * 1e3
a.record()
for _ in range(30):
    e._graph.replay()
b.record(); torch.cuda.synchronize()
t_graph = a.elapsed_time(b) / 30 * 1e3
print(f"native loop: {t_loop:.1f} us/iteration; back-to-back graph replays: {t_graph:.1f} us/iteration")
