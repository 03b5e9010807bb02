"""Dev tool: where one ADMM iteration's time goes (CUDA events per phase), B200 box."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from src.optim.algorithms import ADMMmethod

n, d = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(17)
X = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
ws = torch.zeros(d, dtype=torch.float64, device=dev); ws[:10] = torch.randn(10, generator=g, dtype=torch.float64, device=dev)
y = torch.sign(X @ ws + 0.1 * torch.randn(n, generator=g, dtype=torch.float64, device=dev)); y[y == 0] = 1
s = ADMMmethod(X.cpu().numpy(), y.cpu().numpy().reshape(-1, 1), "superquantile", "binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=1000, tol=1e-6)
del X
e = s.engine
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
tot = {"z": 0.0, "w": 0.0, "dual": 0.0, "wall": 0.0, "passes": 0, "polls": 0}
for it in range(33):
    t0 = time.perf_counter()
    a = ev(); e.z_step(s.rho); b = ev()
    p0, q0 = e.fista_stats["passes"], e.fista_stats["polls"]
    s._w_subproblem_device(); c = ev()
    pf, df = e.dual_step(s.rho); dd = ev()
    s._w = e.w_host.numpy().reshape(-1, 1).copy()
    s.rho = np.min((s.rho * (1.02 if pf > 1e-2 else 1.07), 217 * s.num_feature))
    torch.cuda.synchronize()
    if it >= 3:
        tot["z"] += a.elapsed_time(b); tot["w"] += b.elapsed_time(c); tot["dual"] += c.elapsed_time(dd)
        tot["wall"] += (time.perf_counter() - t0) * 1e3
        tot["passes"] += e.fista_stats["passes"] - p0; tot["polls"] += e.fista_stats["polls"] - q0
K = 30
print(f"n={n} d={d} per iteration: z {tot['z']/K:.3f} ms, w {tot['w']/K:.3f} ms ({tot['passes']/K:.1f} passes, {tot['polls']/K:.1f} polls, "
      f"{tot['w']/tot['passes']*1e3:.1f} us/pass), dual {tot['dual']/K:.3f} ms, wall {tot['wall']/K:.3f} ms")
