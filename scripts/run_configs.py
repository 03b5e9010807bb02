"""Dev tool: reduced-size versions of BASELINE configs 2-4 on the B200 box — iterations/s and per-iteration
kernel counters, to see that every (spectrum, loss, regulariser) combination runs on its intended route."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from src.optim.algorithms import ADMMmethod, Optimizer

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
CASES = [
    ("C2 SRM superquantile l1", int(1_000_000 * scale), 1000, "superquantile", [0.8], "binary_cross_entropy", None, dict(l1_reg=0.01), False),
    ("C2' SRM superquantile l2", int(1_000_000 * scale), 1000, "superquantile", [0.8], "binary_cross_entropy", None, dict(l2_reg=0.01), False),
    ("C3 EHRM l2", int(4_000_000 * scale), 500, "ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01), False),
    ("C4 AoRR hinge l2", int(2_000_000 * scale), 200, "aorr", [0.2, 0.8], "hinge", None, dict(l2_reg=1e-4), True),
    ("C4' AoRR BCE l2", int(2_000_000 * scale), 200, "aorr", [0.2, 0.8], "binary_cross_entropy", None, dict(l2_reg=1e-4), True),
    ("ERM l1 (run_SRM.py)", int(600_000 * scale), 1000, "erm", None, "binary_cross_entropy", None, dict(l1_reg=0.01), False),
]
dev = torch.device("cuda")
for tag, n, d, wf, args, loss, B, kw, intercept in CASES:
    g = torch.Generator(device=dev); g.manual_seed(17)
    X = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
    ws = torch.zeros(d, dtype=torch.float64, device=dev); ws[:10] = torch.randn(10, generator=g, dtype=torch.float64, device=dev)
    y = torch.sign(X @ ws + 0.1 * torch.randn(n, generator=g, dtype=torch.float64, device=dev)); y[y == 0] = 1
    Xh = X.cpu().numpy(); yh = y.cpu().numpy().reshape(-1, 1)
    del X, y
    if intercept:
        Xh = np.hstack([Xh, np.ones((n, 1))])
    t0 = time.perf_counter()
    s = ADMMmethod(Xh, yh, wf, loss, B=B, args=args, max_iter=1000, tol=1e-6, **kw)
    torch.cuda.synchronize(); t_build = time.perf_counter() - t0
    e = s.engine
    K = 30
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(5):
            Optimizer.main_loop(s, i, 0.0, False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(5, 5 + K):
            Optimizer.main_loop(s, i, 0.0, False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nseg = __import__("ctypes").c_int32(-1); e.lib.rbl_pav_config(e.h, 0, __import__("ctypes").byref(nseg))
    print(f"{tag:28s} n={n:8d} d={Xh.shape[1]:5d}: {K/dt:8.1f} it/s ({dt/K*1e3:7.3f} ms/it), build {t_build:5.2f} s, "
          f"w_mode {e.w_mode}, pav runs {nseg.value}, active {e.active_stats}, dual {e.dual_stats}, "
          f"graph {'yes' if e._graph is not None else 'no'}, primal {s.primal_feasibility:.2e} lbfgs {s.last_info}", flush=True)
    e.close(); del s, e, Xh
    torch.cuda.empty_cache()
