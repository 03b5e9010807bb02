"""Dev tool: host -> HBM upload + build time of the solver constructor for pinned vs pageable input, with the
threaded staging (rbl_h2d_pageable, 4 / 8 / 16 threads) and with the driver's own pageable path."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200.engine import AdmmEngine

n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 1000
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d))
y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
sig = np.ones(n) / n
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True); Xp.copy_(torch.from_numpy(X))
t0 = time.perf_counter(); Y = np.empty_like(X); np.copyto(Y, X); t1 = time.perf_counter()
print(f"host memcpy (1 thread) {X.nbytes / (t1 - t0) / 1e9:.1f} GB/s", flush=True)
del Y
def run(tag, arr, env):
    for k, v in env.items():
        os.environ[k] = v
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = AdmmEngine(arr, y, "binary_cross_entropy", sig)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        path = getattr(e, "upload_path", "?")
        e.close(); del e
        print(f"{tag:32s} rep {rep}: {dt:.3f} s  ({X.nbytes / dt / 1e9:.1f} GB/s incl. build)  [{path}]", flush=True)
    for k in env:
        del os.environ[k]
run("pinned", Xp.numpy(), {})
X2 = np.array(Xp.numpy())      # what bench.py times: a fresh pageable copy of the pinned array
run("pageable copy of pinned, 8 thr", X2, {"RBL_UPLOAD_THREADS": "8"})
del X2
for th in ("4", "8", "16"):
    run(f"pageable, staging {th} threads", X, {"RBL_UPLOAD_THREADS": th})
run("pageable, driver path", X, {"RBL_PAGEABLE_STAGING": "0"})
