"""Dev tool: phase timing inside the persistent Gram-FISTA kernel (globaltimer stamps of CTA 0), B200 box."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from src.optim.algorithms import ADMMmethod, Optimizer

os.environ["RBL_GRAPH"] = "0"
n, d = 200_000, 1000
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); ws = np.zeros(d); ws[:10] = rng.normal(size=10)
y = np.sign(X @ ws + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
s = ADMMmethod(X, y, "superquantile", "binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=100, tol=1e-9)
e = s.engine
st = torch.zeros(64, dtype=torch.int64, device=e.device)
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(12):
        if i == 11:
            _cabi.check(e.lib.rbl_sort_debug(e.h, st.data_ptr()))
            e.z_step(s.rho)
            _cabi.check(e.lib.rbl_sort_debug(e.h, 0))  # sort done without stamps... keep FISTA stamps only
            _cabi.check(e.lib.rbl_sort_debug(e.h, st.data_ptr()))
            s._w_subproblem_device()
            torch.cuda.synchronize()
            break
        Optimizer.main_loop(s, i, 0.0, False)
t = st.cpu().numpy()
k = int(t[0]); ts = t[1:1 + k].astype(np.float64)
print("stamps:", k, " total %.1f us" % ((ts[-1] - ts[0]) / 1e3))
print("deltas (us):", " ".join("%.1f" % x for x in np.diff(ts) / 1e3))
