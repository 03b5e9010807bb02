"""Dev tool: which route the hinted sort takes along a solve (bucket route vs LSD fallback), B200 box."""
import contextlib, ctypes, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from src.optim.algorithms import ADMMmethod, Optimizer

os.environ["RBL_GRAPH"] = sys.argv[1] if len(sys.argv) > 1 else "0"
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
d = 64
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); ws = np.zeros(d); ws[:10] = rng.normal(size=10)
y = np.sign(X @ ws + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
kw = {"c2": dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.01, args=[0.8]),
      "c3": dict(weight_function="ehrm", loss="binary_cross_entropy", B=-5, l2_reg=0.01),
      "c4": dict(weight_function="aorr", loss="hinge", args=[0.2, 0.8], l2_reg=1e-4)}[cfg]
s = ADMMmethod(X, y, max_iter=100, tol=1e-9, **kw)
e = s.engine
st = (ctypes.c_int32 * 4)()
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(60):
        Optimizer.main_loop(s, i, 0.0, False)
        _cabi.check(e.lib.rbl_sort_stats(e.h, e._stream(), st))
        if i < 15 or i % 5 == 0:
            sys.stderr.write("it %2d: buckets %d route %d largest bucket %d flag %d graph %s\n"
                             % (i, st[0], st[1], st[2], st[3], e._graph is not None))
