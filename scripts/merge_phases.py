"""Dev tool: phase timing inside the few-segment PAV merge kernel (globaltimer stamps), B200 box."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from rbl_b200 import _cabi
from src.optim.algorithms import ADMMmethod, Optimizer

os.environ["RBL_GRAPH"] = "0"
n, d = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); ws = np.zeros(d); ws[:10] = rng.normal(size=10)
y = np.sign(X @ ws + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
s = ADMMmethod(X, y, "superquantile", "binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=100, tol=1e-9)
e = s.engine
st = torch.zeros(64, dtype=torch.int64, device=e.device)
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(31):
        d_ = 0
        if i in (3, 10, 20, 30):
            torch.cuda.synchronize()
            _cabi.check(e.lib.rbl_sort_config(e.h, 1))   # legacy sort: leaves the stamp buffer to the merge kernel
            _cabi.check(e.lib.rbl_sort_debug(e.h, st.data_ptr()))
            e.z_step(s.rho)
            torch.cuda.synchronize()
            _cabi.check(e.lib.rbl_sort_debug(e.h, 0))
            _cabi.check(e.lib.rbl_sort_config(e.h, 0))
            t = st.cpu().numpy(); k = int(t[0]); ts = t[1:1 + k].astype(np.float64)
            sys.stderr.write("it %d: %d stamps, total %.1f us; [offsets scan | (windows | search + run snaps | block value | rest of finish) per merge] = %s\n"
                             % (i, k, (ts[-1] - ts[0]) / 1e3, " ".join("%.1f" % x for x in np.diff(ts) / 1e3)))
            for j in (1, 2):
                hl, hh, al, ah = (int(v) for v in t[40 + 4 * j: 44 + 4 * j])
                sys.stderr.write("    merge %d: guess [%d, %d) answer [%d, %d): moved %d / %d ranks; snapped to whole runs: -%d / +%d\n"
                                 % (j, hl, hh, al, ah, al - hl, ah - hh, int(t[56 + 2 * j]), int(t[57 + 2 * j])))
        Optimizer.main_loop(s, i, 0.0, False)
