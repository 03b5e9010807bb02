import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-for-rank-based-loss_b200")):
    sys.path.insert(0, p)
from oracle import rbl_oracle as O
from rbl_b200.batched import BatchedADMM
from rbl_b200.engine import AdmmEngine
d2 = np.load(os.path.join(ROOT, "tests/golden/data_600x64.npz")); X, y = d2["X"], d2["y"]
regs = [0.3, 0.1, 0.03, 0.01, 0.003, 0.001, 0.0003, 0.05, 0.02, 0.007]
b = BatchedADMM(X, y, "superquantile", "binary_cross_entropy", l1_regs=regs, args=[0.8], max_iter=30, tol=1e-7)
o = O.OracleADMM(X, y, "superquantile", "binary_cross_entropy", l1_reg=0.3, args=[0.8], max_iter=30, tol=1e-7, small_lasso=False)
for it in range(15):
    w, z, lam, rho = b.state(0)
    o.w, o.z, o.lam, o.rho = w.copy(), z.copy(), lam.copy(), rho
    if it == 14:
        # reproduce the w-step inputs
        zo = o.z_step()
        bb = zo + o.lam / o.rho
        lamf = (0.3 / (2 * o.rho * o.n)) * o.n
        wo, info = O.fista(o.w, o.D, bb, lamf, return_info=True)
        e1 = AdmmEngine(X, y, "binary_cross_entropy", np.ones(o.n) / o.n)
        w1, i1 = e1.fista(e1.vec(o.w), e1.vec(bb), lamf)
        print("oracle info", info, "single-gpu info", i1, "rel", np.linalg.norm(w1.cpu().numpy() - wo) / np.linalg.norm(wo), type(lamf), lamf)
    b.step()
    if it == 14:
        w2 = b.state(0)[0]
        print("batched info", b.last_fista_info[0], "rel vs oracle", np.linalg.norm(w2 - wo) / np.linalg.norm(wo), "vs single", np.linalg.norm(w2 - w1.cpu().numpy()) / np.linalg.norm(wo))
        print("z rel", np.linalg.norm(b.state(0)[1] - zo) / np.linalg.norm(zo))
        print(w2[np.abs(w2) > 0], wo[np.abs(wo) > 0], w1.cpu().numpy()[np.abs(wo) > 0])
