"""GPU parity tests of the batched mode (K10): multi-RHS fused pass + per-instance FISTA state machines
and the BatchedADMM driver, against a loop over the single-instance oracle."""
import contextlib
import ctypes
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rbl_oracle as O  # noqa: E402


def _rel(a, b):
    return np.linalg.norm(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n,d,B", [(600, 64, 5), (1000, 201, 8), (3001, 1000, 11), (257, 40, 1)])
def test_batched_fista_matches_oracle_per_instance(n, d, B):
    from rbl_b200 import _cabi
    from rbl_b200.engine import AdmmEngine, _pow_table

    rng = np.random.default_rng(n + d + B)
    X = rng.normal(size=(n, d))
    e = AdmmEngine(X, -np.ones(n), "binary_cross_entropy", np.ones(n) / n)  # D = X
    _cabi.check(e.lib.rbl_batch_create(e.h, B))
    tab = _pow_table(np.float32(2.5))
    _cabi.check(e.lib.rbl_fista_config(e.h, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
    bs = rng.normal(size=(B, n))
    w0 = rng.normal(size=(B, d)) * 0.01
    lam_vals = 10 ** rng.uniform(-1, 2.5, size=B)
    bd, wd = e.vec(bs).reshape(B, n), e.vec(w0).reshape(B, d)
    lams = (ctypes.c_double * B)(*lam_vals)
    flags = (ctypes.c_int32 * B)(*([0] * B))
    s = e._stream
    _cabi.check(e.lib.rbl_fista_batch_begin(e.h, B, wd.data_ptr(), lams, flags, 17.0, 7e-5, 5000, s()))
    done, its, passes = (ctypes.c_int32 * B)(), (ctypes.c_int32 * B)(), (ctypes.c_int32 * B)()
    for _ in range(400):
        _cabi.check(e.lib.rbl_fista_batch_steps(e.h, B, e.D.data_ptr(), bd.data_ptr(), 8, s()))
        _cabi.check(e.lib.rbl_fista_batch_poll(e.h, B, s(), done, its, passes, None))
        if all(done):
            break
    assert all(done)
    W = torch.zeros((B, d), dtype=torch.float64, device=e.device)
    R = torch.zeros((B, n), dtype=torch.float64, device=e.device)
    _cabi.check(e.lib.rbl_fista_batch_result(e.h, B, W.data_ptr(), R.data_ptr(), s()))
    W, R = W.cpu().numpy(), R.cpu().numpy()
    for j in range(B):
        wo, info = O.fista(w0[j], X, bs[j], np.float64(lam_vals[j]), return_info=True)
        assert its[j] == info["iters"], (j, its[j], info)                       # same branch decisions
        assert np.linalg.norm(W[j] - wo) <= 1e-11 * max(np.linalg.norm(wo), 1e-3), j
        rref = bs[j] - X @ wo
        assert np.linalg.norm(R[j] - rref) <= 1e-11 * np.linalg.norm(rref), j  # residual of the accepted iterate
    e.close()


@pytest.mark.parametrize("mode", ["gram", "stream"])
def test_batched_admm_matches_independent_solves(golden_dir, mode):
    """Lockstep: before every ADMM iteration each oracle instance is put in the batched solver's state, then
    both step once.  With identical inner (FISTA) branch decisions — same iteration count, same number of
    line-search trials, same final L — the iterates agree to 1e-9.  When a line-search test `LHS > RHS`
    (e.g. a 1-sparse step on a column whose squared norm sits next to the float32 L grid) or the stop test
    `||dbeta|| < 7e-5` is a rounding-level near-tie, a different summation order may take the other branch;
    the iterate then moves by about the inner tolerance.  Such flips must be rare."""
    from rbl_b200.batched import BatchedADMM

    d2 = np.load(os.path.join(golden_dir, "data_600x64.npz"))
    X, y = d2["X"], d2["y"]
    regs = [0.3, 0.1, 0.03, 0.01, 0.003, 0.001, 0.0003, 0.05, 0.02, 0.007]
    for wf, args in [("superquantile", [0.8]), ("erm", None)]:
        b = BatchedADMM(X, y, wf, "binary_cross_entropy", l1_regs=regs, args=args, max_iter=30, tol=1e-7, mode=mode)
        assert b.mode == mode
        orc = [O.OracleADMM(X, y, wf, "binary_cross_entropy", l1_reg=r, args=args, max_iter=30, tol=1e-7,
                            small_lasso=False) for r in regs]
        flips = 0
        for it in range(30):
            for j, o in enumerate(orc):
                w, z, lam, rho = b.state(j)
                o.w, o.z, o.lam, o.rho = w.copy(), z.copy(), lam.copy(), rho
            b.step()
            for j, o in enumerate(orc):
                o.step()
                w, z, lam, rho = b.state(j)
                bi, bp, bL = b.last_fista_info[j]            # (iters, passes = 1 + trials, L)
                oi, ot, oL = o.last_fista_info               # (iters, trials, L)
                same = (bi == oi) and (bp - 1 == ot) and (bL == oL)
                flips += 0 if same else 1
                tol = 1e-9 if same else 2e-3
                assert _rel(w, o.w) < tol and _rel(z, o.z) < 1e-9, (wf, it, j, _rel(w, o.w), same)
                assert abs(float(rho) - float(o.rho)) <= 1e-15 * float(o.rho)
        assert flips <= 3, flips
        for j, o in enumerate(orc):
            o.w = b.state(j)[0].copy()
            assert abs(b.objective(j) - o.objective()) < 1e-11 * abs(o.objective())
        b.close()


@pytest.mark.parametrize("mode", ["gram", "stream"])
def test_batched_admm_ragged_convergence(mode):
    """instances converge at different iterations; finished ones retire without disturbing the others"""
    from rbl_b200.batched import BatchedADMM

    rng = np.random.default_rng(9)
    n, d = 800, 30
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:4] = rng.normal(size=4)
    y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
    regs = [1.0, 0.5, 0.2, 0.05, 0.01]
    b = BatchedADMM(X, y, "erm", "binary_cross_entropy", l1_regs=regs, max_iter=400, tol=1e-5, mode=mode)
    with contextlib.redirect_stdout(io.StringIO()):
        W = b.main_loop()
    assert b.converged.all() and len(set(b.iters.tolist())) > 1       # ragged
    for j, r in enumerate(regs):
        o = O.OracleADMM(X, y, "erm", "binary_cross_entropy", l1_reg=r, max_iter=400, tol=1e-5, small_lasso=False)
        wo = o.main_loop()
        # free-running to tol 1e-5: end states agree to well below the tolerance (inner-solver near-ties
        # may shift an iterate by ~1e-4 along the way, both runs still converge to the same minimiser)
        assert abs(int(o.iters) - int(b.iters[j])) <= 2, (j, o.iters, b.iters[j])
        assert _rel(W[:, j], wo) < 1e-4, (j, _rel(W[:, j], wo))
        assert abs(b.objective(j) - o.objective()) < 1e-7 * abs(o.objective())
    b.close()


@pytest.mark.parametrize("B_clip", [-5.0, 0.5])
def test_batched_ehrm_makes_the_candidate_choice_per_instance(golden_dir, B_clip):
    """EHRM in batched mode: every instance makes the reference's choice between min(B, prox_a) and max(B, prox_b)
    (PAV_cpt.py:203-226) at each of its z-steps — B = -5: candidate 2, B = 0.5: candidate 1 — in lockstep with
    per-instance oracles at 1e-9; the multi-RHS stream mode refuses EHRM."""
    from rbl_b200.batched import BatchedADMM

    d2 = np.load(os.path.join(golden_dir, "data_600x64.npz"))
    X, y = d2["X"], d2["y"]
    regs = [0.05, 0.01, 0.002]
    b = BatchedADMM(X, y, "ehrm", "binary_cross_entropy", l1_regs=regs, B_clip=B_clip, max_iter=15, tol=1e-7)
    assert b.mode == "gram"
    orc = [O.OracleADMM(X, y, "ehrm", "binary_cross_entropy", l1_reg=r, B=B_clip, max_iter=15, tol=1e-7,
                        small_lasso=False) for r in regs]
    for it in range(15):
        for j, o in enumerate(orc):
            w, z, lam, rho = b.state(j)
            o.w, o.z, o.lam, o.rho = w.copy(), z.copy(), lam.copy(), rho
        b.step()
        for j, o in enumerate(orc):
            o.step()
            w, z, lam, rho = b.state(j)
            assert _rel(z, o.z) < 1e-9, (it, j, _rel(z, o.z))
            assert _rel(w, o.w) < 1e-9 or np.linalg.norm(w - o.w) < 1e-9, (it, j, _rel(w, o.w))
    picked = [c.ehrm_stats for c in b.inst]
    assert all((s["cand1"] > 0) == (B_clip > -1) for s in picked), picked
    b.close()
    with pytest.raises(ValueError):
        BatchedADMM(X, y, "ehrm", "binary_cross_entropy", l1_regs=regs, B_clip=B_clip, mode="stream")
