"""GPU parity tests at the solver level: the drop-in ADMMmethod / seams against (a) golden vectors
produced by the reference itself (tests/golden, oracle/gen_golden.py) and (b) the CPU oracle run on
the same inputs.  Tolerances stated per assert."""
import contextlib
import ctypes
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rbl_oracle as O  # noqa: E402


@pytest.fixture(autouse=True, params=["gram", "stream"])
def w_mode(request, monkeypatch):
    """every solver-level test runs with the w-step on G = D^T D (default where it applies) and with the
    w-step streaming D per trial"""
    monkeypatch.setenv("RBL_W_MODE", request.param)
    return request.param


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _args(s):
    return None if s == "" else [float(a) if "." in a else int(a) for a in s.split(",")]


def _rel(a, b):
    return np.linalg.norm(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)) / max(np.linalg.norm(b), 1e-300)


def test_zstep_vs_reference_golden_and_oracle(golden_dir):
    from src.optim.algorithms import ADMMmethod

    g = _load(golden_dir, "zstep.npz")
    d = _load(golden_dir, "data_300x40.npz")
    for ci in range(int(g["ncases"])):
        wf, args, loss, B, rho = g[f"c{ci}_meta"]
        args, B, rho = _args(args), (None if B == "" else float(B)), float(rho)
        s = ADMMmethod(d["X"], d["y"], wf, loss, l2_reg=0.01, B=B, args=args)
        s.w = g[f"c{ci}_w"]
        s.lagrangian = g[f"c{ci}_lam"]
        s.rho = rho
        z = s.z_subproblem().reshape(-1)
        o = O.OracleADMM(d["X"], d["y"], wf, loss, l2_reg=0.01, B=B, args=args)
        o.w, o.lam, o.rho = g[f"c{ci}_w"].copy(), g[f"c{ci}_lam"].copy(), rho
        zo = o.z_step()
        # vs oracle: same exact prox, both to machine precision
        assert _rel(z, zo) < 1e-12, (ci, wf, loss)
        # vs the reference's own output: limited by ITS inner tolerances (Newton 1e-6 / 1e-4,
        # hinge bisection with early exit) — see tests/test_oracle.py
        assert _rel(z, g[f"c{ci}_ref_z"]) < (1e-4 if loss == "hinge" else 5e-9), (ci, wf, loss)
        s.engine.close()


@pytest.mark.parametrize("persistent", [True, False])
def test_fista_vs_reference_golden_and_oracle(golden_dir, w_mode, persistent, monkeypatch):
    from src.util.fast_lasso import FISTA

    if not persistent:
        if w_mode != "gram":
            pytest.skip("the persistent kernel is a Gram-mode feature")
        monkeypatch.setenv("RBL_GRAM_PERSISTENT", "0")  # one launch per line-search trial (what large d falls back to)

    g = _load(golden_dir, "fista.npz")
    d = _load(golden_dir, "data_300x40.npz")
    D = -d["y"] * d["X"]
    for lam in (0.5, 20.0, 300.0):
        for lam_v, key in ((lam, f"ref_f64_{lam}"), (np.float64(lam), f"ref_f64_np_{lam}")):
            w, info = FISTA(g["w0"], D, g["b"], lam_v, np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000,
                            dtype=torch.float64, return_info=True)
            wo, oinfo = O.fista(g["w0"], D, g["b"], lam_v, return_info=True)
            assert info["iters"] == oinfo["iters"], (lam, info, oinfo)          # same branch decisions
            assert info["L"] == oinfo["L"]
            # one fused pass per trial (affine recombination of g and r) vs three matvecs: same iterates
            # up to fp64 rounding
            assert np.linalg.norm(w - wo) <= 1e-11 * max(np.linalg.norm(wo), 1e-3), lam
            assert np.linalg.norm(w - g[key]) <= 1e-11 * max(np.linalg.norm(g[key]), 1e-3), lam
            # stream: one pass over D per trial (+1 initial); gram: one sweep over G per trial
            # (the persistent kernel carries up to 8 line-search candidates per sweep)
            if w_mode == "gram":
                # the oracle counts matvecs: 2 per iteration + 1 per line-search trial
                assert info["passes"] <= info["trials"] == oinfo["passes"] - 2 * oinfo["iters"]
            else:
                assert info["passes"] == 1 + info["trials"]


def test_l2_step_vs_reference_golden(golden_dir):
    from src.util.w_LBFGS import w_solver

    g = _load(golden_dir, "l2step.npz")
    d = _load(golden_dir, "data_300x40.npz")
    D = -d["y"] * d["X"]
    for rho in (1e-5, 1e-2, 1.0):
        w = w_solver(2, g["w0"].reshape(-1, 1), g["z"].reshape(-1, 1), g["lam"].reshape(-1, 1), rho, None, D, 0.01)
        # same scipy L-BFGS-B driver; f/g from the fused pass differ from D@w / DTD@w by rounding only
        assert _rel(w, g[f"ref_{rho}"]) < 1e-9, rho


def test_objective_vs_reference_golden(golden_dir):
    from src.optim.objective import rankbasedObjective

    g = _load(golden_dir, "objective.npz")
    d = _load(golden_dir, "data_300x40.npz")
    for wf, args, loss, B, kw in [("erm", None, "binary_cross_entropy", None, dict(l1_reg=0.01)),
                                  ("superquantile", [0.8], "binary_cross_entropy", None, dict(l2_reg=0.01)),
                                  ("aorr", [0.2, 0.8], "hinge", None, dict(l2_reg=1e-4)),
                                  ("ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01)),
                                  ("esrm", [1.5], "hinge", None, dict(l1_reg=0.1))]:
        ob = rankbasedObjective(torch.from_numpy(d["X"]), torch.from_numpy(d["y"]), wf, loss, kw.get("l2_reg"),
                                kw.get("l1_reg"), B, None, args)
        v = ob.get_arrogate_loss(torch.from_numpy(g["w"].reshape(-1, 1)))
        ref = float(g[f"ref_{wf}_{loss}"])
        assert abs(v - ref) < 1e-12 * abs(ref), (wf, v, ref)
        ob.problem.close()


@pytest.mark.parametrize("fname", ["trajectory.npz", "trajectory_extra.npz"])
def test_trajectories_vs_oracle_and_reference(golden_dir, fname):
    """40 ADMM iterations through the drop-in class vs the oracle (free-running, same inputs) and the
    reference's recorded iterates (oracle/gen_golden.py: main and trajectories_extra)."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    g = _load(golden_dir, fname)
    d1 = _load(golden_dir, "data_300x40.npz")
    d2 = _load(golden_dir, "data_600x64.npz")
    tags = sorted({k[:-5] for k in g.files if k.endswith("_meta")})
    for tag in tags:
        # (erm_l1, sq_l1: 300 x 40 + l1 is the reference's sklearn-Lasso branch, algorithms.py:194-197 — coordinate
        # descent on the device and in the oracle)
        wf, args, loss, B, kw = g[f"{tag}_meta"]
        args, B, kw = _args(args), (None if B == "" else float(B)), eval(kw)
        d = d2 if tag.endswith("_fista") else d1
        s = ADMMmethod(d["X"], d["y"], wf, loss, B=B, args=args, max_iter=40, tol=1e-6, **kw)
        o = O.OracleADMM(d["X"], d["y"], wf, loss, B=B, args=args, max_iter=40, tol=1e-6, **kw)
        for i in range(40):
            with contextlib.redirect_stdout(io.StringIO()):
                Optimizer.main_loop(s, i, 0.0, False)
            o.step()
            ew, ez = _rel(s.w, o.w), _rel(s.z, o.z)
            # north_star: iterates within 1e-9 relative of the float64 reference implementation
            assert ew < 1e-9 and ez < 1e-9, (tag, i + 1, ew, ez)
            assert abs(float(s.rho) - float(o.rho)) <= 1e-15 * float(o.rho)
            if f"{tag}_w_{i+1}" in g.files:
                # vs the reference's own iterates: bounded by ITS inexact inner solvers (test_oracle.py)
                tol = 2e-8 if i + 1 <= 3 else 1e-5
                if wf == "aorr_dc" and i + 1 <= 3:
                    tol = 1e-7  # rho0 = 2e-7: the reference's own z is only ~1e-9 * sigma/rho accurate (test_oracle.py)
                assert _rel(s.w, g[f"{tag}_w_{i+1}"]) < tol, (tag, i + 1)
        obj = s.objective.get_arrogate_loss(torch.from_numpy(s.w).double())
        assert abs(obj - o.objective()) < 1e-9 * abs(obj)
        assert abs(obj - float(g[f"{tag}_obj"])) < 1e-7
        s.engine.close()


def test_driver_style_run_with_store(golden_dir):
    """run_demo.py-style usage: start_store / main_loop(verbose) / final_res, plus error behaviour."""
    from src.optim.algorithms import ADMMmethod

    d = _load(golden_dir, "data_300x40.npz")
    s = ADMMmethod(d["X"], d["y"], "erm", "binary_cross_entropy", l2_reg=1e-4, max_iter=12)
    with pytest.raises(ValueError):
        s.final_res()
    s.start_store(d["X_test"], d["y_test"], "erm", "binary_cross_entropy", l2_reg=1e-4)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        w = s.main_loop(verbose=True)
    assert "iter_num= 0" in buf.getvalue() and "loss=" in buf.getvalue()
    w2, t, tr, te = s.final_res()
    assert w.shape == (40, 1) and len(t) == len(tr) == len(te) == 13 and tr[-1] < tr[0]
    with pytest.raises(ValueError):
        ADMMmethod(d["X"], d["y"], "superquantile", "binary_cross_entropy", l2_reg=1e-4)  # args is None
    with pytest.raises(ValueError):
        ADMMmethod(d["X"], d["y"], "erm", "binary_cross_entropy", l2_reg=1e-4, B=-5)       # B without ehrm
    with pytest.raises(ValueError):
        ADMMmethod(d["X"], d["y"], "erm", "nope", l2_reg=1e-4)
    with pytest.raises(TypeError):
        ADMMmethod(d["X"], d["y"], "erm", "binary_cross_entropy")                          # no regulariser


def test_aorr_and_ehrm_driver_flows_with_test_metrics():
    """run_AoRR_fixed.py:152-156 and run_EHRM.py:36-41 end to end on the drop-in names: solver with an intercept
    column, start_store on the test split with its own spectrum args, main_loop, final_res, calculate_accuracy and
    calculate_statistics — every number against the oracle on the same w."""
    from src.optim.algorithms import ADMMmethod
    from src.util.calculate_acc import calculate_accuracy
    from src.util.fair_metric import calculate_statistics

    rng = np.random.default_rng(77)
    n, nt, d = 900, 600, 9
    Xa = rng.normal(size=(n + nt, d))
    ws = rng.normal(size=d)
    ya = np.where(Xa @ ws + 0.7 * rng.normal(size=n + nt) > 0, 1, -1).reshape(-1, 1)
    grp = (rng.random(n + nt) < 0.4).astype(int)
    X, Xt = np.hstack([Xa[:n], np.ones((n, 1))]), np.hstack([Xa[n:], np.ones((nt, 1))])
    y, yt, gt = ya[:n], ya[n:], grp[n:]
    for wf, loss, args, targs, B in (("aorr_dc", "hinge", [500, 10], [1, 0], None),
                                     ("ehrm", "binary_cross_entropy", None, None, -5)):
        s = ADMMmethod(X, y, wf, loss, l2_reg=1e-4 if wf == "aorr_dc" else 0.01, B=B, args=args, max_iter=25)
        s.start_store(Xt, yt, wf, loss, l2_reg=1e-4 if wf == "aorr_dc" else 0.01, B=B, args=targs)
        with contextlib.redirect_stdout(io.StringIO()):
            s.main_loop(verbose=False)
        w, t, tr, te = s.final_res()
        assert w.shape == (d + 1, 1) and len(t) == len(tr) == len(te) and np.all(np.isfinite(te))
        sig_t = O.spectrum(wf, nt, targs)
        sig_t = sig_t[0] if isinstance(sig_t, tuple) else sig_t
        ref_te = O.objective(-yt * Xt, w, sig_t, loss, l2_reg=1e-4 if wf == "aorr_dc" else 0.01)
        assert abs(te[-1] - ref_te) < 1e-10 * max(1.0, abs(ref_te))
        # the per-iteration train loss reuses the engine's D w (no extra pass over D): same value
        sig = O.spectrum(wf, n, args)
        sig = sig[0] if isinstance(sig, tuple) else sig
        ref_tr = O.objective(-y * X, w, sig, loss, l2_reg=1e-4 if wf == "aorr_dc" else 0.01)
        assert abs(tr[-1] - ref_tr) < 1e-10 * max(1.0, abs(ref_tr))
        from src.optim.objective import rankbasedObjective
        fresh = rankbasedObjective.get_arrogate_loss(s.objective, torch.from_numpy(w).double())  # with its own pass
        assert abs(fresh - tr[-1]) < 1e-12 * max(1.0, abs(fresh))
        acc = calculate_accuracy(w.reshape(-1, 1), Xt, yt, threshold=0.5, loss=loss)
        assert acc == O.calculate_accuracy(w, Xt, yt, 0.5, loss)
        if wf == "ehrm":
            np.testing.assert_allclose(calculate_statistics(w.reshape(-1, 1), Xt, yt, gt),
                                       O.calculate_statistics(w, Xt, yt, gt), rtol=1e-12, atol=1e-13)
        s.engine.close()


def test_full_size_properties():
    """BASELINE config-2 sized z-step (n = 1M): size-independent properties instead of an oracle run."""
    from rbl_b200 import _cabi
    from rbl_b200.engine import AdmmEngine

    n = 1_000_000
    rng = np.random.default_rng(1)
    sig = O.spectrum("superquantile", n, [0.8])
    e = AdmmEngine(np.zeros((n, 2)), np.ones(n), "binary_cross_entropy", sig)
    m = rng.normal(size=n)
    md = e.vec(m)
    _cabi.check(e.lib.rbl_sort_margins(e.h, md.data_ptr(), e.m_sorted.data_ptr(), e.perm.data_ptr(), e._stream()))
    ms = e.m_sorted.cpu().numpy()
    perm = e.perm.cpu().numpy().astype(np.int64)
    assert np.all(np.diff(ms) >= 0)                                  # sortedness
    assert np.array_equal(np.sort(perm), np.arange(n))               # a permutation
    np.testing.assert_array_equal(ms, m[perm])                       # carried indices
    for rho in (1e-5, 1e-2):
        _cabi.check(e.lib.rbl_pav_prox(e.h, 0, e.m_sorted.data_ptr(), rho, e.z_sorted.data_ptr(), e._stream()))
        z = e.z_sorted.cpu().numpy()
        assert np.all(np.diff(z) >= 0)                               # isotonic
        zo = O.pav_prox("binary_cross_entropy", sig, ms, rho)        # the C oracle handles 1M in < 1 s
        assert np.max(np.abs(z - zo)) < 1e-12 * max(1.0, np.max(np.abs(zo)))
    e.close()
    # idempotence / fixed points at full size: with sigma = 0 the prox is the projection of an already sorted vector
    # onto the isotonic cone, i.e. the vector itself, bit for bit, on both PAV routes; the scatter then returns the
    # margins to row order exactly (z == m), so no row is active
    import ctypes
    e0 = AdmmEngine(np.zeros((n, 2)), np.ones(n), "binary_cross_entropy", np.zeros(n))
    for force_tree in (0, 1):
        _cabi.check(e0.lib.rbl_pav_config(e0.h, force_tree, None))
        e0.set_state(w=np.zeros(2), z=np.zeros(n), lam=-m * 0.5)      # margins = D w - lam / rho = m at rho = 0.5
        e0.z_step(0.5)
        np.testing.assert_array_equal(e0.z_sorted.cpu().numpy(), np.sort(m))
        np.testing.assert_array_equal(e0.z.cpu().numpy(), m)
        if e0.w_mode == "gram":
            cnt = ctypes.c_int32(-1)
            _cabi.check(e0.lib.rbl_active_count(e0.h, ctypes.byref(cnt), e0._stream()))
            assert cnt.value == 0
    e0.close()


def test_smooth_admm_vs_oracle_and_reference(golden_dir):
    """smoothADMMmethod on the device path vs the oracle (first 30 iterations, before
    the t -> 0 schedule makes the trajectory chaotic) and the reference's end state (loose, see test_oracle)."""
    from src.optim.algorithms import Optimizer, smoothADMMmethod

    g = _load(golden_dir, "trajectory.npz")
    d = _load(golden_dir, "data_300x40.npz")
    s = smoothADMMmethod(d["X"], d["y"], "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=40, tol=1e-6)
    o = O.OracleSmoothADMM(d["X"], d["y"], "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=40, tol=1e-6)
    for i in range(30):
        with contextlib.redirect_stdout(io.StringIO()):
            Optimizer.main_loop(s, i, 0.0, False)
        o.step()
        if i >= 17:
            s.t = max(s.t * 0.9, 1e-9) % np.power(s.rho, -0.1) * np.power(i, -0.1)
            o.t = max(o.t * 0.9, 1e-9) % np.power(o.rho, -0.1) * np.power(i, -0.1)
        assert _rel(s.w, o.w) < 1e-8 and _rel(s.z, o.z) < 1e-8, (i, _rel(s.w, o.w))
    s.engine.close()
    s2 = smoothADMMmethod(d["X"], d["y"], "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=40, tol=1e-6)
    with contextlib.redirect_stdout(io.StringIO()):
        w = s2.main_loop(verbose=False)
    ref = g["sadmm_erm_l1_w_final"]
    assert abs(float(s2.t) - float(g["sadmm_erm_l1_t_final"])) < 1e-12 * float(s2.t)
    assert _rel(w, ref) < 2e-3
    obj = s2.objective.get_arrogate_loss(torch.from_numpy(s2.w).double())
    assert abs(obj - float(g["sadmm_erm_l1_obj"])) < 5e-5
    s2.engine.close()


@pytest.mark.parametrize("tag,n,d,wf,args,loss,B,kw,intercept", [
    ("C2 twin", 6000, 100, "superquantile", [0.8], "binary_cross_entropy", None, dict(l1_reg=0.01), False),
    ("C2 twin l2", 6000, 100, "superquantile", [0.8], "binary_cross_entropy", None, dict(l2_reg=0.01), False),
    ("C3 twin", 6000, 50, "ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01), False),
    ("C4 twin hinge", 2400, 200, "aorr", [0.2, 0.8], "hinge", None, dict(l2_reg=1e-4), True),
    ("C4 twin bce", 2400, 200, "aorr", [0.2, 0.8], "binary_cross_entropy", None, dict(l2_reg=1e-4), True),
    ("extremile", 3000, 40, "extremile", [2.0], "hinge", None, dict(l1_reg=0.05), False),
    # run_AoRR_fixed.py:109-155: aorr_dc with [k, m] as counts (titanic-like: n ~ 800, k = 500, m = 10), intercept
    ("AoRR fixed hinge", 800, 12, "aorr_dc", [500, 10], "hinge", None, dict(l2_reg=1e-4), True),
    ("AoRR fixed bce", 800, 12, "aorr_dc", [500, 10], "binary_cross_entropy", None, dict(l2_reg=1e-4), True),
])
def test_reduced_size_twins_of_baseline_configs(tag, n, d, wf, args, loss, B, kw, intercept, w_mode):
    """Reduced-n twins of BASELINE configs 2-4 on planted data, 30 ADMM iterations in lockstep
    with the oracle: every iteration starts from the device state and must land within 1e-9 (w, z) unless
    the inner solver took a different branch on a rounding-level near-tie (then ~ the inner tolerance)."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    rng = np.random.default_rng(len(tag) * 1000 + n)
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:6] = rng.normal(size=6)
    y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
    if intercept:  # run_AoRR_ratio.py:40-41 appends a column of ones (odd d: padded leading dimension)
        X = np.hstack([X, np.ones((n, 1))])
    s = ADMMmethod(X, y, wf, loss, B=B, args=args, max_iter=30, tol=1e-9, **kw)
    o = O.OracleADMM(X, y, wf, loss, B=B, args=args, max_iter=30, tol=1e-9, small_lasso=False, **kw)
    flips = []
    for i in range(30):
        o.w, o.z, o.lam, o.rho = s.w.reshape(-1).copy(), s.z.reshape(-1).copy(), s.lagrangian.reshape(-1).copy(), s.rho
        with contextlib.redirect_stdout(io.StringIO()):
            Optimizer.main_loop(s, i, 0.0, False)
        o.step()
        ew, ez = _rel(s.w, o.w), _rel(s.z, o.z)
        assert ez < 1e-9, (tag, i, ez)
        if ew >= 1e-9:
            # allowed ONLY when the inner solver demonstrably took another branch on a rounding-level near-tie: the
            # device's and the oracle's inner iteration / evaluation counts must differ, and the two answers must
            # still agree to the inner solver's own tolerance
            if "l1_reg" in kw:
                hi, hd = (ctypes.c_int32 * 8)(), (ctypes.c_double * 4)()
                s.engine.lib.rbl_fista_poll(s.engine.h, s.engine._stream(), hi, hd)
                dev_info, cpu_info = (int(hi[1]), int(hi[3])), tuple(o.last_fista_info[:2])
            else:
                dev_info, cpu_info = (s.last_info["nit"], s.last_info["nfev"]), o.last_lbfgs_info
            flips.append({"tag": tag, "w_mode": w_mode, "iteration": i, "rel_w": ew, "device": dev_info,
                          "oracle": cpu_info})
            assert dev_info != cpu_info, ("w differs without a recorded branch difference", flips[-1])
            assert ew < 5e-3, (tag, i, ew)
    # the escape hatch is counted and logged (gpurun_out/twin_flips.jsonl when the directory exists)
    print(f"[twin] {tag} ({w_mode}): {len(flips)} inner-solver branch flips in 30 iterations {flips}")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        with open(os.path.join(out_dir, "twin_flips.jsonl"), "a") as f:
            f.write(json.dumps({"tag": tag, "w_mode": w_mode, "flips": flips}, default=str) + "\n")
    assert len(flips) <= 2, (tag, flips)
    obj = s.objective.get_arrogate_loss(torch.from_numpy(s.w).double())
    o.w = s.w.reshape(-1).copy()
    assert abs(obj - o.objective()) < 1e-11 * abs(obj)
    s.engine.close()


def test_native_loop_equals_per_iteration_loop(w_mode):
    """ADMMmethod.main_loop runs the captured iteration graph in the library's native loop (rbl_admm_run: stop test,
    rho schedule, lam in C); stepping the same solve one Optimizer.main_loop call at a time must give the same
    iterates bit for bit, the same rho and the same stopping iteration."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    rng = np.random.default_rng(11)
    n, d = 4000, 60
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:5] = rng.normal(size=5)
    y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
    kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=60,
              tol=1e-3)
    a = ADMMmethod(X, y, **kw)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        wa = a.main_loop(verbose=True)
    b = ADMMmethod(X, y, **kw)
    it_b = None
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(60):
            if Optimizer.main_loop(b, i, 0.0, False):
                it_b = i
                break
    np.testing.assert_array_equal(wa, b.w)
    np.testing.assert_array_equal(a.z, b.z)
    assert float(a.rho) == float(b.rho)
    out = buf.getvalue()
    assert "iter_num= 0 " in out and "iter_num= 10 " in out          # verbose prints every 10th iteration
    if it_b is not None:
        assert f"iter_num= {it_b} " in out and "algorithm converges within tolerance" in out
    if w_mode == "gram":
        assert a.engine._graph is not None and a.engine._graph_replays > 0
    a.engine.close()
    b.engine.close()


@pytest.mark.parametrize("wf,args,loss,kw,n,d", [
    ("superquantile", [0.8], "binary_cross_entropy", dict(l2_reg=0.01), 5000, 80),
    ("aorr", [0.2, 0.8], "hinge", dict(l2_reg=1e-4), 2400, 201),
    ("erm", None, "binary_cross_entropy", dict(l2_reg=1e-4), 3000, 33),
])
def test_native_l2_loop_equals_per_iteration_loop(w_mode, wf, args, loss, kw, n, d):
    """l2 problems: ADMMmethod.main_loop runs [graph: z-step + gradient pass] -> the library's L-BFGS-B -> [graph:
    dual step] in rbl_admm_run_l2; stepping Optimizer.main_loop one iteration at a time (same kernels, same L-BFGS-B,
    interpreter in between) must give the same iterates bit for bit, the same rho and the same stopping iteration —
    and both must track the oracle (scipy's L-BFGS-B) at 1e-9 per step."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    if w_mode != "gram":
        pytest.skip("the native l2 loop is a Gram-mode feature")
    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:5] = rng.normal(size=5)
    y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
    a = ADMMmethod(X, y, wf, loss, args=args, max_iter=45, tol=1e-3, **kw)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        wa = a.main_loop(verbose=True)
    b = ADMMmethod(X, y, wf, loss, args=args, max_iter=45, tol=1e-3, **kw)
    o = O.OracleADMM(X, y, wf, loss, args=args, max_iter=45, tol=1e-3, **kw)
    it_b = None
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(45):
            o.w, o.z, o.lam, o.rho = b.w.reshape(-1).copy(), b.z.reshape(-1).copy(), b.lagrangian.reshape(-1).copy(), b.rho
            done = Optimizer.main_loop(b, i, 0.0, False)
            o.step()
            assert _rel(b.w, o.w) < 1e-9 and _rel(b.z, o.z) < 1e-9, (i, _rel(b.w, o.w), _rel(b.z, o.z))
            if done:
                it_b = i
                break
    np.testing.assert_array_equal(wa, b.w)
    np.testing.assert_array_equal(a.z, b.z)
    assert float(a.rho) == float(b.rho)
    assert a.engine._graph_l2 is not None and getattr(a.engine, "_graph_dual", None) is not None
    assert a.engine._graph_dual_replays > 0 and a.last_info["solver"].startswith("librbl_b200")
    out = buf.getvalue()
    assert "iter_num= 0 " in out and "iter_num= 10 " in out
    if it_b is not None:
        assert f"iter_num= {it_b} " in out and "algorithm converges within tolerance" in out
    a.engine.close()
    b.engine.close()


def test_full_size_gram_route_equals_stream_route(w_mode, monkeypatch):
    """BASELINE config-2 row count (n = 1M; d = 32 keeps it small): eight ADMM iterations with the w-step on
    G = D^T D, the active-row gradient gather, the sparse dual pass and the replayed iteration graph must give the
    iterates of the formulation that streams all of D for every FISTA trial — the size-independent property that
    ties the fast route to the straightforward one at full n."""
    if w_mode != "gram":
        pytest.skip("runs both routes itself")
    from src.optim.algorithms import ADMMmethod, Optimizer

    rng = np.random.default_rng(3)
    n, d = 1_000_000, 32
    X = rng.standard_normal(size=(n, d))
    ws = np.zeros(d)
    ws[:4] = rng.normal(size=4)
    y = np.sign(X @ ws + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.01, args=[0.8], max_iter=8,
              tol=1e-12)
    res = {}
    monkeypatch.setenv("RBL_SPARSE_CAP", "16")  # default d // 16 = 2 would keep this small d on the dense dual pass
    for mode in ("gram", "stream"):
        monkeypatch.setenv("RBL_W_MODE", mode)
        s = ADMMmethod(X, y, **kw)
        assert s.engine.w_mode == mode
        with contextlib.redirect_stdout(io.StringIO()):
            for i in range(8):
                Optimizer.main_loop(s, i, 0.0, False)
        res[mode] = (s.w.copy(), s.z.copy(), float(s.rho), dict(s.engine.active_stats), dict(s.engine.dual_stats),
                     s.engine._graph is not None)
        s.engine.close()
    (wg, zg, rg, act, dual, graphed), (wst, zst, rst, _, _, _) = res["gram"], res["stream"]
    assert _rel(wg, wst) < 1e-9 and _rel(zg, zst) < 1e-9 and rg == rst
    assert graphed and act["gathered"] >= 4 and dual["sparse"] >= 4     # the fast route really ran
    assert act["rows"] < 0.8 * act["calls"] * n                        # and read fewer rows than full passes


@pytest.mark.parametrize("n,d", [(33, 1), (50, 7), (200, 3), (500, 60), (90, 60), (501, 1), (600, 7), (1000, 129),
                                 (700, 255), (5000, 2)])
def test_odd_and_tiny_shapes_in_lockstep_with_the_oracle(n, d, w_mode):
    """odd d (padded leading dimension), d = 1, n barely above 2d, tiny n: 12 ADMM iterations of the l1 path in
    lockstep with the oracle, both w-step formulations (persistent FISTA grid = min(#SMs, d) CTAs).  Shapes with
    n <= 500 and d <= 60 take the reference's small-problem branch (algorithms.py:194-197, scikit-learn's Lasso
    coordinate descent — rbl_lasso_cd_gram on the device, its restatement in the oracle), the others FISTA."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    rng = np.random.default_rng(n * 7 + d)
    X = rng.normal(size=(n, d))
    ws = rng.normal(size=d) * (rng.random(d) < 0.5)
    y = np.sign(X @ ws + 0.3 * rng.normal(size=n) + 1e-12).reshape(-1, 1)
    kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.02, args=[0.7], max_iter=12,
              tol=1e-12)
    s = ADMMmethod(X, y, **kw)
    o = O.OracleADMM(X, y, **kw)
    for i in range(12):
        o.w, o.z, o.lam, o.rho = s.w.reshape(-1).copy(), s.z.reshape(-1).copy(), s.lagrangian.reshape(-1).copy(), s.rho
        with contextlib.redirect_stdout(io.StringIO()):
            Optimizer.main_loop(s, i, 0.0, False)
        o.step()
        assert _rel(s.z, o.z) < 1e-9, (n, d, i)
        assert _rel(s.w, o.w) < 1e-9 or np.linalg.norm(s.w.reshape(-1) - o.w) < 1e-9, (n, d, i, _rel(s.w, o.w))
        if n <= 500 and d <= 60:
            assert s.last_info.get("mode") == "lasso_cd" and s.engine.lasso_cd_info()[0] == o.last_cd_sweeps
    s.engine.close()


@pytest.mark.parametrize("n,d", [(5000, 2050), (9000, 4096)])
def test_wide_problems_in_lockstep_with_the_oracle(n, d, w_mode):
    """d = 2050 (persistent FISTA with fewer candidates per sweep, G rows read from L2) and d = 4096 (state does not
    fit in shared memory: one launch per line-search trial): 6 iterations in lockstep with the oracle."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    rng = np.random.default_rng(d)
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:8] = rng.normal(size=8)
    y = np.sign(X @ ws + 0.1 * rng.normal(size=n)).reshape(-1, 1)
    kw = dict(weight_function="superquantile", loss="binary_cross_entropy", l1_reg=0.05, args=[0.8], max_iter=6,
              tol=1e-12)
    s = ADMMmethod(X, y, **kw)
    assert s.engine.w_mode == w_mode
    o = O.OracleADMM(X, y, small_lasso=False, **kw)
    for i in range(6):
        o.w, o.z, o.lam, o.rho = s.w.reshape(-1).copy(), s.z.reshape(-1).copy(), s.lagrangian.reshape(-1).copy(), s.rho
        with contextlib.redirect_stdout(io.StringIO()):
            Optimizer.main_loop(s, i, 0.0, False)
        o.step()
        assert _rel(s.z, o.z) < 1e-9, (n, d, i)
        assert _rel(s.w, o.w) < 1e-9 or np.linalg.norm(s.w.reshape(-1) - o.w) < 1e-9, (n, d, i, _rel(s.w, o.w))
    if w_mode == "gram":
        assert s.engine._persistent == (d <= 2050)
    s.engine.close()


def test_ehrm_candidate_choice_vs_reference_golden_and_oracle(golden_dir, w_mode):
    """a5: the reference's all-or-nothing choice between min(B, prox_{sigma_a}) and max(B, prox_{sigma_b})
    (PAV_cpt.py:222-226) made on the device from the two sums of rbl_ehrm_candidate_sums.  (a) the PAV_solver_CPT
    mirror over the reference-generated (B, rho, scale) grid where either candidate wins: same choice and z within
    1e-12 of the oracle, within the reference's own Newton tolerance of its output; (b) whole solves with B = 0.5
    and B = -1 (candidate 1 wins) in free-running lockstep with the oracle at 1e-9."""
    from src.optim.algorithms import ADMMmethod, Optimizer
    from src.util.PAV_cpt import PAV_solver_CPT

    g = _load(golden_dir, "zstep_ehrm.npz")
    sa, sb = g["sigma_a"], g["sigma_b"]
    won = {1: 0, 2: 0}
    if w_mode == "gram":  # the seam does not depend on the w-step formulation: once is enough
        for k in range(int(g["ncases"])):
            B, rho, _ = g[f"k{k}_par"]
            m = g[f"k{k}_m"]
            solver = PAV_solver_CPT(sa, sb, B, m, rho)
            z = solver.get_opt()
            zo, choice = O.ehrm_pav(sa, sb, B, m, rho, return_choice=True)
            f1, f2 = O.ehrm_candidate_sums(sa, sb, B, m, rho)
            assert solver.selected == choice, (k, solver.fvals, (f1, f2))
            assert abs(solver.fvals[0] - f1) <= 1e-12 * abs(f1) and abs(solver.fvals[1] - f2) <= 1e-12 * abs(f2)
            assert np.max(np.abs(z - zo)) <= 1e-12 * max(1.0, np.max(np.abs(zo))), (k, B, rho)
            assert np.max(np.abs(z - g[f"k{k}_ref_z"])) < 2e-8 * max(1.0, np.max(np.abs(zo))), (k, B, rho)
            won[choice] += 1
        assert won[1] >= 10 and won[2] >= 10
    d = _load(golden_dir, "data_300x40.npz")
    for tag in ("ehrm_B05_l2", "ehrm_Bm1_l2"):
        B = float(g[f"{tag}_B"])
        s = ADMMmethod(d["X"], d["y"], "ehrm", "binary_cross_entropy", B=B, l2_reg=0.01, max_iter=40, tol=1e-6)
        o = O.OracleADMM(d["X"], d["y"], "ehrm", "binary_cross_entropy", B=B, l2_reg=0.01, max_iter=40, tol=1e-6)
        for i in range(40):
            with contextlib.redirect_stdout(io.StringIO()):
                Optimizer.main_loop(s, i, 0.0, False)
            o.step()
            assert _rel(s.w, o.w) < 1e-9 and _rel(s.z, o.z) < 1e-9, (tag, i + 1, _rel(s.w, o.w), _rel(s.z, o.z))
            if f"{tag}_w_{i+1}" in g.files:
                assert _rel(s.w, g[f"{tag}_w_{i+1}"]) < (2e-8 if i + 1 <= 3 else 1e-5), (tag, i + 1)
        st = s.engine.ehrm_stats
        assert st["cand1"] > 0, (tag, st)                    # candidate 1 really ran on the device
        obj = s.objective.get_arrogate_loss(torch.from_numpy(s.w).double())
        assert abs(obj - float(g[f"{tag}_obj"])) < 1e-7
        s.engine.close()


def _config1_data(g):
    """run_SRM.py:21-36 — regenerated with scikit-learn's seeded generator and checked against the checksum stored
    with the reference's trajectory (tests/golden/c1_trajectory.npz)"""
    from sklearn.model_selection import train_test_split
    from src.util.load_data import get_data

    X, y = get_data("synthetic", num_row=10000, num_feature=1000, seed=17)
    Xtr, _, ytr, _ = train_test_split(X, y, test_size=0.4, random_state=17)
    cs = np.array([float(Xtr.sum()), float(np.abs(Xtr).sum()), float((Xtr * Xtr).sum()), float(Xtr[::7, ::11].sum())])
    if not (np.allclose(cs, g["x_checksum"], rtol=1e-12, atol=1e-9) and float(ytr.sum()) == float(g["y_sum"])):
        pytest.skip("scikit-learn's generator gives other data here than where the golden was made")
    return Xtr, ytr


def test_config1_solve_vs_reference_trajectory(golden_dir, w_mode):
    """BASELINE configs[0] at its stated size, against the reference as shipped (+ float64 FISTA): 6000 x 1000
    ERM / BCE / l1 = 0.01, tol 1e-6 — the reference's own 126-iteration trajectory (oracle/gen_golden.py::config1).
    Free-running: same stopping iteration, residual norms and rho per iteration, iterates within the reference's
    inner-solver tolerance, objective to 1e-9."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    g = _load(golden_dir, "c1_trajectory.npz")
    X, y = _config1_data(g)
    s = ADMMmethod(X, y, "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=200, tol=1e-6)
    n_ref = int(g["iterations"])
    it = None
    for i in range(200):
        rho_used = float(s.rho)
        with contextlib.redirect_stdout(io.StringIO()):
            done = Optimizer.main_loop(s, i, 0.0, False)
        if i < n_ref:
            assert abs(rho_used - float(g["rho"][i])) <= 1e-12 * rho_used, i
            # the reference's z comes from a Newton iteration stopped at ||delta|| < 1e-6 (individual_solver.py:103):
            # its primal residual is reproduced to ~1e-8 on most iterations, 6e-4 at worst; its w to 3e-11
            assert abs(s.primal_feasibility - float(g["primal"][i])) <= 2e-3 * float(g["primal"][i]), i
        if f"w_{i+1}" in g.files:
            assert _rel(s.w, g[f"w_{i+1}"]) < 1e-8, (i + 1, _rel(s.w, g[f"w_{i+1}"]))
        if done:
            it = i + 1
            break
    assert it == n_ref, (it, n_ref)
    obj = s.objective.get_arrogate_loss(torch.from_numpy(s.w).double())
    assert abs(obj - float(g["objective"])) < 1e-10 * abs(obj), (obj, float(g["objective"]))
    s.engine.close()


def test_config1_lockstep_with_oracle(golden_dir, w_mode):
    """configs[0] again, first 40 iterations in free-running lockstep with the oracle (north_star: 1e-9)."""
    from src.optim.algorithms import ADMMmethod, Optimizer

    g = _load(golden_dir, "c1_trajectory.npz")
    X, y = _config1_data(g)
    s = ADMMmethod(X, y, "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=200, tol=1e-6)
    o = O.OracleADMM(X, y, "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=200, tol=1e-6)
    for i in range(40):
        with contextlib.redirect_stdout(io.StringIO()):
            Optimizer.main_loop(s, i, 0.0, False)
        o.step()
        assert _rel(s.w, o.w) < 1e-9 and _rel(s.z, o.z) < 1e-9, (i + 1, _rel(s.w, o.w), _rel(s.z, o.z))
        assert float(s.rho) == float(o.rho)
    s.engine.close()


@pytest.mark.parametrize("kind", ["aorr_ratio_l2", "srm_spectra_l1", "ehrm_B_l2"])
def test_sweep_over_a_shared_design_matrix(w_mode, kind):
    """ADMMmethod(_share=first): the sweeps of run_AoRR_ratio.py:32-46 (four (k, m) ratios, hinge, l2), of the SRM
    drivers (spectra x l1) and an EHRM B grid, all over ONE device-resident D, G = D^T D and D^T.  Every shared
    solver must give, bit for bit, what a solver built from (X, y) on its own gives, hold the same D / G storage as
    the first, and track the oracle; the solvers are advanced interleaved to show they do not share state."""
    from src.optim.algorithms import ADMMmethod

    if w_mode != "gram":
        pytest.skip("sharing G is a Gram-mode feature")
    rng = np.random.default_rng(7)
    n, d = 2600, 61
    X = rng.normal(size=(n, d))
    ws = np.zeros(d)
    ws[:6] = rng.normal(size=6)
    y = np.sign(X @ ws + 0.2 * rng.normal(size=n)).reshape(-1, 1)
    if kind == "aorr_ratio_l2":
        X = np.hstack([X, np.ones((n, 1))])  # the driver's intercept column (run_AoRR_ratio.py:26)
        cfgs = [dict(weight_function="aorr", loss="hinge", args=a, l2_reg=1e-4)
                for a in ([0.1, 0.9], [0.2, 0.8], [0.3, 0.7], [0.05, 0.6])]
    elif kind == "srm_spectra_l1":
        cfgs = [dict(weight_function="superquantile", loss="binary_cross_entropy", args=[0.8], l1_reg=0.01),
                dict(weight_function="extremile", loss="binary_cross_entropy", args=[2.5], l1_reg=0.003),
                dict(weight_function="esrm", loss="hinge", args=[2.0], l1_reg=0.02),
                dict(weight_function="erm", loss="binary_cross_entropy", l2_reg=1e-3)]
    else:
        cfgs = [dict(weight_function="ehrm", loss="binary_cross_entropy", B=b, l2_reg=0.01) for b in (-5, -1, 0.5)]
    iters = 24
    first = ADMMmethod(X, y, max_iter=iters, tol=1e-12, **cfgs[0])
    shared = [first] + [ADMMmethod(None, None, max_iter=iters, tol=1e-12, _share=first, **c) for c in cfgs[1:]]
    for s in shared[1:]:
        assert s.engine.D.data_ptr() == first.engine.D.data_ptr()
        assert s.engine.gram().data_ptr() == first.engine.gram().data_ptr()
        assert s.num_row == n and s.num_feature == X.shape[1]
    with contextlib.redirect_stdout(io.StringIO()):
        for lo in range(0, iters, 8):           # interleaved: 8 iterations of each solver in turn
            for s in shared:
                s.advance(lo, 8)
    for c, s in zip(cfgs, shared):
        own = ADMMmethod(X, y, max_iter=iters, tol=1e-12, **c)
        with contextlib.redirect_stdout(io.StringIO()):
            for lo in range(0, iters, 8):
                own.advance(lo, 8)
        np.testing.assert_array_equal(s.w, own.w)
        np.testing.assert_array_equal(s.z, own.z)
        np.testing.assert_array_equal(s.lagrangian, own.lagrangian)
        assert float(s.rho) == float(own.rho)
        o = O.OracleADMM(X, y, max_iter=iters, tol=1e-12, **c)
        for _ in range(iters):
            o.step()
        assert _rel(s.w, o.w) < 1e-7 and _rel(s.z, o.z) < 1e-7, (c, _rel(s.w, o.w), _rel(s.z, o.z))
        own.engine.close()
    for s in reversed(shared):
        s.engine.close()


def test_fista_session_reuses_the_uploaded_matrix(golden_dir, w_mode):
    """FistaSession (extension of the fast_lasso seam for callers that loop over one X): same results as FISTA(),
    bit for bit, for several (beta, y, lam) on the same matrix, with one engine / one upload behind them."""
    from src.util.fast_lasso import FISTA, FistaSession

    g = _load(golden_dir, "fista.npz")
    d = _load(golden_dir, "data_300x40.npz")
    D = -d["y"] * d["X"]
    rng = np.random.default_rng(4)
    cases = [(g["w0"], g["b"], np.float64(0.5)), (g["w0"] * 0.0, g["b"] + 0.1 * rng.normal(size=g["b"].shape), 20.0),
             (rng.normal(size=g["w0"].shape) * 0.01, g["b"], np.float64(300.0))]
    for dtype in (torch.float64, torch.float32):
        with FistaSession(D, dtype=dtype) as sess:
            eng = sess.engine
            for beta, b, lam in cases:
                w1, i1 = sess(beta, b, lam, np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000, return_info=True)
                w2, i2 = FISTA(beta, D, b, lam, np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000, dtype=dtype,
                               return_info=True)
                np.testing.assert_array_equal(w1, w2)
                assert w1.dtype == (np.float32 if dtype == torch.float32 else np.float64)
                assert i1["iters"] == i2["iters"] and i1["L"] == i2["L"]
                assert sess.engine is eng
        with pytest.raises(RuntimeError):
            sess(cases[0][0], cases[0][1], cases[0][2], np.float32(17), np.float32(2.5))
