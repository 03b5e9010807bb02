"""Generate tests/golden/*.npz from the *reference itself* (shimmed import, oracle/ref_shim.py).

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

TEST INFRASTRUCTURE ONLY.  The reference ships no tests or golden vectors for this path, so
these fixtures are the pin: every array under `ref_*` keys is an output of the unmodified
reference code (plus the two-line get_opt shim), produced with numpy 2.3.5 / scipy 1.18.1 /
scikit-learn 1.9.0 / torch 2.11.0 on CPU.  Inputs are stored next to the outputs so the tests
need neither the reference nor scikit-learn's generator to be bit-stable.
"""
import os
import time
import warnings

import numpy as np

warnings.simplefilter("ignore")

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    import torch
    from sklearn.model_selection import train_test_split

    from oracle import ref_shim

    ns = ref_shim.load(fista_dtype=torch.float64)
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- dataset: load_data.py:101-116 + the drivers' split (run_SRM.py:26) -------------
    X, y = ns.load_data.get_data("synthetic", num_row=500, num_feature=40, seed=17)
    Xtr, Xte, ytr, yte = train_test_split(X, y, test_size=0.4, random_state=17)
    n, d = Xtr.shape
    D = -ytr * Xtr
    np.savez_compressed(os.path.join(OUT, "data_300x40.npz"), X=Xtr, y=ytr, X_test=Xte, y_test=yte)

    # ---- spectra: objective.py:97-187 ---------------------------------------------------
    spec = {}
    for name, args in [("erm", None), ("extremile", [2.0]), ("superquantile", [0.8]), ("superquantile", [0.5]),
                       ("esrm", [1.5]), ("aorr", [0.2, 0.8]), ("aorr", [0.1, 0.9]), ("aorr_dc", [60, 20]),
                       ("ehrm", None)]:
        for nn in (97, 300):
            wf = ns.objective.get_weights(name, args)
            key = f"{name}|{'' if args is None else ','.join(map(str, args))}|{nn}"
            if isinstance(wf, tuple):
                spec[key + "|a"] = wf[0](nn).numpy()
                spec[key + "|b"] = wf[1](nn).numpy()
            else:
                spec[key] = wf(nn).numpy()
    np.savez_compressed(os.path.join(OUT, "spectra.npz"), **spec)

    # ---- z-step: algorithms.py:88-106 (margins -> sort -> PAV prox -> scatter) ----------
    zc = {}
    cases = [
        ("erm", None, "binary_cross_entropy", None, 1e-3),
        ("erm", None, "binary_cross_entropy", None, 1e-5),
        ("superquantile", [0.8], "binary_cross_entropy", None, 1e-5),
        ("superquantile", [0.8], "binary_cross_entropy", None, 1e-2),
        ("extremile", [2.0], "binary_cross_entropy", None, 1e-3),
        ("esrm", [1.5], "binary_cross_entropy", None, 1e-4),
        ("aorr", [0.2, 0.8], "binary_cross_entropy", None, 1e-3),
        ("aorr", [0.2, 0.8], "hinge", None, 1e-2),       # reference bisection: loose pin only
        ("superquantile", [0.8], "hinge", None, 1.0),     # idem
        ("ehrm", None, "binary_cross_entropy", -5, 1e-3),
        ("ehrm", None, "binary_cross_entropy", -5, 1e-1),
    ]
    for ci, (wf, args, loss, B, rho) in enumerate(cases):
        s = ns.algorithms.ADMMmethod(Xtr, ytr, wf, loss, l2_reg=0.01, B=B, args=args)
        s.w = rng.normal(size=(d, 1)) * 0.3
        s.lagrangian = rng.normal(size=(n, 1)) * rho
        s.rho = rho
        z = s.z_subproblem().reshape(-1)
        zc[f"c{ci}_w"] = s.w.reshape(-1)
        zc[f"c{ci}_lam"] = s.lagrangian.reshape(-1)
        zc[f"c{ci}_ref_z"] = z
        zc[f"c{ci}_meta"] = np.array([wf, "" if args is None else ",".join(map(str, args)), loss,
                                      "" if B is None else str(B), repr(rho)])
    zc["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "zstep.npz"), **zc)

    # ---- element prox: individual_solver.py:112-130 -------------------------------------
    m = np.sort(rng.normal(size=200) * 3)
    sg = np.abs(rng.normal(size=200)) / 200
    pr = {"m": m, "sigma": sg}
    for rho in (1e-4, 1e-2, 1.0):
        pr[f"ref_bce_{rho}"] = ns.individual_solver.individual_solver("binary_cross_entropy", sg, rho, m)
        pr[f"ref_hinge_{rho}"] = ns.individual_solver.individual_solver("hinge", sg, rho, m)
    np.savez_compressed(os.path.join(OUT, "prox.npz"), **pr)

    # ---- FISTA: fast_lasso.py:22-69 as called at algorithms.py:199-201 -------------------
    fc = {"b": rng.normal(size=n), "w0": rng.normal(size=d) * 0.01}
    for lam in (0.5, 20.0, 300.0):
        fc[f"ref_f64_{lam}"] = ns.FISTA(fc["w0"], D, fc["b"], lam, np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000)
        fc[f"ref_f64_np_{lam}"] = ns.FISTA(fc["w0"], D, fc["b"], np.float64(lam), np.float32(17), np.float32(2.5), tol=7e-5, max_iter=5000)
        fc[f"ref_f32_{lam}"] = ns._orig_fista(fc["w0"], D, fc["b"], lam, np.float32(17), np.float32(2.5), tol=7e-5,
                                              max_iter=5000, dtype=torch.float32)
        torch.set_default_dtype(torch.float32)
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, "fista.npz"), **fc)

    # ---- l2 w-step: w_LBFGS.py:31-53 -----------------------------------------------------
    lc = {"z": rng.normal(size=n), "lam": rng.normal(size=n) * 1e-3, "w0": rng.normal(size=d) * 0.01}
    DTD = D.T @ D
    for rho in (1e-5, 1e-2, 1.0):
        lc[f"ref_{rho}"] = ns.w_lbfgs.w_solver(2, lc["w0"].reshape(-1, 1), lc["z"].reshape(-1, 1),
                                               lc["lam"].reshape(-1, 1), rho, DTD, D, 0.01).reshape(-1)
    np.savez_compressed(os.path.join(OUT, "l2step.npz"), **lc)

    # ---- objective: objective.py:71-87 ---------------------------------------------------
    oc = {"w": rng.normal(size=d) * 0.2}
    for wf, args, loss, B, kw in [("erm", None, "binary_cross_entropy", None, dict(l1_reg=0.01)),
                                  ("superquantile", [0.8], "binary_cross_entropy", None, dict(l2_reg=0.01)),
                                  ("aorr", [0.2, 0.8], "hinge", None, dict(l2_reg=1e-4)),
                                  ("ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01)),
                                  ("esrm", [1.5], "hinge", None, dict(l1_reg=0.1))]:
        ob = ns.objective.rankbasedObjective(torch.from_numpy(Xtr.copy()), torch.from_numpy(ytr.copy()), wf, loss,
                                             kw.get("l2_reg"), kw.get("l1_reg"), B, None, args)
        oc[f"ref_{wf}_{loss}"] = np.array(ob.get_arrogate_loss(torch.from_numpy(oc["w"].reshape(-1, 1))))
    np.savez_compressed(os.path.join(OUT, "objective.npz"), **oc)

    # ---- trajectories: ADMMmethod.main_loop, algorithms.py:119-164,209-216 ---------------
    tr = {}
    runs = [
        ("erm_l1", "erm", None, "binary_cross_entropy", None, dict(l1_reg=0.01)),
        ("erm_l2", "erm", None, "binary_cross_entropy", None, dict(l2_reg=1e-4)),
        ("sq_l1", "superquantile", [0.8], "binary_cross_entropy", None, dict(l1_reg=0.01)),
        ("sq_l2", "superquantile", [0.8], "binary_cross_entropy", None, dict(l2_reg=0.01)),
        ("extremile_l2", "extremile", [2.0], "binary_cross_entropy", None, dict(l2_reg=0.01)),
        ("aorr_bce_l2", "aorr", [0.2, 0.8], "binary_cross_entropy", None, dict(l2_reg=1e-4)),
        ("ehrm_l2", "ehrm", None, "binary_cross_entropy", -5, dict(l2_reg=0.01)),
    ]
    # l1 runs on 300x40 take the reference's sklearn-Lasso branch (n<=500 and d<=60,
    # algorithms.py:194-197); the FISTA branch needs a bigger problem: 600x64.
    X2, y2 = ns.load_data.get_data("synthetic", num_row=1000, num_feature=64, seed=17)
    X2tr, _, y2tr, _ = train_test_split(X2, y2, test_size=0.4, random_state=17)
    np.savez_compressed(os.path.join(OUT, "data_600x64.npz"), X=X2tr, y=y2tr)
    runs += [("erm_l1_fista", "erm", None, "binary_cross_entropy", None, dict(l1_reg=0.01)),
             ("sq_l1_fista", "superquantile", [0.8], "binary_cross_entropy", None, dict(l1_reg=0.01))]
    snaps = (1, 2, 3, 10, 40)
    for tag, wf, args, loss, B, kw in runs:
        Xa, ya = (X2tr, y2tr) if tag.endswith("_fista") else (Xtr, ytr)
        s = ns.algorithms.ADMMmethod(Xa, ya, wf, loss, B=B, args=args, max_iter=40, tol=1e-6, **kw)
        t0 = time.time()
        for i in range(40):
            with ref_shim.quiet():
                done = ns.algorithms.Optimizer.main_loop(s, i, t0, False)
            if (i + 1) in snaps:
                tr[f"{tag}_w_{i+1}"] = np.asarray(s.w, dtype=np.float64).reshape(-1)
                tr[f"{tag}_z_{i+1}"] = s.z.reshape(-1)
                tr[f"{tag}_lam_{i+1}"] = s.lagrangian.reshape(-1)
                tr[f"{tag}_rho_{i+1}"] = np.array(float(s.rho))
            if done:
                break
        tr[f"{tag}_obj"] = np.array(s.objective.get_arrogate_loss(torch.from_numpy(s.w).double()))
        tr[f"{tag}_meta"] = np.array([wf, "" if args is None else ",".join(map(str, args)), loss,
                                      "" if B is None else str(B), repr(kw)])
        print(tag, "iters", i + 1, "obj", float(tr[f"{tag}_obj"]))
    # ---- smoothADMMmethod (algorithms.py:223-263): full run incl. t schedule and final soft-threshold
    sm = ns.algorithms.smoothADMMmethod(Xtr, ytr, "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=40, tol=1e-6)
    with ref_shim.quiet():
        w_s = sm.main_loop(verbose=False)
    tr["sadmm_erm_l1_w_final"] = np.asarray(w_s, dtype=np.float64).reshape(-1)
    tr["sadmm_erm_l1_t_final"] = np.array(float(sm.t))
    tr["sadmm_erm_l1_obj"] = np.array(sm.objective.get_arrogate_loss(torch.from_numpy(sm.w).double()))
    print("sadmm", float(tr["sadmm_erm_l1_obj"]), float(sm.t))
    np.savez_compressed(os.path.join(OUT, "trajectory.npz"), **tr)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))




def metrics():
    """tests/golden/metrics.npz: calculate_acc.py / fair_metric.py of the reference on seeded inputs
    (`python -m oracle.gen_golden metrics`; does not touch the other fixtures)."""
    import importlib
    import sys

    from oracle import ref_shim

    ref_shim.load()
    sys.path.insert(0, ref_shim.REF_ROOT)
    try:
        acc_mod = importlib.import_module("src.util.calculate_acc")
        fair_mod = importlib.import_module("src.util.fair_metric")
    finally:
        sys.path.remove(ref_shim.REF_ROOT)
    rng = np.random.default_rng(20261019)
    n, d = 700, 24
    X = rng.normal(size=(n, d))
    w_true = rng.normal(size=d)
    y = np.where(X @ w_true + 0.8 * rng.normal(size=n) > 0, 1, -1).reshape(-1, 1)
    group = (rng.random(n) < 0.35 + 0.2 * (y.reshape(-1) > 0)).astype(int)   # group membership correlated with y
    out = {"X": X, "y": y, "group": group}
    for k, scale in enumerate((1.0, 0.2, 3.0)):
        w = (w_true * scale + 0.5 * rng.normal(size=d)).reshape(-1, 1)
        out[f"w{k}"] = w
        for thr in (0.5, 0.3, 0.8):
            out[f"ref_acc_bce_{k}_{thr}"] = np.array(acc_mod.calculate_accuracy(w, X, y, threshold=thr))
            out[f"ref_stats_{k}_{thr}"] = np.array(fair_mod.calculate_statistics(w, X, y, group, threshold=thr),
                                                   dtype=np.float64)
        out[f"ref_acc_hinge_{k}"] = np.array(acc_mod.calculate_accuracy(w, X, y, loss="hinge"))
    out["nw"] = np.array(3)
    # split_group.py:3-25 (run_EHRM.py:25): row ids recovered by splitting X = arange(n)
    sys.path.insert(0, ref_shim.REF_ROOT)
    try:
        split_mod = importlib.import_module("src.util.split_group")
    finally:
        sys.path.remove(ref_shim.REF_ROOT)
    ids = np.arange(n).reshape(-1, 1)
    tr, te, ytr, yte, gtr, gte = split_mod.train_test_split_group(ids, y, group, test_size=0.4, random_state=17)
    out["ref_split_train_ids"], out["ref_split_test_ids"] = tr.reshape(-1), te.reshape(-1)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics.npz", os.path.getsize(os.path.join(OUT, "metrics.npz")))


def trajectories_extra():
    """tests/golden/trajectory_extra.npz: more forty-iteration runs of the reference's ADMMmethod.main_loop on the
    datasets already committed (`python -m oracle.gen_golden extra`; the other fixtures are not touched):
    the spectra the first set does not cover in a full loop (esrm, aorr_dc) and extremile on the FISTA branch."""
    import torch

    from oracle import ref_shim

    ns = ref_shim.load(fista_dtype=torch.float64)
    d1 = np.load(os.path.join(OUT, "data_300x40.npz"))
    d2 = np.load(os.path.join(OUT, "data_600x64.npz"))
    runs = [
        ("esrm_l2", "esrm", [1.5], "binary_cross_entropy", None, dict(l2_reg=0.01)),
        ("aorr_dc_bce_l2", "aorr_dc", [150, 30], "binary_cross_entropy", None, dict(l2_reg=1e-4)),
        ("sq05_l2", "superquantile", [0.5], "binary_cross_entropy", None, dict(l2_reg=0.01)),
        ("extremile_l1_fista", "extremile", [2.0], "binary_cross_entropy", None, dict(l1_reg=0.01)),
        ("esrm_l1_fista", "esrm", [1.5], "binary_cross_entropy", None, dict(l1_reg=0.01)),
    ]
    snaps = (1, 2, 3, 10, 40)
    tr = {}
    for tag, wf, args, loss, B, kw in runs:
        dd = d2 if tag.endswith("_fista") else d1
        s = ns.algorithms.ADMMmethod(dd["X"], dd["y"], wf, loss, B=B, args=args, max_iter=40, tol=1e-6, **kw)
        t0 = time.time()
        for i in range(40):
            with ref_shim.quiet():
                done = ns.algorithms.Optimizer.main_loop(s, i, t0, False)
            if (i + 1) in snaps:
                tr[f"{tag}_w_{i+1}"] = np.asarray(s.w, dtype=np.float64).reshape(-1)
                tr[f"{tag}_z_{i+1}"] = s.z.reshape(-1)
                tr[f"{tag}_lam_{i+1}"] = s.lagrangian.reshape(-1)
                tr[f"{tag}_rho_{i+1}"] = np.array(float(s.rho))
            if done:
                break
        tr[f"{tag}_obj"] = np.array(s.objective.get_arrogate_loss(torch.from_numpy(s.w).double()))
        tr[f"{tag}_meta"] = np.array([wf, "" if args is None else ",".join(map(str, args)), loss,
                                      "" if B is None else str(B), repr(kw)])
        print(tag, "iters", i + 1, "obj", float(tr[f"{tag}_obj"]))
    np.savez_compressed(os.path.join(OUT, "trajectory_extra.npz"), **tr)
    print("trajectory_extra.npz", os.path.getsize(os.path.join(OUT, "trajectory_extra.npz")))


def ehrm_select():
    """tests/golden/zstep_ehrm.npz (`python -m oracle.gen_golden ehrm`): the reference's EHRM z-step where its
    all-or-nothing candidate choice (PAV_cpt.py:222-226) goes either way.  (a) PAV_solver_CPT.get_opt on sorted
    margins for a grid of (B, rho, margin scale); (b) 40-iteration ADMMmethod trajectories with B = 0.5 and B = -1
    (candidate 1 wins at least part of the way) next to the shipped B = -5 of trajectory.npz."""
    import torch

    from oracle import ref_shim

    ns = ref_shim.load(fista_dtype=torch.float64)
    rng = np.random.default_rng(20261020)
    n = 400
    sa = ns.objective.get_weights("ehrm", None)[0](n).numpy().reshape(-1)
    sb = ns.objective.get_weights("ehrm", None)[1](n).numpy().reshape(-1)
    out = {"sigma_a": sa, "sigma_b": sb}
    k = 0
    for B in (-5.0, -1.0, 0.0, 0.5, 2.0, 5.0):
        for rho in (1e-4, 1e-2, 1.0):
            for scale in (0.01, 1.0, 5.0):
                m = np.sort(rng.normal(size=n) * scale)
                solver = ns.pav_cpt.PAV_solver_CPT(sa, sb, B, m, rho)
                out[f"k{k}_m"] = m
                out[f"k{k}_par"] = np.array([B, rho, scale])
                out[f"k{k}_ref_z"] = np.asarray(solver._orig_get_opt(), dtype=np.float64)
                k += 1
    out["ncases"] = np.array(k)
    d1 = np.load(os.path.join(OUT, "data_300x40.npz"))
    snaps = (1, 2, 3, 10, 40)
    for tag, B in (("ehrm_B05_l2", 0.5), ("ehrm_Bm1_l2", -1.0)):
        s = ns.algorithms.ADMMmethod(d1["X"], d1["y"], "ehrm", "binary_cross_entropy", B=B, l2_reg=0.01, max_iter=40,
                                     tol=1e-6)
        t0 = time.time()
        for i in range(40):
            with ref_shim.quiet():
                done = ns.algorithms.Optimizer.main_loop(s, i, t0, False)
            if (i + 1) in snaps:
                out[f"{tag}_w_{i+1}"] = np.asarray(s.w, dtype=np.float64).reshape(-1)
                out[f"{tag}_z_{i+1}"] = s.z.reshape(-1)
                out[f"{tag}_rho_{i+1}"] = np.array(float(s.rho))
            if done:
                break
        out[f"{tag}_obj"] = np.array(s.objective.get_arrogate_loss(torch.from_numpy(s.w).double()))
        out[f"{tag}_B"] = np.array(B)
        print(tag, "iters", i + 1, "obj", float(out[f"{tag}_obj"]))
    np.savez_compressed(os.path.join(OUT, "zstep_ehrm.npz"), **out)
    print("zstep_ehrm.npz", os.path.getsize(os.path.join(OUT, "zstep_ehrm.npz")))


def config1():
    """tests/golden/c1_trajectory.npz (`python -m oracle.gen_golden c1`): BASELINE configs[0] exactly, run by the
    reference itself — run_SRM.py:21-36: get_data("synthetic", 10000, 1000, seed=17), 60/40 split (random_state=17)
    -> 6000 x 1000, ERM, BCE, l1_reg = 0.01, ADMMmethod(tol=1e-6), float64 FISTA.  Stored: w at selected iterations,
    the residual norms and rho of EVERY iteration, the final objective, the reference's wall time here, and a
    checksum of the data (the tests regenerate X with scikit-learn's seeded generator and verify the checksum; the
    48 MB matrix itself is not committed)."""
    import torch
    from sklearn.model_selection import train_test_split

    from oracle import ref_shim

    ns = ref_shim.load(fista_dtype=torch.float64)
    X, y = ns.load_data.get_data("synthetic", num_row=10000, num_feature=1000, seed=17)
    Xtr, _, ytr, _ = train_test_split(X, y, test_size=0.4, random_state=17)
    out = {"x_checksum": np.array([float(Xtr.sum()), float(np.abs(Xtr).sum()), float((Xtr * Xtr).sum()),
                                   float(Xtr[::7, ::11].sum())]),
           "y_sum": np.array(float(ytr.sum())), "shape": np.array(Xtr.shape)}
    s = ns.algorithms.ADMMmethod(Xtr, ytr, "erm", "binary_cross_entropy", l1_reg=0.01, max_iter=200, tol=1e-6)
    snaps = (1, 2, 3, 5, 10, 20, 40, 80, 120)
    prim, dual, rhos = [], [], []
    t0 = time.time()
    for i in range(200):
        rho_used = float(s.rho)
        with ref_shim.quiet():
            import contextlib
            import io
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                w_prev = s.w.copy()
                done = ns.algorithms.Optimizer.main_loop(s, i, t0, False)
        Dw = s.D @ s.w
        # the multiplier was updated with the old rho: z - Dw = (lam_new - lam_old)/rho; recompute the norms directly
        prim.append(float(np.linalg.norm(s.z - Dw)))
        dual.append(float(np.linalg.norm(s.w - w_prev)))
        rhos.append(rho_used)
        if (i + 1) in snaps or done:
            out[f"w_{i+1}"] = np.asarray(s.w, dtype=np.float64).reshape(-1)
        if done:
            break
    wall = time.time() - t0
    out["iterations"] = np.array(i + 1)
    out["primal"], out["dual"], out["rho"] = np.array(prim), np.array(dual), np.array(rhos)
    out["w_final"] = np.asarray(s.w, dtype=np.float64).reshape(-1)
    out["objective"] = np.array(s.objective.get_arrogate_loss(torch.from_numpy(s.w).double()))
    out["ref_wall_s"] = np.array(wall)
    out["ref_cores"] = np.array(os.cpu_count() or 1)
    print("c1: iterations", i + 1, "objective", float(out["objective"]), "wall", wall)
    np.savez_compressed(os.path.join(OUT, "c1_trajectory.npz"), **out)
    print("c1_trajectory.npz", os.path.getsize(os.path.join(OUT, "c1_trajectory.npz")))


if __name__ == "__main__":
    import sys as _sys

    if "metrics" in _sys.argv[1:]:
        metrics()
    elif "extra" in _sys.argv[1:]:
        trajectories_extra()
    elif "ehrm" in _sys.argv[1:]:
        ehrm_select()
    elif "c1" in _sys.argv[1:]:
        config1()
    else:
        main(), metrics(), trajectories_extra(), ehrm_select(), config1()
