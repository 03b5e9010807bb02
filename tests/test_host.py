"""CPU: host-side logic — C-ABI surface, shard arithmetic, the CPU emulation of the device PAV (same
header as the kernels), loud failure without a GPU, and the N>1 collective plumbing on gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import rbl_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "admm-for-rank-based-loss_b200")


def test_cabi_library_loads_and_exports_every_declared_symbol():
    from rbl_b200 import _cabi, build

    build.build()
    hdr = open(os.path.join(ROOT, "include", "rbl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rbl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(build.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rbl_b200.h but not exported"
    assert sorted(_cabi.exported_symbols()) == declared      # the ctypes binding covers the whole header
    assert _cabi.load().rbl_version() == _cabi.ABI_VERSION == 2


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rbl_b200 import RblError
    from src.optim.algorithms import ADMMmethod

    with pytest.raises(RblError):
        ADMMmethod(np.zeros((8, 2)), np.ones((8, 1)), l1_reg=0.1)
    # the steps either side of the loop have no CPU path either
    from rbl_b200 import ingest
    from src.util.calculate_acc import calculate_accuracy
    from src.util.fair_metric import calculate_statistics
    from src.util.load_data import get_data

    with pytest.raises(RblError):
        calculate_accuracy(np.zeros((2, 1)), np.zeros((8, 2)), np.ones((8, 1)))
    with pytest.raises(RblError):
        calculate_statistics(np.zeros((2, 1)), np.zeros((8, 2)), np.ones((8, 1)), np.zeros(8, dtype=int))
    with pytest.raises(RblError):
        ingest.to_device_padded(np.zeros((8, 3)))
    with pytest.raises(RblError):
        get_data("synthetic", num_row=50, num_feature=20, seed=1, device="cuda")
    with pytest.raises(ValueError, match="is not supported"):   # argument errors come first, as in the reference
        calculate_accuracy(np.zeros((2, 1)), np.zeros((8, 2)), np.ones((8, 1)), loss="nope")
    X, y = get_data("synthetic", num_row=50, num_feature=20, seed=1)   # the reference's host recipe still works
    assert X.shape == (50, 20) and set(np.unique(y)) == {-1, 1} and abs(X.mean()) < 1e-12


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "libpav_oracle" not in src and "rbl_oracle" not in src, f"{f} links the oracle"


def test_shard_bounds_cover_rows_exactly():
    from rbl_b200.engine import shard_bounds

    for n in (1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _emul_lib():
    path = os.path.join(ROOT, "tests", "native", "libpav_emul.so")
    src = os.path.join(ROOT, "tests", "native", "pav_emul.cpp")
    hdrs = [os.path.join(PKG, "csrc", h) for h in ("pav_core.h", "prox_core.h", "lbfgs_core.h")]
    if not os.path.exists(path) or any(os.path.getmtime(p) > os.path.getmtime(path) for p in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(PKG, "csrc"),
                               "-o", path, src])
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.emul_pav.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_int, dp,
                             ctypes.POINTER(ctypes.c_int64)]
    lib.emul_pav_fewseg.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_int, dp,
                                    ctypes.POINTER(ctypes.c_int64)]
    ip = ctypes.POINTER(ctypes.c_int64)
    lib.emul_pav_fewseg_hinted.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_int, dp, ip,
                                           ip, ip]
    lib.emul_key.restype = ctypes.c_uint64
    lib.emul_key.argtypes = [ctypes.c_double]
    lib.emul_unkey.restype = ctypes.c_double
    lib.emul_unkey.argtypes = [ctypes.c_uint64]
    return lib


def test_device_pav_logic_matches_stack_pav_on_cpu():
    """csrc/pav_core.h (the code the kernels run) emulated on the host vs the oracle's stack PAV."""
    lib = _emul_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(11)
    for trial in range(1500):
        n = int(rng.integers(1, 400))
        loss = ["binary_cross_entropy", "hinge"][trial % 2]
        kind = trial % 5
        m = np.sort(rng.normal(size=n) * 10 ** rng.uniform(-3, 1.5))
        if kind == 0:
            sig = np.abs(rng.normal(size=n)) * (rng.random(n) > 0.3)
        elif kind == 1:
            sig = O.spectrum("superquantile", n, [rng.uniform(0.1, 0.95)])
        elif kind == 2:
            sig = O.spectrum("aorr", n, [0.2, 0.8]) if n >= 5 else np.ones(n) / n
        elif kind == 3:
            sig = O.spectrum("extremile", n, [rng.uniform(1, 4)])
        else:
            sig = O.spectrum("ehrm", n)[1]
        if trial % 7 == 0:
            m = np.round(m, 1)
        rho = 10 ** rng.uniform(-6, 1)
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        zo = O.pav_prox(loss, sig, m, rho)
        z = np.empty_like(m)
        lib.emul_pav(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m.ctypes.data_as(dp), rho, int(rng.integers(1, 8)),
                     z.ctypes.data_as(dp), None)
        assert np.max(np.abs(z - zo)) <= 1e-13 * max(1.0, np.max(np.abs(zo))), (trial, n, loss, kind)
        assert np.all(np.diff(z) >= 0)


def test_few_segment_pav_route_matches_stack_pav_on_cpu():
    """The few-segment route of the device PAV (level-0 ranges = runs of non-increasing sigma, merges over a lazily
    overlaid value array, csrc/pav_kernels.cu) emulated on the host from the same header vs the oracle's stack PAV
    — spectra with few runs (erm, superquantile, aorr, step functions) and, for the logic, arbitrary ones."""
    lib = _emul_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(23)
    for trial in range(1200):
        n = int(rng.integers(1, 500))
        loss = ["binary_cross_entropy", "hinge"][trial % 2]
        kind = trial % 6
        m = np.sort(rng.normal(size=n) * 10 ** rng.uniform(-3, 1.5))
        if kind == 0:
            sig = np.ones(n) / n
        elif kind == 1:
            sig = O.spectrum("superquantile", n, [rng.uniform(0.05, 0.97)])
        elif kind == 2:
            sig = O.spectrum("aorr", n, [0.2, 0.8]) if n >= 5 else np.ones(n) / n
        elif kind == 3:  # a few random steps up and down
            sig = np.repeat(np.abs(rng.normal(size=5)) * (rng.random(5) > 0.3), -(-n // 5))[:n]
        elif kind == 4:  # non-increasing pieces with a few jumps up
            sig = np.sort(np.abs(rng.normal(size=n)))[::-1].copy()
            for cut in rng.integers(0, n, size=3):
                sig[cut:] += abs(rng.normal())
        else:            # many runs: the route is not used on the device here, but the logic must still hold
            sig = O.spectrum("extremile", n, [rng.uniform(1, 4)])
        if trial % 7 == 0:
            m = np.round(m, 1)
        rho = 10 ** rng.uniform(-6, 1)
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        zo = O.pav_prox(loss, sig, m, rho)
        z = np.empty_like(m)
        runs = ctypes.c_int64(0)
        lib.emul_pav_fewseg(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m.ctypes.data_as(dp), rho,
                            int(rng.integers(1, 8)), z.ctypes.data_as(dp), ctypes.byref(runs))
        assert np.max(np.abs(z - zo)) <= 1e-13 * max(1.0, np.max(np.abs(zo))), (trial, n, loss, kind, runs.value)
        assert np.all(np.diff(z) >= 0)
        if kind in (0, 1, 2):
            assert runs.value <= 3, (kind, runs.value)
        # warm-started searches (the pooled blocks of a previous z-step as guesses): exact, shifted, far-off and
        # nonsense guesses must all give the bit-identical result — a guess only moves the first round of probes
        nm = max(int(runs.value) - 1, 1)
        ip = ctypes.POINTER(ctypes.c_int64)
        blocks = np.full(2 * nm, -1, dtype=np.int64)
        chunk = int(rng.integers(1, 8))
        z1 = np.empty_like(m)
        lib.emul_pav_fewseg_hinted(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m.ctypes.data_as(dp), rho, chunk,
                                   z1.ctypes.data_as(dp), None, None, blocks.ctypes.data_as(ip))
        np.testing.assert_array_equal(z1, z)
        for mode in range(4):
            if mode == 0:
                h = blocks.copy()
            elif mode == 1:
                h = blocks + rng.integers(-3, 4, size=blocks.size)
            elif mode == 2:
                h = blocks + rng.integers(-n, n + 1, size=blocks.size)
            else:
                h = rng.integers(-5, n + 5, size=blocks.size)
            h = np.ascontiguousarray(h, dtype=np.int64)
            z2 = np.empty_like(m)
            lib.emul_set_hint_stride(8 if (trial + mode) % 2 else 32)   # both probe spacings of the device kernel
            lib.emul_pav_fewseg_hinted(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m.ctypes.data_as(dp), rho, chunk,
                                       z2.ctypes.data_as(dp), None, h.ctypes.data_as(ip), None)
            np.testing.assert_array_equal(z2, z, err_msg=str((trial, n, loss, kind, mode)))
        lib.emul_set_hint_stride(32)


def test_warm_started_merge_search_at_scale_on_cpu():
    """the warm-started few-segment merge at n ~ 1e5 (probe offsets up to 2^15 around the guess really spread out):
    guesses taken from a neighbouring problem (margins perturbed as between two ADMM iterations), shifted by up to
    +-5000 ranks, or absent — bit-identical pooled blocks every time, equal to the oracle's stack PAV."""
    lib = _emul_lib()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
    rng = np.random.default_rng(99)
    for trial in range(12):
        n = int(rng.integers(80_000, 250_000))
        loss = ["binary_cross_entropy", "hinge"][trial % 2]
        sig = (O.spectrum("superquantile", n, [rng.uniform(0.5, 0.95)]) if trial % 3 else
               O.spectrum("aorr", n, [0.2, 0.8]))
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        rho = 10 ** rng.uniform(-5, -1)
        m0 = np.sort(rng.normal(size=n))
        m1 = np.sort(m0 * 1.02 + 0.01 * rng.normal(size=n))      # the next iteration's margins
        nm = 3
        prev = np.full(2 * nm, -1, dtype=np.int64)
        z0 = np.empty(n)
        lib.emul_pav_fewseg_hinted(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m0.ctypes.data_as(dp), rho, 10,
                                   z0.ctypes.data_as(dp), None, None, prev.ctypes.data_as(ip))
        zo = O.pav_prox(loss, sig, m1, rho)
        ref = np.empty(n)
        lib.emul_pav_fewseg_hinted(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m1.ctypes.data_as(dp), rho, 10,
                                   ref.ctypes.data_as(dp), None, None, None)
        assert np.max(np.abs(ref - zo)) <= 1e-13 * max(1.0, np.max(np.abs(zo)))
        for stride in (32, 8):                # the device kernel's two probe spacings (pav_hint_pos)
            lib.emul_set_hint_stride(stride)
            for shift in (0, 7, -60, -300, 5000, -5000):
                h = np.ascontiguousarray(np.where(prev >= 0, prev + shift, -1), dtype=np.int64)
                z = np.empty(n)
                lib.emul_pav_fewseg_hinted(O.LOSS_IDS[loss], n, sig.ctypes.data_as(dp), m1.ctypes.data_as(dp), rho,
                                           10, z.ctypes.data_as(dp), None, h.ctypes.data_as(ip), None)
                np.testing.assert_array_equal(z, ref, err_msg=str((trial, n, loss, shift, stride)))
        lib.emul_set_hint_stride(32)


def test_radix_key_transform_is_order_preserving():
    lib = _emul_lib()
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.normal(size=2000) * 10 ** rng.uniform(-300, 300, size=2000),
                        [0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, np.nan]])
    keys = np.array([lib.emul_key(float(v)) for v in x], dtype=np.uint64)
    order = np.argsort(keys, kind="stable")
    ref = np.argsort(x, kind="stable")          # numpy: NaN last, -0.0 == +0.0
    np.testing.assert_array_equal(order, ref)
    back = np.array([lib.emul_unkey(int(k)) for k in keys])
    np.testing.assert_array_equal(back[:-1], x[:-1] + 0.0)
    assert np.isnan(back[-1])


def test_row_sharded_collectives_on_gloo():
    """world_size = 2 on CPU/gloo: the sharded data flow of engine.py (all-gather of margins, all-reduce
    of [D^T r, ||r||^2] partials, replicated sort+PAV, local scatter) reproduces the unsharded oracle."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", PYTHONPATH=ROOT + os.pathsep + PKG)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert all("OK" in o for o in outs), outs


def test_fair_statistics_host_formulas_match_reference(golden_dir):
    """rbl_b200.metrics.statistics_from_counts (the host half of calculate_statistics) on confusion counts produced by
    the oracle, against the reference's own outputs in tests/golden/metrics.npz (fair_metric.py:13-40)."""
    from oracle import rbl_oracle as O
    from rbl_b200.metrics import statistics_from_counts

    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    X, y, grp = g["X"], g["y"], g["group"]
    n = X.shape[0]
    for k in range(int(g["nw"])):
        for thr in (0.5, 0.3, 0.8):
            counts, sb, sbl = O.confusion_by_group(g[f"w{k}"], X, y, grp, thr)
            c = np.zeros(16)
            c[1], c[2:14], c[14], c[15] = n, counts.reshape(-1), sb, sbl
            np.testing.assert_allclose(statistics_from_counts(c), g[f"ref_stats_{k}_{thr}"], rtol=1e-13, atol=1e-14)
    # empty classes: nan / inf as in the reference, no exception
    c = np.zeros(16)
    c[1], c[2], c[8], c[14], c[15] = 4, 2, 2, 4.0, 0.0
    spd, di, *_ = statistics_from_counts(c)
    assert spd == 0.0 and di == np.inf


def test_split_group_matches_reference(golden_dir):
    """src/util/split_group.py mirror against the row ids the reference's own function produced
    (tests/golden/metrics.npz, oracle/gen_golden.py::metrics)."""
    from src.util.split_group import train_test_split_group

    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    X, y, grp = g["X"], g["y"], g["group"]
    Xtr, Xte, ytr, yte, gtr, gte = train_test_split_group(X, y, grp, test_size=0.4, random_state=17)
    tr, te = g["ref_split_train_ids"], g["ref_split_test_ids"]
    assert len(te) == int(len(X) * 0.4) and len(tr) + len(te) == len(X)
    np.testing.assert_array_equal(Xtr, X[tr])
    np.testing.assert_array_equal(Xte, X[te])
    np.testing.assert_array_equal(ytr, y[tr])
    np.testing.assert_array_equal(gte, grp[te])


def test_product_spectra_equal_reference_bit_for_bit(golden_dir):
    """objective.py:97-187 mirrored in rbl_b200/spectra.py: every family against the reference's own vectors
    (tests/golden/spectra.npz), and the CPT pair at a large n against the oracle's Python-float loop — the
    differences distort((i+1)/n) - distort(i/n) cancel ~log10(n) digits, so only the same scalar libm pow calls
    (rbl_cpt_weights) reproduce them exactly; a vectorised pow was 3e-9 off at n = 4M."""
    from oracle import rbl_oracle as O
    from rbl_b200 import spectra as S

    g = np.load(os.path.join(golden_dir, "spectra.npz"))
    for key in g.files:
        parts = key.split("|")
        name, n = parts[0], int(parts[2])
        args = None if parts[1] == "" else [float(a) if "." in a else int(a) for a in parts[1].split(",")]
        wf = S.get_weights(name, args)
        s = (wf[0] if parts[3] == "a" else wf[1])(n) if name == "ehrm" else wf(n)
        s = np.asarray(s, dtype=np.float64).reshape(-1)
        if name == "ehrm":
            np.testing.assert_array_equal(s, g[key], err_msg=key)
        else:
            np.testing.assert_allclose(s, g[key], rtol=1e-14, atol=1e-18, err_msg=key)
    n = 200_003
    a, b = O.spectrum("ehrm", n)
    wf = S.get_weights("ehrm", None)
    np.testing.assert_array_equal(np.asarray(wf[0](n)).reshape(-1), a)
    np.testing.assert_array_equal(np.asarray(wf[1](n)).reshape(-1), b)


def test_library_lbfgsb_tracks_scipy():
    """csrc/lbfgs_core.h — the L-BFGS-B the library runs for the l2 / smoothed-l1 w-steps (w_LBFGS.py:48-62 call
    scipy.optimize.minimize(method='L-BFGS-B', options={'maxiter': 1000})) — against the installed scipy on the
    reference's own problem classes and on a non-quadratic that exercises the More'-Thuente line search: same
    iteration and evaluation counts, iterates equal to rounding (the compact representation scipy uses and the
    two-loop recursion here are the same map in exact arithmetic)."""
    from scipy.optimize import minimize

    lib = _emul_lib()
    FG = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_double))
    lib.emul_lbfgs.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), FG, ctypes.c_int, ctypes.c_int,
                               ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                               ctypes.POINTER(ctypes.c_double)]
    rng = np.random.default_rng(31)

    def run_both(fun, x0):
        n = x0.size

        def cb(xp, fp, gp):
            x = np.ctypeslib.as_array(xp, shape=(n,))
            f, g = fun(x)
            fp[0] = f
            np.ctypeslib.as_array(gp, shape=(n,))[:] = g
            return 0

        x = x0.copy()
        nit, nfev, fo = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
        st = lib.emul_lbfgs(n, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), FG(cb), 10, 1000, ctypes.byref(nit),
                            ctypes.byref(nfev), ctypes.byref(fo))
        ref = minimize(fun, x0, jac=True, method="L-BFGS-B", options={"maxiter": 1000})
        return x, nit.value, nfev.value, st, ref

    cases = []
    for trial in range(12):   # the l2 w-step: rho/2 ||D w - b||^2 + reg/2 ||w||^2 with the reference's tiny rho
        n, d = int(rng.integers(50, 400)), int(rng.integers(2, 60))
        D, b = rng.normal(size=(n, d)), rng.normal(size=n)
        rho, reg = 10 ** rng.uniform(-5, 0), 10 ** rng.uniform(-4, -1)
        cases.append((lambda w, D=D, b=b, rho=rho, reg=reg: (0.5 * rho * float((D @ w - b) @ (D @ w - b)) + 0.5 * reg * float(w @ w),
                                                              rho * (D.T @ (D @ w - b)) + reg * w), rng.normal(size=d) * 0.1))
    for trial in range(6):    # the smoothed-l1 w-step of smoothADMMmethod (w_LBFGS.py:11-28)
        n, d = int(rng.integers(80, 300)), int(rng.integers(5, 40))
        D, b = rng.normal(size=(n, d)), rng.normal(size=n)
        rho, reg, t = 10 ** rng.uniform(-4, 0), 0.01, 10 ** rng.uniform(-3, 0)

        def huber(w, D=D, b=b, rho=rho, reg=reg, t=t):
            r = D @ w - b
            small = np.abs(w) <= t
            f = 0.5 * rho * float(r @ r) + 0.25 * reg * float(np.sum(w[small] ** 2)) / t + 0.5 * reg * float(np.sum(np.abs(w[~small]) - 0.5 * t))
            return f, rho * (D.T @ r) + np.where(small, 0.5 * reg * w / t, 0.5 * reg * np.sign(w))
        cases.append((huber, rng.normal(size=d) * 0.1))

    def rosen(x):             # line search with real work to do
        f = float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
        g = np.zeros_like(x)
        g[:-1] = -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        g[1:] += 200 * (x[1:] - x[:-1] ** 2)
        return f, g
    cases.append((rosen, np.full(8, -1.2)))
    cases.append((rosen, rng.normal(size=20)))
    for k, (fun, x0) in enumerate(cases):
        x, nit, nfev, st, ref = run_both(fun, x0)
        assert st == 0 and ref.success, (k, st, ref.message)
        assert (nit, nfev) == (ref.nit, ref.nfev), (k, nit, nfev, ref.nit, ref.nfev)
        # the reference's problem classes: equal to rounding; the 100-iteration non-convex runs amplify the rounding
        # differences of the two (mathematically identical) direction formulas, the control flow stays identical
        tol = 1e-9 if k < 18 else 1e-6
        assert np.linalg.norm(x - ref.x) <= tol * max(1.0, np.linalg.norm(ref.x)), (k, np.linalg.norm(x - ref.x))


def test_rank_order_partition_bucket_search():
    """the bucket search of ss_partition_ranked_kernel (restated in tests/emul/partition_ranked.py): guess from the
    previous rank, gallop + bisection from the guess — equal to the full bisection for any guess, with duplicated
    splitters and keys equal to splitters"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "partition_ranked", os.path.join(os.path.dirname(__file__), "emul", "partition_ranked.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.run_cases()


def test_bucket_network_key_only_with_tie_fallback():
    """the hinted sort's per-bucket network (restated in tests/emul/bucket_network.py): keys-only comparator, re-sort
    with (key, index) when the bucket holds equal keys, A / B merge positions — stable order for distinct keys,
    heavy ties, constant keys and a single tie, over the bucket loads that exercise every A / B split"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "bucket_network", os.path.join(os.path.dirname(__file__), "emul", "bucket_network.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.run_cases()
