"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) — each compares the sm_100a path with
the CPU oracle (oracle/rbl_oracle.py) or numpy on the same seeded inputs.  Tolerances are stated at
each assert.  Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rbl_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def E():
    from rbl_b200 import _cabi, engine

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return engine, _cabi


def _mk(E, X, y=None, loss="binary_cross_entropy", sigma=None, clip=None):
    engine, _ = E
    n = X.shape[0]
    if y is None:
        y = -np.ones(n)  # D = X
    if sigma is None:
        sigma = np.ones(n) / n
    return engine.AdmmEngine(X, y, loss, sigma, clip=clip)


SHAPES = [(1000, 200), (777, 201), (6000, 1000), (300, 40), (50, 7), (3000, 1500), (1500, 3000), (257, 5000),
          (64, 8192), (100003, 64)]


@pytest.mark.parametrize("n,d", SHAPES)
def test_matvec_and_fused_pass(E, n, d):
    rng = np.random.default_rng(n * 31 + d)
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    e = _mk(E, X, y)
    D = -y[:, None] * X
    # D itself is bit-exact (one multiply and a sign flip)
    np.testing.assert_array_equal(e.D[:, :d].cpu().numpy(), D)
    assert float(e.D[:, d:].abs().sum()) == 0.0
    x = rng.normal(size=d)
    b = rng.normal(size=n)
    xd, bd = e.vec(x), e.vec(b)
    out = e.matvec(xd).cpu().numpy()
    ref = D @ x
    scale = np.abs(D) @ np.abs(x)
    # fp64 dot products in a different summation order: error <= ~d*eps*sum|terms|
    assert np.max(np.abs(out - ref) / scale) < 1e-14
    _, cabi = E
    cabi.check(e.lib.rbl_fused_pass(e.h, e.D.data_ptr(), xd.data_ptr(), bd.data_ptr(), e.r.data_ptr(),
                                    e.red.data_ptr(), e._stream()))
    r = e.r.cpu().numpy()
    red = e.red.cpu().numpy()
    rref = b - ref
    assert np.max(np.abs(r - rref) / (np.abs(b) + scale)) < 1e-14
    gref = D.T @ rref
    gscale = np.abs(D).T @ np.abs(rref)
    assert np.max(np.abs(red[:d] - gref) / gscale) < 1e-13
    assert abs(red[d] - rref @ rref) < 1e-13 * (rref @ rref)
    e.close()


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 100003, 1 << 20, 4_000_037])
def test_sort_bit_exact_vs_stable_argsort(E, n):
    rng = np.random.default_rng(n)
    e = _mk(E, np.zeros((n, 2)))
    cases = {"normal": rng.normal(size=n), "ties": np.round(rng.normal(size=n), 1),
             "mixed": np.concatenate([rng.normal(size=n - n // 2) * 1e-300, rng.normal(size=n // 2) * 1e300])}
    sp = rng.normal(size=n)
    if n >= 31:
        sp[:8] = [0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, 1.0, -1.0]
        sp[8:16] = sp[:8]
    cases["special"] = sp
    for name, m in cases.items():
        md = e.vec(m)
        ref = np.argsort(m, kind="stable")
        # both implementations: the persistent cooperative kernel (default) and the three-launches-per-pass one
        for legacy in (0, 1):
            E[1].check(e.lib.rbl_sort_config(e.h, legacy))
            e.perm.zero_()
            e.m_sorted.zero_()
            E[1].check(e.lib.rbl_sort_margins(e.h, md.data_ptr(), e.m_sorted.data_ptr(), e.perm.data_ptr(),
                                              e._stream()))
            perm = e.perm.cpu().numpy()
            ms = e.m_sorted.cpu().numpy()
            np.testing.assert_array_equal(perm, ref, err_msg=f"{name} legacy={legacy}")         # bit-exact permutation
            np.testing.assert_array_equal(ms, m[ref] + 0.0, err_msg=f"{name} legacy={legacy}")  # -0.0 -> +0.0
    e.close()


@pytest.mark.parametrize("n", [3000, 70001, 1 << 20, 3_000_017, 1 << 22, (1 << 22) + 5])
def test_splitter_sort_bit_exact_and_falls_back(E, n):
    """rbl_sort_margins_near: buckets from the previous rank order + per-bucket shared-memory sort must give the
    same bit-exact stable permutation as the LSD sort — with a good hint (same data, moved / shifted / scaled data),
    with heavy ties, and with useless hints, where the device-side fallback to the LSD sort runs."""
    _, cabi = E
    rng = np.random.default_rng(n)
    e = _mk(E, np.zeros((n, 2)))
    nseg = ctypes.c_int32(0)
    base = rng.normal(size=n)

    st = (ctypes.c_int32 * 4)()

    def sort_near(m, hint_perm, route=None):
        md, hd = e.vec(m), torch.from_numpy(np.ascontiguousarray(hint_perm, dtype=np.int32)).to(e.device)
        e.perm.zero_()
        e.m_sorted.zero_()
        cabi.check(e.lib.rbl_sort_margins_near(e.h, md.data_ptr(), hd.data_ptr(), e.m_sorted.data_ptr(),
                                               e.perm.data_ptr(), e._stream()))
        ref = np.argsort(m, kind="stable")
        np.testing.assert_array_equal(e.perm.cpu().numpy(), ref)
        np.testing.assert_array_equal(e.m_sorted.cpu().numpy(), m[ref] + 0.0)
        cabi.check(e.lib.rbl_sort_stats(e.h, e._stream(), st))
        assert st[3] == 0
        assert (st[0] > 0) == (n <= (1 << 22))     # larger sorts always take the LSD route
        if route is not None and st[0] > 0:
            assert st[1] == route, (route, list(st))

    hint = np.argsort(base, kind="stable")
    sort_near(base, hint, 1)                                           # exact hint: the bucket route
    assert st[0] == 0 or st[2] <= -(-n // st[0]) + 1                   # perfectly balanced buckets
    sort_near(base + 1e-4 * rng.normal(size=n), hint, 1)               # the margins moved a little (ranks by ~100)
    sort_near(base * 1e3 + 5e3, hint, 1)                               # shifted and scaled: same ranks, still fast
    sort_near(-base, hint, 1)                                          # reversed order: splitters get sorted
    sort_near(base + 1e-2 * rng.normal(size=n), hint)                  # moved a lot: either route, same answer
    sort_near(np.round(base, 2), hint)                                 # ~600 distinct values: long runs of ties
    sort_near(base, np.arange(n))                                      # useless hint (identity)
    overflow = 2 if n > 8192 else None                                 # (a few thousand keys fit in ONE bucket)
    sort_near(base, np.zeros(n, dtype=np.int32), overflow)             # degenerate hint (all row 0) -> LSD
    sort_near(np.full(n, 0.25), hint, overflow)                        # all keys equal -> LSD
    for _ in range(5):                                                 # after an overflow the bucket route pauses
        sort_near(base, hint)                                          # for 4 calls (still exact, via LSD) ...
    sort_near(base, hint, 1)                                           # ... and then works again
    # in place, as the engine calls it: the hint IS the output buffer of the previous call
    md = e.vec(base + 2e-3 * rng.normal(size=n))
    cabi.check(e.lib.rbl_sort_margins_near(e.h, md.data_ptr(), e.perm.data_ptr(), e.m_sorted.data_ptr(),
                                           e.perm.data_ptr(), e._stream()))
    mm = md.cpu().numpy()
    np.testing.assert_array_equal(e.perm.cpu().numpy(), np.argsort(mm, kind="stable"))
    # ... which is the case where the partition walks the keys in the PREVIOUS RANK ORDER (bucket guessed from the
    # rank, verified against the splitters, gallop + bisection when the row moved further): a chain of in-place
    # calls over drifting, jumping, reversed, tied and constant data, each checked bit for bit; rbl_sort_config
    # bit 2 (row-order partition) must give the same
    for cfg in (0, 4):
        cabi.check(e.lib.rbl_sort_config(e.h, cfg))
        cur = base.copy()
        amp = 50.0 / n   # a few per cent of a bucket's width: the regime of consecutive ADMM iterations
        drift = ("drift", lambda c: c + amp * rng.normal(size=n))
        chain = [drift] * 7     # (the first calls may still sit in the pause an earlier overflow started)
        chain += [("jump", lambda c: c + 0.5 * rng.normal(size=n)), ("reverse", lambda c: -c),
                  ("ties", lambda c: np.round(c, 2)), drift,
                  ("scale", lambda c: 1e-30 * c), ("signs", lambda c: np.where(rng.random(n) < 0.5, 0.0, -0.0)),
                  ("constant", lambda c: np.full(n, -3.5)), ("back", lambda c: base + 0.0)]
        chain += [drift] * 11
        if cfg == 4:            # the row-order partition has had its cases above: a short chain
            chain = [drift] * 6 + chain[8:11] + [drift] * 6   # (the ties may overflow a bucket: 4 paused calls)
        routes = []
        for name, f in chain:
            cur = f(cur)
            md = e.vec(cur)
            cabi.check(e.lib.rbl_sort_margins_near(e.h, md.data_ptr(), e.perm.data_ptr(), e.m_sorted.data_ptr(),
                                                   e.perm.data_ptr(), e._stream()))
            ref = np.argsort(cur, kind="stable")
            np.testing.assert_array_equal(e.perm.cpu().numpy(), ref, err_msg=f"{name} cfg={cfg}")
            np.testing.assert_array_equal(e.m_sorted.cpu().numpy(), cur[ref] + 0.0, err_msg=f"{name} cfg={cfg}")
            cabi.check(e.lib.rbl_sort_stats(e.h, e._stream(), st))
            assert st[3] == 0
            routes.append(int(st[1]))
        if st[0] > 0:
            assert routes[5] == 1 and routes[6] == 1 and routes[-1] == 1, routes   # bucket route on drifting data
            if n > 8192 and cfg == 0:
                assert routes[13] == 2, routes                                      # constant keys: LSD fallback
    cabi.check(e.lib.rbl_sort_config(e.h, 0))
    e.close()


SPECTRA = [("erm", None), ("superquantile", [0.8]), ("aorr", [0.2, 0.8]), ("extremile", [2.5]), ("esrm", [1.5]),
           ("ehrm", None)]


@pytest.mark.parametrize("loss", ["binary_cross_entropy", "hinge"])
@pytest.mark.parametrize("n", [1, 2, 3, 1023, 1024, 1025, 5000, 200000])
def test_pav_matches_oracle(E, loss, n):
    rng = np.random.default_rng(n + (7 if loss == "hinge" else 0))
    for wf, args in SPECTRA:
        if wf == "aorr" and n < 5:
            continue
        sig = O.spectrum(wf, n, args)
        sig = sig[1] if isinstance(sig, tuple) else sig
        for rho, scale in [(1e-5, 1e-3), (1e-3, 1.0), (1.0, 3.0)]:
            m = np.sort(rng.normal(size=n) * scale)
            if rho == 1e-3 and loss == "hinge":
                m = np.round(m, 2)  # ties, many blocks on the kink
            e = _mk(E, np.zeros((n, 2)), loss=loss, sigma=sig)
            e.m_sorted.copy_(e.vec(m))
            zo = O.pav_prox(loss, sig, m, rho)
            nseg = ctypes.c_int32(-1)
            # both routes to the same unique minimiser: the general merge tree, and (when the spectrum has few
            # runs of non-increasing sigma: erm, superquantile, aorr) the few-segment path
            for force_tree in (1, 0):
                E[1].check(e.lib.rbl_pav_config(e.h, force_tree, ctypes.byref(nseg)))
                e.z_sorted.zero_()
                E[1].check(e.lib.rbl_pav_prox(e.h, O.LOSS_IDS[loss], e.m_sorted.data_ptr(), rho,
                                              e.z_sorted.data_ptr(), e._stream()))
                z = e.z_sorted.cpu().numpy()
                # both solved to machine precision; block sums in double-double
                assert np.max(np.abs(z - zo)) <= 1e-12 * max(1.0, np.max(np.abs(zo))), (wf, rho, n, force_tree)
                assert np.all(np.diff(z) >= 0)
            if wf in ("erm", "superquantile", "aorr") and n >= 5:
                assert 1 <= nseg.value <= 3, (wf, nseg.value)
            e.close()


@pytest.mark.parametrize("loss", ["binary_cross_entropy", "hinge"])
@pytest.mark.parametrize("wf,args", [("superquantile", [0.8]), ("aorr", [0.2, 0.8]), ("superquantile", [0.5])])
def test_few_segment_merge_warm_start_is_only_a_guess(E, loss, wf, args):
    """A sequence of related z-steps (margins drift and rescale as along a solve, then jump): the warm-started
    few-segment merge (guesses = the previous call's pooled blocks) must return the bit-identical z_sorted of the
    cold search every time, and both must equal the oracle's stack PAV."""
    n = 300_000
    rng = np.random.default_rng(len(wf) + n)
    sig = O.spectrum(wf, n, args)
    warm, cold = _mk(E, np.zeros((n, 2)), loss=loss, sigma=sig), _mk(E, np.zeros((n, 2)), loss=loss, sigma=sig)
    E[1].check(cold.lib.rbl_pav_config(cold.h, 2, None))      # bit 1: no warm start
    m = np.sort(rng.normal(size=n))
    rho = 1e-5
    for it in range(14):
        if it == 10:
            m = np.sort(rng.normal(size=n) * 5 + 3)            # a jump: the guesses are useless
        else:
            m = np.sort(m * (1.0 + 0.05 * rng.normal()) + 0.02 * rng.normal(size=n))
        rho *= 1.3
        zs = []
        for e in (warm, cold):
            e.m_sorted.copy_(e.vec(m))
            E[1].check(e.lib.rbl_pav_prox(e.h, O.LOSS_IDS[loss], e.m_sorted.data_ptr(), rho, e.z_sorted.data_ptr(),
                                          e._stream()))
            zs.append(e.z_sorted.cpu().numpy())
        np.testing.assert_array_equal(zs[0], zs[1], err_msg=str((wf, loss, it)))
        zo = O.pav_prox(loss, sig, m, rho)
        assert np.max(np.abs(zs[0] - zo)) <= 1e-12 * max(1.0, np.max(np.abs(zo))), (wf, loss, it)
    warm.close()
    cold.close()


def test_prox_elementwise(E):
    rng = np.random.default_rng(3)
    n = 10000
    m = rng.normal(size=n) * 10 ** rng.uniform(-2, 2, size=n)
    s = 10 ** rng.uniform(-8, 0, size=n) * (rng.random(n) > 0.1)
    from src.util.individual_solver import individual_solver

    for rho in (1e-6, 1e-2, 10.0):
        for loss in ("binary_cross_entropy", "hinge"):
            z = individual_solver(loss, s, rho, m)
            zo = O.prox_vec(loss, s, m, rho)
            assert np.max(np.abs(z - zo) / np.maximum(1.0, np.abs(m))) < 1e-13, (loss, rho)


def test_objective_matches_oracle(E):
    rng = np.random.default_rng(5)
    n, d = 5000, 33
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0).reshape(-1, 1)
    w = rng.normal(size=d) * 0.3
    from src.optim.objective import rankbasedObjective

    for wf, args, loss, kw in [("erm", None, "binary_cross_entropy", dict(l1_reg=0.01)),
                               ("superquantile", [0.8], "binary_cross_entropy", dict(l2_reg=0.01)),
                               ("aorr", [0.2, 0.8], "hinge", dict(l2_reg=1e-4)),
                               ("ehrm", None, "binary_cross_entropy", dict(l2_reg=0.01))]:
        ob = rankbasedObjective(torch.from_numpy(X), torch.from_numpy(y), wf, loss, kw.get("l2_reg"),
                                kw.get("l1_reg"), -5 if wf == "ehrm" else None, None, args)
        v = ob.get_arrogate_loss(torch.from_numpy(w.reshape(-1, 1)))
        sig = O.spectrum(wf, n, args)
        sig = sig[0] if isinstance(sig, tuple) else sig
        vo = O.objective(-y * X, w, sig, loss, **kw)
        assert abs(v - vo) < 1e-12 * abs(vo), (wf, v, vo)  # fp64 sums in a different order
        ob.problem.close()


@pytest.mark.parametrize("n,d", [(3000, 40), (2400, 201), (5000, 1000), (70000, 130), (900, 64), (4097, 513)])
def test_gram_build_eval_and_dual_pass(E, n, d):
    """G = D^T D on the FP64 tensor cores vs numpy; the Gram identities behind the w-step
    (D^T(b - D w) = g0 - G (w - w0), ||b - D w||^2 = ss0 - 2 dw.g0 + dw.G dw); the dual update fused into
    the D w pass (algorithms.py:132-136)."""
    _, cabi = E
    rng = np.random.default_rng(n + 7 * d)
    X = rng.normal(size=(n, d))
    e = _mk(E, X)
    G = e.gram()[:, :d].cpu().numpy()
    ref = X.T @ X
    # fp64 dot products of length n in a different order: |err| <= ~ n eps sum|terms|; observed ~1e-15 relative
    assert np.max(np.abs(G - ref)) < 1e-13 * np.max(np.abs(ref))
    np.testing.assert_array_equal(G, G.T)                      # exactly symmetric
    assert float(e.G[:, d:].abs().sum()) == 0.0                # padding column stays zero
    w0, w, b = rng.normal(size=d), rng.normal(size=d), rng.normal(size=n)
    w0d, wd, bd = e.vec(w0), e.vec(w0 + 1e-3 * w), e.vec(b)
    e._pass_at(w0d, bd)
    cabi.check(e.lib.rbl_gram_eval(e.h, e.G.data_ptr(), w0d.data_ptr(), e.red0.data_ptr(), wd.data_ptr(),
                                   e.red1.data_ptr(), e._stream()))
    red = e.red1.cpu().numpy()
    r = b - X @ (w0 + 1e-3 * w)
    gref = X.T @ r
    assert np.max(np.abs(red[:d] - gref)) < 1e-12 * np.max(np.abs(np.abs(X).T @ np.abs(r)))
    assert abs(red[d] - r @ r) < 1e-13 * (r @ r)
    # dual pass: dense w takes the streaming pass; a sparse w (exact zeros, as the l1 w-step leaves them) takes
    # the sector-gather kernel — both chosen on the device from nnz(w) vs the cap
    for nnz, cap in ((d, max(1, d // 16)), (min(d, 5), max(5, d // 16)), (min(d, 13), 16), (min(d, 29), 32), (0, 4),
                     (min(d, 40), 64)):
        wv = np.zeros(d)
        sup = rng.choice(d, size=nnz, replace=False)
        wv[sup] = rng.normal(size=nnz)
        z, lam = rng.normal(size=n), rng.normal(size=n)
        wd2, zd, lamd = e.vec(wv), e.vec(z), e.vec(lam)
        Dw = torch.empty(n, dtype=torch.float64, device=e.device)
        # every second case through the transposed copy of D (coalesced column reads)
        use_t = (nnz % 2 == 1)
        if use_t and e.Dt is None:
            e._build_transpose()
            np.testing.assert_array_equal(e.Dt.cpu().numpy(), X.T)
        cabi.check(e.lib.rbl_dual_pass(e.h, e.D.data_ptr(), e.Dt.data_ptr() if use_t else 0, wd2.data_ptr(),
                                       w0d.data_ptr(), zd.data_ptr(), Dw.data_ptr(), lamd.data_ptr(), 0.37, cap, 0,
                                       e._out4.data_ptr(), 0, e._stream()))
        dw = X @ wv
        scale = np.abs(X) @ np.abs(wv) + 1e-300
        assert np.max(np.abs(Dw.cpu().numpy() - dw) / scale) < 1e-14, (nnz, cap)
        assert np.max(np.abs(lamd.cpu().numpy() - (lam + 0.37 * (z - dw)))) < 1e-13 * max(1.0, np.max(scale))
        o = e._out4.cpu().numpy()
        assert abs(o[0] - np.sum((z - dw) ** 2)) < 1e-13 * o[0]
        assert abs(o[1] - np.sum((wv - w0) ** 2)) < 1e-13 * o[1]
        assert abs(o[2] - wv @ wv) <= 1e-13 * o[2] and abs(o[3] - np.abs(wv).sum()) <= 1e-13 * o[3]
        assert int(o[4]) == nnz and bool(o[5]) == (nnz <= cap), (nnz, cap, o[4:6])
    e.close()


@pytest.mark.parametrize("wf,args,loss,n,d,frac", [
    ("superquantile", [0.8], "binary_cross_entropy", 20000, 200, 0.75),   # 20% of the rows active -> gather
    ("superquantile", [0.8], "binary_cross_entropy", 20000, 200, 0.05),   # same state, forced dense
    ("aorr", [0.2, 0.8], "hinge", 5000, 101, 0.75),                       # hinge: margins below the kink drop out
    ("erm", None, "binary_cross_entropy", 3000, 64, 0.75),                # every row active -> streaming pass
    ("ehrm", None, "binary_cross_entropy", 4000, 50, 0.75),               # clip at B makes rows active
])
def test_active_row_gradient_pass(E, wf, args, loss, n, d, frac):
    """After a z-step, b - D w = z - m vanishes on every row the prox left alone; rbl_grad_pass over the active
    rows must give the same [D^T (b - D w), ||b - D w||^2] as the full fused pass (rounding only)."""
    _, cabi = E
    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    sig = O.spectrum(wf, n, args)
    sig = sig[1] if isinstance(sig, tuple) else sig
    e = _mk(E, X, y, loss=loss, sigma=sig, clip=-5.0 if wf == "ehrm" else None)
    e.active_dense_frac = frac
    rho = 0.05
    e.set_state(w=rng.normal(size=d) * 0.05, z=np.zeros(n), lam=rng.normal(size=n) * 0.01)
    e.z_step(rho)
    if e.w_mode != "gram":
        pytest.skip("active rows are a Gram-mode feature")
    cnt = ctypes.c_int32(0)
    cabi.check(e.lib.rbl_active_count(e.h, ctypes.byref(cnt), e._stream()))
    z, m = e.z.cpu().numpy(), e.m.cpu().numpy()
    assert cnt.value == int(np.count_nonzero(z != m))
    if wf == "superquantile":
        assert cnt.value < 0.35 * n          # the sigma = 0 ranks outside the pooled block are untouched
    e.gram()
    e._pass_at(e.w, e.b, use_active=True)
    red_a = e.red0.cpu().numpy().copy()
    e._pass_at(e.w, e.b, use_active=False)
    red_f = e.red0.cpu().numpy().copy()
    D = e.D[:, :d].cpu().numpy()
    delta = z - m
    gref = D.T @ delta
    scale = np.abs(D).T @ np.abs(delta) + 1e-300
    assert np.max(np.abs(red_a[:d] - gref) / scale) < 1e-14
    assert np.max(np.abs(red_f[:d] - gref) / np.maximum(scale, np.abs(D).T @ np.abs(e.b.cpu().numpy()) * 1e-2)) < 1e-13
    assert abs(red_a[d] - delta @ delta) <= 1e-13 * (delta @ delta)
    e.close()


def test_pipelined_upload_builds_the_same_design_and_gram(E, monkeypatch):
    """Host arrays are uploaded chunk by chunk while D = -y*X and G = D^T D are built behind the copy
    (engine._upload_pipelined); D must be bit-identical to the one-shot build and G equal up to the summation
    order of the chunks."""
    engine, _ = E
    rng = np.random.default_rng(8)
    n, d = 21001, 65   # odd d: padded leading dimension; ragged last chunk
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    monkeypatch.setenv("RBL_W_MODE", "gram")
    monkeypatch.setenv("RBL_PIPELINE_ROWS", "4096")
    a = engine.AdmmEngine(X, y, "binary_cross_entropy", np.ones(n) / n)
    assert getattr(a, "gram_during_upload", False)
    monkeypatch.setenv("RBL_PIPELINE", "0")
    b = engine.AdmmEngine(X, y, "binary_cross_entropy", np.ones(n) / n)
    assert not getattr(b, "gram_during_upload", False)
    np.testing.assert_array_equal(a.D.cpu().numpy(), b.D.cpu().numpy())
    np.testing.assert_array_equal(a.D[:, :d].cpu().numpy(), -y[:, None] * X)
    Ga, Gb = a.gram()[:, :d].cpu().numpy(), b.gram()[:, :d].cpu().numpy()
    ref = X.T @ X
    assert np.max(np.abs(Ga - ref)) < 1e-13 * np.max(np.abs(ref)) and np.max(np.abs(Gb - ref)) < 1e-13 * np.max(np.abs(ref))
    np.testing.assert_array_equal(Ga, Ga.T)
    assert float(a.G[:, d:].abs().sum()) == 0.0
    a.close()
    b.close()
