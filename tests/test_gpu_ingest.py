"""GPU parity: device-side data ingest (rbl_standardize_columns, rbl_gather_rows through the C ABI) against
scikit-learn's own `preprocessing.scale` and `train_test_split` — the third-party calls the reference's
load_data.py:115 / run_SRM.py:26 make."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d", [(1, 2), (5, 1), (1000, 7), (4097, 64), (30001, 200), (20000, 1001), (300, 5000)])
def test_standardize_matches_sklearn_scale(n, d):
    from sklearn import preprocessing

    from rbl_b200 import ingest

    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d)) * rng.uniform(0.1, 30.0, size=d) + rng.normal(size=d) * 5.0
    if d > 2:
        X[:, 1] = 3.25          # constant column: std 0 -> scale stays 1, result exactly 0
    ref = preprocessing.scale(X)
    Xp, dd = ingest.to_device_padded(X)
    mean, scale = ingest.standardize_(Xp, dd)
    got = Xp[:, :d].cpu().numpy()
    # same formulas, other summation order: mean and std agree to ~1e-15 relative, the quotient to ~1e-14 absolute
    np.testing.assert_allclose(mean.cpu().numpy(), X.mean(axis=0), rtol=1e-13, atol=1e-14)
    sd = X.std(axis=0)
    sd[sd < 10 * np.finfo(float).eps] = 1.0
    np.testing.assert_allclose(scale.cpu().numpy(), sd, rtol=1e-13)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-12)
    if d > 2:
        assert np.all(got[:, 1] == 0.0)
    if Xp.shape[1] != d:
        assert torch.all(Xp[:, d] == 0)          # the padding column stays zero
    # idempotent up to rounding: standardised data has mean 0 and std 1
    m2, s2 = ingest.standardize_(Xp, dd)
    keep = np.ones(d, dtype=bool)
    if d > 2:
        keep[1] = False
    if n > 1:
        assert np.max(np.abs(m2.cpu().numpy())) < 1e-13 and np.max(np.abs(s2.cpu().numpy()[keep] - 1)) < 1e-12


def test_split_matches_sklearn_and_feeds_the_solver():
    from sklearn.model_selection import train_test_split

    from rbl_b200 import ingest
    from src.optim.algorithms import ADMMmethod
    from src.util.load_data import get_data

    X, y = get_data("synthetic", num_row=3000, num_feature=41, seed=17)               # host path (reference recipe)
    Xd, yd = get_data("synthetic", num_row=3000, num_feature=41, seed=17, device="cuda")  # device standardisation
    assert np.array_equal(y, yd)
    np.testing.assert_allclose(Xd.cpu().numpy(), X, rtol=0, atol=2e-12)
    Xtr, Xte, ytr, yte = train_test_split(X, y, test_size=0.4, random_state=17)
    Xp, d = ingest.to_device_padded(X)
    got = ingest.train_test_split_device(Xp, y, test_size=0.4, random_state=17, d=d)
    np.testing.assert_array_equal(got[0][:, :d].cpu().numpy(), Xtr)     # same rows, bit for bit
    np.testing.assert_array_equal(got[1][:, :d].cpu().numpy(), Xte)
    np.testing.assert_array_equal(got[2], ytr)
    np.testing.assert_array_equal(got[3], yte)
    with pytest.raises(IndexError):
        ingest.split_rows(Xp, np.array([0, 3000]))
    # the device tensors go straight into the solver: same iterates as from the host arrays
    s1 = ADMMmethod(Xtr, ytr, "superquantile", "binary_cross_entropy", l2_reg=0.01, args=[0.8], max_iter=8)
    s2 = ADMMmethod(got[0][:, :d], ytr, "superquantile", "binary_cross_entropy", l2_reg=0.01, args=[0.8], max_iter=8)
    w1, w2 = s1.main_loop(verbose=False), s2.main_loop(verbose=False)
    np.testing.assert_array_equal(w1, w2)
    s1.engine.close()
    s2.engine.close()


def test_standardize_full_size_properties():
    """2 M x 500 (8 GB): column means 0 and stds 1 after the pass, to rounding; untouched padding not applicable."""
    from rbl_b200 import ingest

    n, d = 2_000_000, 500
    gen = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=gen) * 7.0 + 2.5
    ingest.standardize_(X)
    assert float(X.mean(dim=0).abs().max()) < 1e-12
    assert float((X.std(dim=0, unbiased=False) - 1).abs().max()) < 1e-12


def test_split_group_on_device_rows_match_host_split():
    """run_EHRM.py:25 flow with a device-resident X: same rows as the host split, labels and groups alongside."""
    from src.util.split_group import train_test_split_group

    rng = np.random.default_rng(5)
    n, d = 1201, 13
    X = rng.normal(size=(n, d))
    y = np.where(rng.random(n) > 0.5, 1, -1).reshape(-1, 1)
    grp = (rng.random(n) < 0.3).astype(int)
    host = train_test_split_group(X, y, grp, test_size=0.4, random_state=17)
    dev = train_test_split_group(torch.from_numpy(X).cuda(), y, grp, test_size=0.4, random_state=17)
    np.testing.assert_array_equal(dev[0].cpu().numpy(), host[0])
    np.testing.assert_array_equal(dev[1].cpu().numpy(), host[1])
    for a, b in zip(dev[2:], host[2:]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("nbytes,threads", [(8, 1), (1 << 20, 3), ((8 << 20) - 8, 4), ((8 << 20) + 8, 4), (100_000_008, 8),
                                            (300_000_000, 16)])
def test_pageable_upload_is_exact(nbytes, threads):
    """rbl_h2d_pageable: a pageable host array staged by several host threads through pinned slots lands on the
    device byte for byte, for sizes around the 8 MB chunking, more threads than chunks, and repeated calls that
    reuse the slots while earlier transfers may still be in flight."""
    import ctypes

    from rbl_b200 import _cabi

    lib = _cabi.load()
    rng = np.random.default_rng(nbytes % 1000 + threads)
    n = nbytes // 8
    dst = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for rep in range(3):
        src = rng.standard_normal(n)                      # ordinary numpy memory: pageable
        _cabi.check(lib.rbl_h2d_pageable(0, dst.data_ptr(), src.ctypes.data, n * 8, threads, stream))
        got = dst.cpu().numpy()                           # same stream: ordered after the upload
        np.testing.assert_array_equal(got, src)


def test_solver_from_pageable_equals_solver_from_pinned():
    """the pipelined constructor (chunks of X staged, turned into rows of D and accumulated into G while the next
    chunk travels) gives bit-identical D and G whether X is pageable (threaded staging) or pinned (direct DMA)."""
    from rbl_b200.engine import AdmmEngine

    rng = np.random.default_rng(8)
    n, d = 70_000, 120
    X = rng.standard_normal((n, d))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
    Xp.copy_(torch.from_numpy(X))
    import os
    os.environ["RBL_PIPELINE_ROWS"] = "9000"              # force several chunks at this size
    try:
        a = AdmmEngine(X, y, "binary_cross_entropy", np.ones(n) / n)
        b = AdmmEngine(Xp.numpy(), y, "binary_cross_entropy", np.ones(n) / n)
    finally:
        del os.environ["RBL_PIPELINE_ROWS"]
    assert a.upload_path.startswith("pageable") and b.upload_path.startswith("pinned")
    assert torch.equal(a.D, b.D)
    if a.G is not None:
        assert torch.equal(a.gram(), b.gram())
    np.testing.assert_array_equal(a.D[:, :d].cpu().numpy(), -y[:, None] * X)
    a.close()
    b.close()
