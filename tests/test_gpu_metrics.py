"""GPU parity: test-set metrics (rbl_test_metrics through the C ABI and the src/util mirrors) against the
reference's own outputs (tests/golden/metrics.npz, made by oracle/gen_golden.py::metrics from
calculate_acc.py / fair_metric.py) and against the oracle restatement on seeded inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rbl_oracle as O  # noqa: E402


def _clear_of_threshold(w, X, thr):
    """the counts are exact integers provided no probability sits within rounding of the threshold"""
    return np.min(np.abs(O.class_probs(w, X) - thr)) > 1e-9


def test_metrics_match_reference_golden(golden_dir):
    from src.util.calculate_acc import calculate_accuracy
    from src.util.fair_metric import calculate_statistics

    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    X, y, grp = g["X"], g["y"], g["group"]
    for k in range(int(g["nw"])):
        w = g[f"w{k}"]
        for thr in (0.5, 0.3, 0.8):
            assert _clear_of_threshold(w, X, thr)
            # integer counts over the same rows: exact
            assert calculate_accuracy(w, X, y, threshold=thr) == float(g[f"ref_acc_bce_{k}_{thr}"])
            # rates: ratios of exact integers; TI: n-term sums in another order (1e-13)
            np.testing.assert_allclose(calculate_statistics(w, X, y, grp, threshold=thr), g[f"ref_stats_{k}_{thr}"],
                                       rtol=1e-13, atol=1e-14)
        assert calculate_accuracy(w, X, y, loss="hinge") == float(g[f"ref_acc_hinge_{k}"])
    with pytest.raises(ValueError, match="is not supported"):
        calculate_accuracy(g["w0"], X, y, loss="multinomial_cross_entropy")


@pytest.mark.parametrize("n,d", [(1, 1), (7, 3), (1000, 201), (4097, 64), (50000, 1000), (300, 7001), (20000, 37),
                                 (5001, 16), (5003, 17), (5000, 33), (4999, 130), (3001, 256), (3000, 258),
                                 (2000, 8200)])
def test_metrics_counts_match_oracle(n, d):
    from rbl_b200.metrics import DeviceTestSet

    rng = np.random.default_rng(n * 13 + d)
    X = rng.normal(size=(n, d))
    wt = rng.normal(size=d) / np.sqrt(d)
    y = np.where(X @ wt + 0.5 * rng.normal(size=n) > 0, 1.0, -1.0)
    grp = (rng.random(n) < 0.4).astype(np.int64)
    ts = DeviceTestSet(X, y.reshape(-1, 1), group=grp)
    for scale, thr in ((1.0, 0.5), (4.0, 0.35), (0.05, 0.5)):
        w = wt * scale + 0.1 * rng.normal(size=d) / np.sqrt(d)
        if not _clear_of_threshold(w, X, thr):
            continue
        c = ts.counts(w, thr)
        ref, sb, sbl = O.confusion_by_group(w, X, y, grp, thr)
        np.testing.assert_array_equal(c[2:14].reshape(2, 6), ref)                     # exact integers
        assert c[1] == n and c[0] == round(O.calculate_accuracy(w, X, y, thr) * n)
        np.testing.assert_allclose([c[14], c[15]], [sb, sbl], rtol=1e-12, atol=1e-12)  # summation order only
        c2 = ts.counts(w, thr)
        np.testing.assert_array_equal(c, c2)                                           # fixed-order reduction
        assert ts.accuracy(w, loss="hinge") == O.calculate_accuracy(w, X, y, loss="hinge")
    # no group vector: everything lands in group 0
    c = DeviceTestSet(X, y).counts(wt)
    assert c[2] == n and c[8] == 0


def test_metrics_extreme_scores_and_labels():
    """saturated probabilities (b = 0 gives nan in TI exactly as numpy does), labels other than +-1 and group ids
    other than 0 / 1 follow the reference's masks."""
    from rbl_b200.metrics import DeviceTestSet, statistics_from_counts

    X = np.array([[800.0], [-800.0], [800.0], [-800.0], [0.5], [-0.5]])
    w = np.array([1.0])
    y = np.array([1.0, 1.0, -1.0, -1.0, 1.0, -1.0])
    grp = np.array([0, 1, 0, 1, 2, 1])
    c = DeviceTestSet(X, y, group=grp).counts(w)
    ref, sb, sbl = O.confusion_by_group(w, X, y, grp)
    np.testing.assert_array_equal(c[2:14].reshape(2, 6), ref)
    assert np.isnan(c[15]) and np.isnan(sbl)          # row 1: prob = 0, y = 1 -> b = 0 -> 0 * log 0
    got, want = statistics_from_counts(c), O.statistics_from_counts(ref, sb, sbl, len(y))
    np.testing.assert_allclose(got, want, rtol=1e-14, equal_nan=True)


def test_metrics_full_size_linearity():
    """size-independent properties at a test set the CPU oracle would take minutes on (400k x 1000, 3.2 GB):
    group counts partition the rows, confusion cells partition each group, flipping every label swaps TP<->FP and
    FN<->TN, and threshold 0 / 1+ predict all-positive / all-negative."""
    from rbl_b200.metrics import DeviceTestSet

    n, d = 400_000, 1000
    gen = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=gen)
    w = torch.randn(d, dtype=torch.float64, device="cuda", generator=gen) / d ** 0.5
    y = torch.where(X @ w + 0.3 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen) > 0, 1.0, -1.0)
    grp = (torch.rand(n, device="cuda", generator=gen) < 0.3).to(torch.int32)
    ts, ts_flip = DeviceTestSet(X, y, group=grp), DeviceTestSet(X, -y, group=grp)
    c, f = ts.counts(w), ts_flip.counts(w)
    g, gf = c[2:14].reshape(2, 6), f[2:14].reshape(2, 6)
    assert g[:, 0].sum() == n and np.all(g[:, 2:].sum(axis=1) == g[:, 0]) and np.all(g[:, 2] + g[:, 5] == g[:, 1])
    np.testing.assert_array_equal(gf[:, [0, 1, 5, 4, 3, 2]], g)
    assert c[0] + f[0] == n                              # a prediction matches y or -y
    pred = (X @ w >= 0)
    assert abs(c[0] - float((torch.where(pred, 1.0, -1.0) == y).sum())) <= 2   # rows within rounding of 0 may flip
    allpos, allneg = ts.counts(w, threshold=0.0), ts.counts(w, threshold=1.5)
    assert allpos[3] + allpos[9] == n and allneg[3] + allneg[9] == 0
