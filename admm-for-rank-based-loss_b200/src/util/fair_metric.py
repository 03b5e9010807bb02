"""Drop-in for the reference's src/util/fair_metric.py: calculate_statistics -> (SPD, DI, EOD, AOD, TI, FNRD), the
group-fairness numbers run_EHRM.py:41 prints.  One pass over X_test on the B200 gives the per-group confusion
counts and the Theil sums; the ratios are formed on the host.  Pass a DeviceTestSet as `X_test` (built with
group=...) to keep the test set resident between calls."""
from rbl_b200.metrics import DeviceTestSet


def calculate_statistics(w, X_test, label_test=None, group_test=None, threshold=0.5):
    ts = X_test if isinstance(X_test, DeviceTestSet) else DeviceTestSet(X_test, label_test, group=group_test)
    return ts.statistics(w, threshold)
