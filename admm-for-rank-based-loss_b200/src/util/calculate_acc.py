"""Drop-in for the reference's src/util/calculate_acc.py: same name, arguments, return value and error; the pass
over X_test runs on the B200 (rbl_b200.metrics.DeviceTestSet).  Pass a DeviceTestSet as `X_test` to keep the test
set resident between calls (per-iteration curves) instead of uploading it every time."""
from rbl_b200.metrics import DeviceTestSet, _loss_id


def calculate_accuracy(w, X_test, y_test=None, threshold=0.5, loss='binary_cross_entropy'):
    _loss_id(loss)  # ValueError for an unknown loss before any upload (reference :17-18)
    ts = X_test if isinstance(X_test, DeviceTestSet) else DeviceTestSet(X_test, y_test)
    return ts.accuracy(w, threshold, loss)
