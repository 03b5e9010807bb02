"""Drop-in for the reference's src/util/split_group.py (run_EHRM.py:25): a shuffled train / test split that carries
the group labels along.  The permutation is numpy's legacy global stream (`np.random.seed(random_state)` +
`np.random.permutation`), exactly as there; the first int(n * test_size) shuffled rows are the test set.  Host arrays
are indexed on the host as in the reference; a device-resident X (torch CUDA tensor, e.g. from
`get_data(..., device=...)`) has its rows gathered on the device (`rbl_b200.ingest.split_rows`)."""
import numpy as np


def train_test_split_group(X, y, group, test_size=0.2, random_state=None):
    n_samples, n_features = X.shape
    if random_state is not None:
        np.random.seed(random_state)
    shuffled_index = np.random.permutation(n_samples)
    n_test_samples = int(n_samples * test_size)
    test_index = shuffled_index[:n_test_samples]
    train_index = shuffled_index[n_test_samples:]
    if isinstance(X, np.ndarray):
        X_train, X_test = X[train_index], X[test_index]
    else:
        from rbl_b200 import ingest

        Xp, d = ingest.to_device_padded(X)
        X_train, X_test = ingest.split_rows(Xp, train_index, d)[:, :d], ingest.split_rows(Xp, test_index, d)[:, :d]
    y, group = np.asarray(y), np.asarray(group)
    return X_train, X_test, y[train_index], y[test_index], group[train_index], group[test_index]
