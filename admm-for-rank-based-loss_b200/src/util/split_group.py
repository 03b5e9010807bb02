"""Drop-in for the reference's src/util/split_group.py (run_EHRM.py:25): a shuffled train / test split that carries
the group labels along.  The permutation is numpy's legacy global stream (`np.random.seed(random_state)` then
`np.random.permutation(n)`), exactly as there; its first int(n * test_size) entries are the test rows, the rest
the training rows.  Host arrays are indexed on the host as in the reference; a device-resident X (torch CUDA tensor,
e.g. from `get_data(..., device=...)`) has its rows gathered on the device (`rbl_b200.ingest.split_rows`)."""
import numpy as np


def _take_rows(X, rows):
    if isinstance(X, np.ndarray):
        return X[rows]
    from rbl_b200 import ingest

    Xp, d = ingest.to_device_padded(X)
    return ingest.split_rows(Xp, rows, d)[:, :d]


def train_test_split_group(X, y, group, test_size=0.2, random_state=None):
    n = X.shape[0]
    if random_state is not None:
        np.random.seed(random_state)
    order = np.random.permutation(n)
    cut = int(n * test_size)
    held_out, kept = order[:cut], order[cut:]
    y, group = np.asarray(y), np.asarray(group)
    return (_take_rows(X, kept), _take_rows(X, held_out), y[kept], y[held_out], group[kept], group[held_out])
