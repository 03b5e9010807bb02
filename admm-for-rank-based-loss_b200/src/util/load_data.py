"""Drop-in for the `synthetic` recipe of src/util/load_data.py:101-116 (the CSV / UCI / UTKFace loaders of the
reference are out of scope).

`get_data(...)` returns host arrays exactly as the reference does.  With `device=...` (an extension) the matrix is
uploaded once after scikit-learn's generator and standardised ON the device (`rbl_b200.ingest.standardize_`), and a
device tensor is returned instead — `ADMMmethod`, `DeviceTestSet` and `train_test_split_device` take it without a
copy back to the host."""
from sklearn import preprocessing
from sklearn.datasets import make_classification


def get_data(dataname, num_row=None, num_feature=None, seed=17, device=None):
    if dataname == "synthetic":
        if num_row is None or num_feature is None:
            raise ValueError("Number of samples and features should be specified for synthetic data!")
        (X, label) = make_classification(n_samples=num_row, n_features=num_feature, n_classes=2, random_state=seed)
        label[label == 0] = -1
        label = label.reshape((-1, 1))
    else:
        raise ValueError(
            f"Unrecognized data '{dataname}'! Options: ['synthetic'] (other loaders are out of scope here)"
        )
    if device is not None:
        from rbl_b200 import ingest

        Xp, d = ingest.to_device_padded(X, device)
        ingest.standardize_(Xp, d)
        return Xp[:, :d], label
    X = preprocessing.scale(X)
    return X, label
