"""Drop-in for the `synthetic` recipe of src/util/load_data.py:101-116 (data ingest is outside the hot
path; the CSV / UCI / UTKFace loaders of the reference are out of scope)."""
from sklearn import preprocessing
from sklearn.datasets import make_classification


def get_data(dataname, num_row=None, num_feature=None, seed=None):
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
    X = preprocessing.scale(X)
    return X, label
