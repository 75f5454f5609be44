import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def u1():
    """(trainset, testset) of the reference's tests/u1_ml100k_{train,test} fixture via our own loader."""
    import surprise_b200 as sb
    from surprise_b200.model_selection import PredefinedKFold
    data = sb.Dataset.load_from_folds([(os.path.join(GOLDEN, "u1_ml100k_train"),
                                        os.path.join(GOLDEN, "u1_ml100k_test"))], sb.Reader("ml-100k"))
    return next(PredefinedKFold().split(data))


@pytest.fixture(scope="session")
def u1_golden():
    with open(os.path.join(GOLDEN, "u1_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def u1_arrays():
    return dict(np.load(os.path.join(GOLDEN, "u1_arrays.npz")))


@pytest.fixture(scope="session")
def toy():
    return dict(np.load(os.path.join(GOLDEN, "toy_sims.npz")))


@pytest.fixture(scope="session")
def floats():
    return dict(np.load(os.path.join(GOLDEN, "float_sims.npz")))


def inner_pairs(trainset, testset):
    """raw testset -> (iu, ii) int32 arrays with -1 for unknown ids."""
    iu = np.full(len(testset), -1, dtype=np.int32)
    ii = np.full(len(testset), -1, dtype=np.int32)
    for k, (uid, iid, _) in enumerate(testset):
        try:
            iu[k] = trainset.to_inner_uid(uid)
        except ValueError:
            pass
        try:
            ii[k] = trainset.to_inner_iid(iid)
        except ValueError:
            pass
    return iu, ii


def rmse_mae(est, testset, trainset, default):
    """accuracy.rmse / mae semantics on raw estimates (clip to the scale, offset 0 for ml-100k)."""
    lo, hi = trainset.rating_scale
    e = np.clip(np.asarray(est, dtype=np.float64) - trainset.offset, lo, hi)
    r = np.array([t[2] for t in testset]) - trainset.offset
    return float(np.sqrt(np.mean((r - e) ** 2))), float(np.mean(np.abs(r - e)))
