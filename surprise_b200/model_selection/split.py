"""Cross-validation iterators used by the hot-path tests (reference: model_selection/split.py):
PredefinedKFold (:654-685), KFold (:53-122), train_test_split (:543-600)."""
import numpy as np

from ..utils import get_rng


class PredefinedKFold(object):
    """Folds come from Dataset.load_from_folds()."""

    def split(self, data):
        self.n_splits = len(data.folds_files)
        for raw_train, raw_test in data.raw_folds():
            yield data.construct_trainset(raw_train), data.construct_testset(raw_test)

    def get_n_folds(self):
        return self.n_splits


class KFold(object):
    def __init__(self, n_splits=5, random_state=None, shuffle=True):
        self.n_splits, self.random_state, self.shuffle = n_splits, random_state, shuffle

    def split(self, data):
        n = len(data.raw_ratings)
        if self.n_splits > n or self.n_splits < 2:
            raise ValueError("Incorrect value for n_splits={0}. Must be >=2 and less than the number of ratings"
                             .format(n))
        idx = np.arange(n)
        if self.shuffle:
            get_rng(self.random_state).shuffle(idx)
        start = stop = 0
        for fold in range(self.n_splits):
            start = stop
            stop += n // self.n_splits + (1 if fold < n % self.n_splits else 0)
            test = [data.raw_ratings[k] for k in idx[start:stop]]
            train = [data.raw_ratings[k] for k in np.concatenate((idx[:start], idx[stop:]))]
            yield data.construct_trainset(train), data.construct_testset(test)

    def get_n_folds(self):
        return self.n_splits


def train_test_split(data, test_size=.2, random_state=None, shuffle=True):
    n = len(data.raw_ratings)
    n_test = int(np.ceil(test_size * n)) if isinstance(test_size, float) else int(test_size)
    idx = np.arange(n)
    if shuffle:
        idx = get_rng(random_state).permutation(n)
    test = [data.raw_ratings[k] for k in idx[:n_test]]
    train = [data.raw_ratings[k] for k in idx[n_test:]]
    return data.construct_trainset(train), data.construct_testset(test)
