"""cross_validate: the caller of the hot path (reference: model_selection/validation.py:29-142,
fit_and_score :683-769), single process -- fit/test run on the GPU, fold-level joblib fan-out would
only multiply CUDA contexts."""
import time

import numpy as np

from .. import accuracy
from .split import KFold


def fit_and_score(algo, trainset, testset, measures, return_train_measures=False):
    t0 = time.time()
    algo.fit(trainset)
    fit_time = time.time() - t0
    t0 = time.time()
    predictions = algo.test(testset)
    test_time = time.time() - t0
    train_predictions = algo.test(trainset.build_testset()) if return_train_measures else None
    test_m, train_m = {}, {}
    for m in measures:
        f = getattr(accuracy, m.lower())
        test_m[m] = f(predictions, verbose=0)
        if return_train_measures:
            train_m[m] = f(train_predictions, verbose=0)
    return test_m, train_m, fit_time, test_time


def cross_validate(algo, data, measures=["rmse", "mae"], cv=None, return_train_measures=False, n_jobs=1,
                   pre_dispatch="2*n_jobs", verbose=False):
    measures = [m.lower() for m in measures]
    cv = KFold(n_splits=5) if cv is None else (KFold(n_splits=cv) if isinstance(cv, int) else cv)
    outs = [fit_and_score(algo, tr, te, measures, return_train_measures) for tr, te in cv.split(data)]
    ret = {}
    for m in measures:
        ret["test_" + m] = np.array([o[0][m] for o in outs])
        if return_train_measures:
            ret["train_" + m] = np.array([o[1][m] for o in outs])
    ret["fit_time"] = tuple(o[2] for o in outs)
    ret["test_time"] = tuple(o[3] for o in outs)
    if verbose:
        for k, v in ret.items():
            print(k, v)
    return ret
