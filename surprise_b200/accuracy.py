"""rmse / mae / fcp over a list of Prediction tuples (reference: surprise/accuracy.py:22-143)."""
from collections import defaultdict

import numpy as np


def _need(predictions):
    if not predictions:
        raise ValueError("Prediction list is empty.")


def rmse(predictions, verbose=True):
    _need(predictions)
    mse = np.mean([float((true_r - est) ** 2) for (_, _, true_r, est, _) in predictions])
    out = np.sqrt(mse)
    if verbose:
        print("RMSE: {0:1.4f}".format(out))
    return out


def mae(predictions, verbose=True):
    _need(predictions)
    out = np.mean([float(abs(true_r - est)) for (_, _, true_r, est, _) in predictions])
    if verbose:
        print("MAE:  {0:1.4f}".format(out))
    return out


def fcp(predictions, verbose=True):
    _need(predictions)
    per_user = defaultdict(list)
    for u0, _, r0, est, _ in predictions:
        per_user[u0].append((est, r0))
    nc, nd = {}, {}
    for u0, prefs in per_user.items():
        c = d = 0
        for esi, ri in prefs:
            for esj, rj in prefs:
                if esi > esj and ri > rj:
                    c += 1
                if esi >= esj and ri < rj:
                    d += 1
        nc[u0], nd[u0] = c, d
    c_mean = np.mean(list(nc.values())) if nc else 0
    d_mean = np.mean(list(nd.values())) if nd else 0
    try:
        out = c_mean / (c_mean + d_mean)
    except ZeroDivisionError:
        raise ValueError("cannot compute fcp on this list of prediction. Does every user have at least two "
                         "predictions?")
    if verbose:
        print("FCP:  {0:1.4f}".format(out))
    return out
