"""CPU ground truth for BASELINE configs[1] (SVD f=100, 20 epochs, random_state=0 on the ml-1M-shaped synthetic
ratings of bench.py): the oracle's sequential SVD.sgd (matrix_factorization.pyx:241-262 order) and, where
oracle/_ref is present, the compiled reference itself on the same trainset; writes the held-out RMSE / MAE to
tests/golden/svd_c2_oracle_rmse.json.  CPU only (~10 s for the C port, ~2 min for the Cython reference).
usage: python tools/oracle_svd_c2.py [noref]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from surprise_b200 import synth  # noqa: E402
from surprise_b200.trainset import Trainset  # noqa: E402

F, EPOCHS, SEED = 100, 20, 0
d = synth.shaped("ml-1m", seed=SEED)
u, i, r = d["train"]
ts = Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
uu, ii, rr = ts.coo()
mu = float(ts.global_mean)
rng = np.random.RandomState(SEED)
pu0 = rng.normal(0, .1, (ts.n_users, F)); qi0 = rng.normal(0, .1, (ts.n_items, F))
tu, ti, tr = d["test"]


def score(pu, qi, bu, bi):
    est, imp = oracle.mf_estimate(tu, ti, True, mu, pu, qi, bu, bi)
    e = np.clip(est, 1, 5)
    return float(np.sqrt(np.mean((e - tr) ** 2))), float(np.mean(np.abs(e - tr)))


t0 = time.time()
pu, qi, bu, bi = oracle.svd_sgd(uu, ii, rr, pu0, qi0, EPOCHS, True, mu, *([.005] * 4), *([.02] * 4))
dt = time.time() - t0
rmse, mae = score(pu, qi, bu, bi)
out = {"generator": "tools/oracle_svd_c2.py: synth.shaped('ml-1m', seed=0), SVD f=100 20 epochs RandomState(0) init, "
                    "held-out = the synthetic test split (clip to [1, 5])",
       "n_users": ts.n_users, "n_items": ts.n_items, "n_ratings": int(len(rr)),
       "oracle_heldout_rmse": rmse, "oracle_heldout_mae": mae, "oracle_cpu_s": dt}
print(json.dumps(out), flush=True)
if "noref" not in sys.argv[1:]:
    try:
        ref = oracle.import_reference()
        sys.path.insert(0, ROOT)
        import bench
        rts, kept = bench.reference_trainset(ref, ts)
        algo = ref.SVD(n_factors=F, n_epochs=EPOCHS, random_state=SEED)
        ref.AlgoBase.fit(algo, rts)
        t0 = time.time()
        algo.sgd(rts)
        out["reference_cpu_s"] = time.time() - t0
        out["reference_heldout_rmse"], out["reference_heldout_mae"] = score(
            np.asarray(algo.pu), np.asarray(algo.qi), np.asarray(algo.bu), np.asarray(algo.bi))
        out["reference_equals_oracle_bitwise"] = bool(np.array_equal(np.asarray(algo.pu), pu)
                                                      and np.array_equal(np.asarray(algo.qi), qi))
        print(json.dumps(out), flush=True)
    except ImportError as e:
        out["reference"] = "unavailable: %r" % (e,)
with open(os.path.join(ROOT, "tests", "golden", "svd_c2_oracle_rmse.json"), "w") as fh:
    json.dump(out, fh, indent=1)
