"""pearson_baseline build at the ml-20M shape for several plane-band budgets (SB2_SIM_PLANE_GIB)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import surprise_b200 as sb
from surprise_b200 import similarities as sims, synth
d = synth.shaped("ml-20m", seed=0)
u, i, r = d["train"]
ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
algo = sb.KNNBaseline(sim_options={"name": "pearson_baseline", "user_based": False})
sb.AlgoBase.fit(algo, ts)
bu, bi = algo.compute_baselines()
yr = ts.user_csr()
inp = sims.upload_inputs("pearson_baseline", ts.n_items, yr, bi, bu)
kw = dict(global_mean=float(ts.global_mean), x_biases=bi, y_biases=bu, shrinkage=100)
for gib in (48, 24, 12, 6, 3, 24):
    os.environ["SB2_SIM_PLANE_GIB"] = str(gib)
    ts_ = []
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = sims.build_device("pearson_baseline", ts.n_items, yr, 1, inputs=inp, **kw)
        torch.cuda.synchronize(); ts_.append(time.perf_counter() - t0); del out
    print("plane budget %4.0f GiB: %.3f s / %.3f s" % (gib, ts_[0], ts_[1]), flush=True)
