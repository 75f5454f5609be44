"""General (sparse, fp64, reference-order) similarity path against the tensor-core digit path at the ml-20M shape.
usage: python tools/time_general_sim.py [scale=1.0]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import surprise_b200 as sb  # noqa: E402
from surprise_b200 import similarities as sims, synth  # noqa: E402

scale = float(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("scale=")), 1.0))
d = synth.shaped("ml-20m", seed=0, scale=scale)
u, i, r = d["train"]
ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
algo = sb.KNNBaseline(sim_options={"name": "pearson_baseline", "user_based": False})
sb.AlgoBase.fit(algo, ts)
bu, bi = algo.compute_baselines()
yr = ts.user_csr()
n_x = ts.n_items
visits = float(np.sum(np.diff(yr[0]).astype(np.float64) ** 2))
print("shape %d x %d, %d ratings, pair visits %.3g" % (ts.n_users, n_x, ts.n_ratings, visits), flush=True)
inp = sims.upload_inputs("pearson_baseline", n_x, yr, bi, bu)
kw = dict(global_mean=float(ts.global_mean), x_biases=bi, y_biases=bu, shrinkage=100)
res = {}
for mode in ("digit", "general"):
    os.environ["SB2_SIM_PATH"] = mode
    for kind in ("cosine", "pearson_baseline"):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            s = sims.build_device(kind, n_x, yr, 1, inputs=inp, **(kw if kind == "pearson_baseline" else {}))
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if rep == 1:
                res[(mode, kind)] = s[:2048].cpu().numpy()
            del s
        print("%-8s %-17s %.3f s  (%.3g pair visits/s)" % (mode, kind, dt, visits / dt), flush=True)
print("cosine rows equal bit for bit:", np.array_equal(res[("digit", "cosine")], res[("general", "cosine")]))
print("pearson_baseline max abs diff (digit vs general = reference arithmetic): %.3g"
      % float(np.nanmax(np.abs(res[("digit", "pearson_baseline")] - res[("general", "pearson_baseline")]))))
