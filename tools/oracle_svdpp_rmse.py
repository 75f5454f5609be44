"""CPU ground truth for BASELINE configs[3]: the oracle's sequential SVD++ (matrix_factorization.pyx:301-504
order) on the synthetic ml-10M-shaped workload of tools/bench_configs.py c4; prints held-out RMSE.
usage: python tools/oracle_svdpp_rmse.py [scale=1.0] [epochs=20]   (CPU only; ~1 h at scale 1)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from surprise_b200 import synth
from surprise_b200.trainset import Trainset
scale = float(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("scale=")), 1.0))
epochs = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("epochs=")), 20))
d = synth.shaped("ml-10m", seed=0, scale=scale)
u, i, r = d["train"]
ts = Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
uu, ii, rr = ts.coo()
uptr, uidx, _ = ts.user_csr()
f = 20
rng = np.random.RandomState(0)
pu = rng.normal(0, .1, (ts.n_users, f)); qi = rng.normal(0, .1, (ts.n_items, f)); yj = rng.normal(0, .1, (ts.n_items, f))
t0 = time.time()
pu, qi, yj, bu, bi = oracle.svdpp_sgd(uu, ii, rr, uptr, uidx, pu, qi, yj, epochs, float(ts.global_mean),
                                      .007, .007, .007, .007, .007, .02, .02, .02, .02, .02)
dt = time.time() - t0
tu, ti, tr = d["test"]
est = oracle.mf_estimate(tu, ti, True, float(ts.global_mean), pu, qi, bu, bi, yj=yj, u_ptr=uptr, ui_idx=uidx)
if isinstance(est, tuple): est = est[0]
rmse = float(np.sqrt(np.mean((np.clip(est, 0.5, 5) - tr) ** 2)))
out = {"scale": scale, "epochs": epochs, "n_ratings": int(len(rr)), "oracle_svdpp_heldout_rmse": rmse, "cpu_s": dt}
print(json.dumps(out), flush=True)
json.dump(out, open(os.path.join(ROOT, "tools", "out", "oracle_svdpp_scale%g.json" % scale), "w"))
