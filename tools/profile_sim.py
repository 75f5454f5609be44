"""Similarity build at a tensor-core-relevant shape (for ncu): cosine, item-based, half-star ratings.
usage: python tools/profile_sim.py [n_items n_users n_ratings kind]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from surprise_b200 import similarities as sims, synth  # noqa: E402
from surprise_b200.trainset import Trainset  # noqa: E402

n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_users = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
n_ratings = int(sys.argv[3]) if len(sys.argv) > 3 else 4_000_000
kind = sys.argv[4] if len(sys.argv) > 4 else "cosine"
d = synth.ratings(n_users, n_items, n_ratings, step=0.5, seed=1, holdout=0.0)
u, i, r = d["train"]
ts = Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
yr = ts.user_csr()
kw = {}
if kind == "pearson_baseline":
    rng = np.random.RandomState(0)
    kw = dict(global_mean=float(ts.global_mean), x_biases=rng.normal(0, .4, ts.n_items), y_biases=rng.normal(0, .4, ts.n_users))
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sims.build_device(kind, ts.n_items, yr, 1, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
n_x, n_y = ts.n_items, ts.n_users
G = 3 if kind == "pearson" else 2
print("kind=%s n_x=%d n_y=%d nnz=%d build %.2f ms (incl. H2D of the CSR); algorithmic %.1f TOP/s (2*G*n_x^2*n_y, G=%d)"
      % (kind, n_x, n_y, len(r), dt * 1e3, 2 * G * n_x * n_x * n_y / dt / 1e12, G))
print("checksum", float(out.sum()))
