"""torchrun --nproc-per-node N tools/multi_gpu_check.py : sharded similarity build == single-GPU build (bit-exact),
ring SVD RMSE == single-GPU RMSE (to 0.005).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import surprise_b200 as sb
from surprise_b200 import distributed as D, similarities as sims, synth

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
d = synth.ratings(20000, 6000, 1_500_000, step=0.5, seed=9, holdout=0.0)
u, i, r = d["train"]
ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
yr = ts.user_csr()
out = {"world": world}
for kind in ("cosine", "pearson"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, _, full = D.sim_build_sharded(dist if world > 1 else None, kind, ts.n_items, yr, 1, gather=True)
    torch.cuda.synchronize(); out[kind + "_sharded_s"] = time.perf_counter() - t0
    ref = sims.build_device(kind, ts.n_items, yr, 1)
    out[kind + "_bit_exact"] = bool(torch.equal(full, ref))
# k-NN estimates with the matrix left sharded == single-GPU estimates on the full matrix
rng = np.random.RandomState(5)
px = rng.randint(-1, ts.n_items, 50000).astype(np.int32); py = rng.randint(0, ts.n_users, 50000).astype(np.int32)
b, e, block = D.sim_build_sharded(dist if world > 1 else None, "msd", ts.n_items, yr, 1)
got = D.knn_predict_sharded(dist if world > 1 else None, block, b, e, ts.n_items, px, py, yr, 40, 1)
full = sims.build_device("msd", ts.n_items, yr, 1)
ref = D.knn_predict_sharded(None, full, 0, ts.n_items, ts.n_items, px, py, yr, 40, 1)
out["knn_sharded_bit_exact"] = bool(all(np.array_equal(a, c) for a, c in zip(got, ref)))
del block, full
# NMF sharded over ranks must be bit-identical to the single-GPU fit
import ctypes as C
from surprise_b200 import _native as nat
d2 = synth.ratings(30000, 4000, 2_000_000, seed=3, holdout=0.0)
u2, i2, r2 = d2["train"]
ts2 = sb.Trainset.from_coo(u2, i2, r2, d2["n_users"], d2["n_items"])
uu, ii, rr = ts2.coo()
rng = np.random.RandomState(0)
pu0 = rng.uniform(0, 1, (ts2.n_users, 15)); qi0 = rng.uniform(0, 1, (ts2.n_items, 15))
for biased in (0, 1):
    prm = nat.NmfParams(n_factors=15, n_epochs=5, biased=biased, reserved=0, global_mean=float(ts2.global_mean), reg_pu=.06,
                        reg_qi=.06, reg_bu=.02, reg_bi=.02, lr_bu=.005, lr_bi=.005)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got = D.nmf_fit_sharded(dist if world > 1 else None, ts2.n_users, ts2.n_items, uu, ii, rr, prm, pu0, qi0)
    torch.cuda.synchronize(); out["nmf_sharded_s_biased%d" % biased] = time.perf_counter() - t0
    ref = D.nmf_fit_sharded(None, ts2.n_users, ts2.n_items, uu, ii, rr, prm, pu0, qi0)
    out["nmf_bit_exact_biased%d" % biased] = bool(all(np.array_equal(a, b) for a, b in zip(got, ref)))
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
