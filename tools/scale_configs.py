"""BASELINE configs[2] / [4] across the GPUs of one box: run under torchrun (or plain python for N=1).
  c3: pearson_baseline item-item similarity at the ml-20M shape, symmetric row-block sharding (each rank computes
      1/N of the upper-triangular tiles, one NCCL exchange of the transposed blocks); output stays sharded.
  c5: NMF f=15 at the Netflix shape, accumulators sharded by user / item range, all-gather of the factors per epoch.
Time = max over ranks of the wall clock between two barriers (device synchronised).  One JSON line on rank 0.
usage: torchrun --nproc-per-node N tools/scale_configs.py [c3] [c5] [scale=1.0] [epochs=50]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat, distributed as D, synth  # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
which = [a for a in sys.argv[1:] if a.startswith("c")] or ["c3", "c5"]
scale = float(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("scale=")), 1.0))
epochs = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("epochs=")), 50))
dd = dist if world > 1 else None
out = {"n_gpus": world, "scale": scale}


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt) if best is None else min(best, float(dt))
        del r
    return best


if "c3" in which:
    d = synth.shaped("ml-20m", seed=0, scale=scale)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
    algo = sb.KNNBaseline(sim_options={"name": "pearson_baseline", "user_based": False})
    sb.AlgoBase.fit(algo, ts)
    bu, bi = algo.compute_baselines()
    yr = ts.user_csr()
    kw = dict(global_mean=float(ts.global_mean), x_biases=bi, y_biases=bu, shrinkage=100)
    from surprise_b200 import similarities as sims
    inp = sims.upload_inputs("pearson_baseline", ts.n_items, yr, bi, bu)     # ratings + baselines resident in HBM
    t = timed(lambda: D.sim_build_sharded(dd, "pearson_baseline", ts.n_items, yr, 1, inputs=inp, **kw), reps=3)
    out["c3_pearson_baseline_build_s"] = t
    out["c3_shape"] = "%d x %d, %d ratings" % (ts.n_users, ts.n_items, ts.n_ratings)
    t = timed(lambda: D.sim_build_sharded(dd, "cosine", ts.n_items, yr, 1, inputs=inp), reps=3)
    out["c3_cosine_build_s"] = t
    t = timed(lambda: D.sim_build_sharded(dd, "pearson_baseline", ts.n_items, yr, 1, **kw), reps=1)
    out["c3_pearson_baseline_build_s_from_host_csr"] = t
    del inp
    del ts, yr

if "c5" in which:
    import ctypes as C
    d = synth.shaped("netflix", seed=0, scale=scale)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    uu, ii, rr = ts.coo()
    rng = np.random.RandomState(0)
    pu0 = rng.uniform(0, 1, (ts.n_users, 15)); qi0 = rng.uniform(0, 1, (ts.n_items, 15))
    prm = nat.NmfParams(n_factors=15, n_epochs=epochs, biased=0, reserved=0, global_mean=0.0, reg_pu=.06, reg_qi=.06,
                        reg_bu=.02, reg_bi=.02, lr_bu=.005, lr_bi=.005)
    st = {}
    t = timed(lambda: D.nmf_fit_sharded(dd, ts.n_users, ts.n_items, uu, ii, rr, prm, pu0, qi0, stats=st), reps=2)
    ep = torch.tensor([st["epochs_s"]], device="cuda")
    if world > 1:
        dist.all_reduce(ep, op=dist.ReduceOp.MAX)
    out["c5_nmf_fit_s_incl_upload_and_plan"] = t
    out["c5_nmf_epochs_s"] = float(ep)
    out["c5_visits_per_s"] = len(rr) * epochs / float(ep)
    out["c5_shape"] = "%d x %d, %d ratings, %d epochs" % (ts.n_users, ts.n_items, len(rr), epochs)

if rank == 0:
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "scale_n%d.json" % world), "w") as fh:
        json.dump(out, fh, indent=1)
if world > 1:
    dist.destroy_process_group()
