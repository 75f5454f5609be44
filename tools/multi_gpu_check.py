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
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
