"""Throughput of the batched estimate kernels at the ml-1M shape (device-resident inputs, CUDA-event timed):
knn_predict_kernel (KNNBasic / KNNBaseline, k=40) and mf_predict_kernel (SVD f=100), with the achieved fraction of the
measured HBM peak on their algorithmic bytes.  Writes gpurun_out/predict.json."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat, similarities as sims, synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
d = synth.shaped("ml-1m", seed=0)
u, i, r = d["train"]
ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
lib = nat.lib()
out = {"shape": "%d users x %d items, %d ratings" % (ts.n_users, ts.n_items, ts.n_ratings)}
rng = np.random.RandomState(0)
n_pairs = 2_000_000
pu_ = rng.randint(0, ts.n_users, n_pairs).astype(np.int32)
pi_ = rng.randint(0, ts.n_items, n_pairs).astype(np.int32)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


# ---- k-NN, item-based: x = item, y = user; neighbours = the user's ratings ---------------------------------------
yr = ts.user_csr()
sim = sims.build_device("msd", ts.n_items, yr, 1)
d_ptr, d_idx, d_val = nat.to_dev(yr[0], np.int64), nat.to_dev(yr[1], np.int32), nat.to_dev(yr[2], np.float64)
d_x, d_y = nat.to_dev(pi_, np.int32), nat.to_dev(pu_, np.int32)
est = nat.empty_dev((n_pairs,), np.float64); ak = nat.empty_dev((n_pairs,), np.int32); imp = nat.empty_dev((n_pairs,), np.uint8)
bx = nat.to_dev(rng.normal(0, .3, ts.n_items), np.float64); by = nat.to_dev(rng.normal(0, .3, ts.n_users), np.float64)
lens = np.diff(yr[0])[pu_]
bytes_pair = float(np.mean(lens)) * (8 + 4 + 8) + 8 + 8 + 13     # sim gather + idx + r per neighbour, ids, outputs
for mode, name in ((0, "knn_basic"), (2, "knn_baseline")):
    t = timeit(lambda: nat.check(lib.sb2_knn_predict_dev(n_pairs, nat.ptr(d_x), nat.ptr(d_y), ts.n_items, nat.ptr(sim), ts.n_items,
                                                         nat.ptr(d_ptr), nat.ptr(d_idx), nat.ptr(d_val), 40, 1, mode,
                                                         float(ts.global_mean), nat.ptr(bx), nat.ptr(by), nat.ptr(est), nat.ptr(ak),
                                                         nat.ptr(imp), nat.stream())))
    out[name] = {"pairs_per_s": n_pairs / t, "mean_neighbour_list": float(np.mean(lens)), "algorithmic_bytes_per_pair": bytes_pair,
                 "achieved_GBs": bytes_pair * n_pairs / t / 1e9, "frac_of_measured_hbm": bytes_pair * n_pairs / t / 1e9 / PEAK}

# ---- factor model ---------------------------------------------------------------------------------------------------
f = 100
pu = nat.to_dev(rng.normal(0, .1, (ts.n_users, f)), np.float64); qi = nat.to_dev(rng.normal(0, .1, (ts.n_items, f)), np.float64)
bu = nat.to_dev(rng.normal(0, .1, ts.n_users), np.float64); bi = nat.to_dev(rng.normal(0, .1, ts.n_items), np.float64)
d_u, d_i = nat.to_dev(pu_, np.int32), nat.to_dev(pi_, np.int32)
t = timeit(lambda: nat.check(lib.sb2_mf_predict_dev(n_pairs, nat.ptr(d_u), nat.ptr(d_i), f, 1, float(ts.global_mean), nat.ptr(pu),
                                                    nat.ptr(qi), nat.ptr(bu), nat.ptr(bi), None, None, None, nat.ptr(est), nat.ptr(imp),
                                                    nat.stream())))
b = 2 * f * 8 + 16 + 8 + 9
out["mf_predict_f100"] = {"pairs_per_s": n_pairs / t, "algorithmic_bytes_per_pair": b, "achieved_GBs": b * n_pairs / t / 1e9,
                          "frac_of_measured_hbm": b * n_pairs / t / 1e9 / PEAK,
                          "note": "factor matrices (7.8 MB) are L2-resident: the gathers are served by L2, not HBM"}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "predict.json"), "w"), indent=1)
