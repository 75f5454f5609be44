"""BASELINE.json configs[2..4] at full shape on one B200: wall-clock, throughput, achieved fraction of the
kernel's bound, and a parity spot-check against the oracle.  Writes gpurun_out/configs.json.
usage: python tools/bench_configs.py [c3] [c4] [c5] [scale=1.0]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import oracle  # noqa: E402
import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat, similarities as sims, synth  # noqa: E402

which = [a for a in sys.argv[1:] if a.startswith("c")] or ["c3", "c4", "c5"]
scale = float(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("scale=")), 1.0))
PEAK_HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = {}


def sync_time(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


if "c3" in which:
    # KNNBaseline pearson_baseline item-item, ml-20M shape (138k x 27k, 20M half-star ratings)
    t0 = time.perf_counter()
    d = synth.shaped("ml-20m", seed=0, scale=scale)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
    print("c3 data %.1fs: %d users %d items %d ratings" % (time.perf_counter() - t0, ts.n_users, ts.n_items, ts.n_ratings), flush=True)
    algo = sb.KNNBaseline(sim_options={"name": "pearson_baseline", "user_based": False})
    sb.AlgoBase.fit(algo, ts)
    t_als, (bu, bi) = sync_time(algo.compute_baselines)
    yr = ts.user_csr()
    d_yr = [nat.to_dev(yr[0], np.int64), nat.to_dev(yr[1], np.int32), nat.to_dev(yr[2], np.float64)]
    d_bx, d_by = nat.to_dev(bi, np.float64), nat.to_dev(bu, np.float64)
    n_x, n_y = ts.n_items, ts.n_users
    sim = nat.empty_dev((n_x, n_x), np.float64)

    def build():
        nat.check(nat.lib().sb2_sim_build_dev(3, n_x, n_y, *[nat.ptr(t) for t in d_yr], len(yr[2]), 2, 1,
                                              float(ts.global_mean), nat.ptr(d_bx), nat.ptr(d_by), 100.0, 0, n_x,
                                              nat.ptr(sim), nat.stream()))
    t1, _ = sync_time(build)
    t2, _ = sync_time(build)
    rng = np.random.RandomState(0)
    pi, pj = rng.randint(0, n_x, 20000), rng.randint(0, n_x, 20000)
    o = np.lexsort((u, i))
    xptr = np.concatenate(([0], np.cumsum(np.bincount(i, minlength=n_x)))).astype(np.int64)
    want = oracle.similarity_pairs("pearson_baseline", pi, pj, xptr, u[o], r[o], 1, float(ts.global_mean), bi, bu, 100.0)
    got = sim[torch.as_tensor(pi, device="cuda"), torch.as_tensor(pj, device="cuda")].cpu().numpy()
    err = float(np.nanmax(np.abs(got - want)))
    tcos, _ = sync_time(lambda: nat.check(nat.lib().sb2_sim_build_dev(0, n_x, n_y, *[nat.ptr(t) for t in d_yr], len(yr[2]), 2, 1,
                                                                      0.0, None, None, 100.0, 0, n_x, nat.ptr(sim), nat.stream())))
    want = oracle.similarity_pairs("cosine", pi, pj, xptr, u[o], r[o], 1)
    got = sim[torch.as_tensor(pi, device="cuda"), torch.as_tensor(pj, device="cuda")].cpu().numpy()
    flops = 2 * 2 * n_x * n_x * n_y
    out["c3"] = {"workload": "pearson_baseline item-item sim, %dx%d, %d ratings (ml-20M shape x%.2f)" % (n_y, n_x, len(r), scale),
                 "baseline_als_s": t_als, "sim_build_s_first": t1, "sim_build_s": t2, "algorithmic_ops": flops,
                 "algorithmic_TOPs": flops / t2 / 1e12, "max_abs_err_vs_oracle_20000_sampled_pairs": err,
                 "cosine_build_s": tcos, "cosine_TOPs": flops / tcos / 1e12,
                 "cosine_sampled_pairs_bit_exact": bool(np.array_equal(got, want))}
    print(json.dumps(out["c3"]), flush=True)
    del sim, d_yr
    torch.cuda.empty_cache()

if "c4" in which:
    # SVD++ f=20, 20 epochs, ml-10M shape (72k x 10.7k, 10M half-star ratings)
    t0 = time.perf_counter()
    d = synth.shaped("ml-10m", seed=0, scale=scale)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
    print("c4 data %.1fs: %d users %d items %d ratings" % (time.perf_counter() - t0, ts.n_users, ts.n_items, ts.n_ratings), flush=True)
    algo = sb.SVDpp(random_state=0)
    t1, _ = sync_time(lambda: algo.fit(ts))
    t2, _ = sync_time(lambda: algo.fit(ts))
    tu, ti, tr = d["test"]
    est, _ = algo._estimate_batch(tu, ti)
    rmse = float(np.sqrt(np.mean((np.clip(est, 0.5, 5) - tr) ** 2)))
    svd = sb.SVD(n_factors=20, lr_all=.007, random_state=0)
    t3, _ = sync_time(lambda: svd.fit(ts))
    est2, _ = svd._estimate_batch(tu, ti)
    gold = os.path.join(ROOT, "tests", "golden", "svdpp_oracle_rmse.json")
    oracle_rmse = {r_["scale"]: r_ for r_ in json.load(open(gold))["runs"]}.get(scale) if os.path.exists(gold) else None
    out["c4"] = {"oracle_heldout_rmse": oracle_rmse and oracle_rmse["oracle_svdpp_heldout_rmse"],
                 "oracle_c_port_cpu_s": oracle_rmse and oracle_rmse["cpu_s"],
                 "workload": "SVD++ f=20 20 epochs, %dx%d, %d ratings (ml-10M shape x%.2f)" % (ts.n_users, ts.n_items, len(r), scale),
                 "fit_s_first": t1, "fit_s": t2, "updates_per_s": len(r) * 20 / t2, "heldout_rmse": rmse,
                 "svd_f20_fit_s": t3, "svd_f20_updates_per_s": len(r) * 20 / t3,
                 "svd_f20_heldout_rmse": float(np.sqrt(np.mean((np.clip(est2, 0.5, 5) - tr) ** 2)))}
    print(json.dumps(out["c4"]), flush=True)

if "c5" in which:
    # NMF f=15, 50 epochs, Netflix shape (480k x 17.7k, 100M ratings)
    t0 = time.perf_counter()
    d = synth.shaped("netflix", seed=0, scale=scale)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    print("c5 data %.1fs: %d users %d items %d ratings" % (time.perf_counter() - t0, ts.n_users, ts.n_items, ts.n_ratings), flush=True)
    uu, ii, rr = ts.coo()
    f = 15
    rng = np.random.RandomState(0)
    pu0 = rng.uniform(0, 1, (ts.n_users, f)); qi0 = rng.uniform(0, 1, (ts.n_items, f))
    d_u, d_i, d_r = nat.to_dev(uu, np.int32), nat.to_dev(ii, np.int32), nat.to_dev(rr, np.float64)
    d_bu, d_bi = nat.empty_dev((ts.n_users,), np.float64), nat.empty_dev((ts.n_items,), np.float64)

    def fit(n_epochs):
        d_pu, d_qi = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
        prm = nat.NmfParams(n_factors=f, n_epochs=n_epochs, biased=0, reserved=0, global_mean=0.0, reg_pu=.06, reg_qi=.06,
                            reg_bu=.02, reg_bi=.02, lr_bu=.005, lr_bi=.005)
        dt, _ = sync_time(lambda: nat.check(nat.lib().sb2_nmf_fit_dev(ts.n_users, ts.n_items, len(rr), nat.ptr(d_u), nat.ptr(d_i),
                                                                      nat.ptr(d_r), C.byref(prm), nat.ptr(d_pu), nat.ptr(d_qi),
                                                                      nat.ptr(d_bu), nat.ptr(d_bi), nat.stream())))
        return dt, d_pu, d_qi
    t2e, d_pu, d_qi = fit(2)
    t0 = time.perf_counter()
    want = oracle.nmf_sgd(ts.n_users, ts.n_items, uu, ii, rr, np.diff(ts.user_csr()[0]), np.diff(ts.item_csr()[0]), pu0, qi0, 2,
                          False, 0.0, .06, .06, .02, .02, .005, .005)
    t_cpu = time.perf_counter() - t0
    exact = bool(np.array_equal(d_pu.cpu().numpy(), want[0]) and np.array_equal(d_qi.cpu().numpy(), want[1]))
    t50, _, _ = fit(50)
    visits = len(rr) * 50
    out["c5"] = {"workload": "NMF f=15 50 epochs, %dx%d, %d ratings (Netflix shape x%.2f)" % (ts.n_users, ts.n_items, len(rr), scale),
                 "fit_s": t50, "rating_visits_per_s": visits / t50, "algorithmic_bytes_per_visit": 16 * f + 40,
                 "achieved_GBs": (16 * f + 40) * visits / t50 / 1e9, "frac_of_measured_hbm": (16 * f + 40) * visits / t50 / 1e9 / PEAK_HBM,
                 "two_epochs_bit_exact_vs_oracle": exact, "oracle_c_port_2_epochs_s": t_cpu,
                 "oracle_visits_per_s": len(rr) * 2 / t_cpu}
    print(json.dumps(out["c5"]), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
prev = {}
p = os.path.join(ROOT, "gpurun_out", "configs.json")
if os.path.exists(p):
    prev = json.load(open(p))
prev.update(out)
json.dump(prev, open(p, "w"), indent=1)
