"""The multi-rank DSGD ring (SVD / SVD++ sharded by user, item blocks handed rank -> rank inside the kernel).

On ONE GPU all ranks of a ring run as ONE kernel launch (sb2_svd_ring_run_local: n x B co-resident CTAs, CTA b
working as CTA b % B of rank b / B with that rank's arguments): the same kernel body exchanges item blocks through
the very same code path as over NVLink (stores through "peer" pointers + system-scope flags + credits), so the
schedule and the mailboxes are exercised by a single-GPU test run -- without ever issuing kernels that wait for one
another as separate launches.  With >= 2 GPUs the same fits run as real processes over NCCL + cudaIpc
(tests marked with the device-count skip).  Parity: held-out RMSE / MAE within 0.005 of the sequential oracle, and
a conflict-free input (order-independent SGD) reproduced to fp32 rounding.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat  # noqa: E402
from surprise_b200 import distributed as D  # noqa: E402
from surprise_b200 import synth  # noqa: E402

RMSE_TOL = 0.005


def _prm(f, epochs, mu, lr, reg, yj=False, biased=1):
    return nat.SgdParams(n_factors=f, n_epochs=epochs, biased=biased, reserved=0, global_mean=mu, lr_bu=lr, lr_bi=lr,
                         lr_pu=lr, lr_qi=lr, lr_yj=lr if yj else 0., reg_bu=reg, reg_bi=reg, reg_pu=reg, reg_qi=reg,
                         reg_yj=reg if yj else 0.)


def virtual_ring_fit(world, n_users, n_items, u, i, r, prm, pu0, qi0, yj0=None, ur_csr=None, n_epochs=None):
    """All ranks of a ring in this process, one stream per rank, on the current device.
    Returns (pu, qi, bu, bi[, yj]) assembled from the ranks' own rows."""
    import torch
    lib = nat.lib()
    n_epochs = prm.n_epochs if n_epochs is None else n_epochs
    with_yj = yj0 is not None
    f = prm.n_factors
    d_u, d_i, d_r = nat.to_dev(u, np.int32), nat.to_dev(i, np.int32), nat.to_dev(r, np.float64)
    d_up = nat.to_dev(ur_csr[0], np.int64) if with_yj else None
    d_ui = nat.to_dev(ur_csr[1], np.int32) if with_yj else None
    d_pu0, d_qi0 = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
    d_yj0 = nat.to_dev(yj0, np.float64) if with_yj else None
    plans = []
    try:
        for g in range(world):
            plan = C.c_void_p()
            nat.check(lib.sb2_svd_ring_create_dev(n_users, n_items, len(r), nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r),
                                                  C.byref(prm), int(with_yj), nat.ptr(d_up), nat.ptr(d_ui), g, world,
                                                  nat.stream(), C.byref(plan)))
            plans.append(plan)
        for g in range(world):
            left, right = D.ring_neighbours(g, world)
            if world > 1:
                nat.check(lib.sb2_svd_ring_connect_local(plans[g], plans[left], plans[right]))
            nat.check(lib.sb2_svd_plan_reset_dev(plans[g], nat.ptr(d_pu0), nat.ptr(d_qi0), nat.ptr(d_yj0), nat.stream()))
        torch.cuda.synchronize()
        arr = (C.c_void_p * world)(*[p.value for p in plans])
        if not with_yj:
            nat.check(lib.sb2_svd_ring_run_local(arr, world, n_epochs, nat.stream()))
        else:
            stride = C.c_int()
            lib.sb2_svd_ring_info(plans[0], None, None, None, C.byref(stride))
            xch = [torch.empty((n_items, stride.value + 1), dtype=torch.float32, device="cuda") for _ in range(world)]
            xarr = (C.c_void_p * world)(*[t.data_ptr() for t in xch])
            for _ in range(n_epochs):
                nat.check(lib.sb2_svd_ring_epoch_local(arr, world, xarr, nat.stream()))
                total = torch.stack(xch).sum(0)          # what the NCCL all-reduce does across processes
                for g in range(world):
                    nat.check(lib.sb2_svd_ring_epoch_dev(plans[g], 1, nat.ptr(total), nat.stream()))
        torch.cuda.synchronize()
        parts = {k: [] for k in ("pu", "qi", "bu", "bi")}
        yj = None
        for g in range(world):
            nat.check(lib.sb2_svd_plan_status(plans[g], nat.stream()))
            nu, ni = C.c_int64(), C.c_int64()
            lib.sb2_svd_ring_info(plans[g], C.byref(nu), C.byref(ni), None, None)
            assert nu.value == D.local_rows(n_users, g, world) and ni.value == D.local_rows(n_items, g, world)
            t_pu = nat.empty_dev((nu.value, f), np.float64); t_qi = nat.empty_dev((ni.value, f), np.float64)
            t_bu = nat.empty_dev((nu.value,), np.float64); t_bi = nat.empty_dev((ni.value,), np.float64)
            t_yj = nat.empty_dev((n_items, f), np.float64) if with_yj else None
            nat.check(lib.sb2_svd_plan_read_dev(plans[g], nat.ptr(t_pu), nat.ptr(t_qi), nat.ptr(t_bu), nat.ptr(t_bi),
                                                nat.ptr(t_yj), nat.stream()))
            torch.cuda.synchronize()
            for k, t in (("pu", t_pu), ("qi", t_qi), ("bu", t_bu), ("bi", t_bi)):
                parts[k].append(t)
            if with_yj:
                if yj is not None:
                    assert torch.equal(yj, t_yj)     # y_j is replicated: identical on every rank
                yj = t_yj
        out = [D.interleave_rows(parts[k], n_users if k in ("pu", "bu") else n_items, world).cpu().numpy()
               for k in ("pu", "qi", "bu", "bi")]
        if with_yj:
            out.append(yj.cpu().numpy())
        return tuple(out)
    finally:
        import torch
        torch.cuda.synchronize()
        for p in plans:
            lib.sb2_svd_plan_destroy(p)


def _scores(est, tr, lo=1.0):
    e = np.clip(est, lo, 5.0)
    return float(np.sqrt(np.mean((e - tr) ** 2))), float(np.mean(np.abs(e - tr)))


@pytest.mark.parametrize("world", (1, 2, 3))
def test_virtual_ring_svd_rmse_vs_oracle(world, monkeypatch):
    monkeypatch.setenv("SB2_DSGD_BLOCKS", "32")
    d = synth.ratings(2000, 1200, 150_000, seed=12)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    uu, ii, rr = ts.coo()
    tu, ti, tr_ = d["test"]
    mu, f, ep = float(ts.global_mean), 24, 8
    rng = np.random.RandomState(1)
    pu0 = rng.normal(0, .1, (ts.n_users, f)); qi0 = rng.normal(0, .1, (ts.n_items, f))
    got = virtual_ring_fit(world, ts.n_users, ts.n_items, uu, ii, rr, _prm(f, ep, mu, .005, .02), pu0, qi0)
    w = oracle.svd_sgd(uu, ii, rr, pu0, qi0, ep, True, mu, *([.005] * 4), *([.02] * 4))
    want, _ = oracle.mf_estimate(tu, ti, True, mu, *w)
    est, _ = oracle.mf_estimate(tu, ti, True, mu, *got)
    (rw, mw), (rg, mg) = _scores(want, tr_), _scores(est, tr_)
    assert abs(rw - rg) <= RMSE_TOL and abs(mw - mg) <= RMSE_TOL, (world, rw, rg)


@pytest.mark.parametrize("world,blocks", [(2, 32), (3, 16), (4, 8), (2, 1)])
def test_virtual_ring_conflict_free_matches_oracle(world, blocks, monkeypatch):
    """A permutation matrix of ratings makes SGD order-independent: every rating is updated exactly once per epoch
    whatever the schedule, so any block lost, duplicated or overwritten on its way round the ranks shows up as a
    factor that differs from the sequential oracle by more than fp32 rounding.  (2, 1): one CTA per rank, no
    clusters -- every hand-off is a rank hand-off.)"""
    monkeypatch.setenv("SB2_DSGD_BLOCKS", str(blocks))
    n, f, ep = 997, 12, 3
    rng = np.random.RandomState(2)
    u = np.arange(n, dtype=np.int32); i = rng.permutation(n).astype(np.int32)
    r = rng.randint(1, 6, n).astype(np.float64)
    mu = float(np.mean(r))
    pu0 = rng.normal(0, .1, (n, f)); qi0 = rng.normal(0, .1, (n, f))
    for biased in (1, 0):
        got = virtual_ring_fit(world, n, n, u, i, r, _prm(f, ep, mu, .005, .02, biased=biased), pu0, qi0)
        want = oracle.svd_sgd(u, i, r, pu0, qi0, ep, bool(biased), mu, *([.005] * 4), *([.02] * 4))
        for a, b, name in zip(got, want, ("pu", "qi", "bu", "bi")):
            assert np.allclose(a, b, rtol=0, atol=2e-6), (world, biased, name, float(np.abs(a - b).max()))


@pytest.mark.parametrize("world", (1, 2, 3))
def test_virtual_ring_svdpp_rmse_vs_oracle(world, monkeypatch):
    monkeypatch.setenv("SB2_DSGD_BLOCKS", "32")
    d = synth.ratings(3000, 300, 200_000, seed=5)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    uu, ii, rr = ts.coo()
    ptr, idx, _ = ts.user_csr()
    tu, ti, tr_ = d["test"]
    mu, f, ep = float(ts.global_mean), 20, 8
    rng = np.random.RandomState(0)
    pu0 = rng.normal(0, .1, (ts.n_users, f)); qi0 = rng.normal(0, .1, (ts.n_items, f)); yj0 = rng.normal(0, .1, (ts.n_items, f))
    pu, qi, bu, bi, yj = virtual_ring_fit(world, ts.n_users, ts.n_items, uu, ii, rr, _prm(f, ep, mu, .007, .02, yj=True),
                                          pu0, qi0, yj0, (ptr, idx))
    w = oracle.svdpp_sgd(uu, ii, rr, ptr, idx, pu0, qi0, yj0, ep, mu, *([.007] * 5), *([.02] * 5))
    want, _ = oracle.mf_estimate(tu, ti, True, mu, w[0], w[1], w[3], w[4], w[2], ptr, idx)
    est, _ = oracle.mf_estimate(tu, ti, True, mu, pu, qi, bu, bi, yj, ptr, idx)
    (rw, mw), (rg, mg) = _scores(want, tr_), _scores(est, tr_)
    assert abs(rw - rg) <= RMSE_TOL and abs(mw - mg) <= RMSE_TOL, (world, rw, rg)


def test_ring_wait_deadline_is_an_error_not_a_hang():
    """A rank whose neighbours never run: its kernel gives up after the bounded wait and the status call reports
    SB2_ERR_CUDA (NativeError) instead of the GPU hanging."""
    import torch
    lib = nat.lib()
    d = synth.ratings(300, 200, 5_000, seed=1)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    uu, ii, rr = ts.coo()
    f = 8
    prm = _prm(f, 2, float(ts.global_mean), .005, .02)
    d_u, d_i, d_r = nat.to_dev(uu, np.int32), nat.to_dev(ii, np.int32), nat.to_dev(rr, np.float64)
    rng = np.random.RandomState(0)
    d_pu0 = nat.to_dev(rng.normal(0, .1, (ts.n_users, f)), np.float64)
    d_qi0 = nat.to_dev(rng.normal(0, .1, (ts.n_items, f)), np.float64)
    plans = []
    for g in range(2):
        plan = C.c_void_p()
        nat.check(lib.sb2_svd_ring_create_dev(ts.n_users, ts.n_items, len(rr), nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r),
                                              C.byref(prm), 0, None, None, g, 2, nat.stream(), C.byref(plan)))
        plans.append(plan)
    try:
        with pytest.raises(ValueError):      # not connected yet
            nat.check(lib.sb2_svd_plan_run(plans[0], 2, nat.stream()))
        for g in range(2):
            nat.check(lib.sb2_svd_ring_connect_local(plans[g], plans[1 - g], plans[1 - g]))
            nat.check(lib.sb2_svd_plan_reset_dev(plans[g], nat.ptr(d_pu0), nat.ptr(d_qi0), None, nat.stream()))
        nat.check(lib.sb2_svd_plan_run(plans[0], 2, nat.stream()))   # rank 1 never launches
        with pytest.raises(nat.NativeError):
            nat.check(lib.sb2_svd_plan_status(plans[0], nat.stream()))
    finally:
        torch.cuda.synchronize()
        for p in plans:
            lib.sb2_svd_plan_destroy(p)


# ---- real processes, one per GPU (needs >= 2 GPUs: `gpurun --gpus 2`) -----------------------------------------
def _proc_worker(rank, world, port, out_dir, algo_name):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = synth.ratings(3000, 1500, 200_000, seed=7)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    algo = sb.SVD(n_factors=40, n_epochs=10, random_state=3) if algo_name == "svd" else sb.SVDpp(n_epochs=8, random_state=3)
    D.fit_sharded(algo, ts, dist)
    tu, ti, _ = d["test"]
    est, _ = algo._estimate_batch(tu, ti)
    np.savez(os.path.join(out_dir, "%s_r%d.npz" % (algo_name, rank)), est=est, pu=algo.pu, qi=algo.qi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("algo_name", ("svd", "svdpp"))
def test_ring_processes_over_nvlink(tmp_path, algo_name):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_proc_worker, args=(world, port, str(tmp_path), algo_name), nprocs=world, join=True)
    d = synth.ratings(3000, 1500, 200_000, seed=7)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    tu, ti, tr_ = d["test"]
    outs = [np.load(os.path.join(str(tmp_path), "%s_r%d.npz" % (algo_name, g))) for g in range(world)]
    for o in outs[1:]:      # every rank ends up with the same model
        assert np.array_equal(o["pu"], outs[0]["pu"]) and np.array_equal(o["qi"], outs[0]["qi"])
    single = (sb.SVD(n_factors=40, n_epochs=10, random_state=3) if algo_name == "svd"
              else sb.SVDpp(n_epochs=8, random_state=3)).fit(ts)
    est1, _ = single._estimate_batch(tu, ti)
    uu, ii, rr = ts.coo()
    mu = float(ts.global_mean)
    rng = np.random.RandomState(3)
    if algo_name == "svd":
        pu0 = rng.normal(0, .1, (ts.n_users, 40)); qi0 = rng.normal(0, .1, (ts.n_items, 40))
        w = oracle.svd_sgd(uu, ii, rr, pu0, qi0, 10, True, mu, *([.005] * 4), *([.02] * 4))
        want, _ = oracle.mf_estimate(tu, ti, True, mu, *w)
    else:
        ptr, idx, _ = ts.user_csr()
        pu0 = rng.normal(0, .1, (ts.n_users, 20)); qi0 = rng.normal(0, .1, (ts.n_items, 20)); yj0 = rng.normal(0, .1, (ts.n_items, 20))
        w = oracle.svdpp_sgd(uu, ii, rr, ptr, idx, pu0, qi0, yj0, 8, mu, *([.007] * 5), *([.02] * 5))
        want, _ = oracle.mf_estimate(tu, ti, True, mu, w[0], w[1], w[3], w[4], w[2], ptr, idx)
    rw, mw = _scores(want, tr_)
    rg, mg = _scores(outs[0]["est"], tr_)
    r1, _ = _scores(est1, tr_)
    assert abs(rw - rg) <= RMSE_TOL and abs(mw - mg) <= RMSE_TOL, (rw, rg, r1)


# ---- the paths that shard without a per-step exchange, as real processes --------------------------------------------
def _sharded_worker(rank, world, port, out_dir):
    import json
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from surprise_b200 import similarities as sims
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = {}
    d = synth.ratings(9000, 3000, 500_000, step=0.5, seed=9, holdout=0.0)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
    yr = ts.user_csr()
    mu = float(ts.global_mean)
    rng = np.random.RandomState(1)
    bx, by = rng.normal(0, .3, ts.n_items), rng.normal(0, .3, ts.n_users)
    for path in ("digit", "general"):
        os.environ["SB2_SIM_PATH"] = path
        for kind in ("cosine", "pearson", "pearson_baseline"):
            kw = dict(global_mean=mu, x_biases=bx, y_biases=by, shrinkage=100) if kind == "pearson_baseline" else {}
            _, _, full = D.sim_build_sharded(dist, kind, ts.n_items, yr, 1, gather=True, **kw)   # CSR broadcast from rank 0
            ref = sims.build_device(kind, ts.n_items, yr, 1, **kw)
            out["sim_%s_%s" % (path, kind)] = bool(torch.equal(full, ref))
    os.environ.pop("SB2_SIM_PATH")
    # k-NN estimates with the matrix left sharded == estimates on the full matrix
    px = rng.randint(-1, ts.n_items, 20000).astype(np.int32); py = rng.randint(0, ts.n_users, 20000).astype(np.int32)
    b, e, block = D.sim_build_sharded(dist, "msd", ts.n_items, yr, 1)
    got = D.knn_predict_sharded(dist, block, b, e, ts.n_items, px, py, yr, 40, 1)
    full = sims.build_device("msd", ts.n_items, yr, 1)
    ref = D.knn_predict_sharded(None, full, 0, ts.n_items, ts.n_items, px, py, yr, 40, 1)
    out["knn_sharded"] = bool(all(np.array_equal(a, c) for a, c in zip(got, ref)))
    # NMF: accumulators sharded, in-place block all-gather per epoch: bit-identical to the single-GPU fit
    d2 = synth.ratings(10001, 1999, 600_000, seed=3, holdout=0.0)
    u2, i2, r2 = d2["train"]
    ts2 = sb.Trainset.from_coo(u2, i2, r2, d2["n_users"], d2["n_items"])
    uu, ii, rr = ts2.coo()
    pu0 = rng.uniform(0, 1, (ts2.n_users, 15)); qi0 = rng.uniform(0, 1, (ts2.n_items, 15))
    for biased in (0, 1):
        prm = nat.NmfParams(n_factors=15, n_epochs=4, biased=biased, reserved=0, global_mean=float(ts2.global_mean),
                            reg_pu=.06, reg_qi=.06, reg_bu=.02, reg_bi=.02, lr_bu=.005, lr_bi=.005)
        got = D.nmf_fit_sharded(dist, ts2.n_users, ts2.n_items, uu, ii, rr, prm, pu0, qi0)
        ref = D.nmf_fit_sharded(None, ts2.n_users, ts2.n_items, uu, ii, rr, prm, pu0, qi0)
        out["nmf_biased%d" % biased] = bool(all(np.array_equal(a, c) for a, c in zip(got, ref)))
    # ALS baselines: segment ranges per rank, in-place all-gather between the two passes of an epoch
    got = D.baseline_als_sharded(dist, ts2, 10, 15.0, 10.0)
    ref = oracle.baseline_als(ts2.n_users, ts2.n_items, *ts2.user_csr(), *ts2.item_csr(), float(ts2.global_mean))
    out["baseline_als"] = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]))
    with open(os.path.join(out_dir, "sharded_r%d.json" % rank), "w") as fh:
        json.dump(out, fh)
    dist.barrier()
    dist.destroy_process_group()


def test_baseline_als_range_passes_bit_exact(u1, u1_golden):
    """sb2_baseline_als_pass_dev (the unit of the multi-rank ALS) driven by baseline_als_sharded without a process
    group: sha256(bu), sha256(bi) = the reference's goldens on the fixture."""
    import hashlib
    ts, _ = u1
    bu, bi = D.baseline_als_sharded(None, ts, 10, 15.0, 10.0)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(bu) == u1_golden["baseline_als"]["bu_sha256"] and sha(bi) == u1_golden["baseline_als"]["bi_sha256"]


def test_sharded_similarity_knn_nmf_processes(tmp_path):
    """Row-sharded similarity build (both implementations, rating CSR broadcast over NCCL, symmetric shards + transpose
    exchange), k-NN estimates on the sharded matrix, NMF with sharded accumulators, ALS baselines by segment ranges:
    every rank's result must equal the single-GPU result / the oracle bit for bit."""
    import json
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    port = 36500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for g in range(world):
        with open(os.path.join(str(tmp_path), "sharded_r%d.json" % g)) as fh:
            res = json.load(fh)
        assert all(res.values()), (g, res)
