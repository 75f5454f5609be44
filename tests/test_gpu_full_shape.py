"""BASELINE.json configs[1], [2], [4] at their FULL shapes, and the host-buffer forms of the C-ABI (the calls the
bench's `e2e` figure times), against the oracle.

  config 2  SVD f=100 x 20 epochs on the ml-1M-shaped synthetic ratings: held-out RMSE / MAE within 0.005 of
            tests/golden/svd_c2_oracle_rmse.json (tools/oracle_svd_c2.py: the C oracle AND the compiled reference,
            which agree bit for bit there), through the Python API, sb2_svd_fit (host form) and the plan API.
  config 3  27k x 138k item-item pearson_baseline + cosine: 20 000 sampled pairs against oracle.similarity_pairs
            (cosine bit-exact, pearson_baseline 1e-9 absolute), symmetry / unit diagonal on the device.
  config 5  NMF f=15 at the Netflix shape: two epochs bit-identical to the oracle.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat  # noqa: E402
from surprise_b200 import similarities as sims  # noqa: E402
from surprise_b200 import synth  # noqa: E402

RMSE_TOL = 0.005
PB_ATOL = 1e-9


def _sgd_prm(f, epochs, mu, lr=.005, reg=.02, lr_yj=0., reg_yj=0., biased=1):
    return nat.SgdParams(n_factors=f, n_epochs=epochs, biased=biased, reserved=0, global_mean=mu, lr_bu=lr, lr_bi=lr,
                         lr_pu=lr, lr_qi=lr, lr_yj=lr_yj, reg_bu=reg, reg_bi=reg, reg_pu=reg, reg_qi=reg, reg_yj=reg_yj)


def _scores(est, tr, lo=1.0):
    e = np.clip(est, lo, 5.0)
    return float(np.sqrt(np.mean((e - tr) ** 2))), float(np.mean(np.abs(e - tr)))


# ---- config 2 ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2():
    d = synth.shaped("ml-1m", seed=0)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    with open(os.path.join(GOLDEN, "svd_c2_oracle_rmse.json")) as fh:
        gold = json.load(fh)
    assert (ts.n_users, ts.n_items, ts.n_ratings) == (gold["n_users"], gold["n_items"], gold["n_ratings"])
    return d, ts, gold


def test_c2_svd_full_shape_python_api(c2):
    d, ts, gold = c2
    algo = sb.SVD(n_factors=100, n_epochs=20, random_state=0).fit(ts)
    tu, ti, tr = d["test"]
    est, _ = oracle.mf_estimate(tu, ti, True, float(ts.global_mean), algo.pu, algo.qi, algo.bu, algo.bi)
    rmse, mae = _scores(est, tr)
    assert abs(rmse - gold["oracle_heldout_rmse"]) <= RMSE_TOL, (rmse, gold["oracle_heldout_rmse"])
    assert abs(mae - gold["oracle_heldout_mae"]) <= RMSE_TOL, (mae, gold["oracle_heldout_mae"])
    if "reference_heldout_rmse" in gold:   # the compiled reference itself, same trainset
        assert abs(rmse - gold["reference_heldout_rmse"]) <= RMSE_TOL


def test_c2_svd_host_form_and_plan_api(c2):
    """sb2_svd_fit (host buffers: what bench.py's e2e times) and the resident plan API (what bench.py's value
    times) on the full config: both within 0.005 of the golden, and bit-identical to each other (the schedule is
    deterministic)."""
    import torch
    d, ts, gold = c2
    lib = nat.lib()
    uu, ii, rr = (np.ascontiguousarray(a) for a in ts.coo())
    nu, ni, n, f = ts.n_users, ts.n_items, len(rr), 100
    mu = float(ts.global_mean)
    prm = _sgd_prm(f, 20, mu)
    rng = np.random.RandomState(0)
    pu0 = rng.normal(0, .1, (nu, f)); qi0 = rng.normal(0, .1, (ni, f))
    tu, ti, tr = d["test"]
    # host form
    pu, qi = pu0.copy(), qi0.copy()
    bu, bi = np.empty(nu), np.empty(ni)
    nat.check(lib.sb2_svd_fit(nu, ni, n, nat.hptr(uu), nat.hptr(ii), nat.hptr(rr), C.byref(prm), nat.hptr(pu),
                              nat.hptr(qi), nat.hptr(bu), nat.hptr(bi)))
    est, _ = oracle.mf_estimate(tu, ti, True, mu, pu, qi, bu, bi)
    rmse, mae = _scores(est, tr)
    assert abs(rmse - gold["oracle_heldout_rmse"]) <= RMSE_TOL and abs(mae - gold["oracle_heldout_mae"]) <= RMSE_TOL
    # plan API, device-resident inputs
    d_u, d_i, d_r = nat.to_dev(uu, np.int32), nat.to_dev(ii, np.int32), nat.to_dev(rr, np.float64)
    d_pu0, d_qi0 = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
    d_pu, d_qi = torch.empty_like(d_pu0), torch.empty_like(d_qi0)
    d_bu, d_bi = nat.empty_dev((nu,), np.float64), nat.empty_dev((ni,), np.float64)
    plan = C.c_void_p()
    nat.check(lib.sb2_svd_plan_create_dev(nu, ni, n, nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r), C.byref(prm), 0, None,
                                          None, nat.stream(), C.byref(plan)))
    try:
        for _ in range(2):   # reset must restore the initial state: the second fit repeats the first
            nat.check(lib.sb2_svd_plan_reset_dev(plan, nat.ptr(d_pu0), nat.ptr(d_qi0), None, nat.stream()))
            nat.check(lib.sb2_svd_plan_run(plan, 20, nat.stream()))
            nat.check(lib.sb2_svd_plan_read_dev(plan, nat.ptr(d_pu), nat.ptr(d_qi), nat.ptr(d_bu), nat.ptr(d_bi), None,
                                                nat.stream()))
            torch.cuda.synchronize()
            assert np.array_equal(d_pu.cpu().numpy(), pu) and np.array_equal(d_qi.cpu().numpy(), qi)
            assert np.array_equal(d_bu.cpu().numpy(), bu) and np.array_equal(d_bi.cpu().numpy(), bi)
        assert lib.sb2_svd_plan_bytes_per_update(plan) == 2 * (2 * f + 2) * 4 + 12
        # host read-back form of the plan
        pu2, qi2 = np.empty((nu, f)), np.empty((ni, f))
        bu2, bi2 = np.empty(nu), np.empty(ni)
        nat.check(lib.sb2_svd_plan_read(plan, nat.hptr(pu2), nat.hptr(qi2), nat.hptr(bu2), nat.hptr(bi2), None))
        assert np.array_equal(pu2, pu) and np.array_equal(bi2, bi)
    finally:
        lib.sb2_svd_plan_destroy(plan)


# ---- host forms of the other fits ---------------------------------------------------------------------------
def test_svdpp_host_form_vs_oracle():
    """sb2_svdpp_fit with host buffers: held-out RMSE / MAE within 0.005 of the per-rating oracle, and the same
    factors as the Python API (which goes through the device form)."""
    d = synth.ratings(3000, 300, 200_000, seed=5)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    uu, ii, rr = (np.ascontiguousarray(a) for a in ts.coo())
    ptr, idx, _ = ts.user_csr()
    ptr, idx = np.ascontiguousarray(ptr, dtype=np.int64), np.ascontiguousarray(idx, dtype=np.int32)
    tu, ti, tr = d["test"]
    mu, f, ep = float(ts.global_mean), 20, 10
    rng = np.random.RandomState(0)
    pu0 = rng.normal(0, .1, (ts.n_users, f)); qi0 = rng.normal(0, .1, (ts.n_items, f)); yj0 = rng.normal(0, .1, (ts.n_items, f))
    prm = _sgd_prm(f, ep, mu, lr=.007, reg=.02, lr_yj=.007, reg_yj=.02)
    pu, qi, yj = pu0.copy(), qi0.copy(), yj0.copy()
    bu, bi = np.empty(ts.n_users), np.empty(ts.n_items)
    nat.check(nat.lib().sb2_svdpp_fit(ts.n_users, ts.n_items, len(rr), nat.hptr(uu), nat.hptr(ii), nat.hptr(rr),
                                      nat.hptr(ptr), nat.hptr(idx), C.byref(prm), nat.hptr(pu), nat.hptr(qi),
                                      nat.hptr(yj), nat.hptr(bu), nat.hptr(bi)))
    w = oracle.svdpp_sgd(uu, ii, rr, ptr, idx, pu0, qi0, yj0, ep, mu, *([.007] * 5), *([.02] * 5))
    want, _ = oracle.mf_estimate(tu, ti, True, mu, w[0], w[1], w[3], w[4], w[2], ptr, idx)
    got, _ = oracle.mf_estimate(tu, ti, True, mu, pu, qi, bu, bi, yj, ptr, idx)
    (rw, mw), (rg, mg) = _scores(want, tr), _scores(got, tr)
    assert abs(rw - rg) <= RMSE_TOL and abs(mw - mg) <= RMSE_TOL, (rw, rg)
    algo = sb.SVDpp(n_epochs=ep, random_state=0).fit(ts)
    assert np.array_equal(algo.pu, pu) and np.array_equal(algo.yj, yj) and np.array_equal(algo.bi, bi)
    with pytest.raises(ValueError):
        nat.check(nat.lib().sb2_svdpp_fit(ts.n_users, ts.n_items, len(rr), nat.hptr(uu), nat.hptr(ii), nat.hptr(rr),
                                          None, None, C.byref(prm), nat.hptr(pu), nat.hptr(qi), nat.hptr(yj),
                                          nat.hptr(bu), nat.hptr(bi)))


def test_sim_build_host_form_bit_exact(u1, u1_golden, u1_arrays):
    """sb2_sim_build with host buffers on the reference's fixture: sha256 of all four matrices = the goldens the
    compiled reference produced (pearson_baseline: 1e-9 absolute against the golden matrix), plus a row shard."""
    import hashlib
    ts, _ = u1
    ptr, idx, val = ts.user_csr()     # item-based: y = user
    ptr, idx, val = (np.ascontiguousarray(ptr, dtype=np.int64), np.ascontiguousarray(idx, dtype=np.int32),
                     np.ascontiguousarray(val, dtype=np.float64))
    n_x, n_y = ts.n_items, ts.n_users
    mu = float(ts.global_mean)
    bu, bi = oracle.baseline_als(ts.n_users, ts.n_items, *ts.user_csr(), *ts.item_csr(), mu)
    lib = nat.lib()
    for kind, code in (("cosine", 0), ("msd", 1), ("pearson", 2), ("pearson_baseline", 3)):
        out = np.empty((n_x, n_x))
        nat.check(lib.sb2_sim_build(code, n_x, n_y, nat.hptr(ptr), nat.hptr(idx), nat.hptr(val), len(val), 1, 1, mu,
                                    nat.hptr(bi), nat.hptr(bu), 100.0, 0, n_x, nat.hptr(out)))
        if kind == "pearson_baseline":
            want = oracle.similarity(kind, n_x, ptr, idx, val, 1, mu, bi, bu, 100.0)
            assert np.allclose(out, want, rtol=0, atol=PB_ATOL)
        else:
            assert hashlib.sha256(out.tobytes()).hexdigest() == u1_golden["sims"]["%s_item_ms1" % kind]["sha256"], kind
        if kind == "cosine":
            part = np.empty((512 - 256, n_x))
            nat.check(lib.sb2_sim_build(code, n_x, n_y, nat.hptr(ptr), nat.hptr(idx), nat.hptr(val), len(val), 1, 1, mu,
                                        None, None, 100.0, 256, 512, nat.hptr(part)))
            assert np.array_equal(part, out[256:512])
    with pytest.raises(ValueError):
        nat.check(lib.sb2_sim_build(3, n_x, n_y, nat.hptr(ptr), nat.hptr(idx), nat.hptr(val), len(val), 1, 1, mu,
                                    None, None, 100.0, 0, n_x, nat.hptr(out)))


def test_baselines_host_forms_bit_exact(u1, u1_golden):
    """sb2_baseline_als / sb2_baseline_sgd with host buffers: sha256(bu), sha256(bi) = the reference's goldens."""
    import hashlib
    ts, _ = u1
    up, ui, ur = (np.ascontiguousarray(a) for a in ts.user_csr())
    ip, iu, ir = (np.ascontiguousarray(a) for a in ts.item_csr())
    uu, ii, rr = (np.ascontiguousarray(a) for a in ts.coo())
    mu = float(ts.global_mean)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    bu, bi = np.empty(ts.n_users), np.empty(ts.n_items)
    nat.check(nat.lib().sb2_baseline_als(ts.n_users, ts.n_items, nat.hptr(up), nat.hptr(ui), nat.hptr(ur), nat.hptr(ip),
                                         nat.hptr(iu), nat.hptr(ir), mu, 10, 15.0, 10.0, nat.hptr(bu), nat.hptr(bi)))
    assert sha(bu) == u1_golden["baseline_als"]["bu_sha256"] and sha(bi) == u1_golden["baseline_als"]["bi_sha256"]
    nat.check(nat.lib().sb2_baseline_sgd(ts.n_users, ts.n_items, len(rr), nat.hptr(uu), nat.hptr(ii), nat.hptr(rr), mu,
                                         20, 0.02, 0.005, nat.hptr(bu), nat.hptr(bi)))
    assert sha(bu) == u1_golden["baseline_sgd"]["bu_sha256"] and sha(bi) == u1_golden["baseline_sgd"]["bi_sha256"]


# ---- config 3 ---------------------------------------------------------------------------------------------
def test_c3_full_shape_similarities_sampled_pairs(monkeypatch):
    """27k x 138k, 20M half-star ratings, item-item: the whole pearson_baseline and cosine matrices are built on
    the device (5.8 GB each); 20 000 random entries are recomputed by the oracle from the two sparse rows."""
    import torch
    d = synth.shaped("ml-20m", seed=0)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"], (0.5, 5.0), 0)
    n_x, mu = ts.n_items, float(ts.global_mean)
    algo = sb.KNNBaseline(sim_options={"name": "pearson_baseline", "user_based": False})
    sb.AlgoBase.fit(algo, ts)
    bu, bi = algo.compute_baselines()          # ALS on the device (bit-exact vs the oracle: test_u1_baselines_bit_exact)
    yr = ts.user_csr()
    rng = np.random.RandomState(0)
    pi, pj = rng.randint(0, n_x, 20000), rng.randint(0, n_x, 20000)
    pj[:200] = pi[:200]                        # some diagonal entries too
    o = np.lexsort((u, i))
    xptr = np.concatenate(([0], np.cumsum(np.bincount(i, minlength=n_x)))).astype(np.int64)
    d_pi, d_pj = torch.as_tensor(pi, device="cuda"), torch.as_tensor(pj, device="cuda")
    inp = sims.upload_inputs("pearson_baseline", n_x, yr, bi, bu)
    want = {kind: oracle.similarity_pairs(kind, pi, pj, xptr, u[o], r[o], 1, mu, bi, bu, 100.0)
            for kind in ("pearson_baseline", "cosine")}
    for path in ("digit", "general"):     # int8 tensor-core contractions / the reference's loop nest in fp64
        monkeypatch.setenv("SB2_SIM_PATH", path)
        for kind in ("pearson_baseline", "cosine"):
            kw = dict(global_mean=mu, x_biases=bi, y_biases=bu, shrinkage=100) if kind == "pearson_baseline" else {}
            sim = sims.build_device(kind, n_x, yr, 1, inputs=inp, **kw)
            got = sim[d_pi, d_pj].cpu().numpy()
            # symmetry and unit diagonal on a 4096-row band of the device matrix
            band = sim[:4096, :4096]
            assert bool(torch.equal(band, band.t())) and bool(torch.all(torch.diagonal(sim) == 1))
            del sim, band
            torch.cuda.empty_cache()
            w = want[kind].copy()
            w[pi == pj] = 1.0
            if kind == "pearson_baseline" and path == "digit":
                assert np.allclose(got, w, rtol=0, atol=PB_ATOL), float(np.nanmax(np.abs(got - w)))
            else:   # exact integer sums (cosine) / the reference's own fp64 arithmetic (general path): same bits
                assert np.array_equal(got, w), (path, kind, float(np.nanmax(np.abs(got - w))))


# ---- config 5 ---------------------------------------------------------------------------------------------
def test_c5_netflix_shape_nmf_two_epochs_bit_exact():
    """480k x 17.7k, 10^8 ratings, f=15: two epochs on the device, every bit of pu and qi equal to the oracle's."""
    d = synth.shaped("netflix", seed=0)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    del d
    uu, ii, rr = ts.coo()
    f = 15
    rng = np.random.RandomState(0)
    pu0 = rng.uniform(0, 1, (ts.n_users, f)); qi0 = rng.uniform(0, 1, (ts.n_items, f))
    d_u, d_i, d_r = nat.to_dev(uu, np.int32), nat.to_dev(ii, np.int32), nat.to_dev(rr, np.float64)
    d_pu, d_qi = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
    d_bu, d_bi = nat.empty_dev((ts.n_users,), np.float64), nat.empty_dev((ts.n_items,), np.float64)
    prm = nat.NmfParams(n_factors=f, n_epochs=2, biased=0, reserved=0, global_mean=0.0, reg_pu=.06, reg_qi=.06,
                        reg_bu=.02, reg_bi=.02, lr_bu=.005, lr_bi=.005)
    nat.check(nat.lib().sb2_nmf_fit_dev(ts.n_users, ts.n_items, len(rr), nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r),
                                        C.byref(prm), nat.ptr(d_pu), nat.ptr(d_qi), nat.ptr(d_bu), nat.ptr(d_bi),
                                        nat.stream()))
    want = oracle.nmf_sgd(ts.n_users, ts.n_items, uu, ii, rr, np.diff(ts.user_csr()[0]), np.diff(ts.item_csr()[0]),
                          pu0, qi0, 2, False, 0.0, .06, .06, .02, .02, .005, .005)
    assert np.array_equal(d_pu.cpu().numpy(), want[0]) and np.array_equal(d_qi.cpu().numpy(), want[1])
