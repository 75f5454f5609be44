"""Pins the C oracle (oracle/surprise_oracle.c) to the golden vectors generated from the compiled,
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import inner_pairs, rmse_mae


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


KINDS = ("cosine", "msd", "pearson", "pearson_baseline")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ms", (1, 4))
def test_toy_similarities_bit_exact(toy, kind, ms):
    sim = oracle.similarity(kind, 8, toy["y_ptr"], toy["x_idx"], toy["r"], ms, float(toy["global_mean"]),
                            toy["bx"], toy["by"], 100.0)
    assert np.array_equal(sim, toy["%s_%d" % (kind, ms)])


def test_toy_known_answers(toy):
    """The value-pinning asserts of the reference's tests/test_similarities.py."""
    a = (toy["y_ptr"], toy["x_idx"], toy["r"])
    cos = oracle.similarity("cosine", 8, *a, 1)
    assert cos[0, 1] == 1 and cos[0, 2] == 1 and cos[3, 4] == 1 and cos[0, 3] == 0 and cos[0, 4] == 0
    dot56 = 1 * 1.5 + 3 * 3.5 + 2 * 2.5
    assert cos[5, 6] == dot56 / ((1 ** 2 + 3 ** 2 + 2 ** 2) * (1.5 ** 2 + 3.5 ** 2 + 2.5 ** 2)) ** 0.5
    msd = oracle.similarity("msd", 8, *a, 1)
    assert msd[0, 1] == 1 and msd[3, 4] == .5 and msd[0, 3] == 0
    pe = oracle.similarity("pearson", 8, *a, 1)
    assert pe[0, 1] == 1 and pe[3, 4] == 0 and pe[2, 3] == 0 and pe[5, 6] == 1 and pe[2, 5] > 0
    mean6 = (1.5 + 3.5 + 2.5) / 3
    var6 = (1.5 - mean6) ** 2 + (3.5 - mean6) ** 2 + (2.5 - mean6) ** 2
    mean7 = (3 + 2 + 2.5) / 3
    var7 = (3 - mean7) ** 2 + (2 - mean7) ** 2 + (2.5 - mean7) ** 2
    num = sum([(1.5 - mean6) * (3 - mean7), (3.5 - mean6) * (2 - mean7), (2.5 - mean6) * (2.5 - mean7)])
    assert pe[6, 7] == num / (var6 * var7) ** 0.5
    for sim in (cos, msd, pe):
        assert np.array_equal(sim, sim.T) and np.all(np.diag(sim) == 1)
    cos4 = oracle.similarity("cosine", 8, *a, 4)
    for i in range(8):
        for j in range(i + 1, 8):
            if i != 1 and j != 2:
                assert cos4[i, j] == 0


def test_toy_shrinkage_variants(toy):
    a = (toy["y_ptr"], toy["x_idx"], toy["r"])
    s0 = oracle.similarity("pearson_baseline", 8, *a, 1, 3.0, toy["bx"], toy["by"], 0.0)
    assert np.array_equal(s0, toy["pearson_baseline_shr0"])
    s7 = oracle.similarity("pearson_baseline", 8, *a, 2, 3.0, toy["bx"], toy["by"], 7.5)
    assert np.array_equal(s7, toy["pearson_baseline_shr7p5"])


def test_msd_zero_division():
    # freq == 0 with min_support <= 0: the reference raises ZeroDivisionError (cdivision off)
    ptr = np.array([0, 1, 2]); idx = np.array([0, 1]); r = np.array([3.0, 4.0])
    with pytest.raises(ZeroDivisionError):
        oracle.similarity("msd", 2, ptr, idx, r, 0)


@pytest.mark.parametrize("orient", ("item", "user"))
def test_u1_similarities_sha(u1, u1_golden, u1_arrays, orient):
    trainset, _ = u1
    ub = orient == "user"
    n_x = trainset.n_users if ub else trainset.n_items
    ptr, idx, val = trainset.item_csr() if ub else trainset.user_csr()
    bu, bi = u1_arrays["als_bu"], u1_arrays["als_bi"]
    bx, by = (bu, bi) if ub else (bi, bu)
    for kind in KINDS:
        sim = oracle.similarity(kind, n_x, ptr, idx, val, 1, float(trainset.global_mean), bx, by, 100.0)
        g = u1_golden["sims"]["%s_%s_ms1" % (kind, orient)]
        assert sha(sim) == g["sha256"], kind
        pi, pj = u1_arrays["pairs_%s_i" % orient], u1_arrays["pairs_%s_j" % orient]
        assert np.array_equal(sim[pi, pj], u1_arrays["sim_%s_%s_ms1" % (kind, orient)])
    sim3 = oracle.similarity("cosine", n_x, ptr, idx, val, 3)
    assert sha(sim3) == u1_golden["sims"]["cosine_%s_ms3" % orient]["sha256"]


def test_u1_sampled_pairs_entry(u1, u1_arrays):
    """orc_similarity_pairs (used for sampled checks at shapes where the dense oracle cannot run)."""
    trainset, _ = u1
    # x rows (items) listing (y = user, r) with y ascending == the order the reference's outer loop meets them
    u, i, r = trainset.coo()
    o = np.lexsort((u, i))
    ptr = np.concatenate(([0], np.cumsum(np.bincount(i, minlength=trainset.n_items)))).astype(np.int64)
    idx, val = u[o], r[o]
    pi, pj = u1_arrays["pairs_item_i"], u1_arrays["pairs_item_j"]
    bi, bu = u1_arrays["als_bi"], u1_arrays["als_bu"]
    for kind in KINDS:
        got = oracle.similarity_pairs(kind, pi, pj, ptr, idx, val, 1, float(trainset.global_mean), bi, bu, 100.0)
        assert np.array_equal(got, u1_arrays["sim_%s_item_ms1" % kind]), kind


def test_u1_global_mean(u1, u1_golden):
    assert repr(float(u1[0].global_mean)) == u1_golden["global_mean"]
    assert (u1[0].n_users, u1[0].n_items, u1[0].n_ratings) == (904, 1187, 8000)


def test_u1_baselines(u1, u1_golden, u1_arrays):
    ts, _ = u1
    bu, bi = oracle.baseline_als(ts.n_users, ts.n_items, *ts.user_csr(), *ts.item_csr(), float(ts.global_mean))
    assert sha(bu) == u1_golden["baseline_als"]["bu_sha256"] and sha(bi) == u1_golden["baseline_als"]["bi_sha256"]
    u, i, r = ts.coo()
    bu, bi = oracle.baseline_sgd(ts.n_users, ts.n_items, u, i, r, float(ts.global_mean))
    assert sha(bu) == u1_golden["baseline_sgd"]["bu_sha256"] and sha(bi) == u1_golden["baseline_sgd"]["bi_sha256"]


def test_u1_nmf_bit_exact(u1, u1_golden, u1_arrays):
    ts, testset = u1
    u, i, r = ts.coo()
    n_ur, n_ir = np.diff(ts.user_csr()[0]), np.diff(ts.item_csr()[0])
    rng = np.random.RandomState(0)
    pu0 = rng.uniform(0, 1, size=(ts.n_users, 15)); qi0 = rng.uniform(0, 1, size=(ts.n_items, 15))
    pu, qi, bu, bi = oracle.nmf_sgd(ts.n_users, ts.n_items, u, i, r, n_ur, n_ir, pu0, qi0, 50, False,
                                    float(ts.global_mean), .06, .06, .02, .02, .005, .005)
    g = u1_golden["algos"]["NMF_rs0"]
    assert sha(pu) == g["pu_sha256"] and sha(qi) == g["qi_sha256"]
    assert np.array_equal(pu, u1_arrays["NMF_rs0_pu"])
    iu, ii = inner_pairs(ts, testset)
    est, imp = oracle.mf_estimate(iu, ii, False, float(ts.global_mean), pu, qi, bu, bi)
    est = np.where(imp > 0, float(ts.global_mean), est)
    rm, ma = rmse_mae(est, testset, ts, float(ts.global_mean))
    assert abs(rm - float(g["rmse"])) < 1e-12 and abs(ma - float(g["mae"])) < 1e-12
    # biased variant: sequential bias recursion
    rng = np.random.RandomState(0)
    pu0 = rng.uniform(0, 1, size=(ts.n_users, 15)); qi0 = rng.uniform(0, 1, size=(ts.n_items, 15))
    _, _, bub, _ = oracle.nmf_sgd(ts.n_users, ts.n_items, u, i, r, n_ur, n_ir, pu0, qi0, 50, True,
                                  float(ts.global_mean), .06, .06, .02, .02, .005, .005)
    assert np.array_equal(bub, u1_arrays["NMF_rs0_biased_bu"])
    rng = np.random.RandomState(3)
    pu0 = rng.uniform(0, 1, size=(ts.n_users, 7)); qi0 = rng.uniform(0, 1, size=(ts.n_items, 7))
    pu, qi, _, _ = oracle.nmf_sgd(ts.n_users, ts.n_items, u, i, r, n_ur, n_ir, pu0, qi0, 3, False, 0.0, .1, .02,
                                  .02, .02, .005, .005)
    g = u1_golden["algos"]["NMF_rs3_f7_e3"]
    assert sha(pu) == g["pu_sha256"] and sha(qi) == g["qi_sha256"]


def test_u1_svd_bit_exact(u1, u1_golden, u1_arrays):
    ts, testset = u1
    u, i, r = ts.coo()
    rng = np.random.RandomState(0)
    pu0 = rng.normal(0, .1, (ts.n_users, 8)); qi0 = rng.normal(0, .1, (ts.n_items, 8))
    pu, qi, bu, bi = oracle.svd_sgd(u, i, r, pu0, qi0, 1, True, float(ts.global_mean), *([.005] * 4), *([.02] * 4))
    for name, a in (("pu", pu), ("qi", qi), ("bu", bu), ("bi", bi)):
        assert np.array_equal(a, u1_arrays["SVD_rs0_f8_e1_" + name]), name
    rng = np.random.RandomState(0)
    pu0 = rng.normal(0, .1, (ts.n_users, 100)); qi0 = rng.normal(0, .1, (ts.n_items, 100))
    pu, qi, bu, bi = oracle.svd_sgd(u, i, r, pu0, qi0, 20, True, float(ts.global_mean), *([.005] * 4), *([.02] * 4))
    assert sha(pu) == u1_golden["algos"]["SVD_rs0"]["pu_sha256"]
    iu, ii = inner_pairs(ts, testset)
    est, _ = oracle.mf_estimate(iu, ii, True, float(ts.global_mean), pu, qi, bu, bi)
    rm, ma = rmse_mae(est, testset, ts, float(ts.global_mean))
    assert abs(rm - float(u1_golden["algos"]["SVD_rs0"]["rmse"])) < 1e-9


def test_u1_svdpp_bit_exact(u1, u1_arrays):
    ts, _ = u1
    u, i, r = ts.coo()
    ptr, idx, _ = ts.user_csr()
    rng = np.random.RandomState(0)
    pu0 = rng.normal(0, .1, (ts.n_users, 4)); qi0 = rng.normal(0, .1, (ts.n_items, 4))
    yj0 = rng.normal(0, .1, (ts.n_items, 4))
    pu, qi, yj, bu, bi = oracle.svdpp_sgd(u, i, r, ptr, idx, pu0, qi0, yj0, 1, float(ts.global_mean), *([.007] * 5),
                                          *([.02] * 5))
    assert np.array_equal(pu, u1_arrays["SVDpp_rs0_f4_e1_pu"])
    assert np.array_equal(qi, u1_arrays["SVDpp_rs0_f4_e1_qi"])
    assert np.array_equal(yj, u1_arrays["SVDpp_rs0_f4_e1_yj"])


@pytest.mark.parametrize("orient", ("item", "user"))
def test_u1_knn_estimates_bit_exact(u1, u1_golden, u1_arrays, orient):
    ts, testset = u1
    ub = orient == "user"
    iu, ii = inner_pairs(ts, testset)
    x, y = (iu, ii) if ub else (ii, iu)
    n_x = ts.n_users if ub else ts.n_items
    ptr, idx, val = ts.item_csr() if ub else ts.user_csr()
    mu = float(ts.global_mean)
    sim = oracle.similarity("cosine", n_x, ptr, idx, val, 1)
    est, ak, imp = oracle.knn_estimate(x, y, sim, ptr, idx, val, 40, 1)
    tag = "KNNBasic_cosine_%s_ms1" % orient
    lo, hi = ts.rating_scale
    est = np.clip(np.where(imp > 0, mu, est), lo, hi)  # Prediction.est is clipped (algo_base.py:166-169)
    assert np.array_equal(est, u1_arrays[tag + "_est"])
    assert np.array_equal(np.where(imp > 0, -1, ak), u1_arrays[tag + "_actual_k"])
    assert int(imp.sum()) == u1_golden["algos"][tag]["n_impossible"]
    bu, bi = u1_arrays["als_bu"], u1_arrays["als_bi"]
    bx, by = (bu, bi) if ub else (bi, bu)
    sim = oracle.similarity("pearson_baseline", n_x, ptr, idx, val, 1, mu, bx, by, 100.0)
    est, ak, imp = oracle.knn_estimate(x, y, sim, ptr, idx, val, 40, 1, 1 if ub else 2, mu, bx, by)
    tag = "KNNBaseline_pb_" + orient
    assert np.array_equal(np.clip(est, lo, hi), u1_arrays[tag + "_est"])
    assert np.array_equal(ak, u1_arrays[tag + "_actual_k"])
    sim = oracle.similarity("msd", n_x, ptr, idx, val, 1)
    est, ak, imp = oracle.knn_estimate(x, y, sim, ptr, idx, val, 10, 3, 1 if ub else 2, mu, bx, by)
    tag = "KNNBaseline_msd_k10_mk3_" + orient
    assert np.array_equal(np.clip(est, lo, hi), u1_arrays[tag + "_est"])
    assert np.array_equal(ak, u1_arrays[tag + "_actual_k"])


@pytest.mark.parametrize("orient", ("item", "user"))
def test_u1_knn_means_zscore_bit_exact(u1, u1_arrays, orient):
    """KNNWithMeans / KNNWithZScore (knns.py:126-208, :312-403): the "next" rows of SURVEY.md section 8f."""
    from surprise_b200.prediction_algorithms.knns import _row_stats
    ts, testset = u1
    ub = orient == "user"
    iu, ii = inner_pairs(ts, testset)
    x, y = (iu, ii) if ub else (ii, iu)
    n_x = ts.n_users if ub else ts.n_items
    ptr, idx, val = ts.item_csr() if ub else ts.user_csr()
    mu = float(ts.global_mean)
    lo, hi = ts.rating_scale
    means, sigmas = _row_stats(ts, ub, True)
    sigmas = np.where(sigmas == 0.0, np.std(ts.user_csr()[2]), sigmas)
    sim = oracle.similarity("msd", n_x, ptr, idx, val, 1)
    est, ak, imp = oracle.knn_estimate(x, y, sim, ptr, idx, val, 40, 1, 3, mu, means, None)
    tag = "KNNWithMeans_msd_" + orient
    assert np.array_equal(np.clip(np.where(imp > 0, mu, est), lo, hi), u1_arrays[tag + "_est"])
    assert np.array_equal(np.where(imp > 0, -1, ak), u1_arrays[tag + "_actual_k"])
    sim = oracle.similarity("pearson", n_x, ptr, idx, val, 1)
    est, ak, imp = oracle.knn_estimate(x, y, sim, ptr, idx, val, 20, 2, 4, mu, means, sigmas)
    tag = "KNNWithZScore_pearson_k20_mk2_" + orient
    assert np.array_equal(np.clip(np.where(imp > 0, mu, est), lo, hi), u1_arrays[tag + "_est"])
    assert np.array_equal(np.where(imp > 0, -1, ak), u1_arrays[tag + "_actual_k"])


def test_float_ratings_similarities(floats):
    """Jester-style float ratings: the oracle must reproduce the reference bit for bit here too."""
    import surprise_b200 as sb
    from surprise_b200.dataset import Dataset
    ds = Dataset.load_from_arrays(floats["uid"], floats["iid"], floats["rating"], sb.Reader(rating_scale=(-10, 10)))
    ts = ds.build_full_trainset()
    assert repr(float(ts.global_mean)) == repr(float(floats["global_mean"]))
    bu, bi = oracle.baseline_als(ts.n_users, ts.n_items, *ts.user_csr(), *ts.item_csr(), float(ts.global_mean))
    assert np.array_equal(bu, floats["als_bu"]) and np.array_equal(bi, floats["als_bi"])
    for ub, o in ((False, "item"), (True, "user")):
        n_x = ts.n_users if ub else ts.n_items
        ptr, idx, val = ts.item_csr() if ub else ts.user_csr()
        bx, by = (bu, bi) if ub else (bi, bu)
        for kind in KINDS:
            sim = oracle.similarity(kind, n_x, ptr, idx, val, 2, float(ts.global_mean), bx, by, 100.0)
            assert np.array_equal(sim, floats["sim_%s_%s" % (kind, o)], equal_nan=True), (kind, o)


def _nan0(a):
    return np.where(np.isnan(a), 0.0, a)


def test_u1_slope_one_bit_exact(u1, u1_golden, u1_arrays):
    """SlopeOne.fit / estimate (slope_one.pyx:44-97) on the fixture: freq, dev, user means, predictions."""
    import hashlib
    ts, testset = u1
    ptr, idx, val = ts.user_csr()
    freq, dev = oracle.slope_one_fit(ts.n_items, ptr, idx, val)
    g = u1_golden["slope_one"]
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(freq) == g["freq_sha256"] and sha(_nan0(dev)) == g["dev_nan0_sha256"]
    assert int(np.isnan(dev).sum()) == g["n_nan"]
    mean = np.array([np.mean(val[ptr[u]:ptr[u + 1]].tolist()) for u in range(ts.n_users)])
    assert np.array_equal(mean, u1_arrays["SlopeOne_user_mean"])
    iu, ii = inner_pairs(ts, testset)
    est, imp = oracle.slope_one_estimate(iu, ii, freq, dev, ptr, idx, mean)
    lo, hi = ts.rating_scale
    assert np.array_equal(np.clip(np.where(imp > 0, ts.global_mean, est), lo, hi), u1_arrays["SlopeOne_est"])
    assert int(imp.sum()) == u1_golden["algos"]["SlopeOne"]["n_impossible"]


def test_slope_one_truncates_ratings(floats):
    """`cdef int r_ui, r_uj` (slope_one.pyx:52): half-star and float ratings are truncated before differencing."""
    import surprise_b200 as sb
    from surprise_b200.dataset import Dataset
    for uid, iid, rat, scale, pre in ((floats["uid"], floats["iid"], floats["rating"], (-10, 10), ""),
                                      (floats["half_uid"], floats["half_iid"], floats["half_rating"], (0.5, 5), "half_")):
        ts = Dataset.load_from_arrays(uid, iid, rat, sb.Reader(rating_scale=scale)).build_full_trainset()
        ptr, idx, val = ts.user_csr()
        freq, dev = oracle.slope_one_fit(ts.n_items, ptr, idx, val)
        assert np.array_equal(freq, floats[pre + "slope_freq"])
        assert np.array_equal(_nan0(dev), _nan0(floats[pre + "slope_dev"]))
        assert np.array_equal(np.isnan(dev), np.isnan(floats[pre + "slope_dev"]))
