"""Generate the golden vectors under tests/golden/ from the COMPILED, UNMODIFIED reference.

Run in the build container (needs oracle/_ref, i.e. /root/reference present at build time):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The outputs are small and committed; the GPU box has no /root/reference, so `-m gpu` tests and
smoke() compare against these files (and against the C oracle, which is itself pinned to them).

What is pinned (reference file:line of the producer in parentheses):
  * toy_sims.npz      -- tests/test_similarities.py:13-20 `yr_global` (n_x=8) through
                         similarities.{cosine,msd,pearson,pearson_baseline} (similarities.pyx:28-361),
                         min_support 1 and 4, full matrices.
  * u1_golden.json    -- tests/u1_ml100k_train -> u1_ml100k_test (904 x 1187, 8000/2000): sha256 of
                         every similarity matrix (both orientations), baselines, NMF factors; RMSE/MAE
                         of KNNBasic / KNNBaseline / SVD / SVDpp / NMF via AlgoBase.fit/test.
  * u1_arrays.npz     -- sampled similarity entries, ALS/SGD baselines, NMF pu/qi (bit-exact target),
                         per-pair estimates + actual_k of the KNN algorithms, SVD/SVDpp/NMF estimates.
  * float_sims.npz    -- a 40 x 60 synthetic Jester-style FLOAT rating set: all four similarity
                         matrices in full + KNNBaseline estimates; SlopeOne freq / dev / estimates on it and on a
                         30 x 25 half-star set (slope_one.pyx:44-97, incl. the C-int truncation of ratings).
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

ref = oracle.import_reference()
from surprise import Dataset, Reader, KNNBasic, KNNBaseline, KNNWithMeans, KNNWithZScore, SVD, SVDpp, NMF, BaselineOnly, SlopeOne, accuracy  # noqa: E402
from surprise import similarities as rsims  # noqa: E402
from surprise.model_selection import PredefinedKFold  # noqa: E402

REF_TESTS = "/root/reference/tests"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- fixture data files (ratings data, not code) -------------------------------------------------
for name in ("u1_ml100k_train", "u1_ml100k_test", "custom_dataset", "custom_train", "custom_test"):
    shutil.copyfile(os.path.join(REF_TESTS, name), os.path.join(HERE, name))

# ---- toy similarities -----------------------------------------------------------------------------
yr_toy = {
    0: [(0, 3), (1, 3), (2, 3), (5, 1), (6, 1.5), (7, 3)],
    1: [(0, 4), (1, 4), (2, 4)],
    2: [(2, 5), (3, 2), (4, 3)],
    3: [(1, 1), (2, 4), (3, 2), (4, 3), (5, 3), (6, 3.5), (7, 2)],
    4: [(1, 5), (2, 1), (5, 2), (6, 2.5), (7, 2.5)],
}
rs = np.random.RandomState(7)
toy_bx = rs.normal(0, 1, 8)
toy_by = rs.normal(0, 1, 5)
toy = {"bx": toy_bx, "by": toy_by, "global_mean": np.float64(3.0)}
ptr, xs, rr = oracle.flatten_yr(yr_toy)
toy.update(y_ptr=ptr, x_idx=xs, r=rr)
for ms in (1, 4):
    toy["cosine_%d" % ms] = rsims.cosine(8, yr_toy, ms)
    toy["msd_%d" % ms] = rsims.msd(8, yr_toy, ms)
    toy["pearson_%d" % ms] = rsims.pearson(8, yr_toy, ms)
    toy["pearson_baseline_%d" % ms] = rsims.pearson_baseline(8, yr_toy, ms, 3, toy_bx, toy_by)
toy["pearson_baseline_shr0"] = rsims.pearson_baseline(8, yr_toy, 1, 3, toy_bx, toy_by, 0)
toy["pearson_baseline_shr7p5"] = rsims.pearson_baseline(8, yr_toy, 2, 3, toy_bx, toy_by, 7.5)
np.savez_compressed(os.path.join(HERE, "toy_sims.npz"), **toy)

# ---- u1 fixture -------------------------------------------------------------------------------------
data = Dataset.load_from_folds([(os.path.join(REF_TESTS, "u1_ml100k_train"),
                                 os.path.join(REF_TESTS, "u1_ml100k_test"))], Reader("ml-100k"))
trainset, testset = next(PredefinedKFold().split(data))
G = {"n_users": trainset.n_users, "n_items": trainset.n_items, "n_ratings": trainset.n_ratings,
     "n_test": len(testset), "global_mean": repr(float(trainset.global_mean)), "sims": {}, "algos": {}}
A = {}
rs = np.random.RandomState(11)


def run(algo, tag, keep_est=True):
    algo.fit(trainset)
    preds = algo.test(testset)
    G["algos"][tag] = {"rmse": repr(float(accuracy.rmse(preds, verbose=False))),
                       "mae": repr(float(accuracy.mae(preds, verbose=False))),
                       "n_impossible": int(sum(p.details["was_impossible"] for p in preds))}
    if keep_est:
        A[tag + "_est"] = np.array([p.est for p in preds])
        A[tag + "_actual_k"] = np.array([p.details.get("actual_k", -1) for p in preds], dtype=np.int32)
    return algo, preds


for ub in (False, True):
    n_x = trainset.n_users if ub else trainset.n_items
    pi = rs.randint(0, n_x, 4000).astype(np.int32)
    pj = rs.randint(0, n_x, 4000).astype(np.int32)
    o = "user" if ub else "item"
    A["pairs_%s_i" % o], A["pairs_%s_j" % o] = pi, pj
    for name in ("cosine", "msd", "pearson", "pearson_baseline"):
        for ms in ((1, 3) if name == "cosine" else (1,)):
            tag = "%s_%s_ms%d" % (name, o, ms)
            algo, _ = run(KNNBasic(sim_options={"name": name, "user_based": ub, "min_support": ms}),
                          "KNNBasic_" + tag)
            s = algo.sim
            G["sims"][tag] = {"sha256": sha(s), "sum": repr(float(s.sum())), "nnz": int(np.count_nonzero(s))}
            A["sim_" + tag] = s[pi, pj]

bo, _ = run(BaselineOnly(), "BaselineOnly_als", keep_est=False)
G["baseline_als"] = {"bu_sha256": sha(bo.bu), "bi_sha256": sha(bo.bi)}
A["als_bu"], A["als_bi"] = bo.bu, bo.bi
bo, _ = run(BaselineOnly(bsl_options={"method": "sgd"}), "BaselineOnly_sgd", keep_est=False)
G["baseline_sgd"] = {"bu_sha256": sha(bo.bu), "bi_sha256": sha(bo.bi)}
A["sgd_bu"], A["sgd_bi"] = bo.bu, bo.bi

for ub in (False, True):
    o = "user" if ub else "item"
    run(KNNBaseline(sim_options={"name": "pearson_baseline", "user_based": ub}), "KNNBaseline_pb_" + o)
    run(KNNBaseline(k=10, min_k=3, sim_options={"name": "msd", "user_based": ub}), "KNNBaseline_msd_k10_mk3_" + o)
run(KNNBasic(k=5, min_k=2, sim_options={"name": "msd", "user_based": True}), "KNNBasic_msd_k5_mk2_user")
for ub in (False, True):
    o = "user" if ub else "item"
    run(KNNWithMeans(sim_options={"name": "msd", "user_based": ub}), "KNNWithMeans_msd_" + o)
    run(KNNWithZScore(k=20, min_k=2, sim_options={"name": "pearson", "user_based": ub}), "KNNWithZScore_pearson_k20_mk2_" + o)

svd, _ = run(SVD(random_state=0), "SVD_rs0")
G["algos"]["SVD_rs0"]["pu_sha256"] = sha(svd.pu)
run(SVD(random_state=0, biased=False), "SVD_rs0_unbiased")
run(SVD(random_state=0, n_factors=20, n_epochs=5), "SVD_rs0_f20_e5")
svd1, _ = run(SVD(random_state=0, n_factors=8, n_epochs=1), "SVD_rs0_f8_e1", keep_est=False)
A["SVD_rs0_f8_e1_pu"], A["SVD_rs0_f8_e1_qi"] = svd1.pu, svd1.qi
A["SVD_rs0_f8_e1_bu"], A["SVD_rs0_f8_e1_bi"] = svd1.bu, svd1.bi
run(SVDpp(random_state=0), "SVDpp_rs0")
pp1, _ = run(SVDpp(random_state=0, n_factors=4, n_epochs=1), "SVDpp_rs0_f4_e1", keep_est=False)
A["SVDpp_rs0_f4_e1_pu"], A["SVDpp_rs0_f4_e1_qi"], A["SVDpp_rs0_f4_e1_yj"] = pp1.pu, pp1.qi, pp1.yj
nmf, _ = run(NMF(random_state=0), "NMF_rs0")
G["algos"]["NMF_rs0"].update(pu_sha256=sha(nmf.pu), qi_sha256=sha(nmf.qi))
A["NMF_rs0_pu"], A["NMF_rs0_qi"] = nmf.pu, nmf.qi
nmfb, _ = run(NMF(random_state=0, biased=True), "NMF_rs0_biased")
A["NMF_rs0_biased_bu"] = nmfb.bu
nmf3, _ = run(NMF(random_state=3, n_factors=7, n_epochs=3, reg_pu=.1, reg_qi=.02), "NMF_rs3_f7_e3", keep_est=False)
G["algos"]["NMF_rs3_f7_e3"].update(pu_sha256=sha(nmf3.pu), qi_sha256=sha(nmf3.qi))

so, _ = run(SlopeOne(), "SlopeOne")
G["slope_one"] = {"freq_sha256": sha(so.freq), "dev_nan0_sha256": sha(np.where(np.isnan(so.dev), 0.0, so.dev)),
                  "n_nan": int(np.isnan(so.dev).sum())}
A["SlopeOne_user_mean"] = np.array(so.user_mean)

with open(os.path.join(HERE, "u1_golden.json"), "w") as fh:
    json.dump(G, fh, indent=1, sort_keys=True)
np.savez_compressed(os.path.join(HERE, "u1_arrays.npz"), **A)

# ---- float ratings (Jester style: scale (-10, 10), offset 11) -------------------------------------
rs = np.random.RandomState(5)
n_u, n_i = 40, 60
mask = rs.rand(n_u, n_i) < 0.35
uu, ii = np.nonzero(mask)
perm = rs.permutation(len(uu))
uu, ii = uu[perm], ii[perm]
rat = np.round(rs.uniform(-10, 10, len(uu)), 2)
import pandas as pd  # noqa: E402

df = pd.DataFrame({"u": uu.astype(np.int32), "i": ii.astype(np.int32), "r": rat})
fdata = Dataset.load_from_df(df, Reader(rating_scale=(-10, 10)))
ftrain = fdata.build_full_trainset()
F = {"uid": uu.astype(np.int32), "iid": ii.astype(np.int32), "rating": rat}
for ub in (False, True):
    o = "user" if ub else "item"
    for name in ("cosine", "msd", "pearson", "pearson_baseline"):
        algo = KNNBasic(sim_options={"name": name, "user_based": ub, "min_support": 2}).fit(ftrain)
        F["sim_%s_%s" % (name, o)] = algo.sim
    kb = KNNBaseline(k=7, sim_options={"name": "pearson_baseline", "user_based": ub, "shrinkage": 10}).fit(ftrain)
    F["knnbaseline_sim_" + o] = kb.sim
    F["als_bu"], F["als_bi"] = kb.bu, kb.bi
    tp = [(int(a), int(b), 0.0) for a in range(0, n_u, 3) for b in range(0, n_i, 7)]
    preds = kb.test(tp)
    F["knnbaseline_est_" + o] = np.array([p.est for p in preds])
    F["knnbaseline_ak_" + o] = np.array([p.details.get("actual_k", -1) for p in preds], dtype=np.int32)
# SlopeOne truncates ratings to C ints (slope_one.pyx:52): pin that on the float set and on half-stars
so = SlopeOne().fit(ftrain)
F["slope_freq"], F["slope_dev"], F["slope_user_mean"] = so.freq, so.dev, np.array(so.user_mean)
preds = so.test(tp)
F["slope_est"] = np.array([p.est for p in preds])
rs = np.random.RandomState(9)
hu, hi = np.nonzero(rs.rand(30, 25) < 0.4)
hp = rs.permutation(len(hu))
hu, hi = hu[hp].astype(np.int32), hi[hp].astype(np.int32)
hr = rs.randint(1, 11, len(hu)) / 2.0
htrain = Dataset.load_from_df(pd.DataFrame({"u": hu, "i": hi, "r": hr}), Reader(rating_scale=(0.5, 5))).build_full_trainset()
so = SlopeOne().fit(htrain)
hp_pairs = [(int(a), int(b), 0.0) for a in range(30) for b in range(0, 25, 3)]
F["half_uid"], F["half_iid"], F["half_rating"] = hu, hi, hr
F["half_slope_freq"], F["half_slope_dev"] = so.freq, so.dev
F["half_slope_est"] = np.array([p.est for p in so.test(hp_pairs)])
F["half_slope_impossible"] = np.array([p.details["was_impossible"] for p in so.test(hp_pairs)])
F["global_mean"] = np.float64(ftrain.global_mean)
np.savez_compressed(os.path.join(HERE, "float_sims.npz"), **F)
print("goldens written to", HERE)
print(json.dumps(G["algos"], indent=1)[:1500])
