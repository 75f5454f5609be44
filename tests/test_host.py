"""CPU-side tests: host data model, the C-ABI library loads and exports every declared symbol, and the
product path fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import surprise_b200 as sb
from surprise_b200 import _native as nat
from surprise_b200 import similarities as sims
from surprise_b200 import synth
from conftest import GOLDEN, ROOT


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "surprise_b200.h")).read()
    declared = set(re.findall(r"\b(sb2_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("sb2_sgd_params")
    lib = C.CDLL(nat.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    assert nat.lib().sb2_version() >= 100


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nat.NativeError):
        sims.cosine(2, {0: [(0, 1.0), (1, 2.0)]}, 1)
    # the C-ABI itself also refuses (host-buffer entry point)
    ptr = np.array([0, 2], dtype=np.int64); idx = np.array([0, 1], dtype=np.int32); r = np.array([1.0, 2.0])
    out = np.zeros((2, 2))
    rc = nat.lib().sb2_sim_build(0, 2, 1, nat.hptr(ptr), nat.hptr(idx), nat.hptr(r), 2, 1, 1, 0.0, None, None, 100.0,
                                 0, 2, nat.hptr(out))
    assert rc == nat.ERR_CUDA and b"no CPU fallback" in nat.lib().sb2_last_error()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "surprise_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src and "liboracle" not in src, f


def test_reader_and_dataset_inner_ids():
    reader = sb.Reader(line_format="user item rating", sep=" ", skip_lines=3, rating_scale=(1, 5))
    data = sb.Dataset.load_from_file(os.path.join(GOLDEN, "custom_dataset"), reader)
    ts = data.build_full_trainset()
    assert ts.n_ratings == len(data.raw_ratings) and ts.rating_scale == (1, 5) and ts.offset == 0
    # first-appearance inner ids, ur / ir list order == file order
    seen_u, seen_i = [], []
    for (u, i, r, _) in data.raw_ratings:
        if u not in seen_u:
            seen_u.append(u)
        if i not in seen_i:
            seen_i.append(i)
    for k, raw in enumerate(seen_u):
        assert ts.to_inner_uid(raw) == k and ts.to_raw_uid(k) == raw
    for k, raw in enumerate(seen_i):
        assert ts.to_inner_iid(raw) == k and ts.to_raw_iid(k) == raw
    flat = [(ts.to_raw_uid(u), ts.to_raw_iid(i), r) for (u, i, r) in ts.all_ratings()]
    by_user = sorted(range(len(data.raw_ratings)), key=lambda k: seen_u.index(data.raw_ratings[k][0]))
    assert flat == [data.raw_ratings[k][:3] for k in by_user]
    assert ts.knows_user(0) and not ts.knows_user(ts.n_users) and not ts.knows_user("UKN__x")
    with pytest.raises(ValueError):
        ts.to_inner_uid("nobody")
    assert float(ts.global_mean) == np.mean([r for (_, _, r) in ts.all_ratings()])
    assert len(ts.build_testset()) == ts.n_ratings
    anti = ts.build_anti_testset()
    assert len(anti) == ts.n_users * ts.n_items - ts.n_ratings
    with pytest.raises(ValueError):
        sb.Reader(name="nope")
    r2 = sb.Reader(rating_scale=(-10, 10))
    assert r2.offset == 11 and r2.parse_line("a b 3.5")[2] == 14.5


def test_trainset_dict_and_array_views_agree(u1):
    ts, _ = u1
    ur, ir = ts.ur, ts.ir
    assert len(ur) == ts.n_users and len(ir) == ts.n_items
    ts2 = sb.Trainset(ur, ir, ts.n_users, ts.n_items, ts.n_ratings, ts.rating_scale, ts.offset,
                      ts._raw2inner_id_users, ts._raw2inner_id_items)
    for a, b in zip(ts.user_csr(), ts2.user_csr()):
        assert np.array_equal(a, b)
    for a, b in zip(ts.item_csr(), ts2.item_csr()):
        assert np.array_equal(a, b)
    assert list(ts.all_ratings()) == list(ts2.all_ratings())
    assert ts2.global_mean == ts.global_mean


def test_load_from_arrays_matches_file_loader(u1):
    ts, _ = u1
    rows = [l.split("\t") for l in open(os.path.join(GOLDEN, "u1_ml100k_train"))]
    uid = np.array([r[0] for r in rows]); iid = np.array([r[1] for r in rows]); rat = np.array([float(r[2]) for r in rows])
    ts2 = sb.Dataset.load_from_arrays(uid, iid, rat, sb.Reader("ml-100k")).build_full_trainset()
    for a, b in zip(ts.user_csr() + ts.item_csr(), ts2.user_csr() + ts2.item_csr()):
        assert np.array_equal(a, b)
    assert ts2.to_inner_uid(rows[0][0]) == 0


def test_rating_denominator():
    assert sims.rating_denominator([1, 2, 5]) == 1
    assert sims.rating_denominator([0.5, 4.5]) == 2
    assert sims.rating_denominator([1.25]) == 4
    assert sims.rating_denominator([7.82, 11.0]) == 100
    with pytest.raises(ValueError):
        sims.rating_denominator([np.pi])


def test_accuracy_known_answers():
    P = sb.Prediction
    preds = [P(None, None, 5, 5, None), P(None, None, 4, 4, None)]
    assert sb.accuracy.rmse(preds, verbose=False) == 0 and sb.accuracy.mae(preds, verbose=False) == 0
    preds = [P(None, None, 0, 2, None), P(None, None, 0, 2, None)]
    assert sb.accuracy.rmse(preds, verbose=False) == 2 and sb.accuracy.mae(preds, verbose=False) == 2
    with pytest.raises(ValueError):
        sb.accuracy.rmse([])


def test_constructor_defaults_match_reference():
    s = sb.SVD()
    assert (s.n_factors, s.n_epochs, s.biased, s.lr_pu, s.reg_qi) == (100, 20, True, .005, .02)
    s = sb.SVD(lr_all=.1, lr_pu=.3, reg_all=.5, reg_bi=.7)
    assert (s.lr_bu, s.lr_pu, s.reg_bu, s.reg_bi) == (.1, .3, .5, .7)
    p = sb.SVDpp()
    assert (p.n_factors, p.n_epochs, p.lr_yj, p.reg_yj) == (20, 20, .007, .02)
    n = sb.NMF()
    assert (n.n_factors, n.n_epochs, n.biased, n.reg_pu, n.lr_bu) == (15, 50, False, .06, .005)
    k = sb.KNNBasic()
    assert (k.k, k.min_k, k.sim_options["user_based"]) == (40, 1, True)
    with pytest.raises(ValueError):
        sb.NMF(init_low=-1)
    with pytest.raises(ValueError):
        sb.utils.get_rng("bad") if hasattr(sb, "utils") else __import__("surprise_b200.utils").utils.get_rng("bad")


def test_train_fit_shim():
    """reference tests/test_train2fit.py: old-style train() algorithms still work through fit()."""
    class Old(sb.AlgoBase):
        def __init__(self):
            sb.AlgoBase.__init__(self)
            self.cnt = -1

        def train(self, trainset):
            sb.AlgoBase.train(self, trainset)
            self.cnt += 1
            self.bu = "x"

        def estimate(self, u, i):
            return self.cnt

    reader = sb.Reader(line_format="user item rating", sep=" ", skip_lines=3, rating_scale=(1, 5))
    ts = sb.Dataset.load_from_file(os.path.join(GOLDEN, "custom_dataset"), reader).build_full_trainset()
    with pytest.warns(UserWarning):
        a = Old()
    with pytest.warns(UserWarning):
        a.fit(ts)
    assert a.cnt == 0 and a.predict("user0", "item0").est == 1  # clipped to the scale's lower bound
    with pytest.warns(UserWarning):
        a.train(ts)
    assert a.cnt == 1


def test_synth_is_deterministic_and_compact():
    a = synth.ratings(300, 200, 5000, seed=1)
    b = synth.ratings(300, 200, 5000, seed=1)
    for x, y in zip(a["train"] + a["test"], b["train"] + b["test"]):
        assert np.array_equal(x, y)
    u, i, r = a["train"]
    assert u.max() + 1 == a["n_users"] and len(np.unique(u)) == a["n_users"]
    assert len(np.unique(u.astype(np.int64) * 10**6 + i)) == len(u)
    assert set(np.unique(r)) <= {1., 2., 3., 4., 5.}
    # first-appearance order
    assert np.array_equal(np.unique(u, return_index=True)[1].argsort(), np.arange(a["n_users"]))


def test_trainset_ingest_fast_paths_match_reference_order():
    """Array ingest: the O(N) counting-sort grouping of Trainset.from_coo and the O(N) table for
    integer raw ids must give what the reference's dict construction gives: inner ids in first-appearance order
    (dataset.py:219-234), ur[u] / ir[i] in file order (trainset.py)."""
    import surprise_b200 as sb
    from surprise_b200 import trainset
    rng = np.random.RandomState(3)
    n = 50_000
    raw_u = rng.randint(-40, 3000, n) * 7          # gaps and negative ids
    raw_i = rng.randint(10, 900, n)
    r = rng.randint(1, 6, n).astype(np.float64)
    ts = sb.Dataset.load_from_arrays(raw_u, raw_i, r, sb.Reader(rating_scale=(1, 5))).build_full_trainset()
    # reference semantics restated with dicts
    u2i, i2i, ur, ir = {}, {}, {}, {}
    for a, b, c in zip(raw_u.tolist(), raw_i.tolist(), r.tolist()):
        uu = u2i.setdefault(a, len(u2i)); ii = i2i.setdefault(b, len(i2i))
        ur.setdefault(uu, []).append((ii, c)); ir.setdefault(ii, []).append((uu, c))
    assert ts.n_users == len(u2i) and ts.n_items == len(i2i)
    for k in (0, 1, 5, len(u2i) - 1):
        assert ts.ur[k] == ur[k]
    for k in (0, 3, len(i2i) - 1):
        assert ts.ir[k] == ir[k]
    assert ts.to_inner_uid(raw_u[0]) == 0 and ts.to_raw_uid(0) == raw_u[0]
    # the grouping kernel against numpy's stable argsort
    u, i, rr = ts.coo()
    perm = rng.permutation(len(rr))
    u, i, rr = u[perm], i[perm], rr[perm]
    o = np.argsort(i, kind="stable")
    ptr, other, val = trainset._group_stable(i, u, rr, ts.n_items, np.bincount(i, minlength=ts.n_items))
    assert np.array_equal(other, u[o]) and np.array_equal(val, rr[o]) and ptr[-1] == len(rr)
