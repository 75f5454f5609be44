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
from surprise_b200.trainset import Trainset
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


def test_row_stats_match_per_row_numpy():
    """knns._row_stats stacks rows of equal length; every mean / sigma must carry the bits of np.mean / np.std applied
    to that row alone (what knns.py:168-170 / :362-366 do), including rows longer than numpy's pairwise block (128)."""
    from surprise_b200.prediction_algorithms.knns import _row_stats
    rng = np.random.RandomState(3)
    lens = np.concatenate((rng.randint(1, 12, 300), rng.randint(100, 400, 40), [1, 7, 8, 9, 127, 128, 129, 1000]))
    n_x = len(lens)
    x = np.repeat(np.arange(n_x, dtype=np.int32), lens)
    n = len(x)
    y = np.arange(n, dtype=np.int32) % 50
    # every item id must be used: append one rating per missing id is not needed (50 ids all hit)
    r = rng.randint(1, 11, n) * 0.5
    ts = Trainset.from_coo(x, y, r, n_x, 50)
    means, sigmas = _row_stats(ts, True, True)
    ptr, _, val = ts.user_csr()
    for k in range(n_x):
        row = val[ptr[k]:ptr[k + 1]]
        assert means[k] == np.mean(row) and sigmas[k] == np.std(row), k


def test_shuffle_split_matches_reference_semantics():
    """ShuffleSplit / train_test_split (reference split.py:506-536): trainset = head of the permutation, testset the
    entries that follow; shuffle=False keeps file order (test = tail); sizes validated like the reference.  With
    oracle/_ref present the split is also compared with the reference's own, seed for seed."""
    import surprise_b200 as sb
    from surprise_b200.model_selection import KFold, ShuffleSplit, train_test_split
    n = 57
    uids = np.arange(n) % 11
    iids = (np.arange(n) * 7) % 13
    rs = (np.arange(n) % 5 + 1).astype(np.float64)
    data = sb.Dataset.load_from_arrays(uids, iids, rs, sb.Reader(rating_scale=(1, 5)))
    tr, te = train_test_split(data, test_size=.25, random_state=7)
    perm = np.random.RandomState(7).permutation(n)
    n_test = int(np.ceil(.25 * n)); n_train = n - n_test
    assert tr.n_ratings == n_train and len(te) == n_test
    assert [(u, i) for (u, i, _) in te] == [(int(uids[k]), int(iids[k])) for k in perm[n_train:n_train + n_test]]
    tr, te = train_test_split(data, test_size=10, shuffle=False)
    assert [(u, i) for (u, i, _) in te] == [(int(uids[k]), int(iids[k])) for k in range(n - 10, n)]
    tr, te = train_test_split(data, test_size=5, train_size=20, random_state=1)
    assert tr.n_ratings == 20 and len(te) == 5
    for bad in (dict(test_size=n), dict(train_size=n), dict(test_size=30, train_size=30), dict(test_size=-1)):
        with pytest.raises(ValueError):
            train_test_split(data, **bad)
    folds = list(KFold(n_splits=4, random_state=3).split(data))
    assert sum(len(te) for _, te in folds) == n and all(tr.n_ratings + len(te) == n for tr, te in folds)
    assert len(list(ShuffleSplit(n_splits=3, test_size=.2, random_state=0).split(data))) == 3
    # list-backed dataset: same rows as the array-backed one
    rows = [(int(u), int(i), float(r), None) for u, i, r in zip(uids, iids, rs)]
    data2 = sb.Dataset.load_from_arrays(uids, iids, rs, sb.Reader(rating_scale=(1, 5)))
    data2._arrays, data2.raw_ratings = None, rows
    tr2, te2 = train_test_split(data2, test_size=.25, random_state=7)
    tr1, te1 = train_test_split(data, test_size=.25, random_state=7)
    assert te1 == te2 and np.array_equal(tr1.coo()[2], tr2.coo()[2])
    try:
        import oracle
        ref = oracle.import_reference()
    except ImportError:
        return
    import pandas as pd
    df = pd.DataFrame({"u": uids, "i": iids, "r": rs})
    rdata = ref.Dataset.load_from_df(df, ref.Reader(rating_scale=(1, 5)))
    from surprise.model_selection import train_test_split as ref_split
    for kw in (dict(test_size=.25, random_state=7), dict(test_size=10, shuffle=False)):
        rtr, rte = ref_split(rdata, **kw)
        otr, ote = train_test_split(data, **kw)
        assert [(int(u), int(i), float(r)) for u, i, r in rte] == [(int(u), int(i), float(r)) for u, i, r in ote]
        assert rtr.n_ratings == otr.n_ratings and rtr.n_users == otr.n_users


def test_vectorised_raw_to_inner_lookup():
    ts = Trainset.from_coo(np.array([0, 1, 2, 0], dtype=np.int32), np.array([0, 0, 1, 2], dtype=np.int32),
                           np.array([1., 2, 3, 4]), 3, 3, raw_uids=[10, 5, 77], raw_iids=["a", "b", "c"])
    assert ts.to_inner_uids([5, 77, 10, 3]).tolist() == [1, 2, 0, -1]
    assert ts.to_inner_iids(["c", "zz", "a"]).tolist() == [2, -1, 0]
    assert ts.to_inner_uid(77) == 2
    plain = Trainset.from_coo(np.array([0, 1, 2, 0], dtype=np.int32), np.array([0, 0, 1, 2], dtype=np.int32),
                              np.array([1., 2, 3, 4]), 3, 3)
    assert plain.to_inner_uids([0, 2, 3, -1]).tolist() == [0, 2, -1, -1]
    assert plain.to_inner_iids(np.array([1, 5])).tolist() == [1, -1]


def test_every_environment_knob_of_the_library_is_documented():
    """DESIGN.md section 10a lists every SB2_* variable the library reads (getenv in csrc/)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    knobs = set()
    for f in glob.glob(os.path.join(root, "surprise_b200", "csrc", "*.cu*")):
        knobs |= set(re.findall(r'getenv\("(SB2_[A-Z0-9_]+)"\)', open(f).read()))
    assert knobs, "no getenv found: the scan is broken"
    design = open(os.path.join(root, "DESIGN.md")).read()
    missing = sorted(k for k in knobs if k not in design)
    assert not missing, "undocumented environment knobs: %s" % missing
