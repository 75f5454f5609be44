"""Synthetic ratings of a named shape: unique (u, i) pairs drawn uniformly,
ratings from a planted low-rank model so that held-out RMSE is informative, rounded to the scale's grid.
Inner ids are first-appearance order of the shuffled list, as Dataset.construct_trainset assigns them
(reference dataset.py:219-234).  Everything is seeded; nothing is read from disk."""
import numpy as np

SHAPES = {
    "ml-100k": (943, 1682, 100_000, 1.0),
    "ml-1m": (6040, 3706, 1_000_000, 1.0),      # BASELINE.json configs[1]
    "ml-10m": (72_000, 10_700, 10_000_000, 0.5),
    "ml-20m": (138_000, 27_000, 20_000_000, 0.5),
    "netflix": (480_000, 17_700, 100_000_000, 1.0),
}


def ratings(n_users, n_items, n_ratings, step=1.0, seed=0, holdout=0.1, rank=8, skew=False):
    """Returns dict(train=(u, i, r), test=(u, i, r), n_users, n_items) with compact inner ids for the train
    part (users / items that only occur in the test part get id -1 there = unknown).
    skew=True is the stress variant of SURVEY.md section 8d: Zipf(1.0) item popularity x log-normal user activity
    (MovieLens-like: a few items rated by a large share of the users, a few very active users)."""
    rng = np.random.default_rng(seed)
    n_total = int(round(n_ratings * (1 + holdout)))
    if n_total > 0.5 * n_users * n_items:
        raise ValueError("shape too dense for the unique-pair sampler")
    if skew:
        p_i = 1.0 / np.arange(1, n_items + 1)
        p_i = rng.permutation(p_i / p_i.sum())
        p_u = rng.lognormal(0.0, 1.0, n_users)
        p_u /= p_u.sum()

        def draw(k):
            return rng.choice(n_users, k, p=p_u).astype(np.int64) * n_items + rng.choice(n_items, k, p=p_i)
    else:
        def draw(k):
            return rng.integers(0, n_users * n_items, k, dtype=np.int64)
    def sorted_unique(a):
        # == np.unique(a), which numpy 2.3 evaluates ~50x slower than a plain sort for 10^7..10^8 int64 keys
        a = np.sort(a)
        if len(a) == 0:
            return a
        keep = np.empty(len(a), dtype=bool)
        keep[0] = True
        np.not_equal(a[1:], a[:-1], out=keep[1:])
        return a[keep]

    lin = sorted_unique(draw(int(n_total * 1.08) + 16))
    while len(lin) < n_total:
        lin = sorted_unique(np.concatenate((lin, draw(n_total))))
    lin = rng.permutation(lin)[:n_total]
    u_raw = (lin // n_items).astype(np.int64)
    i_raw = (lin % n_items).astype(np.int64)
    bu = rng.normal(0, .4, n_users)
    bi = rng.normal(0, .4, n_items)
    p = rng.normal(0, .35, (n_users, rank))
    q = rng.normal(0, .35, (n_items, rank))
    # planted model, evaluated in slabs of 4M ratings (the gathered factor rows of 10^8 ratings would be 14 GB)
    r = np.empty(n_total)
    noise = rng.normal(0, .8, n_total)
    for b in range(0, n_total, 1 << 22):
        sl = slice(b, min(n_total, b + (1 << 22)))
        us, is_ = u_raw[sl], i_raw[sl]
        r[sl] = 3.5 + bu[us] + bi[is_] + np.einsum("kf,kf->k", p[us], q[is_]) + noise[sl]
    del noise
    lo = step if step < 1 else 1.0
    r = np.clip(np.round(r / step) * step, lo, 5.0)

    def first_appearance(a, n):
        # rank of every id by its first occurrence in `a` (-1: absent).  O(len(a)): with repeated indices a fancy
        # assignment keeps the LAST value written, so writing the positions back to front leaves the first one
        first = np.full(n, len(a), dtype=np.int64)
        first[a[::-1]] = np.arange(len(a) - 1, -1, -1, dtype=np.int64)
        uniq = np.nonzero(first < len(a))[0]
        order = np.argsort(first[uniq], kind="stable")
        rank_of = np.full(n, -1, dtype=np.int64)
        rank_of[uniq[order]] = np.arange(len(uniq))
        return rank_of

    n_tr = n_ratings
    map_u = first_appearance(u_raw[:n_tr], n_users)
    map_i = first_appearance(i_raw[:n_tr], n_items)
    tr = (map_u[u_raw[:n_tr]].astype(np.int32), map_i[i_raw[:n_tr]].astype(np.int32), r[:n_tr].copy())
    te = (map_u[u_raw[n_tr:]].astype(np.int32), map_i[i_raw[n_tr:]].astype(np.int32), r[n_tr:].copy())
    return dict(train=tr, test=te, n_users=int(map_u.max()) + 1, n_items=int(map_i.max()) + 1)


def shaped(name, seed=0, scale=1.0):
    nu, ni, n, step = SHAPES[name]
    return ratings(int(nu * scale), int(ni * scale), int(n * scale * scale), step=step, seed=seed)


def shaped_cached(name, seed=0, cache_dir=None, with_coo=False):
    """shaped(name, seed) through an on-disk cache of the generated arrays (uncompressed .npz under cache_dir): the
    Netflix shape takes ~40 s to draw and the benchmark needs it once per process and per N.  with_coo=True also
    caches the all_ratings()-ordered COO (u ascending, file order inside a user: what Trainset.coo() returns).
    Concurrent callers may race to write; the file is written to a temporary name and renamed."""
    import os
    if cache_dir is None:
        return shaped(name, seed=seed)
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, "%s_seed%d%s.npz" % (name, seed, "_coo" if with_coo else ""))
    if os.path.exists(path):
        z = np.load(path)
        d = dict(train=(z["u"], z["i"], z["r"]), test=(z["tu"], z["ti"], z["tr"]), n_users=int(z["n_users"]),
                 n_items=int(z["n_items"]))
        if with_coo:
            d["coo"] = (z["cu"], z["ci"], z["cr"])
        return d
    d = shaped(name, seed=seed)
    arrays = dict(u=d["train"][0], i=d["train"][1], r=d["train"][2], tu=d["test"][0], ti=d["test"][1], tr=d["test"][2],
                  n_users=np.int64(d["n_users"]), n_items=np.int64(d["n_items"]))
    if with_coo:
        from .trainset import Trainset
        ts = Trainset.from_coo(d["train"][0], d["train"][1], d["train"][2], d["n_users"], d["n_items"])
        d["coo"] = tuple(np.ascontiguousarray(a) for a in ts.coo())
        arrays.update(cu=d["coo"][0], ci=d["coo"][1], cr=d["coo"][2])
    tmp = path + ".%d.tmp.npz" % os.getpid()
    np.savez(tmp, **arrays)
    os.replace(tmp, path)
    return d
