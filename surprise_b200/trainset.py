"""Array-backed Trainset.

Same public surface as the reference's surprise/trainset.py:11-261 (ur / ir dict-of-lists, inner ids,
knows_*, to_inner_* / to_raw_*, all_ratings, build_testset, build_anti_testset, global_mean), but the
ratings live in flat CSR arrays so that 20M-100M-rating sets never become Python tuples:

    ur CSR : u_ptr[n_users+1], ui_idx[n], u_r[n]   -- ur[u] in list order   (== all_ratings() order)
    ir CSR : i_ptr[n_items+1], iu_idx[n], i_r[n]   -- ir[i] in list order

The kernels' input contract is the iteration ORDER of those lists (inner ids are first-appearance
order, dataset.py:219-234; all_ratings() is u ascending then ur[u] order, trainset.py:180-190), so both
representations are kept consistent: whichever one the Trainset was built from, the other is
materialised lazily and lists exactly the same sequence.
"""
import numpy as np


def _csr_from_dict(d, n):
    ptr = np.zeros(n + 1, dtype=np.int64)
    for k in range(n):
        ptr[k + 1] = len(d[k])
    np.cumsum(ptr, out=ptr)
    idx = np.empty(ptr[-1], dtype=np.int32)
    val = np.empty(ptr[-1], dtype=np.float64)
    for k in range(n):
        o = ptr[k]
        for t, (j, r) in enumerate(d[k]):
            idx[o + t] = j
            val[o + t] = r
    return ptr, idx, val


def _dict_from_csr(ptr, idx, val):
    out = {}
    idx_l, val_l = idx.tolist(), val.tolist()
    for k in range(len(ptr) - 1):
        b, e = int(ptr[k]), int(ptr[k + 1])
        out[k] = list(zip(idx_l[b:e], val_l[b:e]))
    return out


def _group_stable(key, other, val, n, counts):
    """CSR of the ratings grouped by `key`, file order kept inside a group: (ptr int64, other, val).
    A counting sort (scipy's coo_tocsr kernel, O(N)) when available; numpy's stable argsort otherwise
    (2 s per side at 10M ratings)."""
    ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    if len(val) < 2 ** 31 - 1:
        try:
            from scipy.sparse import _sparsetools
            bp = np.empty(n + 1, dtype=np.int32)
            bj = np.empty(len(val), dtype=np.int32)
            bx = np.empty(len(val), dtype=np.float64)
            _sparsetools.coo_tocsr(n, int(other.max()) + 1 if len(other) else 1, len(val), key, other, val, bp, bj, bx)
            return ptr, bj, bx
        except Exception:  # pragma: no cover - scipy missing or its private kernel moved
            pass
    o = np.argsort(key, kind="stable")
    return ptr, other[o], val[o]


class Trainset(object):
    def __init__(self, ur, ir, n_users, n_items, n_ratings, rating_scale, offset, raw2inner_id_users,
                 raw2inner_id_items):
        self._ur, self._ir = ur, ir
        self._ucsr = self._icsr = None
        self.n_users, self.n_items, self.n_ratings = n_users, n_items, n_ratings
        self.rating_scale, self.offset = rating_scale, offset
        self._raw2inner_id_users = raw2inner_id_users
        self._raw2inner_id_items = raw2inner_id_items
        self._inner2raw_id_users = self._inner2raw_id_items = None
        self._global_mean = None

    # -- construction from flat arrays (the fast path; no Python tuples) -----------------------------
    @classmethod
    def from_coo(cls, u, i, r, n_users=None, n_items=None, rating_scale=(1, 5), offset=0, raw_uids=None,
                 raw_iids=None):
        """u, i: inner ids (0..n-1, every id used) in file order; r: ratings already offset."""
        u = np.ascontiguousarray(u, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        r = np.ascontiguousarray(r, dtype=np.float64)
        n_users = int(u.max()) + 1 if n_users is None else n_users
        n_items = int(i.max()) + 1 if n_items is None else n_items
        ts = cls(None, None, n_users, n_items, len(r), rating_scale, offset, None, None)
        cu = np.bincount(u, minlength=n_users)
        ci = np.bincount(i, minlength=n_items)
        if (cu == 0).any() or (ci == 0).any():
            raise ValueError("from_coo needs compact inner ids (every user / item id must have a rating)")
        ts._ucsr = _group_stable(u, i, r, n_users, cu)
        ts._icsr = _group_stable(i, u, r, n_items, ci)
        ts._raw_uids, ts._raw_iids = raw_uids, raw_iids
        return ts

    # -- the two representations -----------------------------------------------------------------------
    @property
    def ur(self):
        if self._ur is None:
            self._ur = _dict_from_csr(*self._ucsr)
        return self._ur

    @property
    def ir(self):
        if self._ir is None:
            self._ir = _dict_from_csr(*self._icsr)
        return self._ir

    def user_csr(self):
        """(u_ptr, ui_idx, u_r): ur flattened in list order."""
        if self._ucsr is None:
            self._ucsr = _csr_from_dict(self._ur, self.n_users)
        return self._ucsr

    def item_csr(self):
        """(i_ptr, iu_idx, i_r): ir flattened in list order."""
        if self._icsr is None:
            self._icsr = _csr_from_dict(self._ir, self.n_items)
        return self._icsr

    def coo(self):
        """(u, i, r) in all_ratings() order."""
        ptr, idx, val = self.user_csr()
        u = np.repeat(np.arange(self.n_users, dtype=np.int32), np.diff(ptr))
        return u, idx, val

    # -- reference API -------------------------------------------------------------------------------------
    def knows_user(self, uid):
        return isinstance(uid, (int, np.integer)) and 0 <= uid < self.n_users

    def knows_item(self, iid):
        return isinstance(iid, (int, np.integer)) and 0 <= iid < self.n_items

    def _raw_maps(self):
        if self._raw2inner_id_users is None:
            ru = getattr(self, "_raw_uids", None)
            ri = getattr(self, "_raw_iids", None)
            self._raw2inner_id_users = ({k: k for k in range(self.n_users)} if ru is None
                                        else {raw: k for k, raw in enumerate(ru)})
            self._raw2inner_id_items = ({k: k for k in range(self.n_items)} if ri is None
                                        else {raw: k for k, raw in enumerate(ri)})

    def _inner_ids_of(self, raw, which):
        """Vectorised raw -> inner id lookup for a whole testset: int32 array, -1 for ids the trainset does not
        know (the reference's 'UKN__' ids, algo_base.py:140-147).  Integer raw ids of an array-backed trainset go
        through a sorted table (np.searchsorted); anything else through the same dict the scalar lookups use."""
        n = self.n_users if which == "u" else self.n_items
        raws = getattr(self, "_raw_uids" if which == "u" else "_raw_iids", None)
        have_dict = (self._raw2inner_id_users if which == "u" else self._raw2inner_id_items) is not None
        arr = np.asarray(raw) if not isinstance(raw, np.ndarray) else raw
        if arr.dtype.kind in "iu" and arr.ndim == 1:
            if raws is None and not have_dict:       # Trainset.from_coo without raw ids: raw == inner
                out = arr.astype(np.int64)
                return np.where((out >= 0) & (out < n), out, -1).astype(np.int32)
            if raws is not None:
                tab = getattr(self, "_raw_table_" + which, None)
                if tab is None:
                    ra = np.asarray(raws)
                    if ra.dtype.kind in "iu":
                        order = np.argsort(ra, kind="stable")
                        tab = (ra[order], order.astype(np.int32))
                    else:
                        tab = False
                    setattr(self, "_raw_table_" + which, tab)
                if tab is not False:
                    keys, inner = tab
                    pos = np.minimum(np.searchsorted(keys, arr), len(keys) - 1)
                    return np.where(keys[pos] == arr, inner[pos], -1).astype(np.int32)
        self._raw_maps()
        d = self._raw2inner_id_users if which == "u" else self._raw2inner_id_items
        seq = raw.tolist() if isinstance(raw, np.ndarray) else raw
        return np.fromiter((d.get(x, -1) for x in seq), dtype=np.int32, count=len(seq))

    def to_inner_uids(self, raw_uids):
        return self._inner_ids_of(raw_uids, "u")

    def to_inner_iids(self, raw_iids):
        return self._inner_ids_of(raw_iids, "i")

    def to_inner_uid(self, ruid):
        self._raw_maps()
        try:
            return self._raw2inner_id_users[ruid]
        except KeyError:
            raise ValueError("User " + str(ruid) + " is not part of the trainset.")

    def to_inner_iid(self, riid):
        self._raw_maps()
        try:
            return self._raw2inner_id_items[riid]
        except KeyError:
            raise ValueError("Item " + str(riid) + " is not part of the trainset.")

    def to_raw_uid(self, iuid):
        self._raw_maps()
        if self._inner2raw_id_users is None:
            self._inner2raw_id_users = {v: k for k, v in self._raw2inner_id_users.items()}
        try:
            return self._inner2raw_id_users[iuid]
        except KeyError:
            raise ValueError(str(iuid) + " is not a valid inner id.")

    def to_raw_iid(self, iiid):
        self._raw_maps()
        if self._inner2raw_id_items is None:
            self._inner2raw_id_items = {v: k for k, v in self._raw2inner_id_items.items()}
        try:
            return self._inner2raw_id_items[iiid]
        except KeyError:
            raise ValueError(str(iiid) + " is not a valid inner id.")

    def all_ratings(self):
        u, i, r = self.coo()
        for t in zip(u.tolist(), i.tolist(), r.tolist()):
            yield t

    def all_users(self):
        return range(self.n_users)

    def all_items(self):
        return range(self.n_items)

    def build_testset(self):
        return [(self.to_raw_uid(u), self.to_raw_iid(i), r) for (u, i, r) in self.all_ratings()]

    def build_anti_testset(self, fill=None):
        fill = self.global_mean if fill is None else float(fill)
        ptr, idx, _ = self.user_csr()
        out = []
        for u in self.all_users():
            seen = set(idx[ptr[u]:ptr[u + 1]].tolist())
            ru = self.to_raw_uid(u)
            out += [(ru, self.to_raw_iid(i), fill) for i in self.all_items() if i not in seen]
        return out

    @property
    def global_mean(self):
        """np.mean over the ratings in all_ratings() order -- the same pairwise reduction over the same
        sequence as trainset.py:257-259, hence the same bits."""
        if self._global_mean is None:
            self._global_mean = np.mean(self.user_csr()[2])
        return self._global_mean
