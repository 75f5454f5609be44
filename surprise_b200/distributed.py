"""Multi-GPU paths (one process per GPU, torch.distributed for the plumbing): SVD / SVD++ ring, row-sharded
similarity build and k-NN estimates, sharded NMF.

SVD / SVD++ (DSGD strata across ranks).  With P ranks, user u belongs to rank u % P (local row u // P) and item i to
item super-block i % P (local row i // P).  An epoch is P sub-epochs; in sub-epoch E (counted from the start of the
fit) rank g updates the ratings whose user it owns and whose item lies in super-block (g + E) % P, then that
super-block's (qi, bi) rows move to rank g - 1.  The user rows never move.  This is the rank-level repetition of
the CTA-level ring inside the kernel (csrc/sgd.cu) and it runs INSIDE the kernel too: the CTA that finishes an item
block stores it into the left neighbour's buffer through a peer mapping of that rank's memory (cudaIpc, NVLink) and
publishes it with a system-scope release flag; the neighbour's CTAs poll their own memory.  One persistent launch
per rank covers a whole SVD fit (SVD++: one per epoch, around the all-reduce of the per-item y_j sums).

``partition`` / ``ring_epochs`` / ``make_exchange`` below are the HOST MODEL of that schedule: the gloo tests run it
with the oracle's sequential SGD as the unit of work to pin partitioning and rotation without a GPU.
"""
import numpy as np


def partition(u, i, r, rank, world):
    """Ratings owned by `rank`, split by item super-block.  Returns list over sb of (u_local, i_local, r)."""
    mine = (u % world) == rank
    um, im, rm = u[mine], i[mine], r[mine]
    out = []
    for sb in range(world):
        sel = (im % world) == sb
        out.append(((um[sel] // world).astype(np.int32), (im[sel] // world).astype(np.int32), rm[sel]))
    return out


def local_rows(n, part, world):
    """number of ids in [0, n) congruent to part mod world"""
    return (n - part + world - 1) // world


def ring_epochs(rank, world, n_epochs, held_sb, run_block, exchange):
    """The rotation schedule.  held_sb: super-block this rank holds at entry (== rank).  `exchange()` sends
    the held item block to rank-1 and receives from rank+1 (in place)."""
    sb = held_sb
    for _ in range(n_epochs):
        for _s in range(world):
            run_block(sb)
            if world > 1:
                exchange()
                sb = (sb + 1) % world
    return sb


def make_exchange(dist, rank, world, tensors, scratch):
    """P2P rotation of the item-side tensors: send to rank-1, receive from rank+1."""
    dst, src = (rank - 1) % world, (rank + 1) % world

    def exchange():
        ops = []
        for t, s in zip(tensors, scratch):
            ops.append(dist.P2POp(dist.isend, t, dst))
            ops.append(dist.P2POp(dist.irecv, s, src))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for t, s in zip(tensors, scratch):
            t.copy_(s)
    return exchange


def ring_neighbours(rank, world):
    """(left, right): the rank that receives this rank's item blocks, the rank that delivers them."""
    return (rank - 1) % world, (rank + 1) % world


def interleave_rows(parts, n_total, world):
    """parts[p]: the rows p, p + world, p + 2 world, ... (possibly padded at the end) -> the n_total rows in order."""
    first = parts[0]
    full = first.new_zeros((n_total,) + tuple(first.shape[1:]))
    for p in range(world):
        full[p::world] = parts[p][:local_rows(n_total, p, world)]
    return full


def all_gather_interleaved(dist, local, n_total, rank, world):
    """Every rank contributes its rows (rank, rank + world, ...) of an n_total-row array; returns all rows in order."""
    import torch
    if world == 1:
        return local
    n_max = local_rows(n_total, 0, world)
    pad = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return interleave_rows(parts, n_total, world)


class RingSVD(object):
    """SVD / SVD++ fit sharded over the ranks of an initialised torch.distributed group (module docstring).

    u, i, r: the all_ratings() COO (host arrays or CUDA tensors; every rank passes the same data and keeps the
    ratings of its own users); prm: _native.SgdParams; ur_csr = (u_ptr, ui_idx) makes it SVD++.
    dist=None (or world 1) degenerates to the single-GPU plan.  Collective: every rank must construct, reset,
    run, gather and close together."""

    def __init__(self, dist, u, i, r, n_users, n_items, prm, ur_csr=None):
        import ctypes as C
        from . import _native as nat
        self.nat, self.C, self.dist = nat, C, dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.n_users, self.n_items, self.prm = int(n_users), int(n_items), prm
        self.f = prm.n_factors
        self.with_yj = ur_csr is not None
        lib = nat.lib()
        torch = nat.torch_cuda()
        d_u, d_i, d_r = nat.to_dev(u, np.int32), nat.to_dev(i, np.int32), nat.to_dev(r, np.float64)
        d_up = d_ui = None
        if self.with_yj:
            d_up, d_ui = nat.to_dev(ur_csr[0], np.int64), nat.to_dev(ur_csr[1], np.int32)
        self.plan = C.c_void_p()
        nat.check(lib.sb2_svd_ring_create_dev(self.n_users, self.n_items, int(d_r.shape[0]), nat.ptr(d_u), nat.ptr(d_i),
                                              nat.ptr(d_r), C.byref(prm), int(self.with_yj), nat.ptr(d_up), nat.ptr(d_ui),
                                              self.rank, self.world, nat.stream(), C.byref(self.plan)))
        nu, ni, nl, stride = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        lib.sb2_svd_ring_info(self.plan, C.byref(nu), C.byref(ni), C.byref(nl), C.byref(stride))
        self.nu_loc, self.ni_loc, self.n_local, self.stride = nu.value, ni.value, nl.value, stride.value
        self._sync = torch.zeros(1, dtype=torch.int32, device=nat.device())
        self._xch = (torch.empty((self.n_items, self.stride + 1), dtype=torch.float32, device=nat.device())
                     if self.with_yj else None)
        if self.world > 1:
            mine = (C.c_ubyte * 64)()
            nat.check(lib.sb2_svd_ring_ipc_handle(self.plan, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine))
            left, right = ring_neighbours(self.rank, self.world)
            nat.check(lib.sb2_svd_ring_connect_ipc(self.plan, (C.c_ubyte * 64).from_buffer_copy(handles[left]),
                                                   (C.c_ubyte * 64).from_buffer_copy(handles[right])))

    def _rendezvous(self):
        """Stream-ordered meeting point of the ranks (a 4-byte all-reduce): everything every rank enqueued before
        it has completed on its device before anything enqueued after it starts anywhere."""
        if self.world > 1:
            self.dist.all_reduce(self._sync)

    def reset(self, pu0, qi0, yj0=None):
        """pu0 / qi0 (/ yj0): the WHOLE float64 initial matrices (host or device); each rank keeps its rows."""
        nat = self.nat
        d_pu, d_qi = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
        d_yj = nat.to_dev(yj0, np.float64) if self.with_yj else None
        self._rendezvous()   # nobody clears its mailboxes while a neighbour's previous fit is still running
        nat.check(nat.lib().sb2_svd_plan_reset_dev(self.plan, nat.ptr(d_pu), nat.ptr(d_qi), nat.ptr(d_yj), nat.stream()))
        self._rendezvous()   # every mailbox is clear before any kernel of the new fit starts

    def run(self, n_epochs):
        nat = self.nat
        lib = nat.lib()
        if not self.with_yj:
            nat.check(lib.sb2_svd_plan_run(self.plan, int(n_epochs), nat.stream()))
            return
        for _ in range(int(n_epochs)):
            nat.check(lib.sb2_svd_ring_epoch_dev(self.plan, 0, nat.ptr(self._xch), nat.stream()))
            if self.world > 1:
                self.dist.all_reduce(self._xch)
            nat.check(lib.sb2_svd_ring_epoch_dev(self.plan, 1, nat.ptr(self._xch), nat.stream()))

    def status(self):
        """Synchronises; raises NativeError if a wait inside the kernel ran into its deadline."""
        self.nat.check(self.nat.lib().sb2_svd_plan_status(self.plan, self.nat.stream()))

    def gather(self):
        """Full (pu, qi, bu, bi[, yj]) float64 numpy arrays on every rank."""
        nat = self.nat
        f = self.f
        pu = nat.empty_dev((self.nu_loc, f), np.float64); qi = nat.empty_dev((self.ni_loc, f), np.float64)
        bu = nat.empty_dev((self.nu_loc,), np.float64); bi = nat.empty_dev((self.ni_loc,), np.float64)
        yj = nat.empty_dev((self.n_items, f), np.float64) if self.with_yj else None
        self._rendezvous()   # the last blocks pushed by the right neighbour have landed
        nat.check(nat.lib().sb2_svd_plan_read_dev(self.plan, nat.ptr(pu), nat.ptr(qi), nat.ptr(bu), nat.ptr(bi), nat.ptr(yj),
                                                  nat.stream()))
        self.status()
        outs = []
        for t, n_tot in ((pu, self.n_users), (qi, self.n_items), (bu, self.n_users), (bi, self.n_items)):
            outs.append(all_gather_interleaved(self.dist, t, n_tot, self.rank, self.world).cpu().numpy())
        if self.with_yj:
            outs.append(yj.cpu().numpy())
        return tuple(outs)

    def close(self):
        if self.plan:
            self._rendezvous()   # no neighbour still maps / writes this rank's buffers
            self.nat.torch_cuda().cuda.synchronize()
            self.nat.lib().sb2_svd_plan_destroy(self.plan)
            self.plan = None


def fit_sharded(algo, trainset, dist):
    """algo: a surprise_b200 SVD or SVDpp instance; fits it over the ranks of `dist` (every rank calls this with the
    same trainset and ends up with the same fitted attributes), so that algo.test() / predict() work as usual.
    Mirrors algo.fit(trainset) (matrix_factorization.pyx:153-156, :413-418)."""
    from .prediction_algorithms.algo_base import AlgoBase
    AlgoBase.fit(algo, trainset)
    pu0, qi0, yj0 = algo._initial_factors(trainset)
    u, i, r = trainset.coo()
    ur = trainset.user_csr()[:2] if yj0 is not None else None
    ring = RingSVD(dist, u, i, r, trainset.n_users, trainset.n_items, algo._sgd_params(trainset), ur_csr=ur)
    try:
        ring.reset(pu0, qi0, yj0)
        ring.run(algo.n_epochs)
        out = ring.gather()
    finally:
        ring.close()
    algo.pu, algo.qi, algo.bu, algo.bi = out[:4]
    if yj0 is not None:
        algo.yj = out[4]
    algo._dev_cache = None
    return algo


# ------------------------------------------------------------------------------------------------------------
# Row-block sharded similarity build.  Every rank packs the (replicated) rating CSR.
#   symmetric=True (default): rank k owns rows [b_k, b_{k+1}) and computes only the columns >= b_k of them
#     (sb2_sim_build_upper_dev: tiles at or above the block diagonal), with the b_k chosen so that every rank gets
#     the same number of upper-triangular tiles -- 1/N of the single-GPU work.  The missing columns [0, b_k) are the
#     transposes of sub-blocks computed by the ranks before k; one round of NCCL send / recv moves them (half the
#     matrix in total, ~3 GB at the ml-20M shape, against ~1 s of tensor-core work).
#   symmetric=False: every rank computes the whole rectangle of its rows (twice the work, no exchange).
# Row blocks are multiples of the 256-row tile of the CTA-pair MMA kernel.
# ------------------------------------------------------------------------------------------------------------
def sim_row_range(n_x, rank, world, tile=256):
    """Contiguous, tile-aligned row range of `rank` (balanced in tiles)."""
    n_tiles = (n_x + tile - 1) // tile
    lo = (n_tiles * rank) // world
    hi = (n_tiles * (rank + 1)) // world
    return min(lo * tile, n_x), min(hi * tile, n_x)


def sim_tri_ranges(n_x, world, tile=256):
    """world contiguous tile-aligned row ranges with (nearly) equal numbers of upper-triangular tiles: row block rb
    costs n_tiles - rb tiles.  Ranges may be empty when there are fewer row blocks than ranks."""
    n_tiles = (n_x + tile - 1) // tile
    total = n_tiles * (n_tiles + 1) // 2
    bounds, acc, rb = [0], 0, 0
    for k in range(1, world):
        target = total * k / world
        while rb < n_tiles and acc + (n_tiles - rb) / 2.0 <= target:
            acc += n_tiles - rb
            rb += 1
        bounds.append(rb)
    bounds.append(n_tiles)
    return [(min(bounds[k] * tile, n_x), min(bounds[k + 1] * tile, n_x)) for k in range(world)]


def sim_exchange_plan(ranges, rank):
    """The transposed sub-blocks `rank` sends and receives in the symmetric build: lists of (peer, rows, cols) in
    the coordinates of the full matrix.  Rank j < k sends sim[rows_j, rows_k]; rank k stores its transpose."""
    lo, hi = ranges[rank]
    sends = [(k, (lo, hi), ranges[k]) for k in range(rank + 1, len(ranges)) if hi > lo and ranges[k][1] > ranges[k][0]]
    recvs = [(j, ranges[j], (lo, hi)) for j in range(rank) if hi > lo and ranges[j][1] > ranges[j][0]]
    return sends, recvs


def sim_exchange(dist, torch, block, ranges, rank):
    """In place: fills the columns [0, row_begin) of `block` (rows ranges[rank] of the symmetric matrix, columns
    >= row_begin already computed) with the transposes of the sub-blocks the earlier ranks computed."""
    b, e = ranges[rank]
    sends, recvs = sim_exchange_plan(ranges, rank)
    ops, bufs = [], []
    for peer, _, (c0, c1) in sends:
        ops.append(dist.P2POp(dist.isend, block[:, c0:c1].contiguous(), peer))
    for peer, (r0, r1), _ in recvs:
        buf = torch.empty((r1 - r0, e - b), dtype=block.dtype, device=block.device)
        bufs.append((buf, r0, r1))
        ops.append(dist.P2POp(dist.irecv, buf, peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, r0, r1 in bufs:
        block[:, r0:r1] = buf.t()


def upload_inputs_broadcast(dist, kind, n_x, yr, x_biases=None, y_biases=None):
    """similarities.upload_inputs for a group of ranks: rank 0 copies the rating CSR (and the baselines) to its GPU and
    the other ranks receive them GPU -> GPU over NCCL ("rating matrix broadcast once"), instead of N
    processes pushing the same 240 MB (ml-20M shape) through the host's PCIe / pageable staging at the same time.
    Every rank passes the same host yr (only its sizes are used off rank 0)."""
    from . import similarities as sims
    nat = sims.nat
    torch = nat.torch_cuda()
    rank = dist.get_rank()
    pb = kind == "pearson_baseline"
    if rank == 0:
        inp = sims.upload_inputs(kind, n_x, yr, x_biases, y_biases)
        meta = torch.tensor([inp["n_y"], inp["nnz"], inp["denom"]], dtype=torch.int64, device=nat.device())
    else:
        meta = torch.zeros(3, dtype=torch.int64, device=nat.device())
    dist.broadcast(meta, 0)
    n_y, nnz, denom = (int(v) for v in meta.tolist())
    if rank != 0:
        inp = dict(n_x=int(n_x), n_y=n_y, nnz=nnz, denom=denom, ptr=nat.empty_dev((n_y + 1,), np.int64),
                   idx=nat.empty_dev((nnz,), np.int32), val=nat.empty_dev((nnz,), np.float64),
                   bx=nat.empty_dev((int(n_x),), np.float64) if pb else None,
                   by=nat.empty_dev((n_y,), np.float64) if pb else None)
    elif pb:   # rank 0 may hold longer bias arrays than n_x / n_y: broadcast exactly what the kernels read
        inp["bx"], inp["by"] = inp["bx"][:int(n_x)].contiguous(), inp["by"][:n_y].contiguous()
    for k in ("ptr", "idx", "val") + (("bx", "by") if pb else ()):
        dist.broadcast(inp[k], 0)
    return inp


def sim_build_sharded(dist, kind, n_x, yr, min_support, gather=False, symmetric=True, **kw):
    """Returns (row_begin, row_end, block) with block a CUDA float64 tensor (row_end-row_begin) x n_x; with
    gather=True every rank also receives the full matrix (all_gather over NCCL), returned instead of the block."""
    from . import similarities as sims
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    symmetric = symmetric and world > 1
    ranges = sim_tri_ranges(n_x, world) if symmetric else [sim_row_range(n_x, r, world) for r in range(world)]
    b, e = ranges[rank]
    torch = sims.nat.torch_cuda()
    if world > 1 and kw.get("inputs") is None:
        kw = dict(kw, inputs=upload_inputs_broadcast(dist, kind, n_x, yr, kw.get("x_biases"), kw.get("y_biases")))
    if e > b:
        block = sims.build_device(kind, n_x, yr, min_support, row_begin=b, row_end=e, upper=symmetric, **kw)
    else:
        block = torch.empty((0, n_x), dtype=torch.float64, device=sims.nat.device())
    if symmetric:
        sim_exchange(dist, torch, block, ranges, rank)
    if not gather or world == 1:
        return b, e, block
    rows_max = max(hi - lo for lo, hi in ranges)
    pad = torch.zeros((rows_max, n_x), dtype=torch.float64, device=block.device)
    pad[:e - b] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(ranges)], dim=0)
    return 0, n_x, full


def knn_predict_sharded(dist, block, row_begin, row_end, n_x, x, y, yr, k, min_k, mode=0, global_mean=0.0, bx=None,
                        by=None):
    """Batched k-NN estimates with the similarity matrix left row-sharded (SURVEY.md 8e "KNN predict"): every rank
    runs sb2_knn_predict_dev on the pairs whose x lies in its rows [row_begin, row_end) of `block`, the per-pair
    results are summed across ranks (each pair is owned by exactly one rank; unknown x < 0 by rank 0).
    x, y: int32 host arrays (identical on every rank); yr: (ptr, idx, val) host CSR.  Returns (est, actual_k,
    impossible) numpy arrays in the order of the input pairs."""
    from . import _native as nat
    torch = nat.torch_cuda()
    rank = dist.get_rank() if dist is not None else 0
    x = np.ascontiguousarray(x, dtype=np.int32); y = np.ascontiguousarray(y, dtype=np.int32)
    own = (x >= row_begin) & (x < row_end)
    if rank == 0:
        own |= x < 0
    sel = np.nonzero(own)[0]
    n = len(sel)
    dev = block.device
    est = torch.zeros(len(x), dtype=torch.float64, device=dev)
    ak = torch.zeros(len(x), dtype=torch.int32, device=dev)
    imp = torch.zeros(len(x), dtype=torch.int32, device=dev)
    if n:
        d_x, d_y = nat.to_dev(x[sel], np.int32), nat.to_dev(y[sel], np.int32)
        d_ptr, d_idx, d_val = nat.to_dev(yr[0], np.int64), nat.to_dev(yr[1], np.int32), nat.to_dev(yr[2], np.float64)
        d_bx = nat.to_dev(bx, np.float64) if bx is not None else None
        d_by = nat.to_dev(by, np.float64) if by is not None else None
        e_l = nat.empty_dev((n,), np.float64); a_l = nat.empty_dev((n,), np.int32); i_l = nat.empty_dev((n,), np.uint8)
        # the kernel addresses sim + x * ld: shift the base so that global row ids index the local block
        base = block.data_ptr() - int(row_begin) * int(n_x) * 8
        nat.check(nat.lib().sb2_knn_predict_dev(n, nat.ptr(d_x), nat.ptr(d_y), n_x, base, n_x, nat.ptr(d_ptr), nat.ptr(d_idx),
                                                nat.ptr(d_val), int(k), int(min_k), int(mode), float(global_mean),
                                                nat.ptr(d_bx), nat.ptr(d_by), nat.ptr(e_l), nat.ptr(a_l), nat.ptr(i_l),
                                                nat.stream()))
        idx = torch.as_tensor(sel, device=dev)
        est[idx] = e_l; ak[idx] = a_l; imp[idx] = i_l.to(torch.int32)
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(est); dist.all_reduce(ak); dist.all_reduce(imp)
    return est.cpu().numpy(), ak.cpu().numpy(), imp.cpu().numpy().astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------
# NMF sharded over ranks, bit-exact ("bit-exact formulation"): rank g evaluates the ordered
# accumulator sums of a contiguous range of users and a contiguous range of items (sb2_nmf_plan_epoch_dev) and
# the new factor rows are all-gathered after every epoch (NCCL; pu is 58 MB at Netflix scale).
# ------------------------------------------------------------------------------------------------------------
def block_ranges(n, world):
    """world contiguous ranges of ceil(n / world) rows each (the last ones shorter or empty): equal-sized blocks are
    what an in-place NCCL all-gather needs."""
    per = (n + world - 1) // world
    return per, [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def nmf_fit_sharded(dist, n_users, n_items, u, i, r, prm, pu0, qi0, stats=None):
    """u, i, r: all_ratings COO (host arrays, identical on every rank); prm: _native.NmfParams; pu0 / qi0: the
    rng.uniform initial factors.  Returns (pu, qi, bu, bi) as float64 numpy arrays, identical on every rank and
    bit-identical to the single-GPU fit.  stats (dict, optional) receives 'epochs_s': the wall clock of the epoch
    loop alone (device-synchronised, barrier on both sides), without upload / plan creation / download.

    Rank g evaluates the ordered accumulator sums of user block g and item block g (equal-sized contiguous blocks);
    after every epoch the new rows are all-gathered IN PLACE (the factor buffers are padded to world x block rows and
    each rank's block is the slice NCCL expects, so no staging copies: one all_gather_into_tensor per matrix)."""
    import ctypes as C
    import time
    from . import _native as nat
    torch = nat.torch_cuda()
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    lib = nat.lib()
    f = prm.n_factors
    d_u, d_i, d_r = nat.to_dev(u, np.int32), nat.to_dev(i, np.int32), nat.to_dev(r, np.float64)
    per_u, ur = block_ranges(n_users, world)
    per_i, ir = block_ranges(n_items, world)
    dev = nat.device()

    def padded(a0, per):
        t = torch.zeros((per * world, f), dtype=torch.float64, device=dev)
        t[:a0.shape[0]] = nat.to_dev(a0, np.float64)
        return t
    pu = [padded(pu0, per_u), torch.zeros((per_u * world, f), dtype=torch.float64, device=dev)]
    qi = [padded(qi0, per_i), torch.zeros((per_i * world, f), dtype=torch.float64, device=dev)]
    bu = torch.zeros(n_users, dtype=torch.float64, device=dev)
    bi = torch.zeros(n_items, dtype=torch.float64, device=dev)
    plan = C.c_void_p()
    nat.check(lib.sb2_nmf_plan_create_dev(n_users, n_items, len(r), nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r),
                                          prm.n_factors, nat.stream(), C.byref(plan)))
    (u0, u1), (i0, i1) = ur[rank], ir[rank]
    try:
        cur = 0
        if stats is not None:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
        for _ in range(prm.n_epochs):
            nat.check(lib.sb2_nmf_plan_epoch_dev(plan, C.byref(prm), nat.ptr(pu[cur]), nat.ptr(qi[cur]),
                                                 nat.ptr(pu[cur ^ 1]), nat.ptr(qi[cur ^ 1]), nat.ptr(bu), nat.ptr(bi),
                                                 u0, u1, i0, i1, nat.stream()))
            cur ^= 1
            if world > 1:
                dist.all_gather_into_tensor(pu[cur], pu[cur][rank * per_u:(rank + 1) * per_u])
                dist.all_gather_into_tensor(qi[cur], qi[cur][rank * per_i:(rank + 1) * per_i])
        nat.check(lib.sb2_nmf_plan_status(plan, nat.stream()))
        if stats is not None:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            stats["epochs_s"] = time.perf_counter() - t0
    finally:
        lib.sb2_nmf_plan_destroy(plan)
    return pu[cur][:n_users].cpu().numpy(), qi[cur][:n_items].cpu().numpy(), bu.cpu().numpy(), bi.cpu().numpy()


def baseline_als_sharded(dist, trainset, n_epochs=10, reg_u=15.0, reg_i=10.0):
    """baseline_als (optimize_baselines.pyx:14-54) over the ranks of `dist` (last row): per epoch the
    item biases of item block g are evaluated on rank g (ordered sums over ir[i], needing all of bu), all-gathered in
    place, then the same for the user biases.  Every ordered sum lives on one rank: (bu, bi) are bit-identical to the
    single-GPU sb2_baseline_als_dev on every rank.  Returns float64 numpy arrays."""
    from . import _native as nat
    torch = nat.torch_cuda()
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    lib = nat.lib()
    up, ui, ur = trainset.user_csr()
    ip, iu, ir = trainset.item_csr()
    d_up, d_ui, d_ur = nat.to_dev(up, np.int64), nat.to_dev(ui, np.int32), nat.to_dev(ur, np.float64)
    d_ip, d_iu, d_ir = nat.to_dev(ip, np.int64), nat.to_dev(iu, np.int32), nat.to_dev(ir, np.float64)
    n_users, n_items, mu = trainset.n_users, trainset.n_items, float(trainset.global_mean)
    per_u, ru = block_ranges(n_users, world)
    per_i, ri = block_ranges(n_items, world)
    dev = nat.device()
    bu = torch.zeros(per_u * world, dtype=torch.float64, device=dev)
    bi = torch.zeros(per_i * world, dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    (u0, u1), (i0, i1) = ru[rank], ri[rank]
    for _ in range(int(n_epochs)):
        nat.check(lib.sb2_baseline_als_pass_dev(i0, i1, nat.ptr(d_ip), nat.ptr(d_iu), nat.ptr(d_ir), nat.ptr(bu), nat.ptr(bi),
                                                mu, float(reg_i), nat.ptr(status), nat.stream()))
        if world > 1:
            dist.all_gather_into_tensor(bi, bi[rank * per_i:(rank + 1) * per_i])
        nat.check(lib.sb2_baseline_als_pass_dev(u0, u1, nat.ptr(d_up), nat.ptr(d_ui), nat.ptr(d_ur), nat.ptr(bi), nat.ptr(bu),
                                                mu, float(reg_u), nat.ptr(status), nat.stream()))
        if world > 1:
            dist.all_gather_into_tensor(bu, bu[rank * per_u:(rank + 1) * per_u])
    if world > 1:
        dist.all_reduce(status, op=dist.ReduceOp.MAX)
    if int(status.item()):
        raise ZeroDivisionError("float division")
    return bu[:n_users].cpu().numpy(), bi[:n_items].cpu().numpy()
