"""Multi-GPU stratified SGD for SVD: one process per GPU, item blocks rotating rank -> rank.

With P ranks, user u belongs to rank u % P (local row u // P) and item i to item super-block i % P
(local row i // P).  An epoch is P sub-epochs; in sub-epoch S rank g runs the single-GPU stratified
kernel over the ratings whose user it owns and whose item lies in super-block (g + S) % P, then hands
that super-block's (qi, bi) rows to rank g - 1 and receives the next one from rank g + 1
(torch.distributed P2P: NCCL over NVLink on GPUs, gloo in the CPU tests).  The user rows never move.
This is the rank-level repetition of the CTA-level ring inside the kernel (csrc/sgd.cu).

``run_block(sb, pu, bu, qi, bi)`` is the unit of work of one sub-epoch; the product passes a closure
over sb2_svd_plan_run, the CPU tests pass a host function so that the partitioning / rotation logic is
covered without a GPU.
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


class RingSVD(object):
    """SVD fit sharded over the ranks of an initialised torch.distributed group (see module docstring).

    u, i, r: the all_ratings() COO on the host (every rank passes the same arrays); prm: _native.SgdParams.
    world == 1 degenerates to the single-GPU plan."""

    def __init__(self, dist, u, i, r, n_users, n_items, prm):
        import ctypes as C
        from . import _native as nat
        self.nat, self.C, self.dist = nat, C, dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.n_users, self.n_items, self.prm = n_users, n_items, prm
        self.f = prm.n_factors
        P, g = self.world, self.rank
        self.nu_loc = local_rows(n_users, g, P)
        self.ni_loc = [local_rows(n_items, sb, P) for sb in range(P)]
        self.n_local = 0
        self.plans = []
        for sb, (ul, il, rl) in enumerate(partition(np.asarray(u), np.asarray(i), np.asarray(r), g, P)):
            ul, il = np.ascontiguousarray(ul), np.ascontiguousarray(il)
            rl = np.ascontiguousarray(rl, dtype=np.float64)
            plan = C.c_void_p()
            nat.check(nat.lib().sb2_svd_plan_create(self.nu_loc, self.ni_loc[sb], len(rl), nat.hptr(ul), nat.hptr(il),
                                                    nat.hptr(rl), C.byref(prm), 0, C.byref(plan)))
            self.plans.append(plan)
            self.n_local += len(rl)
        torch = nat.torch_cuda()
        self.FP = nat.lib().sb2_svd_plan_stride(self.plans[0])
        dev = nat.device()
        ni_max = max(self.ni_loc)
        self.pu = torch.zeros((self.nu_loc, self.FP), dtype=torch.float32, device=dev)
        self.bu = torch.zeros((self.nu_loc,), dtype=torch.float32, device=dev)
        self.qi = torch.zeros((ni_max, self.FP), dtype=torch.float32, device=dev)
        self.bi = torch.zeros((ni_max,), dtype=torch.float32, device=dev)
        self._scratch = (torch.zeros_like(self.qi), torch.zeros_like(self.bi))
        self._exchange = (make_exchange(dist, g, P, (self.qi, self.bi), self._scratch) if P > 1 else (lambda: None))
        self.held = g

    def reset(self, pu0, qi0):
        """pu0 / qi0: full float64 host (or device) init matrices; each rank keeps its rows."""
        torch = self.nat.torch_cuda()
        P, g = self.world, self.rank
        pu0 = torch.as_tensor(pu0)[g::P].to(self.pu.device, dtype=torch.float32)
        qi0 = torch.as_tensor(qi0)[g::P].to(self.pu.device, dtype=torch.float32)
        self.pu.zero_(); self.qi.zero_(); self.bu.zero_(); self.bi.zero_()
        self.pu[:, :self.f] = pu0
        self.qi[:qi0.shape[0], :self.f] = qi0
        self.held = g

    def _run_block(self, sb):
        nat = self.nat
        lib = nat.lib()
        nat.check(lib.sb2_svd_plan_bind_dev(self.plans[sb], nat.ptr(self.pu), nat.ptr(self.qi), nat.ptr(self.bu),
                                            nat.ptr(self.bi)))
        nat.check(lib.sb2_svd_plan_run(self.plans[sb], 1, nat.stream()))

    def run(self, n_epochs):
        self.held = ring_epochs(self.rank, self.world, n_epochs, self.held, self._run_block, self._exchange)

    def gather(self):
        """Full (pu, qi, bu, bi) float64 numpy arrays on every rank."""
        torch = self.nat.torch_cuda()
        P = self.world
        assert self.held == self.rank
        ni_max = max(self.ni_loc)
        nu_max = local_rows(self.n_users, 0, P)
        outs = []
        for t, n_max, n_tot in ((self.pu, nu_max, self.n_users), (self.qi, ni_max, self.n_items),
                                (self.bu, nu_max, self.n_users), (self.bi, ni_max, self.n_items)):
            pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[:min(t.shape[0], n_max)] = t[:n_max]
            if P > 1:
                parts = [torch.zeros_like(pad) for _ in range(P)]
                self.dist.all_gather(parts, pad)
            else:
                parts = [pad]
            full = torch.zeros((n_tot,) + tuple(t.shape[1:]), dtype=torch.float64, device=t.device)
            for p in range(P):
                rows = local_rows(n_tot, p, P)
                full[p::P] = parts[p][:rows].double()
            outs.append(full[..., :self.f].cpu().numpy() if full.dim() == 2 else full.cpu().numpy())
        return tuple(outs)

    def close(self):
        for p in self.plans:
            self.nat.lib().sb2_svd_plan_destroy(p)
        self.plans = []


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
def even_ranges(n, world):
    """world contiguous ranges covering [0, n) with sizes differing by at most one."""
    return [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]


def _all_gather_rows(dist, full, ranges, rank):
    """In-place: every rank contributes rows ranges[rank] of `full` (2-D CUDA tensor) and receives all others."""
    import torch
    world = len(ranges)
    rows_max = max(hi - lo for lo, hi in ranges)
    lo, hi = ranges[rank]
    pad = torch.zeros((rows_max, full.shape[1]), dtype=full.dtype, device=full.device)
    pad[:hi - lo] = full[lo:hi]
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    for r, (a, b) in enumerate(ranges):
        if r != rank:
            full[a:b] = parts[r][:b - a]


def nmf_fit_sharded(dist, n_users, n_items, u, i, r, prm, pu0, qi0, stats=None):
    """u, i, r: all_ratings COO (host arrays, identical on every rank); prm: _native.NmfParams; pu0 / qi0: the
    rng.uniform initial factors.  Returns (pu, qi, bu, bi) as float64 numpy arrays, identical on every rank and
    bit-identical to the single-GPU fit.  stats (dict, optional) receives 'epochs_s': the wall clock of the epoch
    loop alone (device-synchronised, barrier on both sides), without upload / plan creation / download."""
    import ctypes as C
    import time
    from . import _native as nat
    torch = nat.torch_cuda()
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    lib = nat.lib()
    d_u, d_i, d_r = nat.to_dev(u, np.int32), nat.to_dev(i, np.int32), nat.to_dev(r, np.float64)
    pu = [nat.to_dev(pu0, np.float64), None]
    qi = [nat.to_dev(qi0, np.float64), None]
    pu[1], qi[1] = torch.empty_like(pu[0]), torch.empty_like(qi[0])
    bu = torch.zeros(n_users, dtype=torch.float64, device=pu[0].device)
    bi = torch.zeros(n_items, dtype=torch.float64, device=pu[0].device)
    plan = C.c_void_p()
    nat.check(lib.sb2_nmf_plan_create_dev(n_users, n_items, len(r), nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r),
                                          prm.n_factors, nat.stream(), C.byref(plan)))
    ur, ir = even_ranges(n_users, world), even_ranges(n_items, world)
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
                _all_gather_rows(dist, pu[cur], ur, rank)
                _all_gather_rows(dist, qi[cur], ir, rank)
        nat.check(lib.sb2_nmf_plan_status(plan, nat.stream()))
        if stats is not None:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            stats["epochs_s"] = time.perf_counter() - t0
    finally:
        lib.sb2_nmf_plan_destroy(plan)
    return pu[cur].cpu().numpy(), qi[cur].cpu().numpy(), bu.cpu().numpy(), bi.cpu().numpy()
