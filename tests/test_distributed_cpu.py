"""world_size-2 gloo test of the multi-GPU SVD ring's host logic (partitioning, rotation schedule, P2P
exchange).  The per-block unit of work is injected: here it is the oracle's sequential SGD on the block, so
the 2-process run must equal a single-process replay of the same schedule bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from surprise_b200 import distributed as D  # noqa: E402
from surprise_b200 import synth  # noqa: E402

F, EPOCHS, WORLD = 6, 2, 2
HP = dict(lr=.005, reg=.02)


def _data():
    d = synth.ratings(60, 40, 900, seed=4)
    u, i, r = d["train"]
    rng = np.random.RandomState(0)
    return u, i, r, d["n_users"], d["n_items"], rng.normal(0, .1, (d["n_users"], F)), rng.normal(0, .1, (d["n_items"], F))


def _block_update(ul, il, rl, pu, bu, qi, bi, mu):
    import oracle
    # oracle.svd_sgd starts biases at zero; run one epoch "continuing" by folding biases through a tiny wrapper
    import ctypes as C
    lib = oracle.lib()
    d = C.c_double
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    lib.orc_svd_sgd(C.c_int64(len(rl)), C.c_int(F), p(ul, C.c_int32), p(il, C.c_int32), p(rl, C.c_double), C.c_int(1),
                    C.c_int(1), d(mu), d(HP["lr"]), d(HP["lr"]), d(HP["lr"]), d(HP["lr"]), d(HP["reg"]), d(HP["reg"]),
                    d(HP["reg"]), d(HP["reg"]), p(pu, C.c_double), p(qi, C.c_double), p(bu, C.c_double), p(bi, C.c_double))


def _replay():
    u, i, r, nu, ni, pu0, qi0 = _data()
    mu = float(np.mean(r))
    parts = [D.partition(u, i, r, g, WORLD) for g in range(WORLD)]
    pu = [np.ascontiguousarray(pu0[g::WORLD]) for g in range(WORLD)]
    bu = [np.zeros(len(p)) for p in pu]
    qi = [np.ascontiguousarray(qi0[s::WORLD]) for s in range(WORLD)]
    bi = [np.zeros(len(q)) for q in qi]
    for _ in range(EPOCHS):
        for S in range(WORLD):
            for g in range(WORLD):
                sb = (g + S) % WORLD
                ul, il, rl = parts[g][sb]
                _block_update(np.ascontiguousarray(ul), np.ascontiguousarray(il), np.ascontiguousarray(rl), pu[g],
                              bu[g], qi[sb], bi[sb], mu)
    return pu, bu, qi, bi


def _worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    u, i, r, nu, ni, pu0, qi0 = _data()
    mu = float(np.mean(r))
    parts = D.partition(u, i, r, rank, WORLD)
    ni_max = D.local_rows(ni, 0, WORLD)
    pu = torch.from_numpy(np.ascontiguousarray(pu0[rank::WORLD]).copy())
    bu = torch.zeros(pu.shape[0], dtype=torch.float64)
    qi = torch.zeros((ni_max, F), dtype=torch.float64)
    bi = torch.zeros(ni_max, dtype=torch.float64)
    mine = qi0[rank::WORLD]
    qi[:len(mine)] = torch.from_numpy(mine.copy())
    scratch = (torch.zeros_like(qi), torch.zeros_like(bi))
    exchange = D.make_exchange(dist, rank, WORLD, (qi, bi), scratch)

    def run_block(sb):
        ul, il, rl = parts[sb]
        rows = D.local_rows(ni, sb, WORLD)
        q = np.ascontiguousarray(qi.numpy()[:rows]); b = np.ascontiguousarray(bi.numpy()[:rows])
        _block_update(np.ascontiguousarray(ul), np.ascontiguousarray(il), np.ascontiguousarray(rl), pu.numpy(),
                      bu.numpy(), q, b, mu)
        qi[:rows] = torch.from_numpy(q); bi[:rows] = torch.from_numpy(b)

    held = D.ring_epochs(rank, WORLD, EPOCHS, rank, run_block, exchange)
    assert held == rank
    rows = D.local_rows(ni, rank, WORLD)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), pu=pu.numpy(), bu=bu.numpy(), qi=qi.numpy()[:rows],
             bi=bi.numpy()[:rows])
    dist.destroy_process_group()


def test_partition_covers_every_rating_once():
    u, i, r, nu, ni, _, _ = _data()
    seen = 0
    for g in range(3):
        for sb, (ul, il, rl) in enumerate(D.partition(u, i, r, g, 3)):
            seen += len(rl)
            assert np.all(ul < D.local_rows(nu, g, 3)) and np.all(il < D.local_rows(ni, sb, 3))
    assert seen == len(r)
    assert sum(D.local_rows(nu, g, 3) for g in range(3)) == nu


def test_ring_two_ranks_equals_replay(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    pu, bu, qi, bi = _replay()
    for g in range(WORLD):
        got = np.load(os.path.join(str(tmp_path), "r%d.npz" % g))
        assert np.array_equal(got["pu"], pu[g]) and np.array_equal(got["bu"], bu[g])
        assert np.array_equal(got["qi"], qi[g]) and np.array_equal(got["bi"], bi[g])


def test_sim_row_ranges_cover_and_align():
    for n_x, world in ((1187, 2), (27000, 8), (100, 4), (128, 3), (8192, 8)):
        rows = [D.sim_row_range(n_x, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == n_x
        for (lo, hi), (lo2, _) in zip(rows, rows[1:]):
            assert hi == lo2 and lo % 256 == 0 and lo <= hi


def test_block_ranges():
    for n, w in ((10, 3), (480000, 8), (5, 8), (17700, 4), (8, 8)):
        per, rs = D.block_ranges(n, w)
        assert per * w >= n and rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(hi - lo <= per for lo, hi in rs) and all(lo == min(r * per, n) for r, (lo, hi) in enumerate(rs))


def test_sim_tri_ranges_balance_and_exchange_plan():
    """Symmetric sharded similarity build: tile-aligned contiguous ranges with equal upper-triangular tile counts;
    every send has exactly one matching receive."""
    for n_x, world in ((27000, 8), (27000, 2), (3706, 2), (1187, 3), (300, 4), (256, 2)):
        rs = D.sim_tri_ranges(n_x, world)
        assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == n_x
        for (lo, hi), (lo2, _) in zip(rs, rs[1:]):
            assert hi == lo2 and lo % 256 == 0 and lo <= hi
        nt = (n_x + 255) // 256
        cost = [sum(nt - rb for rb in range(lo // 256, (hi + 255) // 256)) for lo, hi in rs]
        assert sum(cost) == nt * (nt + 1) // 2
        if nt >= 8 * world:
            assert max(cost) <= 1.1 * min(cost)
        sends = {(k, peer, rows, cols) for k in range(world) for (peer, rows, cols) in D.sim_exchange_plan(rs, k)[0]}
        recvs = {(peer, k, rows, cols) for k in range(world) for (peer, rows, cols) in D.sim_exchange_plan(rs, k)[1]}
        assert sends == recvs


def _sim_worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=3)
    n = 1187
    rng = np.random.RandomState(1)
    full = rng.rand(n, n)
    full = (full + full.T) / 2
    ranges = D.sim_tri_ranges(n, 3)
    b, e = ranges[rank]
    block = torch.from_numpy(full[b:e].copy())
    block[:, :b] = 0.0                      # what sb2_sim_build_upper_dev leaves untouched
    D.sim_exchange(dist, torch, block, ranges, rank)
    np.save(os.path.join(out_dir, "s%d.npy" % rank), block.numpy())
    dist.destroy_process_group()


def test_sim_exchange_three_ranks(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sim_worker, args=(port, str(tmp_path)), nprocs=3, join=True)
    rng = np.random.RandomState(1)
    full = rng.rand(1187, 1187)
    full = (full + full.T) / 2
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "s%d.npy" % r)) for r in range(3)], axis=0)
    assert np.array_equal(got, full)


def _gather_worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=3)
    n, f = 10, 3                                   # 10 rows over 3 ranks: blocks of 4 (4 + 4 + 2), buffer padded to 12
    truth = torch.arange(n * f, dtype=torch.float64).reshape(n, f)
    per, ranges = D.block_ranges(n, 3)
    lo, hi = ranges[rank]
    full = torch.full((per * 3, f), -1.0, dtype=torch.float64)
    full[lo:hi] = truth[lo:hi]                     # what one NMF epoch leaves on this rank: only its own block
    dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per])   # in place, as nmf_fit_sharded does
    np.save(os.path.join(out_dir, "g%d.npy" % rank), full[:n].numpy())
    dist.destroy_process_group()


def test_nmf_in_place_block_all_gather_three_ranks(tmp_path):
    """The per-epoch exchange of the sharded NMF (uneven last block, padded buffer): every rank ends up with every row."""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(port, str(tmp_path)), nprocs=3, join=True)
    want = np.arange(30, dtype=np.float64).reshape(10, 3)
    for r in range(3):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "g%d.npy" % r)), want)


def _interleave_worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=3)
    n, f = 10, 4                                   # 10 rows over 3 ranks: 4 + 3 + 3 (uneven, padded all-gather)
    truth = torch.arange(n * f, dtype=torch.float64).reshape(n, f)
    mine = truth[rank::3].clone()
    assert mine.shape[0] == D.local_rows(n, rank, 3)
    full = D.all_gather_interleaved(dist, mine, n, rank, 3)
    vec = D.all_gather_interleaved(dist, truth[rank::3, 0].clone(), n, rank, 3)
    handles = [None] * 3
    dist.all_gather_object(handles, bytes([rank]) * 64)      # the IPC-handle exchange of RingSVD.__init__
    left, right = D.ring_neighbours(rank, 3)
    assert handles[left] == bytes([(rank - 1) % 3]) * 64 and handles[right] == bytes([(rank + 1) % 3]) * 64
    np.savez(os.path.join(out_dir, "i%d.npz" % rank), full=full.numpy(), vec=vec.numpy())
    dist.destroy_process_group()


def test_ring_gather_interleaved_three_ranks(tmp_path):
    """What RingSVD.gather does with the ranks' own rows (users g, g + P, ...): padded all-gather + interleave."""
    import torch.multiprocessing as mp
    port = 37500 + (os.getpid() % 2000)
    mp.spawn(_interleave_worker, args=(port, str(tmp_path)), nprocs=3, join=True)
    want = np.arange(40, dtype=np.float64).reshape(10, 4)
    for r in range(3):
        got = np.load(os.path.join(str(tmp_path), "i%d.npz" % r))
        assert np.array_equal(got["full"], want) and np.array_equal(got["vec"], want[:, 0])


def test_ring_schedule_model_matches_kernel_formula():
    """The kernel's schedule: in sub-epoch E rank g holds super-block (g + E) % P and hands it to g - 1; the host model
    (ring_epochs) must visit the same blocks in the same order and end at home after whole epochs."""
    for world in (1, 2, 3, 8):
        for g in range(world):
            seen = []
            held = D.ring_epochs(g, world, 2, g, seen.append, lambda: None)
            assert held == g
            assert seen == [(g + e) % world for e in range(2 * world)]
            left, right = D.ring_neighbours(g, world)
            # the block rank g holds in sub-epoch E is the one its right neighbour held in E - 1
            assert all((right + e - 1) % world == (g + e) % world for e in range(1, 2 * world))
            assert (left + 1) % world == g
