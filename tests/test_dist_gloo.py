"""world_size-2 gloo test (CPU) of the z-sharded ESDF decomposition in fuel_b200/dist.py: slab shapes, the
occupancy exchange (z-slabs -> x-slabs), the exchange of the 2-D partial (x-slabs -> z-slabs), and the final
all-gather.  The device stages are replaced by CPU stand-ins (exact integer EDT passes in numpy) so that the
N>1 host logic is exercised without a GPU; the result must equal the oracle's full-map ESDF."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INF = 0x3FFFFFFF


def edt1d_sq(f):
    """exact 1-D squared-distance lower envelope of f (INF = no site) along axis 0, brute force"""
    n = f.shape[0]
    q = np.arange(n)
    d2 = (q[:, None] - q[None, :]) ** 2  # [q, v]
    fin = f < INF
    big = np.where(fin, f, np.int64(1) << 40).astype(np.int64)
    out = (d2[:, :, None] + big[None, :, :].reshape(1, n, -1)).min(axis=1)
    return np.where(out >= (np.int64(1) << 39), INF, out).reshape(f.shape)


def cpu_zy(occ):
    """z then y sweep on an x-slab [nxl, ny, nz]: squared 2-D distance in every x plane"""
    o = occ.numpy()
    site = (o & 4) != 0  # optimistic: inflate bit
    f = np.where(site, 0, INF).astype(np.int64)
    nx, ny, nz = f.shape
    g = edt1d_sq(np.moveaxis(f, 2, 0).reshape(nz, -1)).reshape(nz, nx, ny)
    g = np.moveaxis(g, 0, 2)
    g = edt1d_sq(np.moveaxis(g, 1, 0).reshape(ny, -1)).reshape(ny, nx, nz)
    g = np.moveaxis(g, 0, 1)
    return torch.from_numpy(g.astype(np.int32))


def cpu_x(part, res):
    """x sweep on a z-slab [nx, ny, nzl] of the 2-D partial -> metres"""
    c = part.numpy().astype(np.int64)
    nx, ny, nzl = c.shape
    g = edt1d_sq(c.reshape(nx, -1)).reshape(nx, ny, nzl)
    out = np.where(g >= INF, np.inf, res * np.sqrt(g.astype(np.float64)))
    return torch.from_numpy(out.astype(np.float32))


def worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fuel_b200.dist import ShardedESDF
    rng = np.random.default_rng(5)  # same map on every rank
    inflate = (rng.random(n) < 0.01).astype(np.uint8)
    occ = torch.from_numpy((inflate << 2) | 1)
    sh = ShardedESDF(n, 0.1, optimistic=True, stage_fns=(cpu_zy, lambda p: cpu_x(p, 0.1)))
    assert sh.z_range() == (rank * n[2] // world, (rank + 1) * n[2] // world)
    slab = sh.shard_occupancy(occ)
    assert slab.shape == (n[0], n[1], n[2] // world) and slab.is_contiguous()
    part = sh.update(slab)
    assert part.shape == (n[0], n[1], n[2] // world)
    full = sh.gather_full(part)
    ret[rank] = full.numpy()
    dist.destroy_process_group()


def test_z_sharded_esdf_two_ranks(orc):
    n = (12, 10, 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(2, port, n, ret), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    inflate = (rng.random(n) < 0.01).astype(np.int8)
    g = orc.make_grid(n, 0.1, (0, 0, 0))
    ref = orc.update_esdf3d(g, inflate, None, [0, 0, 0], np.array(n) - 1, True, False)
    for r in (0, 1):
        got = ret[r]
        assert got.shape == n
        assert np.allclose(got, ref, rtol=1e-6), "rank %d" % r
    assert np.array_equal(ret[0], ret[1])


def split_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fuel_b200.dist import split_batch_run, split_counts
    B = 11  # uneven split
    x = np.arange(B * 3, dtype=np.float64).reshape(B, 3)
    cnt, off = split_counts(B, world)
    assert sum(cnt) == B and max(cnt) - min(cnt) <= 1 and off[-1] == B
    calls = []

    def fn(lo, hi):
        calls.append((lo, hi))
        return x[lo:hi] * 2.0, x[lo:hi, 0] + 1.0, np.full(hi - lo, rank, dtype=np.int32)

    a, b, c = split_batch_run(fn, B)
    assert calls == [(int(off[rank]), int(off[rank + 1]))]
    ret[rank] = (a, b, c)
    dist.destroy_process_group()


def test_batch_split_over_ranks():
    """One planner on G GPUs: the trajectory batch is cut evenly, each rank runs its share, every rank gets the whole
    result in the original order (fuel_b200.dist.split_batch_run)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(split_worker, args=(2, port, ret), nprocs=2, join=True)
    x = np.arange(33, dtype=np.float64).reshape(11, 3)
    for r in (0, 1):
        a, b, c = ret[r]
        assert np.array_equal(a, x * 2.0) and np.array_equal(b, x[:, 0] + 1.0)
        assert list(c) == [0] * 6 + [1] * 5


def cand_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fuel_b200.dist import gather_candidates
    nz = 8
    # rank r owns z in [4r, 4r+4): addresses with (addr % nz) in that range; different counts per rank, one rank may be empty
    rng = np.random.default_rng(3)
    allc = np.sort(rng.choice(200 * nz, size=57, replace=False)).astype(np.int32)
    cls_all = (1 + (allc % 2)).astype(np.uint8)
    if world == 2:
        mine = (allc % nz) // 4 == rank
    addr, cls = gather_candidates(allc[mine], cls_all[mine])
    ret[rank] = (addr, cls)
    dist.destroy_process_group()


def test_candidate_gather_and_merge():
    """The z-sharded frontier sweep: per-rank candidate lists of different lengths are all-gathered and merged into ONE
    list ascending by address (z slabs interleave in a z-fastest address); overlapping z ranges are refused."""
    from fuel_b200.dist import merge_candidates
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(cand_worker, args=(2, port, ret), nprocs=2, join=True)
    rng = np.random.default_rng(3)
    allc = np.sort(rng.choice(200 * 8, size=57, replace=False)).astype(np.int32)
    for r in (0, 1):
        addr, cls = ret[r]
        assert np.array_equal(addr, allc) and np.array_equal(cls, (1 + (allc % 2)).astype(np.uint8))
    import pytest
    with pytest.raises(ValueError):
        merge_candidates([(np.array([1, 5]), np.array([1, 1])), (np.array([5, 9]), np.array([1, 2]))])
