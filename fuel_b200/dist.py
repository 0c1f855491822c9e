"""z-sharded ESDF update across GPUs (BASELINE config 4; DESIGN.md "multi-GPU").

One process per GPU.  The map is sharded on z: rank r owns planes [r*nz/G, (r+1)*nz/G) of every (x,y)
column, stored z-fastest like the reference (`address = x*ny*nzl + y*nzl + z`, sdf_map.h:145-147).  The
squared EDT is separable and exact in integers, so every sweep may run where its lines are whole:

  1. all-to-all of the occupancy byte: z-slabs -> x-slabs (rank r gets whole z lines of its x range)
  2. z records + zy tiles on the x-slab
  3. THE exchange of the 2-D partial (int32): x-slabs -> z-slabs, round by round beside step 2
  4. x tiles on the own z-slab -> this rank's z-slab of distance_buffer_ (metres)
  5. optional all-gather: every rank gets the full ESDF (what a trajectory batch split over ranks samples)

The product path is `fuelgpu_sharded_esdf_*` of the C ABI (fuel_b200/csrc/sharded.cu): NCCL is called from
inside the library, this module only bootstraps the communicator (the 128-byte NCCL id travels over
torch.distributed) and wraps device tensors.  `stage_fns=(zy_fn, x_fn)` replaces the device stages by
caller-supplied ones and the NCCL exchanges by torch.distributed collectives, so that the sharding logic
(slab shapes, both exchanges, the all-gather) is exercised by a world-size-2 gloo test on CPU.

Replaces nothing in the reference (FUEL is single-process); it is the multi-GPU form of
SDFMap::updateESDF3d (plan_env/src/sdf_map.cpp:152-241) over the whole map.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

EDT_INF = _lib.EDT_INF


class Comm:
    """FuelComm (an NCCL communicator owned by libfuelgpu) bootstrapped over a torch.distributed group."""

    def __init__(self, device, group=None):
        self.group = group
        self.nranks = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = int(device)
        L = _lib.lib()
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            _lib.check(L.fuelgpu_comm_get_unique_id(_lib.ptr(uid)))
        t = torch.from_numpy(uid)
        if dist.get_backend(group) == "nccl":
            t = t.to("cuda:%d" % self.device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = t.cpu().numpy().copy()
        h = C.c_void_p()
        _lib.check(L.fuelgpu_comm_init(self.nranks, self.rank, _lib.ptr(uid), self.device, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            _lib.lib().fuelgpu_comm_destroy(self.handle)
            self.handle = None


def _all_to_all_blocks(blocks, group=None):
    """blocks[d] goes to rank d; returns the list received (index = source rank).  all_to_all on NCCL,
    all_gather + pick on gloo (CPU tests)."""
    G = dist.get_world_size(group)
    r = dist.get_rank(group)
    blocks = [b.contiguous() for b in blocks]
    if dist.get_backend(group) == "nccl":
        out = [torch.empty_like(blocks[s]) for s in range(G)]
        dist.all_to_all(out, blocks, group=group)
        return out
    packed = torch.stack(blocks, dim=0)
    parts = [torch.empty_like(packed) for _ in range(G)]
    dist.all_gather(parts, packed, group=group)
    return [parts[s][r] for s in range(G)]


class ShardedESDF:
    def __init__(self, voxel_num, resolution, optimistic=True, group=None, device=None, stage_fns=None):
        """stage_fns = (zy_fn, x_fn): zy_fn(occ[nxl,ny,nz] uint8) -> int32 squared 2-D distance [nxl,ny,nz]
        (EDT_INF = none); x_fn(partial[nx,ny,nzl] int32) -> float32 metres.  Default: the CUDA path."""
        self.n = tuple(int(v) for v in voxel_num)
        self.res = float(resolution)
        self.optimistic = bool(optimistic)
        self.group = group
        self.G = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.n[2] % self.G or self.n[0] % self.G:
            raise ValueError("nx and nz must be divisible by the world size")
        self.nzl = self.n[2] // self.G
        self.nxl = self.n[0] // self.G
        self.stage_fns = stage_fns
        self.handle = None
        self.comm = None
        if stage_fns is None:
            if device is None:
                device = torch.cuda.current_device()
            self.device = int(device)
            self.comm = Comm(self.device, group)
            h = C.c_void_p()
            n3 = (C.c_int32 * 3)(*self.n)
            _lib.check(_lib.lib().fuelgpu_sharded_esdf_create(self.comm.handle, n3, self.res, C.byref(h)))
            self.handle = h

    def close(self):
        if self.handle:
            _lib.lib().fuelgpu_sharded_esdf_destroy(self.handle)
            self.handle = None
        if self.comm:
            self.comm.close()
            self.comm = None

    def z_range(self, rank=None):
        r = self.rank if rank is None else rank
        return r * self.nzl, (r + 1) * self.nzl

    def x_range(self, rank=None):
        r = self.rank if rank is None else rank
        return r * self.nxl, (r + 1) * self.nxl

    def shard_occupancy(self, occ_full):
        """[nx,ny,nz] occupancy byte (any rank-local copy) -> this rank's contiguous z-slab."""
        z0, z1 = self.z_range()
        return occ_full[:, :, z0:z1].contiguous()

    def update(self, occ_slab, out=None):
        """occ_slab: [nx,ny,nzl] uint8 (bits0-1 tri-state, bit2 inflate).  Returns this rank's z-slab of
        distance_buffer_: [nx,ny,nzl] float32 metres (+inf where the map has no site)."""
        if tuple(occ_slab.shape) != (self.n[0], self.n[1], self.nzl):
            raise ValueError("occupancy slab must be [nx,ny,nz/G]")
        if self.stage_fns is not None:
            return self._update_host(occ_slab)
        if out is None:
            out = torch.empty((self.n[0], self.n[1], self.nzl), dtype=torch.float32, device=occ_slab.device)
        st = torch.cuda.current_stream(occ_slab.device).cuda_stream
        rc = _lib.lib().fuelgpu_sharded_esdf_update(self.handle, C.c_void_p(st), C.c_void_p(occ_slab.data_ptr()),
                                                    _lib.ESDF_OPTIMISTIC if self.optimistic else 0,
                                                    C.c_void_p(out.data_ptr()))
        _lib.check(rc)
        return out

    def last_timing(self):
        """device ms of the last update on this rank: occupancy exchange, records + zy tiles (partial exchange
        beside them), wait for the last rounds, x tiles, total"""
        ms = (C.c_float * 5)()
        _lib.check(_lib.lib().fuelgpu_sharded_esdf_last_timing(self.handle, ms))
        return dict(zip(("occ_exchange", "zy", "exchange_wait", "x", "total"), [float(v) for v in ms]))

    def bytes_exchanged(self):
        return int(_lib.lib().fuelgpu_sharded_esdf_bytes_exchanged(self.handle))

    def uses_peer_memory(self):
        """True when the zy tile kernels store the partial straight into the peers' receive buffers (CUDA IPC over
        NVLink) instead of handing it to ncclSend/ncclRecv."""
        return bool(_lib.lib().fuelgpu_sharded_esdf_uses_peer_memory(self.handle))

    # ---- host-orchestrated twin (CPU tests): same decomposition, torch.distributed collectives ----
    def _update_host(self, occ_slab):
        zy_fn, x_fn = self.stage_fns
        G = self.G
        nx, ny, nz = self.n
        # 1. occupancy z-slabs -> x-slabs
        got = _all_to_all_blocks([occ_slab[d * self.nxl:(d + 1) * self.nxl] for d in range(G)], self.group)
        occ_x = torch.cat(got, dim=2)  # [nxl, ny, nz]: z lines assembled from the G chunks
        assert tuple(occ_x.shape) == (self.nxl, ny, nz)
        # 2. zy stage on the x-slab
        part = zy_fn(occ_x)
        # 3. partial x-slabs -> z-slabs
        got = _all_to_all_blocks([part[:, :, d * self.nzl:(d + 1) * self.nzl] for d in range(G)], self.group)
        part_z = torch.cat(got, dim=0)  # [nx, ny, nzl]
        assert tuple(part_z.shape) == (nx, ny, self.nzl)
        # 4. x stage on the z-slab
        return x_fn(part_z)

    def gather_into_map(self, dist_zslab, sdf_map):
        """all-gather the z-slabs and install the full field as `sdf_map`'s distance_buffer_ on this rank's device: the
        one ESDF broadcast of a planner that splits its trajectory batch over the ranks (then no further collective)."""
        nx, ny, nz = self.n
        buf = torch.empty((self.G, nx, ny, self.nzl), dtype=torch.float32, device=dist_zslab.device)
        st = torch.cuda.current_stream(dist_zslab.device).cuda_stream
        L = _lib.lib()
        _lib.check(L.fuelgpu_sharded_esdf_allgather(self.handle, C.c_void_p(st), C.c_void_p(dist_zslab.contiguous().data_ptr()),
                                                    C.c_void_p(buf.data_ptr())))
        torch.cuda.current_stream(dist_zslab.device).synchronize()  # the map's own stream reads buf next
        _lib.check(L.fuelgpu_esdf_set_from_slabs_dev(sdf_map.handle, C.c_void_p(buf.data_ptr()), self.G), sdf_map.handle)
        sdf_map.synchronize()

    def gather_full(self, dist_zslab):
        """all-gather the z-slabs: every rank gets the full [nx,ny,nz] ESDF."""
        nx, ny, nz = self.n
        if self.stage_fns is None:
            buf = torch.empty((self.G, nx, ny, self.nzl), dtype=torch.float32, device=dist_zslab.device)
            st = torch.cuda.current_stream(dist_zslab.device).cuda_stream
            _lib.check(_lib.lib().fuelgpu_sharded_esdf_allgather(self.handle, C.c_void_p(st),
                                                                 C.c_void_p(dist_zslab.contiguous().data_ptr()),
                                                                 C.c_void_p(buf.data_ptr())))
            return buf.permute(1, 2, 0, 3).reshape(nx, ny, nz)
        parts = [torch.empty_like(dist_zslab) for _ in range(self.G)]
        dist.all_gather(parts, dist_zslab.contiguous(), group=self.group)
        return torch.cat(parts, dim=2)


def split_counts(B, G):
    """even split of a batch of B over G ranks: the first B % G ranks take one more"""
    base, extra = divmod(int(B), int(G))
    cnt = [base + (1 if r < extra else 0) for r in range(G)]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return cnt, off


def split_batch_run(fn, B, group=None, device=None):
    """One planner, G GPUs: rank r runs fn(lo, hi) on its share [lo, hi) of a batch of B independent
    trajectories -- fn returns a tuple of numpy arrays whose first axis is hi - lo -- and every rank gets the results
    of the whole batch in the original order.  The only communication is this gather of the (small) results; the ESDF
    must already be on every rank (ShardedESDF.gather_into_map)."""
    G = dist.get_world_size(group)
    r = dist.get_rank(group)
    cnt, off = split_counts(B, G)
    mine = fn(int(off[r]), int(off[r + 1]))
    out = []
    nccl = dist.get_backend(group) == "nccl"
    mx = max(cnt)
    for a in mine:
        a = np.ascontiguousarray(a)
        assert a.shape[0] == cnt[r]
        pad = np.zeros((mx,) + a.shape[1:], dtype=a.dtype)
        pad[:cnt[r]] = a
        t = torch.from_numpy(pad)
        if nccl:
            t = t.to("cuda:%d" % (torch.cuda.current_device() if device is None else device))
        parts = [torch.empty_like(t) for _ in range(G)]
        dist.all_gather(parts, t, group=group)
        out.append(np.concatenate([parts[k][:cnt[k]].cpu().numpy() for k in range(G)], axis=0))
    return tuple(out)


def optimize_batch_split(opt, x, traj_consts, n_pts, cost_function, max_eval, group=None, **kw):
    """BsplineOptimizer.optimizeBatch of a whole batch spread over the ranks of `group` (each rank's optimizer must sit
    on a map holding the full ESDF).  Returns (x_best, f_best, n_eval) of the whole batch on every rank."""
    from ._lib import FuelTrajConst
    x = np.ascontiguousarray(x, dtype=np.float64)

    def run(lo, hi):
        if hi == lo:
            return (np.empty((0, x.shape[1])), np.empty(0), np.empty(0, dtype=np.int32))
        tc = (FuelTrajConst * (hi - lo)).from_buffer(traj_consts, lo * C.sizeof(FuelTrajConst))
        xb, fb, ne = opt.optimizeBatch(x[lo:hi], tc, n_pts, cost_function, max_eval, **kw)
        return xb.copy(), fb.copy(), ne.copy()

    return split_batch_run(run, x.shape[0], group)


def merge_candidates(parts):
    """[(addr, cls), ...] of all ranks -> one list ascending by address (z-slabs interleave in a z-fastest address)"""
    addr = np.concatenate([np.asarray(a, dtype=np.int32) for a, _ in parts])
    cls = np.concatenate([np.asarray(c, dtype=np.uint8) for _, c in parts])
    order = np.argsort(addr, kind="stable")
    addr, cls = addr[order], cls[order]
    if addr.size > 1 and np.any(addr[1:] == addr[:-1]):
        raise ValueError("a voxel was swept by two ranks: the z ranges overlap")
    return addr, cls


def gather_candidates(addr, cls, group=None, device=None):
    """all-gather of the per-rank candidate lists (variable length, KBs) -> the merged list on every rank"""
    G = dist.get_world_size(group)
    nccl = dist.get_backend(group) == "nccl"
    dev = "cuda:%d" % (torch.cuda.current_device() if device is None else device) if nccl else "cpu"
    n = torch.tensor([addr.size], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(G)]
    dist.all_gather(ns, n, group=group)
    ns = [int(v.item()) for v in ns]
    mx = max(max(ns), 1)
    pack = np.zeros((mx, 2), dtype=np.int32)
    pack[:addr.size, 0] = addr
    pack[:addr.size, 1] = cls
    t = torch.from_numpy(pack).to(dev)
    parts = [torch.empty_like(t) for _ in range(G)]
    dist.all_gather(parts, t, group=group)
    out = []
    for k in range(G):
        a = parts[k][:ns[k]].cpu().numpy()
        out.append((a[:, 0].copy(), a[:, 1].astype(np.uint8)))
    return merge_candidates(out)


def exchange_halo_planes(sdf_map, z_lo, z_hi, group=None):
    """One plane of the occupancy byte each way (the 6-neighbour unknown test reaches +-1): every rank
    publishes its two boundary planes, then installs plane z_lo-1 from the rank below and z_hi+1 from the rank above.
    Device buffers; NCCL all_gather of 2*nx*ny bytes per rank."""
    G = dist.get_world_size(group)
    r = dist.get_rank(group)
    nx, ny, nz = sdf_map.shape
    dev = "cuda:%d" % sdf_map.device
    L = _lib.lib()
    mine = torch.empty((2, nx, ny), dtype=torch.uint8, device=dev)
    _lib.check(L.fuelgpu_map_occupancy_plane_dev(sdf_map.handle, int(z_lo), C.c_void_p(mine[0].data_ptr()), 0), sdf_map.handle)
    _lib.check(L.fuelgpu_map_occupancy_plane_dev(sdf_map.handle, int(z_hi), C.c_void_p(mine[1].data_ptr()), 0), sdf_map.handle)
    sdf_map.synchronize()
    parts = [torch.empty_like(mine) for _ in range(G)]
    dist.all_gather(parts, mine, group=group)
    torch.cuda.synchronize(sdf_map.device)
    if r > 0:
        _lib.check(L.fuelgpu_map_occupancy_plane_dev(sdf_map.handle, int(z_lo) - 1, C.c_void_p(parts[r - 1][1].data_ptr()), 1),
                   sdf_map.handle)
    if r < G - 1:
        _lib.check(L.fuelgpu_map_occupancy_plane_dev(sdf_map.handle, int(z_hi) + 1, C.c_void_p(parts[r + 1][0].data_ptr()), 1),
                   sdf_map.handle)
    sdf_map.synchronize()


def search_frontiers_sharded(ff, update_min, update_max, z_lo, z_hi, group=None):
    """The z-sharded frontier search: this rank sweeps its planes, the candidate cells of all ranks are
    gathered and merged, every rank clusters the full list.  Returns the same Frontier list as ff.search_box on one GPU
    holding the whole map.  The caller has installed the halo planes (exchange_halo_planes)."""
    addr, cls = ff.candidates(update_min, update_max, z_lo, z_hi)
    addr, cls = gather_candidates(addr, cls, group, device=ff.edt_env_.sdf_map_.device)
    return ff.search_from_candidates(update_min, update_max, addr, cls)
