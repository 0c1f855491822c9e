"""z-sharded ESDF update across GPUs (BASELINE config 4; DESIGN.md "multi-GPU").

One process per GPU, `torch.distributed` for the plumbing.  The map is sharded on z: rank r
owns planes [r*nz/G, (r+1)*nz/G) of every (x,y) column, stored z-fastest like the reference
(`address = x*ny*nzl + y*nzl + z`).  The squared EDT is separable and exact in integers, so
the sweep order is free:

  1. x and y sweeps on the local z-slab           (fuelgpu_edt_xy_dev; never cross z)
  2. ONE all-to-all: z-slabs -> x-slabs of the 2-D partial (int32).  x is the slowest axis, so
     the block a rank sends to rank s is the contiguous chunk g2[s*nx/G:(s+1)*nx/G].
  3. z sweep over whole columns assembled from the G received chunks, writes metres
     (fuelgpu_edt_z_chunks_dev).  The result is x-sharded: a contiguous chunk of the full volume.
  4. optional all-gather: every rank gets the full ESDF (what the trajectory batch samples).

Replaces nothing in the reference (FUEL is single-process); it is the multi-GPU form of
SDFMap::updateESDF3d (plan_env/src/sdf_map.cpp:152-241) over the whole map.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib

EDT_INF = _lib.EDT_INF


def _gpu_xy(occ_slab, optimistic):
    """int32 squared 2-D distance of a [nx,ny,nzl] uint8 occupancy slab (device tensor)."""
    nx, ny, nzl = occ_slab.shape
    g2 = torch.empty((nx, ny, nzl), dtype=torch.int32, device=occ_slab.device)
    scratch = torch.empty((2, nx, ny, nzl), dtype=torch.int32, device=occ_slab.device)
    st = torch.cuda.current_stream(occ_slab.device).cuda_stream
    rc = _lib.lib().fuelgpu_edt_xy_dev(C.c_void_p(st), C.c_void_p(occ_slab.data_ptr()), nx, ny, nzl,
                                       _lib.ESDF_OPTIMISTIC if optimistic else 0,
                                       C.c_void_p(g2.data_ptr()), C.c_void_p(scratch.data_ptr()))
    _lib.check(rc)
    return g2


def _gpu_z(chunks, resolution):
    """[G,nxl,ny,nzl] int32 chunks -> [nxl,ny,G*nzl] float32 metres."""
    G, nxl, ny, nzl = chunks.shape
    out = torch.empty((nxl, ny, G * nzl), dtype=torch.float32, device=chunks.device)
    scratch = torch.empty((2, nxl, ny, G * nzl), dtype=torch.int32, device=chunks.device)
    st = torch.cuda.current_stream(chunks.device).cuda_stream
    rc = _lib.lib().fuelgpu_edt_z_chunks_dev(C.c_void_p(st), C.c_void_p(chunks.data_ptr()), G, nxl, ny, nzl,
                                             float(resolution), C.c_void_p(out.data_ptr()),
                                             C.c_void_p(scratch.data_ptr()))
    _lib.check(rc)
    return out


def exchange_z_to_x(g2, group=None):
    """The single exchange step: [nx,ny,nzl] (my z-slab, all x) -> [G,nx/G,ny,nzl] (my x-range, the
    z-slab of every rank).  NCCL: one all_to_all_single.  gloo (CPU tests): all_gather + slice,
    same result."""
    G = dist.get_world_size(group)
    r = dist.get_rank(group)
    nx, ny, nzl = g2.shape
    if nx % G:
        raise ValueError("nx must be divisible by the world size")
    nxl = nx // G
    send = g2.contiguous().view(G, nxl, ny, nzl)
    if dist.get_backend(group) == "nccl":
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
        return recv
    parts = [torch.empty_like(g2) for _ in range(G)]
    dist.all_gather(parts, g2.contiguous(), group=group)
    return torch.stack([p.view(G, nxl, ny, nzl)[r] for p in parts], dim=0)


class ShardedESDF:
    def __init__(self, voxel_num, resolution, optimistic=True, group=None, xy_fn=None, z_fn=None):
        """xy_fn / z_fn default to the CUDA entry points; tests inject CPU stand-ins to exercise the
        sharding logic over gloo."""
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
        self.xy_fn = xy_fn or (lambda occ: _gpu_xy(occ, self.optimistic))
        self.z_fn = z_fn or (lambda ch: _gpu_z(ch, self.res))

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

    def update(self, occ_slab):
        """occ_slab: [nx,ny,nzl] uint8 (bits0-1 tri-state, bit2 inflate).  Returns this rank's x-slab of
        distance_buffer_: [nxl,ny,nz] float32 metres (+inf where the map has no site)."""
        if tuple(occ_slab.shape) != (self.n[0], self.n[1], self.nzl):
            raise ValueError("occupancy slab must be [nx,ny,nz/G]")
        g2 = self.xy_fn(occ_slab)
        chunks = exchange_z_to_x(g2, self.group)
        return self.z_fn(chunks)

    def gather_full(self, dist_xslab):
        """all-gather the x-slabs: every rank gets the full [nx,ny,nz] ESDF (x is the slowest axis, so
        the gathered buffer IS the full volume)."""
        full = torch.empty((self.n[0], self.n[1], self.n[2]), dtype=dist_xslab.dtype, device=dist_xslab.device)
        dist.all_gather_into_tensor(full.view(-1), dist_xslab.contiguous().view(-1), group=self.group) \
            if dist.get_backend(self.group) == "nccl" else self._gather_gloo(full, dist_xslab)
        return full

    def _gather_gloo(self, full, part):
        parts = [torch.empty_like(part) for _ in range(self.G)]
        dist.all_gather(parts, part.contiguous(), group=self.group)
        for r, p in enumerate(parts):
            x0, x1 = self.x_range(r)
            full[x0:x1] = p
