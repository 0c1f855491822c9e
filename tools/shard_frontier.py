#!/usr/bin/env python
"""z-sharded frontier search across N GPUs: every rank knows the tri-state of ITS z planes only (the rest
of its occupancy byte is poisoned), receives one halo plane from each neighbour, sweeps its planes, and clusters the merged
candidate list.  Checked against the whole search on a map that holds everything.  Launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
      tools/shard_frontier.py [nx ny nz]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402
from fuel_b200.dist import exchange_halo_planes, search_frontiers_sharded  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = tuple(int(a) for a in args[:3]) if len(args) >= 3 else (256, 256, 64)
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g, inflate = W.random_boxes_map(n=n, seed=11, n_boxes=max(32, int(np.prod(n)) // 32768), ground_idx=3)
tri = W.known_region(g, inflate, seed=7, n_poses=24, radius=3.5)
nzl = n[2] // world
z_lo, z_hi = rank * nzl, (rank + 1) * nzl - 1 if rank < world - 1 else n[2] - 1
# this rank's knowledge: its own planes; everything else poisoned (OCCUPIED would never be frontier nor unknown)
mine = np.full(n, W.OCCUPIED, dtype=np.uint8)
mine[:, :, z_lo:z_hi + 1] = tri[:, :, z_lo:z_hi + 1]
m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, device=local)
m.occupancy_buffer_inflate_[...] = inflate
m.setOccupancyBuffer(tristate=mine)
m.upload()
exchange_halo_planes(m, z_lo, z_hi)
env = fuel_b200.EDTEnvironment()
env.setMap(m)
ff = fuel_b200.FrontierFinder(env)
t0 = time.perf_counter()
got = search_frontiers_sharded(ff, g.origin, g.map_max, z_lo, z_hi)
torch.cuda.synchronize()
t1 = time.perf_counter()
# the whole search on a map that holds everything (every rank does it: the comparison is local)
m2 = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, device=local)
m2.occupancy_buffer_inflate_[...] = inflate
m2.setOccupancyBuffer(tristate=tri)
m2.upload()
env2 = fuel_b200.EDTEnvironment()
env2.setMap(m2)
ff2 = fuel_b200.FrontierFinder(env2)
ref = ff2.search_box(g.origin, g.map_max)
ok = len(got) == len(ref) and len(ref) > 0
if ok:
    for a, b in zip(got, ref):
        ok = ok and np.array_equal(a.cells_addr_, b.cells_addr_) and np.array_equal(a.average_, b.average_) and \
            np.array_equal(a.filtered_cells_, b.filtered_cells_)
ok = ok and np.array_equal(ff.download_flags(), ff2.download_flags())
flag = torch.tensor([1 if ok else 0], device="cuda:%d" % local)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded frontier search %s on %d GPUs: %d clusters, %d cells, wall %.2f ms (sweep of %d planes per rank + gather + "
          "clustering)" % (n, world, len(ref), sum(c.cells_addr_.size for c in ref), 1e3 * (t1 - t0), nzl))
    print("OK" if int(flag.item()) == 1 else "MISMATCH")
m.close()
m2.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
