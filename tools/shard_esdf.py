#!/usr/bin/env python
"""z-sharded ESDF across N GPUs (BASELINE config 4): parity vs the single-GPU kernel path and
device timing (max over ranks).  Launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/shard_esdf.py [nx ny nz] [--check]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402
from fuel_b200.dist import ShardedESDF  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = tuple(int(a) for a in args[:3]) if len(args) >= 3 else (1024, 1024, 256)
check = "--check" in sys.argv
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = "cuda:%d" % local
g, inflate = W.random_boxes_map(n=n, seed=11, n_boxes=4096 if n[0] >= 512 else 64)
sh = ShardedESDF(n, g.res, optimistic=True)
z0, z1 = sh.z_range()
occ = torch.from_numpy(((inflate[:, :, z0:z1] << 2) | 1).astype(np.uint8)).contiguous().to(dev)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for _ in range(2):
    part = sh.update(occ)
torch.cuda.synchronize()
dist.barrier()
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    part = sh.update(occ)
    e1.record(st)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms.append(float(t.item()))
# stage breakdown on this rank
from fuel_b200.dist import exchange_z_to_x  # noqa: E402
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record(st)
g2 = sh.xy_fn(occ)
ev[1].record(st)
ch = exchange_z_to_x(g2)
ev[2].record(st)
part = sh.z_fn(ch)
ev[3].record(st)
torch.cuda.synchronize()
if rank == 0:
    print("rank0 stages ms: xy %.3f  exchange %.3f  z %.3f" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
                                                           ev[2].elapsed_time(ev[3])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
full = sh.gather_full(part)
e1.record(st)
torch.cuda.synchronize()
tg = e0.elapsed_time(e1)
if check and rank == 0:
    # single-GPU path on the whole map (same C ABI the parity tests pin against the oracle)
    m = fuel_b200.SDFMap(n, g.res, g.origin, optimistic=True, device=local)
    m.occupancy_buffer_inflate_[...] = inflate
    m.occupancy_tri_[...] = 1
    m.upload()
    m.updateESDF3d()
    ref = m.download()
    got = full.cpu().numpy()
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.allclose(got[fin], ref[fin], rtol=1e-6, atol=0), np.abs(got[fin] - ref[fin]).max()
    print("sharded == single-GPU ESDF on %s: OK" % (n,))
    m.close()
if rank == 0:
    nvox = n[0] * n[1] * n[2]
    print("sharded ESDF %s on %d GPUs: update ms (max over ranks) %s, all-gather %.3f ms, "
          "%.1f GB/s algorithmic (5 B/voxel, whole job)" %
          (n, world, ["%.3f" % v for v in ms], tg, 5.0 * nvox / (min(ms) * 1e-3) / 1e9))
dist.destroy_process_group()
