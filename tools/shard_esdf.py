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
tm = sh.last_timing()
if rank == 0:
    print("sharded ESDF %s on %d GPUs (%s): ms (max over ranks) %s  best %.3f  stages(rank0) %s  sent/rank %.1f MB"
          % (n, world, "peer-memory stores" if sh.uses_peer_memory() else "ncclSend/Recv",
             ["%.3f" % v for v in ms], min(ms), {k: round(v, 3) for k, v in tm.items()}, sh.bytes_exchanged() / 1e6))
if check:
    full = sh.gather_full(part)
    torch.cuda.synchronize()
    m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, optimistic=True, device=local)
    m.occupancy_buffer_inflate_[...] = inflate
    m.occupancy_tri_[...] = 1
    m.upload()
    m.updateESDF3d()
    ref = m.download().copy()
    m.close()
    got = full.cpu().numpy()
    ok = np.array_equal(np.isinf(got), np.isinf(ref)) and np.allclose(got[np.isfinite(ref)], ref[np.isfinite(ref)],
                                                                      rtol=1e-6, atol=0)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("OK" if int(flag.item()) == 1 else "MISMATCH")
    if int(flag.item()) != 1:
        sh.close()
        dist.destroy_process_group()
        sys.exit(1)
sh.close()
dist.destroy_process_group()
