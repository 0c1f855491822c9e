#!/usr/bin/env python
"""ESDF update on a map with 1024-sample lines along y (and optionally x): the long-line case of the tile kernels.
usage: python tools/esdf_long.py [nx ny nz] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
n = tuple(a[:3]) if len(a) >= 3 else (256, 1024, 64)
reps = a[3] if len(a) > 3 else 4
g, inflate = W.random_boxes_map(n=n, seed=11, n_boxes=max(16, int(np.prod(n)) // 65536))
m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, optimistic=True)
m.occupancy_buffer_inflate_[...] = inflate
m.occupancy_tri_[...] = 1
m.upload()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
m.set_stream(st.cuda_stream)
ms = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    m.updateESDF3d()
    e1.record(st)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("esdf %s ms:" % (n,), ["%.3f" % v for v in ms], "ps/voxel %.2f" % (1e9 * min(ms) / np.prod(n)))
m.close()
