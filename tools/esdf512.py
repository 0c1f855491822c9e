#!/usr/bin/env python
"""Run the 512^3 ESDF rebuild (BASELINE config 3) a few times: the target of ncu captures.
usage: python tools/esdf512.py [V0|V1] [reps] [--nonopt]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in ("V0", "V1") else "V1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
optimistic = "--nonopt" not in sys.argv
g, inflate = W.pillar_map(variant)
m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, optimistic=optimistic)
m.occupancy_buffer_inflate_[...] = inflate
if optimistic:
    m.occupancy_tri_[...] = np.where(inflate == 1, 2, 1).astype(np.uint8)
else:
    m.setOccupancyBuffer(tristate=W.known_region(g, inflate, seed=7, n_poses=64, radius=4.5))
m.upload()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
m.set_stream(st.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for i in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    m.updateESDF3d()
    e1.record(st)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("esdf512 %s optimistic=%s ms:" % (variant, optimistic), ["%.3f" % v for v in ms],
      "GB/s(alg 5B/vox): %.1f" % (5.0 * g.nvox / (min(ms) * 1e-3) / 1e9))
m.close()
