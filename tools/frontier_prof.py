#!/usr/bin/env python
"""Per-phase timing of the single-cluster frontier kernel (needs a FUEL_PROF=1 build)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402

g, inflate = W.office_map()
tri = W.office_known(g, inflate)
m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, optimistic=True)
m.occupancy_buffer_inflate_[...] = inflate
m.setOccupancyBuffer(tristate=tri)
m.upload()
env = fuel_b200.EDTEnvironment()
env.setMap(m)
ff = fuel_b200.FrontierFinder(env)
import time
for it in range(5):
    ff.reset_flags()
    m.synchronize()
    t0 = time.perf_counter()
    out = ff.search_box(g.origin, g.map_max)
    t1 = time.perf_counter()
    print("search_box wall %.1f us, %d clusters, stage %.1f us" % (1e6 * (t1 - t0), len(out), 1e3 * m.last_timing()["frontier"]))
L = fuel_b200.lib()
if hasattr(C.CDLL(fuel_b200._lib.SO), "fuelgpu_debug_frontier_prof"):
    fn = C.CDLL(fuel_b200._lib.SO).fuelgpu_debug_frontier_prof
    buf = (C.c_longlong * 256)()
    fn(m.handle, buf, 256)
    t = np.array(buf[:], dtype=np.int64)
    t = t[t > 0]
    d = np.diff(t)
    print("n stamps", len(t), "total us %.1f" % ((t[-1] - t[0]) / 1e3))
    print("phase deltas (us):", " ".join("%.1f" % (v / 1e3) for v in d))
