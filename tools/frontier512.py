#!/usr/bin/env python
"""Frontier sweep + clustering + split on the 512^3 pillar map (BASELINE config 3): device timing,
size-independent properties, and (with --oracle) bit-exact comparison with the CPU oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402

variant = "V1"
g, inflate = W.pillar_map(variant)
t0 = time.time()
tri = W.known_region(g, inflate, seed=7, n_poses=64, radius=4.5)
print("known region built in %.1fs: unknown/free/occ =" % (time.time() - t0), np.bincount(tri.ravel()))
m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max)
m.occupancy_buffer_inflate_[...] = inflate
m.setOccupancyBuffer(tristate=tri)
m.upload()
env = fuel_b200.EDTEnvironment()
env.setMap(m)
ff = fuel_b200.FrontierFinder(env)
for it in range(3):
    ff.reset_flags()
    m.synchronize()
    t0 = time.perf_counter()
    out = ff.search_box(g.origin, g.map_max)
    t1 = time.perf_counter()
    ncell = sum(c.cells_addr_.size for c in out)
    print("frontier 512^3: wall %.2f ms, device stage %.2f ms, %d clusters, %d cells" %
          (1e3 * (t1 - t0), m.last_timing()["frontier"], len(out), ncell))
fl = ff.download_flags()
# properties: cells disjoint, all flagged, second sweep finds nothing new
allc = np.concatenate([c.cells_addr_ for c in out])
assert np.unique(allc).size == allc.size
assert np.all(fl.ravel()[allc] == 1)
again = ff.search_box(g.origin, g.map_max)
assert again == [] and np.array_equal(ff.download_flags(), fl)
print("properties ok (disjoint clusters, flagged, idempotent)")
if "--oracle" in sys.argv:
    import oracle
    og = oracle.make_grid(g.n, g.res, g.origin, g.box_min, g.box_max)
    ofl = np.zeros(g.n, dtype=np.int8)
    t0 = time.time()
    ref = oracle.frontier_search(og, tri, ofl, g.origin, g.map_max, oracle.frontier_params(cell_order=1))
    print("oracle: %.1fs, %d clusters" % (time.time() - t0, len(ref)))
    assert len(ref) == len(out)
    bad = 0
    for a, b in zip(out, ref):
        if not np.array_equal(a.cells_addr_, b["addr"]):
            bad += 1
    print("clusters differing from the oracle (canonical cell order): %d of %d" % (bad, len(ref)))
    assert np.array_equal(fl, ofl)
    assert bad == 0
m.close()
