#!/usr/bin/env python
"""Timing of the device solver loop vs history length / eval count (office map, B=1024, N=20)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fuel_b200  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402
from fuel_b200._lib import FuelSolveParams  # noqa: E402

g, inflate = W.office_map()
tri = W.office_known(g, inflate)
m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max, optimistic=True)
m.occupancy_buffer_inflate_[...] = inflate
m.setOccupancyBuffer(tristate=tri)
m.upload()
m.updateESDF3d()
env = fuel_b200.EDTEnvironment()
env.setMap(m)
opt = fuel_b200.BsplineOptimizer()
opt.setEnvironment(env)
B, N = 1024, 20
tr = W.make_trajectories(g, inflate, B=B, n_pts=N)
mask = opt.NORMAL_PHASE | opt.MINTIME
x0 = W.pack_x(tr["ctrl"], tr["dt"])
tcs = opt.traj_consts_from_arrays(tr["pt_dist"], tr["dt"], tr["start"], tr["end_pos"])
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
m.set_stream(st.cuda_stream)
d_tc = torch.from_numpy(np.frombuffer(tcs, dtype=np.uint8).copy()).cuda()
d_x0 = torch.from_numpy(x0).cuda()
d_x = torch.empty_like(d_x0)
d_f = torch.empty(B, dtype=torch.float64, device="cuda")
d_n = torch.empty(B, dtype=torch.int32, device="cuda")
L = fuel_b200.lib()
for mm, K in ((6, 64), (3, 64), (1, 64), (6, 16), (6, 128), (8, 64)):
    sp = FuelSolveParams()
    sp.max_eval, sp.lbfgs_m, sp.xtol_rel = K, mm, 0.0
    ms = []
    for it in range(4):
        d_x.copy_(d_x0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = L.fuelgpu_bspline_optimize_batch_dev(m.handle, B, N, mask, C.byref(opt.params_), C.c_void_p(d_tc.data_ptr()),
                                                  C.byref(sp), C.c_void_p(d_x.data_ptr()), C.c_void_p(d_f.data_ptr()),
                                                  C.c_void_p(d_n.data_ptr()))
        assert rc == 0
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print("m=%d K=%3d: %.3f ms  (%.2f us per eval-step)  mean f_best %.3f  mean n_eval %.1f" %
          (mm, K, min(ms), 1e3 * min(ms) / K, float(d_f.mean()), float(d_n.float().mean())))
# plain cost kernel for reference
d_g = torch.empty_like(d_x0)
ms = []
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    L.fuelgpu_bspline_cost_batch_dev(m.handle, B, N, mask, C.byref(opt.params_), C.c_void_p(d_tc.data_ptr()),
                                     C.c_void_p(d_x0.data_ptr()), C.c_void_p(d_f.data_ptr()), C.c_void_p(d_g.data_ptr()))
    e1.record(st)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("cost_batch (faithful) one evaluation of the batch: %.1f us" % (1e3 * min(ms)))
