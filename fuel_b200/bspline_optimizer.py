"""Host-side mirror of fast_planner::BsplineOptimizer over the C ABI.

Mirrors bspline_opt/include/bspline_opt/bspline_optimizer.h:20-145 and
bspline_opt/src/bspline_optimizer.cpp (file:line under /root/reference/fuel_planner/).
`combineCost` / `costFunction` evaluate on the device (fuelgpu_bspline_cost_batch); the
batched forms take B trajectories at once.  NLopt is a third-party dependency of the
reference that is not available here; `optimize()` drives the device-side projected
L-BFGS of fuelgpu_bspline_optimize_batch instead (iterate-level parity with NLopt is
unpinned).
"""
import ctypes as C

import numpy as np

from ._lib import (MAX_PTS, SOLVE_EXACT_EVALS, FuelOptParams, FuelSolveParams, FuelTrajConst, check, lib, ptr)

COST_FAST_EVAL = 1 << 30


class BsplineOptimizer:
    SMOOTHNESS = 1 << 0  # bspline_optimizer.cpp:10-18
    DISTANCE = 1 << 1
    FEASIBILITY = 1 << 2
    START = 1 << 3
    END = 1 << 4
    GUIDE = 1 << 5
    WAYPOINTS = 1 << 6
    VIEWCONS = 1 << 7
    MINTIME = 1 << 8
    GUIDE_PHASE = SMOOTHNESS | GUIDE | START | END  # :20-21
    NORMAL_PHASE = SMOOTHNESS | DISTANCE | FEASIBILITY | START | END  # :22-23

    def __init__(self):
        self.edt_environment_ = None
        self.setParam()
        self.start_state_ = []
        self.end_state_ = []
        self.guide_pts_ = []
        self.waypoints_ = []
        self.waypt_idx_ = []
        self.time_lb_ = -1.0
        self.cost_function_ = 0
        self.best_variable_ = None
        self.min_cost_ = None
        self.iter_num_ = 0

    def setParam(self, ld_smooth=20.0, ld_dist=10.0, ld_feasi=2.0, ld_start=100.0, ld_end=0.5, ld_guide=1.5,
                 ld_waypt=0.3, ld_view=0.0, ld_time=1.0, dist0=0.7, max_vel=2.0, max_acc=2.0,
                 bspline_degree=3, max_iteration_num=(2, 2000, 200, 200),
                 max_iteration_time=(0.0001, 0.005, 0.003, 0.003), wnl=0.0):
        """bspline_optimizer.cpp:25-57; defaults = exploration_manager/launch/algorithm.xml:170-192."""
        p = FuelOptParams()
        p.wnl = wnl
        (p.ld_smooth, p.ld_dist, p.ld_feasi, p.ld_start, p.ld_end, p.ld_guide, p.ld_waypt, p.ld_view,
         p.ld_time, p.dist0, p.max_vel, p.max_acc, p.order) = (ld_smooth, ld_dist, ld_feasi, ld_start, ld_end,
                                                               ld_guide, ld_waypt, ld_view, ld_time, dist0,
                                                               max_vel, max_acc, bspline_degree)
        self.params_ = p
        self.bspline_degree_ = bspline_degree
        self.max_iteration_num_ = list(max_iteration_num)
        self.max_iteration_time_ = list(max_iteration_time)
        self.time_lb_ = -1.0

    def setEnvironment(self, env):
        self.edt_environment_ = env

    def setCostFunction(self, cost_code):
        self.cost_function_ = int(cost_code)

    def setBoundaryStates(self, start, end):
        self.start_state_ = [np.asarray(s, dtype=np.float64) for s in start]
        self.end_state_ = [np.asarray(e, dtype=np.float64) for e in end]

    def setTimeLowerBound(self, lb):
        self.time_lb_ = float(lb)

    def setGuidePath(self, guide_pt):
        self.guide_pts_ = [np.asarray(g, dtype=np.float64) for g in guide_pt]

    def setWaypoints(self, waypts, waypt_idx):
        self.waypoints_ = [np.asarray(w, dtype=np.float64) for w in waypts]
        self.waypt_idx_ = list(waypt_idx)

    def setViewConstraint(self, pt, direction, idx):
        """setViewConstraint(vc) (:91-93): the fields calcViewCost reads (vc.pt_, vc.dir_, vc.idx_)."""
        self.view_cons_ = (np.asarray(pt, dtype=np.float64), np.asarray(direction, dtype=np.float64), int(idx))

    # ---- per-trajectory constants (what optimize() freezes, :116-141) ----------------------
    @staticmethod
    def pt_dist(ctrl):
        """pt_dist_ (:136-140): sum of segment lengths divided by the POINT count."""
        ctrl = np.asarray(ctrl, dtype=np.float64)
        d = 0.0
        for i in range(ctrl.shape[0] - 1):
            d += float(np.sqrt(np.sum((ctrl[i + 1] - ctrl[i]) ** 2)))
        return d / float(ctrl.shape[0])

    @staticmethod
    def fill_traj_const(tc, pt_dist, knot_span, start, end, time_lb=-1.0, guide=None, waypt=None,
                        waypt_idx=None, view=None):
        tc.view_idx = -1
        if view is not None:
            for k in range(3):
                tc.view_pt[k] = float(view[0][k])
                tc.view_dir[k] = float(view[1][k])
            tc.view_idx = int(view[2])
        tc.pt_dist = float(pt_dist)
        tc.knot_span = float(knot_span)
        start = np.asarray(start, dtype=np.float64).reshape(3, 3)
        end = np.asarray(end, dtype=np.float64).reshape(-1, 3)
        for i in range(3):
            for k in range(3):
                tc.start[i][k] = start[i, k]
        tc.n_end = end.shape[0]
        for i in range(end.shape[0]):
            for k in range(3):
                tc.end[i][k] = end[i, k]
        tc.time_lb = float(time_lb)
        tc.n_guide = 0
        tc.n_waypt = 0
        if guide is not None and len(guide):
            guide = np.asarray(guide, dtype=np.float64).reshape(-1, 3)
            tc.n_guide = guide.shape[0]
            for i in range(guide.shape[0]):
                for k in range(3):
                    tc.guide[i][k] = guide[i, k]
        if waypt is not None and len(waypt):
            waypt = np.asarray(waypt, dtype=np.float64).reshape(-1, 3)
            tc.n_waypt = waypt.shape[0]
            for i in range(waypt.shape[0]):
                for k in range(3):
                    tc.waypt[i][k] = waypt[i, k]
                tc.waypt_idx[i] = int(waypt_idx[i])

    @staticmethod
    def traj_consts_from_arrays(pt_dist, knot_span, start, end_pos, time_lb=None):
        """Vectorised builder for a batch: start [B,3,3], end_pos [B,3] (end_state_.size()==1,
        the exploration call sites).  Returns a ctypes array of FuelTrajConst."""
        B = len(pt_dist)
        arr = (FuelTrajConst * B)()
        buf = np.frombuffer(arr, dtype=np.uint8).reshape(B, C.sizeof(FuelTrajConst))
        T = FuelTrajConst

        def put(field, values, dtype):
            off = getattr(T, field).offset
            v = np.ascontiguousarray(values, dtype=dtype).reshape(B, -1)
            w = v.view(np.uint8).reshape(B, -1)
            buf[:, off:off + w.shape[1]] = w

        put("pt_dist", pt_dist, np.float64)
        put("knot_span", knot_span, np.float64)
        put("start", np.asarray(start, dtype=np.float64).reshape(B, 9), np.float64)
        e = np.zeros((B, 9), dtype=np.float64)
        e[:, :3] = np.asarray(end_pos, dtype=np.float64).reshape(B, 3)
        put("end", e, np.float64)
        put("n_end", np.ones(B), np.int32)
        put("time_lb", -np.ones(B) if time_lb is None else time_lb, np.float64)
        put("view_idx", -np.ones(B), np.int32)
        return arr

    def nvar(self, n_pts, mask):
        return 3 * n_pts + (1 if mask & self.MINTIME else 0)

    # ---- cost / gradient --------------------------------------------------------------------
    def combineCostBatch(self, x, traj_consts, n_pts, cost_function=None, fast_eval=False):
        """combineCost (:518-647) for x [B, nvar]; returns (f [B], grad [B, nvar]).  fast_eval=True evaluates with
        the solver loop's evaluator (FUELGPU_COST_FAST_EVAL): what optimizeBatch runs K times per trajectory."""
        mask = self.cost_function_ if cost_function is None else int(cost_function)
        x = np.ascontiguousarray(x, dtype=np.float64)
        B = x.shape[0]
        nvar = self.nvar(n_pts, mask)
        if x.shape[1] != nvar:
            raise ValueError("x must be [B, %d]" % nvar)
        f = np.empty(B, dtype=np.float64)
        g = np.empty((B, nvar), dtype=np.float64)
        h = self.edt_environment_.sdf_map_.handle
        check(lib().fuelgpu_bspline_cost_batch(h, B, n_pts, mask | (COST_FAST_EVAL if fast_eval else 0),
                                               C.byref(self.params_), traj_consts, ptr(x), ptr(f), ptr(g)), h)
        return f, g

    def optimizeBatch(self, x, traj_consts, n_pts, cost_function, max_eval, lbfgs_m=6, xtol_rel=1e-5, out=None,
                      exact_evals=False):
        """The solver loop of optimize() (:165-253) for B trajectories in one persistent kernel.
        x [B, nvar] initial variables (clamped to the box shrunk by 0.1 m on the device, :196-204).
        Returns (x_best [B, nvar], f_best [B], n_eval [B]); `out` = a tuple of such arrays to reuse (like
        the reference's best_variable_ member) instead of allocating new ones per call."""
        mask = int(cost_function)
        B = len(x)
        if out is None:
            out = (np.empty((B, self.nvar(n_pts, mask)), dtype=np.float64), np.empty(B, dtype=np.float64),
                   np.empty(B, dtype=np.int32))
        xw, fb, ne = out
        if np.shape(x)[1] != self.nvar(n_pts, mask) or xw.shape != np.shape(x):
            raise ValueError("x must be [B, %d]" % self.nvar(n_pts, mask))
        np.copyto(xw, x)
        x = xw
        sp = FuelSolveParams()
        sp.max_eval, sp.lbfgs_m, sp.xtol_rel = int(max_eval), int(lbfgs_m), float(xtol_rel)
        sp.flags = SOLVE_EXACT_EVALS if exact_evals else 0
        h = self.edt_environment_.sdf_map_.handle
        check(lib().fuelgpu_bspline_optimize_batch(h, B, n_pts, mask, C.byref(self.params_), traj_consts,
                                                   C.byref(sp), ptr(x), ptr(fb), ptr(ne)), h)
        return x, fb, ne

    def optimizeBatchBegin(self, x, traj_consts, n_pts, cost_function, max_eval, lbfgs_m=6, xtol_rel=1e-5,
                           exact_evals=False):
        """First half of optimizeBatch: stage the inputs and enqueue the solver, return at once
        (fuelgpu_bspline_optimize_batch_begin).  Collect with optimizeBatchEnd()."""
        mask = int(cost_function)
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[1] != self.nvar(n_pts, mask):
            raise ValueError("x must be [B, %d]" % self.nvar(n_pts, mask))
        sp = FuelSolveParams()
        sp.max_eval, sp.lbfgs_m, sp.xtol_rel = int(max_eval), int(lbfgs_m), float(xtol_rel)
        sp.flags = SOLVE_EXACT_EVALS if exact_evals else 0
        h = self.edt_environment_.sdf_map_.handle
        check(lib().fuelgpu_bspline_optimize_batch_begin(h, x.shape[0], n_pts, mask, C.byref(self.params_), traj_consts,
                                                         C.byref(sp), ptr(x)), h)
        self._pending_shape = x.shape

    def optimizeBatchEnd(self, out=None):
        """Second half: wait for the solver and return (x_best, f_best, n_eval)."""
        B, nvar = self._pending_shape
        if out is None:
            out = (np.empty((B, nvar), dtype=np.float64), np.empty(B, dtype=np.float64), np.empty(B, dtype=np.int32))
        h = self.edt_environment_.sdf_map_.handle
        check(lib().fuelgpu_bspline_optimize_batch_end(h, ptr(out[0]), ptr(out[1]), ptr(out[2])), h)
        return out

    def _own_traj_const(self, ctrl, dt):
        tc = (FuelTrajConst * 1)()
        start = np.zeros((3, 3))
        for i, s in enumerate(self.start_state_[:3]):
            start[i] = s
        self.fill_traj_const(tc[0], self.pt_dist_, dt, start, np.asarray(self.end_state_).reshape(-1, 3),
                             self.time_lb_, self.guide_pts_, self.waypoints_, self.waypt_idx_,
                             getattr(self, "view_cons_", None))
        return tc

    def costFunction(self, x):
        """The NLopt trampoline (:693-706): one evaluation, tracks the best x."""
        x = np.asarray(x, dtype=np.float64)
        f, g = self.combineCostBatch(x[None, :], self._tc, self.point_num_)
        self.iter_num_ += 1
        if self.min_cost_ is None or f[0] < self.min_cost_:
            self.min_cost_ = float(f[0])
            self.best_variable_ = x.copy()
        return float(f[0]), g[0]

    def begin(self, points, dt, cost_function):
        """The setup half of optimize(points, dt, cost_function, ...) (:110-155)."""
        if not self.start_state_:
            raise RuntimeError("Initial state undefined!")  # :112-115
        self.control_points_ = np.asarray(points, dtype=np.float64).copy()
        self.knot_span_ = float(dt)
        self.setCostFunction(cost_function)
        self.order_ = self.bspline_degree_
        self.point_num_ = self.control_points_.shape[0]
        if self.point_num_ > MAX_PTS:
            raise ValueError("at most %d control points" % MAX_PTS)
        self.optimize_time_ = bool(self.cost_function_ & self.MINTIME)
        self.variable_num_ = self.nvar(self.point_num_, self.cost_function_)
        self.pt_dist_ = self.pt_dist(self.control_points_)
        self.iter_num_ = 0
        self.min_cost_ = None
        self._tc = self._own_traj_const(self.control_points_, self.knot_span_)

    def initial_variables(self):
        """q of optimize() (:196-204): control points clamped to the box shrunk by 0.1 m."""
        bmin, bmax = self.edt_environment_.sdf_map_.getBox()
        q = np.minimum(np.maximum(self.control_points_, bmin + 0.1), bmax - 0.1).reshape(-1)
        if self.optimize_time_:
            q = np.concatenate([q, [self.knot_span_]])
        return q

    def optimize(self, points, dt, cost_function, max_num_id, max_time_id=None, lbfgs_m=6):
        """optimize() (:110-253) with the solver loop on the device.  Returns (points, dt)."""
        self.begin(points, dt, cost_function)
        x = self.initial_variables()[None, :].copy()
        sp = FuelSolveParams()
        sp.max_eval = int(self.max_iteration_num_[max_num_id])
        sp.lbfgs_m = int(lbfgs_m)
        sp.xtol_rel = 1e-5
        fb = np.zeros(1)
        ne = np.zeros(1, dtype=np.int32)
        h = self.edt_environment_.sdf_map_.handle
        check(lib().fuelgpu_bspline_optimize_batch(h, 1, self.point_num_, self.cost_function_,
                                                   C.byref(self.params_), self._tc, C.byref(sp), ptr(x), ptr(fb),
                                                   ptr(ne)), h)
        self.best_variable_ = x[0].copy()
        self.min_cost_ = float(fb[0])
        self.iter_num_ = int(ne[0])
        pts = x[0, :3 * self.point_num_].reshape(self.point_num_, 3)
        out_dt = float(x[0, -1]) if self.optimize_time_ else self.knot_span_
        self.start_state_ = []  # :161-162
        self.time_lb_ = -1.0
        return pts, out_dt
