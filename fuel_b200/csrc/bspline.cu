// bspline.cu -- batched B-spline cost/gradient (BsplineOptimizer::combineCost) on sm_100a.
//
// Replaces bspline_opt/src/bspline_optimizer.cpp:518-647 (combineCost) and the calc*Cost
// functions it calls (:255-516), evaluated for a batch of trajectories against one ESDF
// (EDTEnvironment::evaluateEDTWithGrad -> SDFMap::getDistWithGrad, sdf_map.cpp:497-536).
// fp64 throughout; products and sums are kept in the reference's order and are not
// contracted into FMAs (-fmad=false for this file), so results agree with the host
// arithmetic to rounding of the fp32 ESDF samples.
#include "common.cuh"

namespace {

constexpr int MAXP = FUELGPU_MAX_PTS;

struct Terms {
  double f;
  double gt;
};

// One trajectory's full combineCost, sequential restatement.  q/g live in local memory
// (thread-per-trajectory variant) -- the warp-cooperative kernel below is the fast path.
__device__ void combine_cost_thread(const Geom& g, const float* __restrict__ dist,
                                    const FuelOptParams& p, const FuelTrajConst& tc, int n, int mask,
                                    const double* __restrict__ x, double* __restrict__ fout,
                                    double* __restrict__ grad) {
  const bool opt_time = (mask & FUELGPU_MINTIME) != 0;
  const int nvar = opt_time ? 3 * n + 1 : 3 * n;
  const double dt = opt_time ? x[nvar - 1] : tc.knot_span;
  double q[MAXP][3];
  double gq[MAXP][3];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) q[i][k] = x[3 * i + k];
  for (int i = 0; i < nvar; ++i) grad[i] = 0.0;
  double f_combine = 0.0;

  if (mask & FUELGPU_SMOOTHNESS) {  // calcSmoothnessCost :255-282
    double cost = 0.0;
    for (int i = 0; i < n; ++i) gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
    for (int i = 0; i < n - 3; i++) {
      double ji[3], tj[3];
      for (int k = 0; k < 3; ++k)
        ji[k] = (q[i + 3][k] - 3 * q[i + 2][k] + 3 * q[i + 1][k] - q[i][k]) / tc.pt_dist;
      cost += ji[0] * ji[0] + ji[1] * ji[1] + ji[2] * ji[2];
      for (int k = 0; k < 3; ++k) tj[k] = 2 * ji[k] / tc.pt_dist;
      for (int k = 0; k < 3; ++k) {
        gq[i + 0][k] += -tj[k];
        gq[i + 1][k] += 3.0 * tj[k];
        gq[i + 2][k] += -3.0 * tj[k];
        gq[i + 3][k] += tj[k];
      }
    }
    f_combine += p.ld_smooth * cost;
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_smooth * gq[i][k];
    if (opt_time) grad[nvar - 1] += p.ld_smooth * 0.0;
  }
  if (mask & FUELGPU_DISTANCE) {  // calcDistanceCost :284-306
    double cost = 0.0;
    for (int i = 0; i < n; i++) {
      double dg[3];
      const double d = dev_dist_with_grad(g, dist, q[i], dg);
      const double nrm = sqrt(dg[0] * dg[0] + dg[1] * dg[1] + dg[2] * dg[2]);
      if (nrm > 1e-4) {
        dg[0] /= nrm;
        dg[1] /= nrm;
        dg[2] /= nrm;
      }
      gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
      if (d < p.dist0) {
        cost += (d - p.dist0) * (d - p.dist0);
        for (int k = 0; k < 3; ++k) gq[i][k] += 2.0 * (d - p.dist0) * dg[k];
      }
    }
    f_combine += p.ld_dist * cost;
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_dist * gq[i][k];
  }
  if (mask & FUELGPU_FEASIBILITY) {  // calcFeasibilityCost :308-353
    double cost = 0.0, gt = 0.0;
    for (int i = 0; i < n; ++i) gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
    const double dt_inv = 1 / dt;
    const double dt_inv2 = dt_inv * dt_inv;
    for (int i = 0; i < n - 1; ++i)
      for (int k = 0; k < 3; ++k) {
        const double vi = (q[i + 1][k] - q[i][k]) * dt_inv;
        const double vd = fabs(vi) - p.max_vel;
        if (vd > 0.0) {
          cost += vd * vd;
          const double sign = vi > 0 ? 1.0 : -1.0;
          const double tmp = 2 * vd * sign * dt_inv;
          gq[i][k] += -tmp;
          gq[i + 1][k] += tmp;
          if (opt_time) gt += tmp * (-vi);
        }
      }
    for (int i = 0; i < n - 2; ++i)
      for (int k = 0; k < 3; ++k) {
        const double ai = (q[i + 2][k] - 2 * q[i + 1][k] + q[i][k]) * dt_inv2;
        const double ad = fabs(ai) - p.max_acc;
        if (ad > 0.0) {
          cost += ad * ad;
          const double sign = ai > 0 ? 1.0 : -1.0;
          const double tmp = 2 * ad * sign * dt_inv2;
          gq[i][k] += tmp;
          gq[i + 1][k] += -2 * tmp;
          gq[i + 2][k] += tmp;
          if (opt_time) gt += tmp * ai * (-2) * dt;
        }
      }
    f_combine += p.ld_feasi * cost;
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_feasi * gq[i][k];
    if (opt_time) grad[nvar - 1] += p.ld_feasi * gt;
  }
  if (mask & FUELGPU_START) {  // calcStartCost :355-391
    double cost = 0.0, gt = 0.0;
    for (int i = 0; i < 3; ++i) gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
    const double w_pos = 10.0;
    double dq[3];
    for (int k = 0; k < 3; ++k) dq[k] = 1 / 6.0 * (q[0][k] + 4 * q[1][k] + q[2][k]) - tc.start[0][k];
    cost += w_pos * (dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
    for (int k = 0; k < 3; ++k) {
      gq[0][k] += w_pos * 2 * dq[k] * (1 / 6.0);
      gq[1][k] += w_pos * 2 * dq[k] * (4 / 6.0);
      gq[2][k] += w_pos * 2 * dq[k] * (1 / 6.0);
    }
    for (int k = 0; k < 3; ++k) dq[k] = 1 / (2 * dt) * (q[2][k] - q[0][k]) - tc.start[1][k];
    cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
    for (int k = 0; k < 3; ++k) {
      gq[0][k] += 2 * dq[k] * (-1.0) / (2 * dt);
      gq[2][k] += 2 * dq[k] * 1.0 / (2 * dt);
    }
    if (opt_time) {
      double d = 0;
      for (int k = 0; k < 3; ++k) d += dq[k] * (q[2][k] - q[0][k]);
      gt += d / (-dt * dt);
    }
    for (int k = 0; k < 3; ++k) dq[k] = 1 / (dt * dt) * (q[0][k] - 2 * q[1][k] + q[2][k]) - tc.start[2][k];
    cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
    for (int k = 0; k < 3; ++k) {
      gq[0][k] += 2 * dq[k] * 1.0 / (dt * dt);
      gq[1][k] += 2 * dq[k] * (-2.0) / (dt * dt);
      gq[2][k] += 2 * dq[k] * 1.0 / (dt * dt);
    }
    if (opt_time) {
      double d = 0;
      for (int k = 0; k < 3; ++k) d += dq[k] * (q[0][k] - 2 * q[1][k] + q[2][k]);
      gt += d / (-dt * dt * dt);
    }
    f_combine += p.ld_start * cost;
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_start * gq[i][k];
    if (opt_time) grad[nvar - 1] += p.ld_start * gt;
  }
  if (mask & FUELGPU_END) {  // calcEndCost :393-431
    double cost = 0.0, gt = 0.0;
    for (int i = n - 3; i < n; ++i) gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
    const double* q_3 = q[n - 3];
    const double* q_2 = q[n - 2];
    const double* q_1 = q[n - 1];
    double dq[3];
    for (int k = 0; k < 3; ++k) dq[k] = 1 / 6.0 * (q_1[k] + 4 * q_2[k] + q_3[k]) - tc.end[0][k];
    cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
    for (int k = 0; k < 3; ++k) {
      gq[n - 1][k] += 2 * dq[k] * (1 / 6.0);
      gq[n - 2][k] += 2 * dq[k] * (4 / 6.0);
      gq[n - 3][k] += 2 * dq[k] * (1 / 6.0);
    }
    if (tc.n_end >= 2) {
      for (int k = 0; k < 3; ++k) dq[k] = 1 / (2 * dt) * (q_1[k] - q_3[k]) - tc.end[1][k];
      cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
      for (int k = 0; k < 3; ++k) {
        gq[n - 1][k] += 2 * dq[k] * 1.0 / (2 * dt);
        gq[n - 3][k] += 2 * dq[k] * (-1.0) / (2 * dt);
      }
      if (opt_time) {
        double d = 0;
        for (int k = 0; k < 3; ++k) d += dq[k] * (q_1[k] - q_3[k]);
        gt += d / (-dt * dt);
      }
    }
    if (tc.n_end == 3) {
      for (int k = 0; k < 3; ++k) dq[k] = 1 / (dt * dt) * (q_1[k] - 2 * q_2[k] + q_3[k]) - tc.end[2][k];
      cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
      for (int k = 0; k < 3; ++k) {
        gq[n - 1][k] += 2 * dq[k] * 1.0 / (dt * dt);
        gq[n - 2][k] += 2 * dq[k] * (-2.0) / (dt * dt);
        gq[n - 3][k] += 2 * dq[k] * 1.0 / (dt * dt);
      }
      if (opt_time) {
        double d = 0;
        for (int k = 0; k < 3; ++k) d += dq[k] * (q_1[k] - 2 * q_2[k] + q_3[k]);
        gt += d / (-dt * dt * dt);
      }
    }
    f_combine += p.ld_end * cost;
    for (int i = n - 3; i < n; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_end * gq[i][k];
    if (opt_time) grad[nvar - 1] += p.ld_end * gt;
  }
  if (mask & FUELGPU_GUIDE) {  // calcGuideCost :462-475
    double cost = 0.0;
    for (int i = 0; i < n; ++i) gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
    const int end_idx = n - p.order;
    for (int i = p.order; i < end_idx; i++) {
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = q[i][k] - tc.guide[i - p.order][k];
      cost += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      for (int k = 0; k < 3; ++k) gq[i][k] += 2 * d[k];
    }
    f_combine += p.ld_guide * cost;
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_guide * gq[i][k];
  }
  if (mask & FUELGPU_WAYPOINTS) {  // calcWaypointsCost :433-457
    double cost = 0.0;
    for (int i = 0; i < n; ++i) gq[i][0] = gq[i][1] = gq[i][2] = 0.0;
    for (int i = 0; i < tc.n_waypt; ++i) {
      const int idx = tc.waypt_idx[i];
      double dq[3];
      for (int k = 0; k < 3; ++k)
        dq[k] = 1 / 6.0 * (q[idx][k] + 4 * q[idx + 1][k] + q[idx + 2][k]) - tc.waypt[i][k];
      cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
      for (int k = 0; k < 3; ++k) {
        gq[idx][k] += dq[k] * (2.0 / 6.0);
        gq[idx + 1][k] += dq[k] * (8.0 / 6.0);
        gq[idx + 2][k] += dq[k] * (2.0 / 6.0);
      }
    }
    f_combine += p.ld_waypt * cost;
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) grad[3 * i + k] += p.ld_waypt * gq[i][k];
  }
  if (mask & FUELGPU_MINTIME) {  // calcTimeCost :504-516
    const double duration = (n - p.order) * dt;
    double cost = duration;
    double gt = (double)(n - p.order);
    if (tc.time_lb > 0 && duration < tc.time_lb) {
      const double w_lb = 10;
      cost += w_lb * (duration - tc.time_lb) * (duration - tc.time_lb);
      gt += w_lb * 2 * (duration - tc.time_lb) * (n - p.order);
    }
    f_combine += p.ld_time * cost;
    grad[nvar - 1] += p.ld_time * gt;
  }
  *fout = f_combine;
}

__global__ void __launch_bounds__(64) cost_batch_thread_kernel(Geom g, const float* __restrict__ dist,
                                                               FuelOptParams p,
                                                               const FuelTrajConst* __restrict__ tc,
                                                               int n, int mask, int B,
                                                               const double* __restrict__ x,
                                                               double* __restrict__ f,
                                                               double* __restrict__ grad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nvar = (mask & FUELGPU_MINTIME) ? 3 * n + 1 : 3 * n;
  combine_cost_thread(g, dist, p, tc[b], n, mask, x + (int64_t)b * nvar, f + b, grad + (int64_t)b * nvar);
}


// =========================================================================================
// Warp-cooperative evaluation: lane i <-> control point i (n <= 32; with MINTIME n <= 31 in
// the optimiser, where lane n carries dt).  Control points stay in registers; neighbours
// come from warp shuffles; the 8 ESDF samples of every control point are gathered by its
// own lane, so one warp has 8*n independent loads in flight (L2-resident map).
// The accumulation order of every gradient row is the reference's loop order, so a row is
// bit-identical to the sequential restatement; only the scalar sums (costs, dt-gradient)
// are warp reductions and may differ in the last bits.
// =========================================================================================
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ double up(double v, int d, int lane) {
  const double r = __shfl_up_sync(0xffffffffu, v, d);
  return lane >= d ? r : 0.0;
}

struct TrajRegs {  // loop-invariant per-trajectory constants, loaded once
  double pt_dist, knot_span, time_lb;
  double start[3][3];
  double end[3][3];
  int n_end, n_guide, n_waypt;
};

__device__ __forceinline__ void load_traj(const FuelTrajConst* __restrict__ tc, TrajRegs& r) {
  r.pt_dist = tc->pt_dist;
  r.knot_span = tc->knot_span;
  r.time_lb = tc->time_lb;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      r.start[i][k] = tc->start[i][k];
      r.end[i][k] = tc->end[i][k];
    }
  r.n_end = tc->n_end;
  r.n_guide = tc->n_guide;
  r.n_waypt = tc->n_waypt;
}

// q[3]: this lane's control point (lanes >= n hold anything finite).  Returns f in every
// lane, this lane's gradient row in gr[3] (zero for lanes >= n) and the dt-gradient in gdt.
// FAST (the solver kernel): divisions by loop-invariant scalars become multiplications by
// reciprocals computed once per evaluation, the ESDF gradient is normalised with rsqrt, and
// the per-term warp reductions are merged into one (cost) + one (dt-gradient).  Same
// mathematics, rounding differs in the last bits; the faithful variant backs cost_batch.
template <bool FAST>
__device__ __forceinline__ void eval_warp(const Geom& g, const float* __restrict__ dist,
                                          const FuelOptParams& p, const TrajRegs& t,
                                          const FuelTrajConst* __restrict__ tc, int n, int mask,
                                          const double q[3], double dt, int lane, double& f_out,
                                          double gr[3], double& gdt) {
  const bool opt_time = (mask & FUELGPU_MINTIME) != 0;
  const bool act = lane < n;
  double f = 0.0;
  gr[0] = gr[1] = gr[2] = 0.0;
  gdt = 0.0;
  double f_lane = 0.0, gdt_lane = 0.0;  // FAST: per-lane partial sums, reduced once at the end
  const double inv_pt = FAST ? 1.0 / t.pt_dist : 0.0;
  const double dt_inv_f = FAST ? 1.0 / dt : 0.0;
  const double inv2dt = 0.5 * dt_inv_f, invdt2 = dt_inv_f * dt_inv_f;

  // neighbours i+1..i+3
  double q1[3], q2[3], q3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    q1[k] = __shfl_down_sync(0xffffffffu, q[k], 1);
    q2[k] = __shfl_down_sync(0xffffffffu, q[k], 2);
    q3[k] = __shfl_down_sync(0xffffffffu, q[k], 3);
  }

  if (mask & FUELGPU_SMOOTHNESS) {  // calcSmoothnessCost :255-282
    const bool v = lane <= n - 4;
    double tj[3], c = 0.0;
    {
      double ji[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double num = q3[k] - 3 * q2[k] + 3 * q1[k] - q[k];
        ji[k] = FAST ? num * inv_pt : num / t.pt_dist;
      }
      c = ji[0] * ji[0] + ji[1] * ji[1] + ji[2] * ji[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) tj[k] = v ? (FAST ? 2 * ji[k] * inv_pt : 2 * ji[k] / t.pt_dist) : 0.0;
      if (!v) c = 0.0;
    }
    if (FAST)
      f_lane += p.ld_smooth * c;
    else
      f += p.ld_smooth * wsum(c);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double t1 = up(tj[k], 1, lane), t2 = up(tj[k], 2, lane), t3 = up(tj[k], 3, lane);
      double gq = 0.0;
      gq += t3;          // i = p-3: gq[i+3] +=  tj
      gq += -3.0 * t2;   // i = p-2: gq[i+2] += -3 tj
      gq += 3.0 * t1;    // i = p-1: gq[i+1] +=  3 tj
      gq += -tj[k];      // i = p  : gq[i]   += -tj
      gr[k] += p.ld_smooth * gq;
    }
  }
  if (mask & FUELGPU_DISTANCE) {  // calcDistanceCost :284-306
    double c = 0.0, gq[3] = { 0.0, 0.0, 0.0 };
    if (act) {
      double dg[3];
      const double d = dev_dist_with_grad(g, dist, q, dg);
      if (FAST) {
        const double n2 = dg[0] * dg[0] + dg[1] * dg[1] + dg[2] * dg[2];
        if (n2 > 1e-8) {
          const double rn = rsqrt(n2);
          dg[0] *= rn;
          dg[1] *= rn;
          dg[2] *= rn;
        }
      } else {
        const double nrm = sqrt(dg[0] * dg[0] + dg[1] * dg[1] + dg[2] * dg[2]);
        if (nrm > 1e-4) {
          dg[0] /= nrm;
          dg[1] /= nrm;
          dg[2] /= nrm;
        }
      }
      if (d < p.dist0) {
        c = (d - p.dist0) * (d - p.dist0);
#pragma unroll
        for (int k = 0; k < 3; ++k) gq[k] += 2.0 * (d - p.dist0) * dg[k];
      }
    }
    if (FAST)
      f_lane += p.ld_dist * c;
    else
      f += p.ld_dist * wsum(c);
#pragma unroll
    for (int k = 0; k < 3; ++k) gr[k] += p.ld_dist * gq[k];
  }
  if (mask & FUELGPU_FEASIBILITY) {  // calcFeasibilityCost :308-353
    const double dt_inv = FAST ? dt_inv_f : 1 / dt;
    const double dt_inv2 = dt_inv * dt_inv;
    double c = 0.0, gtl = 0.0;
    double tv[3], ta[3];
    const bool vv = lane <= n - 2, va = lane <= n - 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      tv[k] = 0.0;
      const double vi = (q1[k] - q[k]) * dt_inv;
      const double vd = fabs(vi) - p.max_vel;
      if (vv && vd > 0.0) {
        c += vd * vd;
        const double sign = vi > 0 ? 1.0 : -1.0;
        tv[k] = 2 * vd * sign * dt_inv;
        if (opt_time) gtl += tv[k] * (-vi);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ta[k] = 0.0;
      const double ai = (q2[k] - 2 * q1[k] + q[k]) * dt_inv2;
      const double ad = fabs(ai) - p.max_acc;
      if (va && ad > 0.0) {
        c += ad * ad;
        const double sign = ai > 0 ? 1.0 : -1.0;
        ta[k] = 2 * ad * sign * dt_inv2;
        if (opt_time) gtl += ta[k] * ai * (-2) * dt;
      }
    }
    double gt = 0.0;
    if (FAST) {
      f_lane += p.ld_feasi * c;
      gdt_lane += p.ld_feasi * gtl;
    } else {
      f += p.ld_feasi * wsum(c);
      gt = wsum(gtl);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v1 = up(tv[k], 1, lane);
      const double a1 = up(ta[k], 1, lane), a2 = up(ta[k], 2, lane);
      double gq = 0.0;
      gq += v1;          // velocity loop, i = p-1: gq[i+1] += tmp
      gq += -tv[k];      //                i = p  : gq[i]   += -tmp
      gq += a2;          // acceleration loop, i = p-2: gq[i+2] += tmp
      gq += -2 * a1;     //                    i = p-1: gq[i+1] += -2 tmp
      gq += ta[k];       //                    i = p  : gq[i]   += tmp
      gr[k] += p.ld_feasi * gq;
    }
    if (opt_time && !FAST) gdt += p.ld_feasi * gt;
  }
  if (mask & FUELGPU_START) {  // calcStartCost :355-391
    double a[3], b[3], c3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[k] = __shfl_sync(0xffffffffu, q[k], 0);
      b[k] = __shfl_sync(0xffffffffu, q[k], 1);
      c3[k] = __shfl_sync(0xffffffffu, q[k], 2);
    }
    const double w_pos = 10.0;
    double cost = 0.0, gt = 0.0, row[3] = { 0.0, 0.0, 0.0 };
    double dq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dq[k] = 1 / 6.0 * (a[k] + 4 * b[k] + c3[k]) - t.start[0][k];
    cost += w_pos * (dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (lane == 0) row[k] += w_pos * 2 * dq[k] * (1 / 6.0);
      if (lane == 1) row[k] += w_pos * 2 * dq[k] * (4 / 6.0);
      if (lane == 2) row[k] += w_pos * 2 * dq[k] * (1 / 6.0);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) dq[k] = (FAST ? inv2dt * (c3[k] - a[k]) : 1 / (2 * dt) * (c3[k] - a[k])) - t.start[1][k];
    cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (lane == 0) row[k] += (FAST ? 2 * dq[k] * (-1.0) * inv2dt : 2 * dq[k] * (-1.0) / (2 * dt));
      if (lane == 2) row[k] += (FAST ? 2 * dq[k] * inv2dt : 2 * dq[k] * 1.0 / (2 * dt));
    }
    if (opt_time) {
      double d = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) d += dq[k] * (c3[k] - a[k]);
      gt += FAST ? -d * invdt2 : d / (-dt * dt);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) dq[k] = (FAST ? invdt2 * (a[k] - 2 * b[k] + c3[k]) : 1 / (dt * dt) * (a[k] - 2 * b[k] + c3[k])) - t.start[2][k];
    cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (lane == 0) row[k] += (FAST ? 2 * dq[k] * invdt2 : 2 * dq[k] * 1.0 / (dt * dt));
      if (lane == 1) row[k] += (FAST ? 2 * dq[k] * (-2.0) * invdt2 : 2 * dq[k] * (-2.0) / (dt * dt));
      if (lane == 2) row[k] += (FAST ? 2 * dq[k] * invdt2 : 2 * dq[k] * 1.0 / (dt * dt));
    }
    if (opt_time) {
      double d = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) d += dq[k] * (a[k] - 2 * b[k] + c3[k]);
      gt += FAST ? -d * invdt2 * dt_inv_f : d / (-dt * dt * dt);
    }
    f += p.ld_start * cost;
    if (lane < 3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) gr[k] += p.ld_start * row[k];
    }
    if (opt_time) gdt += p.ld_start * gt;
  }
  if (mask & FUELGPU_END) {  // calcEndCost :393-431
    double q_3[3], q_2[3], q_1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q_3[k] = __shfl_sync(0xffffffffu, q[k], n - 3);
      q_2[k] = __shfl_sync(0xffffffffu, q[k], n - 2);
      q_1[k] = __shfl_sync(0xffffffffu, q[k], n - 1);
    }
    double cost = 0.0, gt = 0.0, row[3] = { 0.0, 0.0, 0.0 };
    double dq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dq[k] = 1 / 6.0 * (q_1[k] + 4 * q_2[k] + q_3[k]) - t.end[0][k];
    cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (lane == n - 1) row[k] += 2 * dq[k] * (1 / 6.0);
      if (lane == n - 2) row[k] += 2 * dq[k] * (4 / 6.0);
      if (lane == n - 3) row[k] += 2 * dq[k] * (1 / 6.0);
    }
    if (t.n_end >= 2) {
#pragma unroll
      for (int k = 0; k < 3; ++k) dq[k] = (FAST ? inv2dt * (q_1[k] - q_3[k]) : 1 / (2 * dt) * (q_1[k] - q_3[k])) - t.end[1][k];
      cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (lane == n - 1) row[k] += (FAST ? 2 * dq[k] * inv2dt : 2 * dq[k] * 1.0 / (2 * dt));
        if (lane == n - 3) row[k] += (FAST ? 2 * dq[k] * (-1.0) * inv2dt : 2 * dq[k] * (-1.0) / (2 * dt));
      }
      if (opt_time) {
        double d = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) d += dq[k] * (q_1[k] - q_3[k]);
        gt += FAST ? -d * invdt2 : d / (-dt * dt);
      }
    }
    if (t.n_end == 3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) dq[k] = (FAST ? invdt2 * (q_1[k] - 2 * q_2[k] + q_3[k]) : 1 / (dt * dt) * (q_1[k] - 2 * q_2[k] + q_3[k])) - t.end[2][k];
      cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (lane == n - 1) row[k] += (FAST ? 2 * dq[k] * invdt2 : 2 * dq[k] * 1.0 / (dt * dt));
        if (lane == n - 2) row[k] += (FAST ? 2 * dq[k] * (-2.0) * invdt2 : 2 * dq[k] * (-2.0) / (dt * dt));
        if (lane == n - 3) row[k] += (FAST ? 2 * dq[k] * invdt2 : 2 * dq[k] * 1.0 / (dt * dt));
      }
      if (opt_time) {
        double d = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) d += dq[k] * (q_1[k] - 2 * q_2[k] + q_3[k]);
        gt += FAST ? -d * invdt2 * dt_inv_f : d / (-dt * dt * dt);
      }
    }
    f += p.ld_end * cost;
    if (lane >= n - 3 && lane < n) {
#pragma unroll
      for (int k = 0; k < 3; ++k) gr[k] += p.ld_end * row[k];
    }
    if (opt_time) gdt += p.ld_end * gt;
  }
  if (mask & FUELGPU_GUIDE) {  // calcGuideCost :462-475
    double c = 0.0, gq[3] = { 0.0, 0.0, 0.0 };
    if (lane >= p.order && lane < n - p.order) {
      double d[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) d[k] = q[k] - tc->guide[lane - p.order][k];
      c = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) gq[k] += 2 * d[k];
    }
    if (FAST)
      f_lane += p.ld_guide * c;
    else
      f += p.ld_guide * wsum(c);
#pragma unroll
    for (int k = 0; k < 3; ++k) gr[k] += p.ld_guide * gq[k];
  }
  if (mask & FUELGPU_WAYPOINTS) {  // calcWaypointsCost :433-457
    double cost = 0.0, gq[3] = { 0.0, 0.0, 0.0 };
    for (int w = 0; w < t.n_waypt; ++w) {
      const int idx = tc->waypt_idx[w];
      double dq[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double a = __shfl_sync(0xffffffffu, q[k], idx);
        const double b = __shfl_sync(0xffffffffu, q[k], idx + 1);
        const double c = __shfl_sync(0xffffffffu, q[k], idx + 2);
        dq[k] = 1 / 6.0 * (a + 4 * b + c) - tc->waypt[w][k];
      }
      cost += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (lane == idx) gq[k] += dq[k] * (2.0 / 6.0);
        if (lane == idx + 1) gq[k] += dq[k] * (8.0 / 6.0);
        if (lane == idx + 2) gq[k] += dq[k] * (2.0 / 6.0);
      }
    }
    f += p.ld_waypt * cost;
#pragma unroll
    for (int k = 0; k < 3; ++k) gr[k] += p.ld_waypt * gq[k];
  }
  if (mask & FUELGPU_MINTIME) {  // calcTimeCost :504-516
    const double duration = (n - p.order) * dt;
    double cost = duration;
    double gt = (double)(n - p.order);
    if (t.time_lb > 0 && duration < t.time_lb) {
      const double w_lb = 10;
      cost += w_lb * (duration - t.time_lb) * (duration - t.time_lb);
      gt += w_lb * 2 * (duration - t.time_lb) * (n - p.order);
    }
    f += p.ld_time * cost;
    gdt += p.ld_time * gt;
  }
  if (FAST) {
    // one butterfly for both scalars (the two chains interleave)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      f_lane += __shfl_down_sync(0xffffffffu, f_lane, o);
      gdt_lane += __shfl_down_sync(0xffffffffu, gdt_lane, o);
    }
    f += __shfl_sync(0xffffffffu, f_lane, 0);
    if (opt_time) gdt += __shfl_sync(0xffffffffu, gdt_lane, 0);
  }
  if (!act) gr[0] = gr[1] = gr[2] = 0.0;
  f_out = f;
}

constexpr int WPB = 4;  // warps (trajectories) per CTA

__global__ void __launch_bounds__(WPB * 32) cost_batch_warp_kernel(
    Geom g, const float* __restrict__ dist, FuelOptParams p, const FuelTrajConst* __restrict__ tc, int n,
    int mask, int B, const double* __restrict__ x, double* __restrict__ f, double* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (b >= B) return;
  const bool opt_time = (mask & FUELGPU_MINTIME) != 0;
  const int nvar = opt_time ? 3 * n + 1 : 3 * n;
  const double* xb = x + (int64_t)b * nvar;
  TrajRegs t;
  load_traj(tc + b, t);
  double q[3] = { 0.0, 0.0, 0.0 };
  if (lane < n) {
    q[0] = xb[3 * lane];
    q[1] = xb[3 * lane + 1];
    q[2] = xb[3 * lane + 2];
  }
  const double dt = opt_time ? xb[nvar - 1] : t.knot_span;
  double fo, gr[3], gdt;
  eval_warp<false>(g, dist, p, t, tc + b, n, mask, q, dt, lane, fo, gr, gdt);
  double* gb = grad + (int64_t)b * nvar;
  if (lane < n) {
    gb[3 * lane] = gr[0];
    gb[3 * lane + 1] = gr[1];
    gb[3 * lane + 2] = gr[2];
  }
  if (lane == 0) {
    f[b] = fo;
    if (opt_time) gb[nvar - 1] = gdt;
  }
}

// =========================================================================================
// Persistent per-trajectory solver: replaces the NLopt driver loop of
// BsplineOptimizer::optimize() (:165-253) -- clamp to the box shrunk by 0.1 m (:175-204),
// bounds q0 +- 10 m clipped to that box and dt in [0,5] (:206-217), maxeval stop (:170),
// xtol_rel stop (:173), best-x tracking of costFunction (:693-706) -- around a projected
// L-BFGS with Armijo backtracking.  One warp per trajectory for the whole solve; the
// iterate, gradient and search direction live in registers (lane i = control point i,
// lane n = dt), the (s,y) history in shared memory.
// =========================================================================================
constexpr int MAXM = 8;

struct V3 {
  double v[3];
};
__device__ __forceinline__ double dot3(const V3& a, const V3& b) {
  return wsum(a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]);
}

__global__ void __launch_bounds__(WPB * 32) optimize_warp_kernel(
    Geom g, const float* __restrict__ dist, FuelOptParams p, const FuelTrajConst* __restrict__ tc, int n,
    int mask, int B, FuelSolveParams sp, double* __restrict__ x, double* __restrict__ fbest,
    int* __restrict__ neval_out) {
  extern __shared__ double hist[];  // [WPB][2][m][32][3]
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int b = blockIdx.x * WPB + w;
  if (b >= B) return;
  const bool opt_time = (mask & FUELGPU_MINTIME) != 0;
  const int nvar = opt_time ? 3 * n + 1 : 3 * n;
  const int m = sp.lbfgs_m;
  double* S = hist + (size_t)w * 2 * m * 96;
  double* Y = S + (size_t)m * 96;
  double* xb = x + (int64_t)b * nvar;
  TrajRegs t;
  load_traj(tc + b, t);

  // variables of this lane: control point (lane < n), dt in component 0 of lane n
  const bool is_pt = lane < n;
  const bool is_dt = opt_time && lane == n;
  V3 X, lb, ub;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    X.v[k] = 0.0;
    lb.v[k] = 0.0;
    ub.v[k] = 0.0;
    if (is_pt) {
      const double bmin = g.box_mind[k] + 0.1, bmax = g.box_maxd[k] - 0.1;
      double c = xb[3 * lane + k];
      c = fmax(fmin(c, bmax), bmin);  // :199-203
      X.v[k] = c;
      lb.v[k] = fmax(c - 10.0, bmin);  // :208-214
      ub.v[k] = fmin(c + 10.0, bmax);
    }
  }
  if (is_dt) {
    X.v[0] = xb[nvar - 1];
    lb.v[0] = 0.0;  // :215-218
    ub.v[0] = 5.0;
  }

  auto evaluate = [&](const V3& xx, double& fo, V3& go) {
    const double dtv = opt_time ? __shfl_sync(0xffffffffu, xx.v[0], n) : t.knot_span;
    double gr[3], gdt;
    eval_warp<true>(g, dist, p, t, tc + b, n, mask, xx.v, dtv, lane, fo, gr, gdt);
    go.v[0] = is_pt ? gr[0] : (is_dt ? gdt : 0.0);
    go.v[1] = is_pt ? gr[1] : 0.0;
    go.v[2] = is_pt ? gr[2] : 0.0;
  };
  auto store_best = [&](const V3& xx, double fv) {
    if (is_pt) {
      xb[3 * lane] = xx.v[0];
      xb[3 * lane + 1] = xx.v[1];
      xb[3 * lane + 2] = xx.v[2];
    }
    if (is_dt) xb[nvar - 1] = xx.v[0];
    if (lane == 0) fbest[b] = fv;
  };

  double F;
  V3 G;
  evaluate(X, F, G);
  int neval = 1;
  double best = F;
  store_best(X, F);
  // a NaN/inf start cannot be improved on by comparison; treat as +inf
  if (!(best == best)) best = 1.7976931348623157e308;

  int cnt = 0, head = 0;  // history ring: newest at (head-1) mod m
  double rho[MAXM];
#pragma unroll
  for (int j = 0; j < MAXM; ++j) rho[j] = 0.0;

  while (neval < sp.max_eval) {
    // projected gradient
    V3 PG, D;
    bool actv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      actv[k] = (X.v[k] <= lb.v[k] && G.v[k] > 0.0) || (X.v[k] >= ub.v[k] && G.v[k] < 0.0);
      PG.v[k] = actv[k] ? 0.0 : G.v[k];
    }
    const double pgn2 = dot3(PG, PG);
    if (!(pgn2 > 1e-24)) break;
    // two-loop recursion
    V3 Q = PG;
    double alpha[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      alpha[j] = 0.0;
      if (j < cnt) {
        const int slot = (head - 1 - j + 2 * MAXM * m) % m;
        V3 s, y;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          s.v[k] = S[slot * 96 + lane * 3 + k];
          y.v[k] = Y[slot * 96 + lane * 3 + k];
        }
        alpha[j] = rho[slot] * dot3(s, Q);
#pragma unroll
        for (int k = 0; k < 3; ++k) Q.v[k] -= alpha[j] * y.v[k];
      }
    }
    if (cnt > 0) {
      const int slot = (head - 1 + m) % m;
      V3 s, y;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s.v[k] = S[slot * 96 + lane * 3 + k];
        y.v[k] = Y[slot * 96 + lane * 3 + k];
      }
      const double gamma = dot3(s, y) / dot3(y, y);
#pragma unroll
      for (int k = 0; k < 3; ++k) Q.v[k] *= gamma;
    }
#pragma unroll
    for (int j = MAXM - 1; j >= 0; --j) {
      if (j < cnt) {
        const int slot = (head - 1 - j + 2 * MAXM * m) % m;
        V3 s, y;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          s.v[k] = S[slot * 96 + lane * 3 + k];
          y.v[k] = Y[slot * 96 + lane * 3 + k];
        }
        const double beta = rho[slot] * dot3(y, Q);
#pragma unroll
        for (int k = 0; k < 3; ++k) Q.v[k] += s.v[k] * (alpha[j] - beta);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) D.v[k] = actv[k] ? 0.0 : -Q.v[k];
    double gd = dot3(G, D);
    if (!(gd < 0.0)) {  // not a descent direction: restart from steepest descent
#pragma unroll
      for (int k = 0; k < 3; ++k) D.v[k] = -PG.v[k];
      gd = -pgn2;
      cnt = 0;
    }
    double step = cnt == 0 ? fmin(1.0, 1.0 / sqrt(pgn2)) : 1.0;

    // Armijo backtracking on the projected path
    bool accepted = false;
    V3 XN, GN;
    double FN = 0.0;
    while (neval < sp.max_eval) {
#pragma unroll
      for (int k = 0; k < 3; ++k) XN.v[k] = fmax(fmin(X.v[k] + step * D.v[k], ub.v[k]), lb.v[k]);
      evaluate(XN, FN, GN);
      ++neval;
      if (FN < best) {  // costFunction :698-704
        best = FN;
        store_best(XN, FN);
      }
      V3 dx;
#pragma unroll
      for (int k = 0; k < 3; ++k) dx.v[k] = XN.v[k] - X.v[k];
      const double dec = dot3(G, dx);
      if (FN <= F + 1e-4 * dec) {
        accepted = true;
        break;
      }
      step *= 0.5;
      if (step < 1e-12) break;
    }
    if (!accepted) break;
    V3 s, y;
    bool small = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      s.v[k] = XN.v[k] - X.v[k];
      y.v[k] = GN.v[k] - G.v[k];
      small = small && (fabs(s.v[k]) <= sp.xtol_rel * fabs(XN.v[k]));
    }
    const double sy = dot3(s, y);
    const double ss = dot3(s, s), yy = dot3(y, y);
    if (sy > 1e-10 * sqrt(ss * yy)) {
      const int slot = head;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        S[slot * 96 + lane * 3 + k] = s.v[k];
        Y[slot * 96 + lane * 3 + k] = y.v[k];
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < MAXM; ++j)
        if (j == slot) rho[j] = 1.0 / sy;
      head = (head + 1) % m;
      if (cnt < m) ++cnt;
    }
    X = XN;
    F = FN;
    G = GN;
    if (__all_sync(0xffffffffu, small)) break;  // xtol_rel, :173
  }
  if (lane == 0) neval_out[b] = neval;
}

}  // namespace

int bspline_cost_batch_dev_impl(FuelMap* m, int B, int n_pts, int mask, const FuelOptParams* p,
                                const FuelTrajConst* tc_dev, const double* x_dev, double* f_dev,
                                double* grad_dev) {
  if (B <= 0) return 0;
  if (n_pts <= 32) {
    cost_batch_warp_kernel<<<(B + WPB - 1) / WPB, WPB * 32, 0, m->stream>>>(m->g, m->dist, *p, tc_dev, n_pts,
                                                                        mask, B, x_dev, f_dev, grad_dev);
  } else {
    cost_batch_thread_kernel<<<(B + 63) / 64, 64, 0, m->stream>>>(m->g, m->dist, *p, tc_dev, n_pts, mask, B,
                                                                  x_dev, f_dev, grad_dev);
  }
  FUEL_LAUNCHES(m, 1);
  FUEL_CUDA(m, cudaGetLastError());
  return 0;
}

int bspline_optimize_batch_dev_impl(FuelMap* m, int B, int n_pts, int mask, const FuelOptParams* p,
                                    const FuelTrajConst* tc_dev, const FuelSolveParams* sp,
                                    double* x_dev, double* fbest_dev, int32_t* neval_dev) {
  if (B <= 0) return 0;
  const int need = n_pts + ((mask & FUELGPU_MINTIME) ? 1 : 0);
  if (need > 32)
    return fuel_fail(m, FUELGPU_EUNSUPPORTED, "optimize_batch supports at most 32 lanes (n_pts + dt)");
  const size_t smem = (size_t)WPB * 2 * sp->lbfgs_m * 96 * sizeof(double);
  FUEL_CUDA(m, cudaFuncSetAttribute(optimize_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  optimize_warp_kernel<<<(B + WPB - 1) / WPB, WPB * 32, smem, m->stream>>>(m->g, m->dist, *p, tc_dev, n_pts,
                                                                      mask, B, *sp, x_dev, fbest_dev,
                                                                      neval_dev);
  FUEL_LAUNCHES(m, 1);
  FUEL_CUDA(m, cudaGetLastError());
  return 0;
}
