// Drives the reference's own BsplineOptimizer (bspline_opt/src/bspline_optimizer.cpp, compiled unmodified from
// /root/reference against oracle/ref_standin) so tests can compare the oracle's combineCost with the real code.
// optimize() is called with the NLopt stand-in (ref_standin/nlopt.hpp), which evaluates the reference's objective
// -- costFunction -> combineCost -- at the start point and at the probe points, and records what optimize() passed to
// NLopt (clamped start, bounds).  TEST INFRASTRUCTURE ONLY; part of oracle/_ref/libfuel_ref.so.
#include <bspline_opt/bspline_optimizer.h>
#include <nlopt.hpp>
#include <plan_env/map_ros.h>
#include <stdint.h>
#include <string.h>

using namespace fast_planner;

extern "C" {

void* ref_opt_create(void* sdf_map_handle, int32_t n, const char** keys, const double* values) {
  ros::NodeHandle nh;
  for (int i = 0; i < n; ++i) nh.values[keys[i]] = values[i];
  BsplineOptimizer* o = new BsplineOptimizer();
  o->setParam(nh);
  EDTEnvironment::Ptr env(new EDTEnvironment);
  // the map object stays owned by the Python side: aliasing shared_ptr with a no-op deleter
  env->sdf_map_ = std::shared_ptr<SDFMap>((SDFMap*)sdf_map_handle, [](SDFMap*) {});
  o->setEnvironment(env);
  return o;
}
void ref_opt_destroy(void* h) { delete (BsplineOptimizer*)h; }

// setViewConstraint (bspline_optimizer.cpp:91-93): only pt_, dir_, idx_ are read by calcViewCost (:477-502)
void ref_opt_set_view(void* h, const double* pt, const double* dir, int32_t idx) {
  ViewConstraint vc;
  vc.pt_ = Eigen::Vector3d(pt[0], pt[1], pt[2]);
  vc.pc_ = vc.pt_;
  vc.dir_ = Eigen::Vector3d(dir[0], dir[1], dir[2]);
  vc.pcons_ = vc.pt_;
  vc.idx_ = idx;
  ((BsplineOptimizer*)h)->setViewConstraint(vc);
}

// One optimize(points, dt, cost_function, max_num_id, max_time_id) call (bspline_optimizer.cpp:110-163).
//   ctrl [n_pts][3], start [n_start<=3][3], end [n_end<=3][3], guide [n_guide][3], waypts [n_wp][3] + idx, time_lb
//   probes [n_probe][nvar]: extra points at which the objective is evaluated
// out: f[1+n_probe], grad[1+n_probe][nvar], x0[nvar] (the start point after the reference's clamp), lb/ub[nvar]
int32_t ref_opt_evaluate(void* h, int32_t n_pts, const double* ctrl, double dt, int32_t cost_function, const double* start,
                         int32_t n_start, const double* end, int32_t n_end, const double* guide, int32_t n_guide,
                         const double* waypts, const int32_t* waypt_idx, int32_t n_wp, double time_lb,
                         const double* probes, int32_t n_probe, double* f, double* grad, double* x0, double* lb, double* ub) {
  BsplineOptimizer& o = *(BsplineOptimizer*)h;
  auto vecs = [](const double* p, int n) {
    std::vector<Eigen::Vector3d> v;
    for (int i = 0; i < n; ++i) v.emplace_back(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
  };
  o.setBoundaryStates(vecs(start, n_start), vecs(end, n_end));
  if (n_guide > 0) o.setGuidePath(vecs(guide, n_guide));
  if (n_wp > 0) o.setWaypoints(vecs(waypts, n_wp), std::vector<int>(waypt_idx, waypt_idx + n_wp));
  o.setTimeLowerBound(time_lb);
  Eigen::MatrixXd pts(n_pts, 3);
  for (int i = 0; i < n_pts; ++i)
    for (int j = 0; j < 3; ++j) pts(i, j) = ctrl[3 * i + j];
  const bool opt_time = (cost_function & BsplineOptimizer::MINTIME) != 0;
  const int nvar = 3 * n_pts + (opt_time ? 1 : 0);
  nlopt::Recorder& r = nlopt::recorder();
  r.probes.clear();
  for (int p = 0; p < n_probe; ++p) r.probes.emplace_back(probes + (size_t)p * nvar, probes + (size_t)(p + 1) * nvar);
  r.lb.clear();
  r.ub.clear();
  o.optimize(pts, dt, cost_function, 1, 1);
  if ((int)r.f.size() != 1 + n_probe) return -1;
  for (int p = 0; p <= n_probe; ++p) {
    f[p] = r.f[p];
    memcpy(grad + (size_t)p * nvar, r.grad[p].data(), sizeof(double) * nvar);
  }
  memcpy(x0, r.x0.data(), sizeof(double) * nvar);
  if ((int)r.lb.size() == nvar) {
    memcpy(lb, r.lb.data(), sizeof(double) * nvar);
    memcpy(ub, r.ub.data(), sizeof(double) * nvar);
  }
  return nvar;
}

// B independent trajectories, K objective evaluations each (the start point + K-1 probe points), `threads` host threads
// with one BsplineOptimizer each (the reference runs up to 10 optimisers in parallel threads, planner_manager.cpp:444-453).
// The timed CPU baseline of bench.py: the reference's own combineCost, called the way its NLopt loop calls it.
//   ctrl [B][n_pts][3], dt [B], start [B][3][3], end_pos [B][3], probes [B][K-1][nvar]  ->  f [B][K]
int32_t ref_opt_evaluate_batch(void* sdf_map_handle, int32_t n_keys, const char** keys, const double* values, int32_t B,
                               int32_t n_pts, const double* ctrl, const double* dt, int32_t cost_function, const double* start,
                               const double* end_pos, const double* probes, int32_t K, int32_t threads, double* f) {
  const bool opt_time = (cost_function & BsplineOptimizer::MINTIME) != 0;
  const int nvar = 3 * n_pts + (opt_time ? 1 : 0);
  int bad = 0;
#pragma omp parallel num_threads(threads > 0 ? threads : 1) reduction(+ : bad)
  {
    ros::NodeHandle nh;
    for (int i = 0; i < n_keys; ++i) nh.values[keys[i]] = values[i];
    BsplineOptimizer o;
    o.setParam(nh);
    EDTEnvironment::Ptr env(new EDTEnvironment);
    env->sdf_map_ = std::shared_ptr<SDFMap>((SDFMap*)sdf_map_handle, [](SDFMap*) {});
    o.setEnvironment(env);
    nlopt::Recorder& r = nlopt::recorder();
#pragma omp for schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
      std::vector<Eigen::Vector3d> st, en;
      for (int i = 0; i < 3; ++i) st.emplace_back(start[9 * b + 3 * i], start[9 * b + 3 * i + 1], start[9 * b + 3 * i + 2]);
      en.emplace_back(end_pos[3 * b], end_pos[3 * b + 1], end_pos[3 * b + 2]);
      o.setBoundaryStates(st, en);
      Eigen::MatrixXd pts(n_pts, 3);
      for (int i = 0; i < n_pts; ++i)
        for (int j = 0; j < 3; ++j) pts(i, j) = ctrl[((size_t)b * n_pts + i) * 3 + j];
      r.probes.clear();
      for (int p = 0; p < K - 1; ++p) {
        const double* q = probes + ((size_t)b * (K - 1) + p) * nvar;
        r.probes.emplace_back(q, q + nvar);
      }
      double dtb = dt[b];
      o.optimize(pts, dtb, cost_function, 1, 1);
      if ((int)r.f.size() != K) {
        ++bad;
        continue;
      }
      for (int p = 0; p < K; ++p) f[(size_t)b * K + p] = r.f[p];
    }
  }
  return bad;
}

}  // extern "C"
