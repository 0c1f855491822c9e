// Drives the C++ shim classes (include/fuelgpu_shim.hpp) the way FUEL's callers do:
// initMap -> setOccupied/occupancy_buffer_ -> updateESDF3d -> searchFrontiers -> optimize.
// Writes the results to a text file that tests/test_gpu_shim.py compares with the oracle.
#include <cstdio>
#include <cstdlib>

#include "fuelgpu_shim.hpp"

using namespace fast_planner;

int main(int argc, char** argv) {
  const char* out_path = argc > 1 ? argv[1] : "shim_out.txt";
  MapParam mp;
  mp.map_voxel_num_ = Vector3i(48, 40, 24);
  mp.resolution_ = 0.1;
  mp.map_origin_ = Vector3d(-2.4, -2.0, -0.5);
  mp.box_mind_ = Vector3d(-2.2, -1.8, -0.3);
  mp.box_maxd_ = Vector3d(2.2, 1.8, 1.7);
  mp.optimistic_ = true;
  std::shared_ptr<SDFMap> map(new SDFMap);
  try {
    map->initMap(mp);
  } catch (const FuelGpuError& e) {
    std::printf("initMap failed (code %d): %s\n", e.code, e.what());
    return e.code == FUELGPU_ENODEVICE ? 42 : 1;
  }
  // scene (mirrored in tests/test_gpu_shim.py): known free box with an unknown ball in the
  // middle, one occupied wall.  log-odds: unknown = initMap's value, free = clamp_min, occ = 3.0
  const double clamp_min = std::log(0.12 / 0.88);
  for (int x = 0; x < 48; ++x)
    for (int y = 0; y < 40; ++y)
      for (int z = 0; z < 24; ++z) {
        const int a = map->toAddress(x, y, z);
        const bool known = x >= 4 && x < 44 && y >= 4 && y < 36 && z >= 2 && z < 22;
        const int dx = x - 24, dy = y - 20, dz = z - 12;
        const bool ball = dx * dx + dy * dy + 2 * dz * dz < 81;
        const bool wall = x >= 12 && x <= 13 && y >= 8 && y < 30 && z < 18;
        if (known && !ball) map->occupancy_buffer_[a] = wall ? 3.0 : clamp_min;
        if (known && !ball && wall) map->occupancy_buffer_inflate_[a] = 1;
      }
  map->update_min_ = mp.map_origin_;
  map->update_max_ = Vector3d(2.4, 2.0, 1.9);
  map->updateESDF3d();

  std::shared_ptr<EDTEnvironment> env(new EDTEnvironment);
  env->setMap(map);
  FrontierParam fpar;
  fpar.cluster_min_ = 20;
  fpar.cluster_size_xy_ = 1.0;
  FrontierFinder ff(env, fpar);
  ff.searchFrontiers();

  BsplineOptimizer opt;
  opt.setEnvironment(env);
  FuelOptParams p{ 20.0, 10.0, 2.0, 100.0, 0.5, 1.5, 0.3, 0.0, 1.0, 0.7, 2.0, 2.0, 3 };
  const int iters[4] = { 2, 40, 200, 200 };
  opt.setParam(p, iters);
  std::vector<Vector3d> pts;
  for (int i = 0; i < 12; ++i) pts.emplace_back(-1.9 + 0.3 * i, -1.2 + 0.18 * i + ((i % 3) - 1) * 0.1, 0.6 + 0.03 * i);
  double dt = 0.2;
  std::vector<Vector3d> start = { Vector3d((pts[0](0) + 4 * pts[1](0) + pts[2](0)) / 6, (pts[0](1) + 4 * pts[1](1) + pts[2](1)) / 6,
                                           (pts[0](2) + 4 * pts[1](2) + pts[2](2)) / 6),
                                  Vector3d(1.0, 0.6, 0.1), Vector3d(0, 0, 0) };
  std::vector<Vector3d> end = { Vector3d(1.4, 0.8, 0.93) };
  opt.setBoundaryStates(start, end);
  // one combineCost at the initial point, then the solver
  std::vector<double> x, grad;
  for (auto& q : pts)
    for (int k = 0; k < 3; ++k) x.push_back(q(k));
  x.push_back(dt);
  opt.fillTrajConst(pts, dt);
  const int mask = BsplineOptimizer::NORMAL_PHASE | BsplineOptimizer::MINTIME;
  const double f0 = opt.combineCost(x, grad, 12, mask);
  opt.optimize(pts, dt, mask, 1, 1);

  FILE* fo = std::fopen(out_path, "w");
  if (!fo) return 2;
  std::fprintf(fo, "esdf");
  const int probes[6][3] = { { 5, 5, 3 }, { 20, 20, 5 }, { 30, 10, 15 }, { 13, 15, 10 }, { 40, 30, 20 }, { 0, 0, 0 } };
  for (auto& pr : probes) std::fprintf(fo, " %.17g", map->getDistance(Vector3i(pr[0], pr[1], pr[2])));
  Vector3d g;
  const double d = map->getDistWithGrad(Vector3d(0.513, -0.377, 0.642), g);
  std::fprintf(fo, "\nsample %.17g %.17g %.17g %.17g\n", d, g(0), g(1), g(2));
  std::fprintf(fo, "frontiers %zu\n", ff.tmp_frontiers_.size());
  for (auto& f : ff.tmp_frontiers_) {
    std::fprintf(fo, "cluster %zu %zu %.17g %.17g %.17g", f.cells_.size(), f.filtered_cells_.size(), f.average_(0), f.average_(1),
                 f.average_(2));
    long long h = 0;
    for (int a : f.cell_addr_) h = (h * 1000003LL + a) % 2147483647LL;
    std::fprintf(fo, " %lld\n", h);
  }
  std::fprintf(fo, "cost0 %.17g grad0 %.17g %.17g %.17g gdt %.17g\n", f0, grad[0], grad[16], grad[35], grad[36]);
  std::fprintf(fo, "opt %d %.17g %.17g\n", opt.iter_num_, opt.min_cost_, dt);
  std::fclose(fo);
  std::printf("shim smoke ok: %zu frontier clusters, cost %.6g -> %.6g in %d evals\n", ff.tmp_frontiers_.size(), f0,
              opt.min_cost_, opt.iter_num_);
  return 0;
}
