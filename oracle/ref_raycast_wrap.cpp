// Drives the reference's own RayCaster (plan_env/src/raycast.cpp, compiled unmodified from /root/reference against
// oracle/ref_standin) the way SDFMap::inputPointCloud (sdf_map.cpp:307-311) and FrontierFinder::countVisibleCells
// (frontier_finder.cpp:743-751) do: setParams once, input(start, end), nextId until it returns false.
// TEST INFRASTRUCTURE ONLY; built into oracle/_ref/libfuel_ref.so by oracle/Makefile when /root/reference exists.
#include <plan_env/raycast.h>
#include <stdint.h>

extern "C" int32_t ref_raycast_ids(double resolution, const double origin[3], const double start[3], const double end[3],
                                   int32_t* ids /*[max][3]*/, int32_t max) {
  RayCaster rc;
  rc.setParams(resolution, Eigen::Vector3d(origin[0], origin[1], origin[2]));
  rc.input(Eigen::Vector3d(start[0], start[1], start[2]), Eigen::Vector3d(end[0], end[1], end[2]));
  Eigen::Vector3i idx;
  int32_t n = 0;
  while (n < max && rc.nextId(idx)) {
    ids[3 * n] = idx(0), ids[3 * n + 1] = idx(1), ids[3 * n + 2] = idx(2);
    ++n;
  }
  return n;
}

extern "C" double ref_intbound(double s, double ds) { return intbound(s, ds); }
extern "C" double ref_mod(double v, double m) { return mod(v, m); }
