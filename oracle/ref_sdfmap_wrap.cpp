// Drives the reference's own SDFMap (plan_env/src/sdf_map.cpp + raycast.cpp, compiled unmodified from /root/reference
// against oracle/ref_standin) through a C interface so tests can compare the oracle with the real code:
// updateESDF3d, clearAndInflateLocalMap, inputPointCloud, getDistWithGrad.
// TEST INFRASTRUCTURE ONLY; built into oracle/_ref/libfuel_ref.so by oracle/Makefile when /root/reference exists.
#include <plan_env/map_ros.h>
#include <plan_env/sdf_map.h>
#include <stdint.h>
#include <string.h>

using fast_planner::MapROS;
using fast_planner::SDFMap;

extern "C" {

// keys/values: the sdf_map/* ROS parameters (sdf_map.cpp:19-47,80-81)
void* ref_map_create(int32_t n, const char** keys, const double* values) {
  ros::NodeHandle nh;
  for (int i = 0; i < n; ++i) nh.values[keys[i]] = values[i];
  SDFMap* m = new SDFMap();
  m->initMap(nh);
  return m;
}
void ref_map_destroy(void* h) { delete (SDFMap*)h; }

void ref_map_geometry(void* h, int32_t voxel_num[3], double origin[3], double* resolution) {
  SDFMap& m = *(SDFMap*)h;
  for (int i = 0; i < 3; ++i) {
    voxel_num[i] = MapROS::mp(m).map_voxel_num_(i);
    origin[i] = MapROS::mp(m).map_origin_(i);
  }
  *resolution = MapROS::mp(m).resolution_;
}
// raw access to MapData (sdf_map.h:107-125)
double* ref_map_occupancy(void* h) { return MapROS::md(*(SDFMap*)h).occupancy_buffer_.data(); }
char* ref_map_inflate(void* h) { return MapROS::md(*(SDFMap*)h).occupancy_buffer_inflate_.data(); }
double* ref_map_distance(void* h) { return MapROS::md(*(SDFMap*)h).distance_buffer_.data(); }
void ref_map_set_local_bound(void* h, const int32_t lo[3], const int32_t hi[3]) {
  SDFMap& m = *(SDFMap*)h;
  for (int i = 0; i < 3; ++i) {
    MapROS::md(m).local_bound_min_(i) = lo[i];
    MapROS::md(m).local_bound_max_(i) = hi[i];
  }
}
void ref_map_get_local_bound(void* h, int32_t lo[3], int32_t hi[3]) {
  SDFMap& m = *(SDFMap*)h;
  for (int i = 0; i < 3; ++i) {
    lo[i] = MapROS::md(m).local_bound_min_(i);
    hi[i] = MapROS::md(m).local_bound_max_(i);
  }
}
void ref_map_set_modes(void* h, int optimistic, int signed_dist) {
  MapROS::mp(*(SDFMap*)h).optimistic_ = optimistic != 0;
  MapROS::mp(*(SDFMap*)h).signed_dist_ = signed_dist != 0;
}
void ref_map_update_esdf3d(void* h) { ((SDFMap*)h)->updateESDF3d(); }
void ref_map_clear_and_inflate(void* h) { MapROS::clearAndInflate(*(SDFMap*)h); }
void ref_map_input_point_cloud(void* h, const float* xyz, int32_t n, const double cam[3]) {
  pcl::PointCloud<pcl::PointXYZ> cloud;
  cloud.points.resize(n);
  for (int i = 0; i < n; ++i) {
    cloud.points[i].x = xyz[3 * i];
    cloud.points[i].y = xyz[3 * i + 1];
    cloud.points[i].z = xyz[3 * i + 2];
  }
  ((SDFMap*)h)->inputPointCloud(cloud, n, Eigen::Vector3d(cam[0], cam[1], cam[2]));
}
void ref_map_get_updated_box(void* h, double bmin[3], double bmax[3], int reset) {
  Eigen::Vector3d a, b;
  ((SDFMap*)h)->getUpdatedBox(a, b, reset != 0);
  for (int i = 0; i < 3; ++i) bmin[i] = a(i), bmax[i] = b(i);
}
double ref_map_dist_with_grad(void* h, const double pos[3], double grad[3]) {
  Eigen::Vector3d g;
  const double d = ((SDFMap*)h)->getDistWithGrad(Eigen::Vector3d(pos[0], pos[1], pos[2]), g);
  for (int i = 0; i < 3; ++i) grad[i] = g(i);
  return d;
}
int ref_map_get_occupancy(void* h, const int32_t id[3]) { return ((SDFMap*)h)->getOccupancy(Eigen::Vector3i(id[0], id[1], id[2])); }

}  // extern "C"
