// Drives the reference's own FrontierFinder (active_perception/src/frontier_finder.cpp + perception_utils.cpp, compiled
// unmodified from /root/reference against oracle/ref_standin) so tests can compare the oracle's frontier search and
// viewpoint sampling with the real code.  The two third-party algorithms inside that path (pcl::VoxelGrid,
// Eigen::EigenSolver) are the oracle's own reconstructions on both sides (ref_standin/pcl/filters/voxel_grid.h,
// ref_standin/Eigen/Eigenvalues); everything else -- sweep order, expandFrontier, the filters, computeFrontierInfo,
// the recursive split and its list order, sampleViewpoints / countVisibleCells / isNearUnknown, PerceptionUtils -- is
// the reference's compiled code.  TEST INFRASTRUCTURE ONLY; part of oracle/_ref/libfuel_ref.so.
#include <list>
#include <memory>
#include <string>
#include <vector>

#include <plan_env/edt_environment.h>
#include <plan_env/map_ros.h>
#include <plan_env/raycast.h>
#include <plan_env/sdf_map.h>
#include <ros/ros.h>
#include <stdint.h>
#include <string.h>

#include <active_perception/perception_utils.h>
// the search results live in private members (tmp_frontiers_, frontier_flag_, ...): this translation unit -- the test
// wrapper, not the reference sources -- reads them
#define private public
#include <active_perception/frontier_finder.h>
#undef private

using namespace fast_planner;

namespace {
const Frontier& nth(const std::list<Frontier>& l, int i) {
  auto it = l.begin();
  std::advance(it, i);
  return *it;
}
const std::list<Frontier>& which(FrontierFinder& f, int list_id) {
  return list_id == 0 ? f.tmp_frontiers_ : (list_id == 1 ? f.frontiers_ : f.dormant_frontiers_);
}
}  // namespace

extern "C" {

void* ref_ff_create(void* sdf_map_handle, int32_t n, const char** keys, const double* values) {
  ros::NodeHandle nh;
  for (int i = 0; i < n; ++i) nh.values[keys[i]] = values[i];
  EDTEnvironment::Ptr env(new EDTEnvironment);
  env->sdf_map_ = std::shared_ptr<SDFMap>((SDFMap*)sdf_map_handle, [](SDFMap*) {});
  return new FrontierFinder(env, nh);
}
void ref_ff_destroy(void* h) { delete (FrontierFinder*)h; }

// md_->update_min_/max_ as inputPointCloud would leave them (searchFrontiers reads them through getUpdatedBox)
void ref_map_set_updated_box(void* map, const double bmin[3], const double bmax[3]) {
  SDFMap& m = *(SDFMap*)map;
  for (int i = 0; i < 3; ++i) {
    MapROS::md(m).update_min_(i) = bmin[i];
    MapROS::md(m).update_max_(i) = bmax[i];
  }
}

void ref_ff_search(void* h) { ((FrontierFinder*)h)->searchFrontiers(); }
void ref_ff_compute_to_visit(void* h) { ((FrontierFinder*)h)->computeFrontiersToVisit(); }
int32_t ref_ff_is_covered(void* h) { return ((FrontierFinder*)h)->isFrontierCovered() ? 1 : 0; }
char* ref_ff_flags(void* h) { return ((FrontierFinder*)h)->frontier_flag_.data(); }
int32_t ref_ff_removed_ids(void* h, int32_t* out, int32_t max) {  // removed_ids_ of the last searchFrontiers (:75-84)
  const std::vector<int>& r = ((FrontierFinder*)h)->removed_ids_;
  for (size_t i = 0; i < r.size() && (int)i < max; ++i) out[i] = r[i];
  return (int32_t)r.size();
}

// list_id: 0 tmp_frontiers_, 1 frontiers_, 2 dormant_frontiers_
int32_t ref_ff_count(void* h, int32_t list_id) { return (int32_t)which(*(FrontierFinder*)h, list_id).size(); }
void ref_ff_sizes(void* h, int32_t list_id, int32_t i, int32_t* n_cells, int32_t* n_filtered, int32_t* n_views, int32_t* id) {
  const Frontier& f = nth(which(*(FrontierFinder*)h, list_id), i);
  *n_cells = (int32_t)f.cells_.size();
  *n_filtered = (int32_t)f.filtered_cells_.size();
  *n_views = (int32_t)f.viewpoints_.size();
  *id = f.id_;
}
void ref_ff_get(void* h, int32_t list_id, int32_t i, int32_t* cell_addr, double* filtered, double avg[3], double bmin[3],
                double bmax[3], double* view_pos, double* view_yaw, int32_t* view_visib) {
  FrontierFinder& ff = *(FrontierFinder*)h;
  const Frontier& f = nth(which(ff, list_id), i);
  for (size_t k = 0; k < f.cells_.size(); ++k) {
    Eigen::Vector3i idx;
    ff.edt_env_->sdf_map_->posToIndex(f.cells_[k], idx);
    cell_addr[k] = ff.edt_env_->sdf_map_->toAddress(idx);
  }
  for (size_t k = 0; k < f.filtered_cells_.size(); ++k)
    for (int a = 0; a < 3; ++a) filtered[3 * k + a] = f.filtered_cells_[k](a);
  for (int a = 0; a < 3; ++a) avg[a] = f.average_(a), bmin[a] = f.box_min_(a), bmax[a] = f.box_max_(a);
  for (size_t k = 0; k < f.viewpoints_.size(); ++k) {
    for (int a = 0; a < 3; ++a) view_pos[3 * k + a] = f.viewpoints_[k].pos_(a);
    view_yaw[k] = f.viewpoints_[k].yaw_;
    view_visib[k] = f.viewpoints_[k].visib_num_;
  }
}

}  // extern "C"
