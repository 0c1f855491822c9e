// fuelgpu_shim.hpp -- C++ host-side mirror of the reference classes on the hot path, over the C ABI.
//
// Same class names, method names, argument meaning and sentinel behaviour as the reference
// (file:line under /root/reference/fuel_planner/):
//   fast_planner::SDFMap           plan_env/include/plan_env/sdf_map.h:27-84
//   fast_planner::EDTEnvironment   plan_env/include/plan_env/edt_environment.h:21-51
//   fast_planner::FrontierFinder   active_perception/include/active_perception/frontier_finder.h:53-131
//   fast_planner::BsplineOptimizer bspline_opt/include/bspline_opt/bspline_optimizer.h:20-145
// Only the members the hot path needs are mirrored; everything voxel-scale is a
// call into libfuelgpu (hand-written sm_100a CUDA).  The reference uses Eigen::Vector3d/3i and a
// ros::NodeHandle for parameters; define FUELGPU_SHIM_USE_EIGEN before including this header to
// get the Eigen types, otherwise a minimal 3-vector with the same element access is used (Eigen
// is not installed in the build image).  Parameters arrive as plain structs instead of ROS params.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fuelgpu.h"

#ifdef FUELGPU_SHIM_USE_EIGEN
#include <Eigen/Eigen>
namespace fuelgpu_shim {
using Vector3d = Eigen::Vector3d;
using Vector3i = Eigen::Vector3i;
}  // namespace fuelgpu_shim
#else
namespace fuelgpu_shim {
template <typename T>
struct Vec3 {
  T v[3];
  Vec3() : v{ 0, 0, 0 } {}
  Vec3(T a, T b, T c) : v{ a, b, c } {}
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T* data() { return v; }
  const T* data() const { return v; }
};
using Vector3d = Vec3<double>;
using Vector3i = Vec3<int32_t>;
}  // namespace fuelgpu_shim
#endif

namespace fast_planner {
using fuelgpu_shim::Vector3d;
using fuelgpu_shim::Vector3i;

struct FuelGpuError : std::runtime_error {
  int code;
  FuelGpuError(int c, const char* msg) : std::runtime_error(msg ? msg : ""), code(c) {}
};
inline void fuelgpu_check(int rc, const FuelMap* h) {
  if (rc != FUELGPU_OK) throw FuelGpuError(rc, fuelgpu_last_error(h));
}

// ---- SDFMap (sdf_map.h:27-84) --------------------------------------------------------------
struct MapParam {  // the ROS parameters of initMap (sdf_map.cpp:19-47,79-84) as a struct
  Vector3i map_voxel_num_;
  double resolution_ = 0.1;
  Vector3d map_origin_;
  Vector3d map_size_;  // sdf_map/map_size_x,y,z; a zero component means n * resolution (sdf_map.cpp:34-39)
  Vector3d box_mind_, box_maxd_;
  bool optimistic_ = false, signed_dist_ = false;
  double p_min_ = 0.12, p_occ_ = 0.80;
  double default_dist_ = 0.0;
  int device = 0;
};

class SDFMap {
public:
  enum OCCUPANCY { UNKNOWN, FREE, OCCUPIED };
  typedef std::shared_ptr<SDFMap> Ptr;

  SDFMap() {}
  ~SDFMap() {
    if (gpu_) fuelgpu_map_destroy(gpu_);
  }
  SDFMap(const SDFMap&) = delete;
  SDFMap& operator=(const SDFMap&) = delete;

  // initMap (sdf_map.cpp:12-93)
  void initMap(const MapParam& p) {
    mp_ = p;
    resolution_inv_ = 1 / mp_.resolution_;
    auto logit = [](double x) { return std::log(x / (1 - x)); };
    clamp_min_log_ = logit(mp_.p_min_);
    min_occupancy_log_ = logit(mp_.p_occ_);
    const size_t n = (size_t)mp_.map_voxel_num_(0) * mp_.map_voxel_num_(1) * mp_.map_voxel_num_(2);
    occupancy_buffer_.assign(n, clamp_min_log_ - 0.01);  // unknown_flag_, sdf_map.cpp:56,64
    occupancy_buffer_inflate_.assign(n, 0);
    distance_buffer_.assign(n, mp_.default_dist_);
    for (int i = 0; i < 3; ++i) {
      map_max_boundary_(i) = mp_.map_origin_(i) + (mp_.map_size_(i) > 0.0 ? mp_.map_size_(i) : mp_.map_voxel_num_(i) * mp_.resolution_);
      local_bound_min_(i) = 0;
      local_bound_max_(i) = mp_.map_voxel_num_(i) - 1;
    }
    posToIndex(mp_.box_mind_, box_min_);
    posToIndex(mp_.box_maxd_, box_max_);
    FuelGridDesc d;
    for (int i = 0; i < 3; ++i) {
      d.n[i] = mp_.map_voxel_num_(i);
      d.origin[i] = mp_.map_origin_(i);
      d.box_mind[i] = mp_.box_mind_(i);
      d.box_maxd[i] = mp_.box_maxd_(i);
      d.map_size[i] = mp_.map_size_(i);
    }
    d.resolution = mp_.resolution_;
    fuelgpu_check(fuelgpu_map_create(&d, mp_.device, &gpu_), nullptr);
  }

  // index helpers (sdf_map.h:127-192)
  void posToIndex(const Vector3d& pos, Vector3i& id) const {
    for (int i = 0; i < 3; ++i) id(i) = (int32_t)std::floor((pos(i) - mp_.map_origin_(i)) * resolution_inv_);
  }
  void indexToPos(const Vector3i& id, Vector3d& pos) const {
    for (int i = 0; i < 3; ++i) pos(i) = (id(i) + 0.5) * mp_.resolution_ + mp_.map_origin_(i);
  }
  void boundIndex(Vector3i& id) const {
    for (int i = 0; i < 3; ++i) id(i) = std::max(std::min(id(i), mp_.map_voxel_num_(i) - 1), 0);
  }
  int toAddress(const Vector3i& id) const { return toAddress(id[0], id[1], id[2]); }
  int toAddress(int x, int y, int z) const {
    return x * mp_.map_voxel_num_(1) * mp_.map_voxel_num_(2) + y * mp_.map_voxel_num_(2) + z;
  }
  bool isInMap(const Vector3d& pos) const {
    for (int i = 0; i < 3; ++i)
      if (pos(i) < mp_.map_origin_(i) + 1e-4 || pos(i) > map_max_boundary_(i) - 1e-4) return false;
    return true;
  }
  bool isInMap(const Vector3i& idx) const {
    for (int i = 0; i < 3; ++i)
      if (idx(i) < 0 || idx(i) > mp_.map_voxel_num_(i) - 1) return false;
    return true;
  }
  bool isInBox(const Vector3i& id) const {
    for (int i = 0; i < 3; ++i)
      if (id[i] < box_min_[i] || id[i] >= box_max_[i]) return false;
    return true;
  }
  bool isInBox(const Vector3d& pos) const {
    for (int i = 0; i < 3; ++i)
      if (pos[i] <= mp_.box_mind_[i] || pos[i] >= mp_.box_maxd_[i]) return false;
    return true;
  }
  int getOccupancy(const Vector3i& id) const {
    if (!isInMap(id)) return -1;
    const double occ = occupancy_buffer_[toAddress(id)];
    if (occ < clamp_min_log_ - 1e-3) return UNKNOWN;
    if (occ > min_occupancy_log_) return OCCUPIED;
    return FREE;
  }
  int getOccupancy(const Vector3d& pos) const {
    Vector3i id;
    posToIndex(pos, id);
    return getOccupancy(id);
  }
  int getInflateOccupancy(const Vector3i& id) const {
    if (!isInMap(id)) return -1;
    return (int)occupancy_buffer_inflate_[toAddress(id)];
  }
  void setOccupied(const Vector3d& pos, const int& occ = 1) {
    if (!isInMap(pos)) return;
    Vector3i id;
    posToIndex(pos, id);
    occupancy_buffer_inflate_[toAddress(id)] = (int8_t)occ;
  }
  void resetBuffer() {  // sdf_map.cpp:95-99
    std::fill(occupancy_buffer_inflate_.begin(), occupancy_buffer_inflate_.end(), 0);
    std::fill(distance_buffer_.begin(), distance_buffer_.end(), mp_.default_dist_);
    for (int i = 0; i < 3; ++i) {
      local_bound_min_(i) = 0;
      local_bound_max_(i) = mp_.map_voxel_num_(i) - 1;
    }
  }
  double getDistance(const Vector3i& id) const {
    if (!isInMap(id)) return -1;
    return distance_buffer_[toAddress(id)];
  }
  double getDistance(const Vector3d& pos) const {
    Vector3i id;
    posToIndex(pos, id);
    return getDistance(id);
  }
  double getResolution() const { return mp_.resolution_; }
  int getVoxelNum() const { return mp_.map_voxel_num_[0] * mp_.map_voxel_num_[1] * mp_.map_voxel_num_[2]; }
  void getRegion(Vector3d& ori, Vector3d& size) const {
    ori = mp_.map_origin_;
    for (int i = 0; i < 3; ++i) size(i) = mp_.map_voxel_num_(i) * mp_.resolution_;
  }
  void getBox(Vector3d& bmin, Vector3d& bmax) const {
    bmin = mp_.box_mind_;
    bmax = mp_.box_maxd_;
  }
  void getUpdatedBox(Vector3d& bmin, Vector3d& bmax, bool reset = false) {
    if (fused_) {  // the box is accumulated by the device fusion calls (sdf_map.cpp:321-324)
      fuelgpu_check(fuelgpu_map_get_updated_box(gpu_, bmin.data(), bmax.data(), reset ? 1 : 0), gpu_);
      return;
    }
    bmin = update_min_;
    bmax = update_max_;
    if (reset) reset_updated_box_ = true;
  }

  // inputPointCloud (sdf_map.cpp:259-345): fuses one frame into the device-resident log-odds volume and
  // sets local_bound_min_/max_.  `points` = cloud.points.data() of a pcl::PointCloud<pcl::PointXYZ>
  // (point_stride 4) or packed xyz (point_stride 3).  From here on the occupancy lives on the device:
  // updateESDF3d() no longer uploads occupancy_buffer_.
  FuelFusionParams fusion_ = { 0.65, 0.35, 0.12, 0.90, 0.80, 4.5, 0.5 };  // algorithm.xml:39-50
  void inputPointCloud(const float* points, int point_num, const Vector3d& camera_pos, int point_stride = 4) {
    if (point_num == 0) return;
    fuelgpu_check(fuelgpu_map_input_point_cloud(gpu_, points, point_num, point_stride, camera_pos.data(), &fusion_,
                                                local_bound_min_.data(), local_bound_max_.data()),
                  gpu_);
    fused_ = true;
  }
  // MapROS::proessDepthImage + inputPointCloud (map_ros.cpp:139-140,176-215) as one device call; camera_R is
  // camera_q_.toRotationMatrix() in row-major order.  Returns proj_points_cnt.
  FuelCameraParams camera_ = { 387.229248046875, 387.229248046875, 321.04638671875, 243.44969177246094, 1000.0, 5.0, 0.2, 2, 2 };
  int inputDepthImage(const uint16_t* depth, int rows, int cols, const double camera_R[9], const Vector3d& camera_pos) {
    int32_t cnt = 0;
    fuelgpu_check(fuelgpu_map_input_depth_image(gpu_, depth, rows, cols, &camera_, camera_R, camera_pos.data(), &fusion_,
                                                local_bound_min_.data(), local_bound_max_.data(), &cnt),
                  gpu_);
    if (cnt > 0) fused_ = true;
    return cnt;
  }

  // updateESDF3d (sdf_map.cpp:152-241): occupancy H2D for the x-slabs of the local box, the three
  // sweeps on the device, and the fp64 host mirror the scattered getDistance() readers use.
  void updateESDF3d() {
    if (!fused_)
      fuelgpu_check(fuelgpu_map_upload_occupancy(gpu_, occupancy_buffer_inflate_.data(), occupancy_buffer_.data(), nullptr,
                                               clamp_min_log_, min_occupancy_log_, local_bound_min_.data(),
                                               local_bound_max_.data()),
                  gpu_);
    const int flags = (mp_.optimistic_ ? FUELGPU_ESDF_OPTIMISTIC : 0) | (mp_.signed_dist_ ? FUELGPU_ESDF_SIGNED : 0);
    fuelgpu_check(fuelgpu_esdf_update(gpu_, local_bound_min_.data(), local_bound_max_.data(), flags), gpu_);
    fuelgpu_check(fuelgpu_esdf_download(gpu_, local_bound_min_.data(), local_bound_max_.data(), nullptr,
                                        distance_buffer_.data()),
                  gpu_);
  }
  // clearAndInflateLocalMap (sdf_map.cpp:364-472) on the resident occupancy byte (the byte must have been
  // uploaded, e.g. by the previous updateESDF3d); refreshes occupancy_buffer_inflate_
  void clearAndInflateLocalMap(double obstacles_inflation, double virtual_ceil_height) {
    const int inf_step = (int)std::ceil(obstacles_inflation / mp_.resolution_);
    const int ceil_id = virtual_ceil_height > -0.5
                            ? (int)std::floor((virtual_ceil_height - mp_.map_origin_(2)) * resolution_inv_)
                            : -1;
    fuelgpu_check(fuelgpu_map_inflate(gpu_, local_bound_min_.data(), local_bound_max_.data(), inf_step, ceil_id), gpu_);
    fuelgpu_check(fuelgpu_map_download_occupancy(gpu_, occupancy_buffer_inflate_.data(), nullptr), gpu_);
  }
  // getDistWithGrad (sdf_map.cpp:497-536) on the device ESDF
  double getDistWithGrad(const Vector3d& pos, Vector3d& grad) {
    double d = 0;
    fuelgpu_check(fuelgpu_esdf_sample(gpu_, 1, pos.data(), &d, grad.data()), gpu_);
    return d;
  }

  FuelMap* gpu() const { return gpu_; }
  // MapData (sdf_map.h:107-125): the reference's friend MapROS pokes these directly
  std::vector<double> occupancy_buffer_;
  std::vector<int8_t> occupancy_buffer_inflate_;
  std::vector<double> distance_buffer_;
  Vector3i local_bound_min_, local_bound_max_;
  Vector3d update_min_, update_max_;
  bool reset_updated_box_ = true;
  bool fused_ = false;
  MapParam mp_;

private:
  double resolution_inv_ = 10.0, clamp_min_log_ = 0, min_occupancy_log_ = 0;
  Vector3d map_max_boundary_;
  Vector3i box_min_, box_max_;
  FuelMap* gpu_ = nullptr;
};

// ---- EDTEnvironment (edt_environment.h:21-51) --------------------------------------------------
class EDTEnvironment {
public:
  typedef std::shared_ptr<EDTEnvironment> Ptr;
  void setMap(std::shared_ptr<SDFMap>& map) {
    sdf_map_ = map;
    resolution_inv_ = 1 / sdf_map_->getResolution();
  }
  // edt_environment.cpp:78-87: pass-through, `time` unused
  void evaluateEDTWithGrad(const Vector3d& pos, double /*time*/, double& dist, Vector3d& grad) {
    dist = sdf_map_->getDistWithGrad(pos, grad);
  }
  double evaluateCoarseEDT(Vector3d& pos, double /*time*/) { return sdf_map_->getDistance(pos); }
  std::shared_ptr<SDFMap> sdf_map_;

private:
  double resolution_inv_ = 10.0;
};

// ---- FrontierFinder (frontier_finder.h:25-131) ---------------------------------------------------
struct Viewpoint {  // frontier_finder.h:25-31
  Vector3d pos_;
  double yaw_;
  int visib_num_;
};

struct Frontier {  // frontier_finder.h:34-51 (paths_/costs_ belong to out-of-scope stages)
  std::vector<Vector3d> cells_;
  std::vector<Viewpoint> viewpoints_;
  std::vector<Vector3d> filtered_cells_;
  Vector3d average_;
  int id_ = -1;
  Vector3d box_min_, box_max_;
  std::vector<int32_t> cell_addr_;  // toAddress of every cell (what the C ABI speaks)
};

struct FrontierParam {  // frontier_finder.cpp:29-40
  int cluster_min_ = 100;
  double cluster_size_xy_ = 2.0;
  int down_sample_ = 3;
};

class FrontierFinder {
public:
  FrontierFinder(const std::shared_ptr<EDTEnvironment>& edt, const FrontierParam& p) : edt_env_(edt), p_(p) {
    fuelgpu_check(fuelgpu_frontier_reset_flags(gpu()), gpu());  // frontier_flag_ fill, :26-27
  }

  // searchFrontiers (frontier_finder.cpp:54-121)
  void searchFrontiers() {
    tmp_frontiers_.clear();
    Vector3d update_min, update_max;
    edt_env_->sdf_map_->getUpdatedBox(update_min, update_max, true);
    removed_ids_.clear();
    removeChanged(frontiers_, update_min, update_max, true);           // :71-86
    removeChanged(dormant_frontiers_, update_min, update_max, false);  // :87-92
    FuelFrontierParams fp{ p_.cluster_min_, p_.cluster_size_xy_, p_.down_sample_, 0.4 };
    int32_t nc = 0, ncell = 0, nf = 0;
    fuelgpu_check(fuelgpu_frontier_search(gpu(), update_min.data(), update_max.data(), &fp, &nc, &ncell, &nf), gpu());
    std::vector<int32_t> co(nc + 1), ca(ncell), fo(nc + 1);
    std::vector<double> filt(3 * (size_t)nf), avg(3 * (size_t)nc), bmin(3 * (size_t)nc), bmax(3 * (size_t)nc);
    fuelgpu_check(fuelgpu_frontier_fetch(gpu(), co.data(), ca.data(), fo.data(), filt.data(), avg.data(), bmin.data(),
                                         bmax.data()),
                  gpu());
    const SDFMap& m = *edt_env_->sdf_map_;
    const int ny = m.mp_.map_voxel_num_(1), nz = m.mp_.map_voxel_num_(2);
    for (int c = 0; c < nc; ++c) {
      Frontier f;
      for (int i = co[c]; i < co[c + 1]; ++i) {
        const int a = ca[i];
        Vector3i id(a / (ny * nz), (a / nz) % ny, a % nz);
        Vector3d pos;
        m.indexToPos(id, pos);
        f.cells_.push_back(pos);
        f.cell_addr_.push_back(a);
      }
      for (int i = fo[c]; i < fo[c + 1]; ++i) f.filtered_cells_.emplace_back(filt[3 * i], filt[3 * i + 1], filt[3 * i + 2]);
      for (int k = 0; k < 3; ++k) {
        f.average_(k) = avg[3 * c + k];
        f.box_min_(k) = bmin[3 * c + k];
        f.box_max_(k) = bmax[3 * c + k];
      }
      tmp_frontiers_.push_back(f);
    }
  }
  // computeFrontiersToVisit (frontier_finder.cpp:392-423): sampleViewpoints (:662-695) of every new cluster in one
  // device call; clusters without a qualified viewpoint go dormant.
  FuelViewParams view_ = { 1.5, 2.5, 3, 15 * 3.1415926 / 180.0, 0.21, 0.56125, 0.69222, 0.68901, 4.5 };  // algorithm.xml:106-121
  int min_visib_num_ = 15;
  void computeFrontiersToVisit() {
    first_new_ftr_ = frontiers_.end();
    const int n = (int)tmp_frontiers_.size(), nc = fuelgpu_viewpoint_candidate_count(&view_);
    std::vector<int32_t> fo(n + 1, 0);
    std::vector<double> filt, avg;
    int i = 0;
    for (auto& f : tmp_frontiers_) {
      for (auto& c : f.filtered_cells_) filt.insert(filt.end(), { c(0), c(1), c(2) });
      avg.insert(avg.end(), { f.average_(0), f.average_(1), f.average_(2) });
      fo[i + 1] = fo[i] + (int)f.filtered_cells_.size();
      ++i;
    }
    std::vector<double> pos(3 * (size_t)n * nc), yaw((size_t)n * nc);
    std::vector<int32_t> vis((size_t)n * nc);
    if (n > 0)
      fuelgpu_check(fuelgpu_frontier_sample_viewpoints(gpu(), n, fo.data(), filt.data(), avg.data(), &view_, nc, pos.data(),
                                                       yaw.data(), vis.data()),
                    gpu());
    i = 0;
    for (auto& f : tmp_frontiers_) {
      for (int k = 0; k < nc; ++k) {
        const size_t o = (size_t)i * nc + k;
        if (vis[o] > min_visib_num_) {  // :688
          Viewpoint vp;
          for (int d = 0; d < 3; ++d) vp.pos_(d) = pos[3 * o + d];
          vp.yaw_ = yaw[o];
          vp.visib_num_ = vis[o];
          f.viewpoints_.push_back(vp);
        }
      }
      if (!f.viewpoints_.empty()) {
        auto inserted = frontiers_.insert(frontiers_.end(), f);
        std::sort(inserted->viewpoints_.begin(), inserted->viewpoints_.end(),
                  [](const Viewpoint& a, const Viewpoint& b) { return a.visib_num_ > b.visib_num_; });  // :404-406
        if (first_new_ftr_ == frontiers_.end()) first_new_ftr_ = inserted;
      } else
        dormant_frontiers_.push_back(f);
      ++i;
    }
    int idx = 0;
    for (auto& ft : frontiers_) ft.id_ = idx++;
  }
  std::list<Frontier>::iterator first_new_ftr_;

  void getFrontiers(std::vector<std::vector<Vector3d>>& clusters) const {
    clusters.clear();
    for (auto& f : frontiers_) clusters.push_back(f.cells_);
  }
  void getDormantFrontiers(std::vector<std::vector<Vector3d>>& clusters) const {
    clusters.clear();
    for (auto& f : dormant_frontiers_) clusters.push_back(f.cells_);
  }

  std::list<Frontier> frontiers_, dormant_frontiers_, tmp_frontiers_;
  std::vector<int> removed_ids_;

private:
  FuelMap* gpu() const { return edt_env_->sdf_map_->gpu(); }
  static bool haveOverlap(const Vector3d& min1, const Vector3d& max1, const Vector3d& min2, const Vector3d& max2) {
    for (int i = 0; i < 3; ++i) {  // frontier_finder.cpp:353-363
      const double bmin = std::max(min1[i], min2[i]), bmax = std::min(max1[i], max2[i]);
      if (bmin > bmax + 1e-3) return false;
    }
    return true;
  }
  void removeChanged(std::list<Frontier>& ftrs, const Vector3d& umin, const Vector3d& umax, bool record) {
    std::vector<std::list<Frontier>::iterator> cand;
    std::vector<int32_t> offs(1, 0), addr;
    for (auto it = ftrs.begin(); it != ftrs.end(); ++it)
      if (haveOverlap(it->box_min_, it->box_max_, umin, umax)) {
        cand.push_back(it);
        addr.insert(addr.end(), it->cell_addr_.begin(), it->cell_addr_.end());
        offs.push_back((int32_t)addr.size());
      }
    std::vector<uint8_t> changed(cand.size(), 0);
    if (!cand.empty())  // isFrontierChanged (:365-372) for all candidates in one call
      fuelgpu_check(fuelgpu_frontier_is_changed(gpu(), (int32_t)cand.size(), offs.data(), addr.data(), changed.data()), gpu());
    std::vector<int32_t> clear;
    size_t ci = 0;
    int rmv_idx = 0;
    for (auto it = ftrs.begin(); it != ftrs.end();) {
      const bool is_cand = ci < cand.size() && cand[ci] == it;
      if (is_cand && changed[ci]) {
        clear.insert(clear.end(), it->cell_addr_.begin(), it->cell_addr_.end());  // resetFlag (:62-69)
        it = ftrs.erase(it);
        if (record) removed_ids_.push_back(rmv_idx);
      } else {
        ++rmv_idx;
        ++it;
      }
      if (is_cand) ++ci;
    }
    if (!clear.empty()) fuelgpu_check(fuelgpu_frontier_clear_flags(gpu(), (int32_t)clear.size(), clear.data()), gpu());
  }
  std::shared_ptr<EDTEnvironment> edt_env_;
  FrontierParam p_;
};

// ---- BsplineOptimizer (bspline_optimizer.h:20-145) -------------------------------------------------
class BsplineOptimizer {
public:
  static const int SMOOTHNESS = FUELGPU_SMOOTHNESS, DISTANCE = FUELGPU_DISTANCE, FEASIBILITY = FUELGPU_FEASIBILITY,
                   START = FUELGPU_START, END = FUELGPU_END, GUIDE = FUELGPU_GUIDE, WAYPOINTS = FUELGPU_WAYPOINTS,
                   VIEWCONS = FUELGPU_VIEWCONS, MINTIME = FUELGPU_MINTIME;
  static const int GUIDE_PHASE = SMOOTHNESS | GUIDE | START | END;                      // bspline_optimizer.cpp:20-21
  static const int NORMAL_PHASE = SMOOTHNESS | DISTANCE | FEASIBILITY | START | END;    // :22-23
  typedef std::unique_ptr<BsplineOptimizer> Ptr;

  void setEnvironment(const std::shared_ptr<EDTEnvironment>& env) { edt_environment_ = env; }
  void setParam(const FuelOptParams& p, const int max_iteration_num[4]) {  // :25-57
    params_ = p;
    for (int i = 0; i < 4; ++i) max_iteration_num_[i] = max_iteration_num[i];
    time_lb_ = -1;
  }
  void setBoundaryStates(const std::vector<Vector3d>& start, const std::vector<Vector3d>& end) {
    start_state_ = start;
    end_state_ = end;
  }
  void setTimeLowerBound(const double& lb) { time_lb_ = lb; }
  void setGuidePath(const std::vector<Vector3d>& guide_pt) { guide_pts_ = guide_pt; }
  void setWaypoints(const std::vector<Vector3d>& waypts, const std::vector<int>& waypt_idx) {
    waypoints_ = waypts;
    waypt_idx_ = waypt_idx;
  }
  // setViewConstraint (:91-93): the fields of ViewConstraint (traj_visibility.h:18-24) that calcViewCost reads
  void setViewConstraint(const Vector3d& pt, const Vector3d& dir, int idx) {
    view_pt_ = pt;
    view_dir_ = dir;
    view_idx_ = idx;
  }

  // optimize (:110-163): points = N x 3 control points, row-major; dt in/out.  The NLopt driver
  // loop (:165-253) runs on the device (fuelgpu_bspline_optimize_batch, B = 1).
  void optimize(std::vector<Vector3d>& points, double& dt, const int& cost_function, const int& max_num_id,
                const int& /*max_time_id*/) {
    if (start_state_.empty()) throw std::runtime_error("Initial state undefined!");  // :112-115
    const int n = (int)points.size();
    const bool optimize_time = cost_function & MINTIME;
    const int nvar = 3 * n + (optimize_time ? 1 : 0);
    std::vector<double> x(nvar);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) x[3 * i + k] = points[i](k);
    if (optimize_time) x[nvar - 1] = dt;
    fillTrajConst(points, dt);
    FuelSolveParams sp{ max_iteration_num_[max_num_id], 6, 1e-5 };
    double fbest = 0;
    int32_t neval = 0;
    FuelMap* h = edt_environment_->sdf_map_->gpu();
    fuelgpu_check(fuelgpu_bspline_optimize_batch(h, 1, n, cost_function, &params_, &tc_, &sp, x.data(), &fbest, &neval), h);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) points[i](k) = x[3 * i + k];
    if (optimize_time) dt = x[nvar - 1];
    iter_num_ = neval;
    min_cost_ = fbest;
    start_state_.clear();  // :161-162
    time_lb_ = -1;
  }
  // combineCost (:518-647) through the device, the costFunction trampoline (:693-706)
  double combineCost(const std::vector<double>& x, std::vector<double>& grad, int n_pts, int cost_function) {
    double f = 0;
    grad.resize(x.size());
    FuelMap* h = edt_environment_->sdf_map_->gpu();
    fuelgpu_check(fuelgpu_bspline_cost_batch(h, 1, n_pts, cost_function, &params_, &tc_, x.data(), &f, grad.data()), h);
    return f;
  }
  void fillTrajConst(const std::vector<Vector3d>& points, double dt) {
    std::memset(&tc_, 0, sizeof(tc_));
    double d = 0.0;  // pt_dist_ (:136-140)
    for (size_t i = 0; i + 1 < points.size(); ++i) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (points[i + 1](k) - points[i](k)) * (points[i + 1](k) - points[i](k));
      d += std::sqrt(s);
    }
    tc_.pt_dist = d / double(points.size());
    tc_.knot_span = dt;
    for (size_t i = 0; i < start_state_.size() && i < 3; ++i)
      for (int k = 0; k < 3; ++k) tc_.start[i][k] = start_state_[i](k);
    tc_.n_end = (int32_t)std::min<size_t>(end_state_.size(), 3);
    for (int i = 0; i < tc_.n_end; ++i)
      for (int k = 0; k < 3; ++k) tc_.end[i][k] = end_state_[i](k);
    tc_.time_lb = time_lb_;
    tc_.n_guide = (int32_t)std::min<size_t>(guide_pts_.size(), FUELGPU_MAX_PTS);
    for (int i = 0; i < tc_.n_guide; ++i)
      for (int k = 0; k < 3; ++k) tc_.guide[i][k] = guide_pts_[i](k);
    tc_.n_waypt = (int32_t)std::min<size_t>(waypoints_.size(), FUELGPU_MAX_PTS);
    for (int i = 0; i < tc_.n_waypt; ++i) {
      for (int k = 0; k < 3; ++k) tc_.waypt[i][k] = waypoints_[i](k);
      tc_.waypt_idx[i] = waypt_idx_[i];
    }
    tc_.view_idx = view_idx_;
    for (int k = 0; k < 3; ++k) {
      tc_.view_pt[k] = view_pt_(k);
      tc_.view_dir[k] = view_dir_(k);
    }
  }
  int iter_num_ = 0;
  double min_cost_ = 0;

private:
  std::shared_ptr<EDTEnvironment> edt_environment_;
  FuelOptParams params_;
  FuelTrajConst tc_;
  int max_iteration_num_[4] = { 2, 2000, 200, 200 };
  std::vector<Vector3d> start_state_, end_state_, guide_pts_, waypoints_;
  std::vector<int> waypt_idx_;
  Vector3d view_pt_, view_dir_;
  int view_idx_ = -1;
  double time_lb_ = -1;
};

}  // namespace fast_planner
