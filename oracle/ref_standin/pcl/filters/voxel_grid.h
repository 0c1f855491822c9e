// Stand-in for pcl::VoxelGrid<pcl::PointXYZ> as FrontierFinder::downsample uses it (frontier_finder.cpp:757-774).
// PCL is a third-party dependency absent from the image: the filter is the oracle's RECONSTRUCTION of PCL >= 1.7
// VoxelGrid::applyFilter (orc_voxelgrid_f32 in oracle/fuel_oracle.c, "parity unpinned" for this piece) -- the very same
// function the oracle uses (see Eigen/Eigenvalues in this directory).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <stdint.h>

#include <vector>

extern "C" int32_t orc_voxelgrid_f32(const float* pts, int32_t n, float leaf, float* out);

namespace pcl {
template <typename T>
class VoxelGrid;
template <>
class VoxelGrid<PointXYZ> {
public:
  void setInputCloud(const PointCloud<PointXYZ>::Ptr& c) { in_ = c; }
  void setLeafSize(float lx, float /*ly*/, float /*lz*/) { leaf_ = lx; }  // the reference passes one size three times
  void filter(PointCloud<PointXYZ>& out) {
    const int n = (int)in_->points.size();
    std::vector<float> p(3 * (size_t)(n > 0 ? n : 1)), o(3 * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) p[3 * i] = in_->points[i].x, p[3 * i + 1] = in_->points[i].y, p[3 * i + 2] = in_->points[i].z;
    const int m = orc_voxelgrid_f32(p.data(), n, leaf_, o.data());
    out.points.clear();
    for (int i = 0; i < m; ++i) out.points.emplace_back(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
  }

private:
  PointCloud<PointXYZ>::Ptr in_;
  float leaf_ = 0.f;
};
}  // namespace pcl
