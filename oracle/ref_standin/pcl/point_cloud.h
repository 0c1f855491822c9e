// stand-in: pcl::PointCloud<T> as sdf_map.cpp reads it (`points`).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
namespace pcl {
template <typename T>
struct PointCloud {
  std::vector<T> points;
};
}  // namespace pcl
