// stand-in: pcl::PointCloud<T> as the reference uses it (`points`, ::Ptr).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename T>
struct PointCloud {
  std::vector<T> points;
  typedef std::shared_ptr<PointCloud<T>> Ptr;
};
}  // namespace pcl
