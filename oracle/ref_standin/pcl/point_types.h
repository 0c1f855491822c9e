// stand-in: pcl::PointXYZ as sdf_map.cpp reads it (float x, y, z; 16-byte point like PCL's).  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace pcl {
struct PointXYZ {
  float x, y, z, pad_;
};
}  // namespace pcl
