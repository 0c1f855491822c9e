// stand-in: pcl::PointXYZ as the reference reads and builds it (float x, y, z; 16-byte point like PCL's).
// TEST INFRASTRUCTURE ONLY.
#pragma once
namespace pcl {
struct PointXYZ {
  float x, y, z, pad_;
  PointXYZ() : x(0.f), y(0.f), z(0.f), pad_(1.f) {}
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad_(1.f) {}
};
}  // namespace pcl
