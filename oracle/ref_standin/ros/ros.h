// Stand-in for the one ROS facility plan_env/src/sdf_map.cpp uses: NodeHandle::param(key, out, default), fed from a
// table the test wrapper fills.  TEST INFRASTRUCTURE ONLY (see Eigen/Eigen in this directory).
#pragma once
// the real header drags in most of the standard library; the reference relies on that
#include <algorithm>
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define ROS_ERROR(...) do {} while (0)
#define ROS_WARN(...) do {} while (0)
#define ROS_INFO(...) do {} while (0)
#define ROS_INFO_STREAM(x) do {} while (0)
#define ROS_WARN_THROTTLE(...) do {} while (0)

namespace ros {
struct Duration {
  double toSec() const { return 0.0; }
};
struct Time {  // the reference only measures elapsed time for its log lines
  static Time now() { return Time(); }
  Duration operator-(const Time&) const { return Duration(); }
};
class NodeHandle {
public:
  std::map<std::string, double> values;
  template <typename T>
  bool param(const std::string& key, T& out, const T& def) const {
    auto it = values.find(key);
    if (it == values.end()) {
      out = def;
      return false;
    }
    out = static_cast<T>(it->second);
    return true;
  }
};
}  // namespace ros
