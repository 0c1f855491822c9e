// Stand-in for active_perception/graph_node.h: frontier_finder.cpp calls ViewNode::computeCost / searchPath from the
// cost-matrix functions (updateFrontierCostMatrix, getFullCostMatrix, getPathForTour -- A* over the map, out of scope).
// They must link; the tests never run them.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <Eigen/Eigen>
#include <vector>

using Eigen::Vector3d;  // the real header exports these names
using Eigen::Vector3i;

namespace fast_planner {
class ViewNode {
public:
  static double computeCost(const Eigen::Vector3d&, const Eigen::Vector3d&, const double&, const double&,
                            const Eigen::Vector3d&, const double&, std::vector<Eigen::Vector3d>&) {
    return 0.0;
  }
  static double searchPath(const Eigen::Vector3d&, const Eigen::Vector3d&, std::vector<Eigen::Vector3d>&) { return 0.0; }
};
}  // namespace fast_planner
