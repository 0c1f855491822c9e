// Stand-in for active_perception/traj_visibility.h: bspline_optimizer.h needs the ViewConstraint record only
// (traj_visibility.h:17-23); the visibility utility itself is out of scope.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <Eigen/Eigen>
#include <plan_env/edt_environment.h>

namespace fast_planner {
struct ViewConstraint {
  Eigen::Vector3d pt_;     // unknown point along the traj
  Eigen::Vector3d pc_;     // critical view point
  Eigen::Vector3d dir_;    // critical view direction with safe length
  Eigen::Vector3d pcons_;  // pt to add view constraint
  int idx_;                // idx to add view constraint
};
}  // namespace fast_planner
