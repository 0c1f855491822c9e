// Stand-in for plan_env/edt_environment.h.  The real class also carries the moving-obstacle predictor (ROS, out of
// scope); the optimiser uses sdf_map_ and evaluateEDTWithGrad, which in the reference is a pass-through to
// SDFMap::getDistWithGrad (edt_environment.cpp:78-87) -- the real, compiled sdf_map.cpp here.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>

#include <plan_env/sdf_map.h>

namespace fast_planner {
class EDTEnvironment {
public:
  std::shared_ptr<SDFMap> sdf_map_;
  void setMap(std::shared_ptr<SDFMap>& map) { sdf_map_ = map; }
  void evaluateEDTWithGrad(const Eigen::Vector3d& pos, double /*time*/, double& dist, Eigen::Vector3d& grad) {
    dist = sdf_map_->getDistWithGrad(pos, grad);
  }
  typedef std::shared_ptr<EDTEnvironment> Ptr;
};
}  // namespace fast_planner
