// Stand-in for plan_env/map_ros.h (the ROS I/O wrapper, out of scope): SDFMap::initMap creates a MapROS, hands it the
// map and sets local_updated_.  MapROS is a friend of SDFMap (sdf_map.h:79), so this stand-in is also the test
// wrapper's door to the private buffers and to clearAndInflateLocalMap().  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <ros/ros.h>

#include "plan_env/sdf_map.h"

namespace fast_planner {
class MapROS {
public:
  void setMap(SDFMap* map) { map_ = map; }
  void init() {}
  ros::NodeHandle node_;
  bool local_updated_ = false;
  SDFMap* map_ = nullptr;
  // friend access for oracle/ref_sdfmap_wrap.cpp
  static MapParam& mp(SDFMap& m) { return *m.mp_; }
  static MapData& md(SDFMap& m) { return *m.md_; }
  static void clearAndInflate(SDFMap& m) { m.clearAndInflateLocalMap(); }
};
}  // namespace fast_planner
