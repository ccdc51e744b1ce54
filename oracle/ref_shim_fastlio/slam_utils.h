// TEST INFRASTRUCTURE ONLY: stand-in for slam/common/slam_utils.h; the LIO front-end uses one function of it, and only
// on the INS path (laserMapping.cpp:429), which the oracle does not feed.
#pragma once
#include <Eigen/Core>
Eigen::Matrix4d getTransformFromRPYT(double x, double y, double z, double yaw, double pitch, double roll);
