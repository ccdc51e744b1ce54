// TEST INFRASTRUCTURE ONLY: IMU_Processing.hpp includes this PCL header but uses nothing from it.
#pragma once
#include <pcl/point_cloud.h>
