// TEST INFRASTRUCTURE ONLY: the reference includes this PCL header but the code compiled here uses nothing from it.
#pragma once
#include <pcl/point_cloud.h>
