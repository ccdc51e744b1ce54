// TEST INFRASTRUCTURE ONLY: empty stand-in; ivox3d_node.hpp includes it but uses nothing from it.
#pragma once
#include <pcl/point_types.h>
