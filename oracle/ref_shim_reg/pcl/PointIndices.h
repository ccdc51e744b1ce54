// TEST INFRASTRUCTURE ONLY: stand-in for <pcl/PointIndices.h> (included, not used, by information_matrix_calculator.hpp).
#pragma once
#include <vector>
namespace pcl { struct PointIndices { std::vector<int> indices; }; }
