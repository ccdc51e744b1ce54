// TEST INFRASTRUCTURE ONLY: Boost is not installed; hash_combine only decides bucket order in the
// reference's unordered_map of voxels, never a numerical result.
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <typename T> inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
}
}  // namespace boost
