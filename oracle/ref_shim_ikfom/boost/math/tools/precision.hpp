// TEST INFRASTRUCTURE ONLY: mtkmath.hpp:147 needs boost::math::tools::epsilon<T>() == std::numeric_limits<T>::epsilon().
#pragma once
#include <limits>
namespace boost { namespace math { namespace tools {
template <typename T> inline T epsilon() { return std::numeric_limits<T>::epsilon(); }
}}}  // namespace boost::math::tools
