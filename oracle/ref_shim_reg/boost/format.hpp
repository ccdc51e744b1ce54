// TEST INFRASTRUCTURE ONLY: Boost is not installed; the reference uses boost::format only for debug prints.
#pragma once
#include <ostream>
#include <string>
namespace boost {
struct format {
  explicit format(const std::string&) {}
  template <typename T> format& operator%(const T&) { return *this; }
};
inline std::ostream& operator<<(std::ostream& o, const format&) { return o; }
}  // namespace boost
