#pragma once
#include <cstdio>
#include <map>
#include <string>
namespace ros {
struct Time {
  double t = 0;
  Time() {}
  explicit Time(double s) : t(s) {}
  double toSec() const { return t; }
  Time& fromSec(double s) { t = s; return *this; }
};
struct Duration {
  double d;
  explicit Duration(double s) : d(s) {}
};
class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <typename T>
  bool param(const std::string&, T& value, const T& fallback) const { value = fallback; return false; }
};
}  // namespace ros
#define ROS_WARN(...) std::fprintf(stderr, __VA_ARGS__)
