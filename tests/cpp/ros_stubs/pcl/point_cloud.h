#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename T>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  std::vector<T> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
};
}  // namespace pcl
