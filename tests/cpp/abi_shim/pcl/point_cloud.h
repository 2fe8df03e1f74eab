// TEST INFRASTRUCTURE ONLY -- the slice of pcl::PointCloud the drop-in touches (PCL 1.10: Ptr is a shared_ptr; `points`).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#define PCL_VERSION_CALC(a, b, c) ((a)*100000 + (b)*100 + (c))
#define PCL_VERSION PCL_VERSION_CALC(1, 10, 0)
namespace pcl {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
template <typename PointT>
class PointCloud {
public:
  using Ptr = shared_ptr<PointCloud<PointT>>;
  using ConstPtr = shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = static_cast<std::uint32_t>(points.size());
  }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
