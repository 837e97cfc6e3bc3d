// Shim: minimal stand-in for <pcl/point_cloud.h> so the reference's extraction
// sources compile in place without PCL (test infrastructure; not shipped).
// Only the container surface the reference hot path touches is provided.
#ifndef LFX_SHIM_PCL_POINT_CLOUD_H_
#define LFX_SHIM_PCL_POINT_CLOUD_H_
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl
{
template<typename PointT>
class PointCloud
{
public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  using value_type = PointT;
  using iterator = typename std::vector<PointT>::iterator;
  using const_iterator = typename std::vector<PointT>::const_iterator;

  std::vector<PointT> points;
  bool is_dense = true;

  std::size_t size() const {return points.size();}
  bool empty() const {return points.empty();}
  const PointT & at(std::size_t i) const {return points.at(i);}
  PointT & at(std::size_t i) {return points.at(i);}
  const PointT & operator[](std::size_t i) const {return points[i];}
  PointT & operator[](std::size_t i) {return points[i];}
  void push_back(const PointT & p) {points.push_back(p);}
  void reserve(std::size_t n) {points.reserve(n);}
  iterator begin() {return points.begin();}
  iterator end() {return points.end();}
  const_iterator begin() const {return points.begin();}
  const_iterator end() const {return points.end();}
  PointCloud & operator+=(const PointCloud & rhs)
  {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
};
}  // namespace pcl
#endif
