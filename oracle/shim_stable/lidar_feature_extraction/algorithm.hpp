// Shim (index tie-break variant): shadows extraction/include/lidar_feature_extraction/algorithm.hpp
// with an Argsort that is identical in interface but uses std::stable_sort, i.e. ascending
// (value, index) order. BASELINE.json's north_star: "the reference is run with the same
// index tie-break". Reference behaviour being restated: algorithm.hpp:44-71 (ByValue, Argsort).
// Test infrastructure only.
#ifndef LIDAR_FEATURE_EXTRACTION__ALGORITHM_HPP_
#define LIDAR_FEATURE_EXTRACTION__ALGORITHM_HPP_
#include <algorithm>
#include <iterator>
#include <numeric>
#include <vector>
#include "lidar_feature_extraction/iterator.hpp"

template<typename Iterator>
using ValueType = typename std::iterator_traits<Iterator>::value_type;

template<typename Iterator, typename T = ValueType<Iterator>>
std::vector<int> Argsort(const Iterator & values_begin, const Iterator & values_end)
{
  std::vector<int> indices(values_end - values_begin);
  std::iota(indices.begin(), indices.end(), 0);
  std::stable_sort(
    indices.begin(), indices.end(),
    [&](const int l, const int r) {return *(values_begin + l) < *(values_begin + r);});
  return indices;
}
#endif  // LIDAR_FEATURE_EXTRACTION__ALGORITHM_HPP_
