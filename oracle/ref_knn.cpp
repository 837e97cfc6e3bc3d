// Reference-backed oracle for the localization package's neighbour search: the vendored nanoflann
// (/root/reference/localization/thirdparty/nanoflann, v1.4.2) compiled in place, driven exactly like
// KDTreeEigen (localization/src/kdtree.cpp:34-55, localization/include/lidar_feature_localization/kdtree.hpp:43-77).
//
// TEST INFRASTRUCTURE. Only tests/ may load this library; the product path never does.
//
// Eigen is absent from this image, so KDTreeEigenMatrixAdaptor itself cannot be instantiated; it is a thin wrapper
// that hands a row-major double matrix to KDTreeSingleIndexAdaptor<metric_L2, self, DIM = -1, Eigen::Index>
// (nanoflann.hpp: struct KDTreeEigenMatrixAdaptor), which is what is built here over a plain array: same metric
// functor (L2_Adaptor over doubles), same index type (std::ptrdiff_t == Eigen::Index), same leaf size, same
// findNeighbors call with KNNResultSet<double> and SearchParams(10).
#include <cstddef>
#include <cstdint>
#include <vector>

#include <nanoflann.hpp>

namespace
{
struct RowMajor
{
  const double * data;
  std::ptrdiff_t rows, cols;
  using self_t = RowMajor;
  inline std::ptrdiff_t kdtree_get_point_count() const { return rows; }
  inline double kdtree_get_pt(const std::ptrdiff_t idx, size_t dim) const { return data[idx * cols + static_cast<std::ptrdiff_t>(dim)]; }
  template<class BBOX> bool kdtree_get_bbox(BBOX &) const { return false; }
};
using Metric = nanoflann::metric_L2::traits<double, RowMajor, std::ptrdiff_t>::distance_t;
using Index = nanoflann::KDTreeSingleIndexAdaptor<Metric, RowMajor, -1, std::ptrdiff_t>;
}  // namespace

extern "C" {

// map: [n_map][dim] doubles; queries: [n_q][dim]; out_idx [n_q][k] (uint64), out_d2 [n_q][k] squared distances,
// in the order KNNResultSet leaves them (ascending distance). max_leaf_size: 10 in MakeKDTree (kdtree.hpp:75).
int ref_knn(const double * map, std::int64_t n_map, int dim, int max_leaf_size, const double * queries, std::int64_t n_q, int k,
            std::uint64_t * out_idx, double * out_d2)
{
  if (n_map < k || k <= 0) { return 1; }
  RowMajor m{map, static_cast<std::ptrdiff_t>(n_map), dim};
  Index index(dim, m, nanoflann::KDTreeSingleIndexAdaptorParams(max_leaf_size));
  index.buildIndex();
  for (std::int64_t i = 0; i < n_q; i++) {
    std::vector<std::uint64_t> indices(k);
    std::vector<double> distances(k);
    nanoflann::KNNResultSet<double> result(k);
    result.init(reinterpret_cast<size_t *>(&indices[0]), &distances[0]);
    index.findNeighbors(result, queries + i * dim, nanoflann::SearchParams(10));
    for (int j = 0; j < k; j++) { out_idx[i * k + j] = indices[j]; out_d2[i * k + j] = distances[j]; }
  }
  return 0;
}

}  // extern "C"
