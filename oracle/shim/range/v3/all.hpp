// Shim: the three range-v3 facilities the reference hot path uses
// (views::ints, views::transform, to_vector). Test infrastructure only.
#ifndef LFX_SHIM_RANGE_V3_ALL_HPP_
#define LFX_SHIM_RANGE_V3_ALL_HPP_
#include <type_traits>
#include <utility>
#include <vector>
namespace ranges
{
struct to_vector_fn {};
inline constexpr to_vector_fn to_vector{};
namespace views
{
struct ints_view {int first; int last;};
inline ints_view ints(int first, int last) {return ints_view{first, last};}
template<typename F>
struct transform_fn {F f;};
template<typename F>
transform_fn<F> transform(F f) {return transform_fn<F>{std::move(f)};}
template<typename T, typename F>
struct transform_view {std::vector<T> src; F f;};
// lives in views so ADL finds it through transform_fn
template<typename T, typename F>
transform_view<T, F> operator|(const std::vector<T> & src, transform_fn<F> t)
{
  return transform_view<T, F>{src, std::move(t.f)};
}
}  // namespace views
inline std::vector<int> operator|(const views::ints_view & v, to_vector_fn)
{
  std::vector<int> out;
  for (int i = v.first; i < v.last; i++) {out.push_back(i);}
  return out;
}
template<typename T, typename F>
auto operator|(const views::transform_view<T, F> & v, to_vector_fn)
{
  using R = std::decay_t<decltype(v.f(std::declval<const T &>()))>;
  std::vector<R> out;
  out.reserve(v.src.size());
  for (const T & e : v.src) {out.push_back(v.f(e));}
  return out;
}
}  // namespace ranges
#endif
