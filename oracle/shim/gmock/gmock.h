// Shim: the sliver of gmock the reference's extraction tests use (EXPECT_THAT with
// testing::ElementsAre or a plain value), on top of the gtest-1.8.0 that the reference
// vendors inside its nanoflann submodule. Test infrastructure only.
#ifndef LFX_SHIM_GMOCK_H_
#define LFX_SHIM_GMOCK_H_
#include <gtest/gtest.h>
#include <cstddef>
#include <iterator>
#include <tuple>
#include <utility>
namespace testing
{
template<typename ... Ts>
struct ElementsAreMatcher
{
  std::tuple<Ts...> expected;

  template<typename C, std::size_t ... I>
  bool MatchImpl(const C & c, std::index_sequence<I...>) const
  {
    if (static_cast<std::size_t>(std::distance(c.begin(), c.end())) != sizeof...(Ts)) {
      return false;
    }
    auto it = c.begin();
    bool ok = true;
    // element-wise operator== in order
    (void)std::initializer_list<int>{(ok = ok && (*it == std::get<I>(expected)), ++it, 0)...};
    return ok;
  }
  template<typename C>
  bool Match(const C & c) const {return MatchImpl(c, std::index_sequence_for<Ts...>{});}
};
template<typename ... Ts>
ElementsAreMatcher<Ts...> ElementsAre(Ts ... v) {return ElementsAreMatcher<Ts...>{std::make_tuple(v ...)};}

namespace lfx_shim
{
template<typename V, typename ... Ts>
bool Match(const V & v, const ElementsAreMatcher<Ts...> & m) {return m.Match(v);}
template<typename V, typename M>
bool Match(const V & v, const M & m) {return v == static_cast<V>(m);}
}  // namespace lfx_shim
}  // namespace testing
#define EXPECT_THAT(value, matcher) EXPECT_TRUE(::testing::lfx_shim::Match((value), (matcher)))
#define ASSERT_THAT(value, matcher) ASSERT_TRUE(::testing::lfx_shim::Match((value), (matcher)))
#endif
