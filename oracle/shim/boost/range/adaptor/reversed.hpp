// Shim: boost::adaptors::reverse for containers with rbegin/rend. Test infrastructure only.
#ifndef LFX_SHIM_BOOST_REVERSED_HPP_
#define LFX_SHIM_BOOST_REVERSED_HPP_
namespace boost
{
namespace adaptors
{
template<typename C>
struct reversed_view
{
  const C & c;
  auto begin() const {return c.rbegin();}
  auto end() const {return c.rend();}
};
template<typename C>
reversed_view<C> reverse(const C & c) {return reversed_view<C>{c};}
}  // namespace adaptors
}  // namespace boost
#endif
