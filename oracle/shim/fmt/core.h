// Shim: fmt::format with positional-less "{}" substitution (all the reference uses).
// Test infrastructure only.
#ifndef LFX_SHIM_FMT_CORE_H_
#define LFX_SHIM_FMT_CORE_H_
#include <cstdint>
#include <sstream>
#include <string>
namespace fmt
{
namespace detail
{
inline void put(std::ostringstream & os, const std::uint8_t & v) {os << static_cast<unsigned>(v);}
template<typename T>
void put(std::ostringstream & os, const T & v) {os << v;}
inline void format_rest(std::ostringstream & os, const char * s) {os << s;}
template<typename T, typename ... Rest>
void format_rest(std::ostringstream & os, const char * s, const T & v, const Rest & ... rest)
{
  for (; *s; ++s) {
    if (s[0] == '{' && s[1] == '}') {
      put(os, v);
      format_rest(os, s + 2, rest ...);
      return;
    }
    os << *s;
  }
}
}  // namespace detail
template<typename ... Args>
std::string format(const std::string & f, const Args & ... args)
{
  std::ostringstream os;
  detail::format_rest(os, f.c_str(), args ...);
  return os.str();
}
}  // namespace fmt
#endif
