#ifndef RKB_SHIM_BOOST_CSTDINT_HPP
#define RKB_SHIM_BOOST_CSTDINT_HPP
#include <cstdint>
namespace boost {
using std::int8_t;  using std::int16_t;  using std::int32_t;  using std::int64_t;
using std::uint8_t; using std::uint16_t; using std::uint32_t; using std::uint64_t;
using std::intmax_t; using std::uintmax_t;
}
#endif
