// Minimal header-only stand-in for <boost/config.hpp> (test infrastructure: lets the
// unmodified ReaK sources under /root/reference compile without a Boost install).
#ifndef RKB_SHIM_BOOST_CONFIG_HPP
#define RKB_SHIM_BOOST_CONFIG_HPP
#include <algorithm>
#include <cstddef>
#include <utility>
#define BOOST_STATIC_CONSTANT(type, assignment) static const type assignment
#define BOOST_NOEXCEPT_OR_NOTHROW noexcept
#define BOOST_NOEXCEPT noexcept
#define BOOST_CONSTEXPR constexpr
#define BOOST_USING_STD_MIN() using std::min
#define BOOST_USING_STD_MAX() using std::max
#define BOOST_PREVENT_MACRO_SUBSTITUTION
#endif
