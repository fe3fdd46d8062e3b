#ifndef RKB_SHIM_BOOST_MPL_SIZE_T_HPP
#define RKB_SHIM_BOOST_MPL_SIZE_T_HPP
#include <cstddef>
#include <boost/mpl/bool.hpp>
namespace boost { namespace mpl {
template <std::size_t N> struct size_t {
  static const std::size_t value = N;
  typedef integral_c_tag tag;
  typedef size_t type;
  typedef std::size_t value_type;
  typedef size_t<N + 1> next;
  typedef size_t<N - 1> prior;
  constexpr operator std::size_t() const { return N; }
};
template <std::size_t N> const std::size_t size_t<N>::value;
}}
#endif
