#ifndef RKB_SHIM_BOOST_TYPE_TRAITS_HPP
#define RKB_SHIM_BOOST_TYPE_TRAITS_HPP
#include <type_traits>
#include "mpl/bool.hpp"
namespace boost {
#define RKB_SHIM_TRAIT1(name) \
  template <typename T> struct name : mpl::bool_<std::name<T>::value> {};
#define RKB_SHIM_TRAIT2(name) \
  template <typename A, typename B> struct name : mpl::bool_<std::name<A, B>::value> {};
RKB_SHIM_TRAIT2(is_same) RKB_SHIM_TRAIT2(is_convertible) RKB_SHIM_TRAIT2(is_base_of)
RKB_SHIM_TRAIT1(is_const) RKB_SHIM_TRAIT1(is_pointer) RKB_SHIM_TRAIT1(is_reference)
RKB_SHIM_TRAIT1(is_arithmetic) RKB_SHIM_TRAIT1(is_integral) RKB_SHIM_TRAIT1(is_floating_point)
RKB_SHIM_TRAIT1(is_fundamental) RKB_SHIM_TRAIT1(is_scalar) RKB_SHIM_TRAIT1(is_class)
RKB_SHIM_TRAIT1(is_polymorphic) RKB_SHIM_TRAIT1(is_abstract) RKB_SHIM_TRAIT1(is_void)
RKB_SHIM_TRAIT1(is_enum) RKB_SHIM_TRAIT1(is_signed) RKB_SHIM_TRAIT1(is_unsigned)
#undef RKB_SHIM_TRAIT1
#undef RKB_SHIM_TRAIT2
using std::remove_const; using std::remove_reference; using std::remove_cv; using std::remove_pointer;
using std::add_const; using std::add_pointer; using std::add_lvalue_reference;
template <typename T> struct add_reference { typedef typename std::add_lvalue_reference<T>::type type; };
using std::integral_constant; using std::true_type; using std::false_type;
}
#endif
