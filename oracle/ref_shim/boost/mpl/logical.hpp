#ifndef RKB_SHIM_BOOST_MPL_LOGICAL_HPP
#define RKB_SHIM_BOOST_MPL_LOGICAL_HPP
#include "bool.hpp"
namespace boost { namespace mpl {
template <typename... Ts> struct and_;
template <> struct and_<> : true_ {};
template <typename T, typename... Ts> struct and_<T, Ts...>
  : bool_<static_cast<bool>(T::value) && and_<Ts...>::value> {};
template <typename... Ts> struct or_;
template <> struct or_<> : false_ {};
template <typename T, typename... Ts> struct or_<T, Ts...>
  : bool_<static_cast<bool>(T::value) || or_<Ts...>::value> {};
template <typename T> struct not_ : bool_<!static_cast<bool>(T::value)> {};
}}
#endif
