#ifndef RKB_SHIM_BOOST_ENABLE_IF_HPP
#define RKB_SHIM_BOOST_ENABLE_IF_HPP
namespace boost {
template <bool B, typename T = void> struct enable_if_c { typedef T type; };
template <typename T> struct enable_if_c<false, T> {};
template <typename C, typename T = void> struct enable_if : enable_if_c<static_cast<bool>(C::value), T> {};
template <bool B, typename T = void> struct disable_if_c { typedef T type; };
template <typename T> struct disable_if_c<true, T> {};
template <typename C, typename T = void> struct disable_if : disable_if_c<static_cast<bool>(C::value), T> {};
template <bool B, typename T> struct lazy_enable_if_c { typedef typename T::type type; };
template <typename T> struct lazy_enable_if_c<false, T> {};
template <typename C, typename T> struct lazy_enable_if : lazy_enable_if_c<static_cast<bool>(C::value), T> {};
}
#endif
