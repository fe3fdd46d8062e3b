#ifndef RKB_SHIM_BOOST_MPL_COMPARISON_HPP
#define RKB_SHIM_BOOST_MPL_COMPARISON_HPP
#include "bool.hpp"
namespace boost { namespace mpl {
template <typename A, typename B> struct equal_to : bool_<(A::value == B::value)> {};
template <typename A, typename B> struct not_equal_to : bool_<(A::value != B::value)> {};
template <typename A, typename B> struct less : bool_<(A::value < B::value)> {};
template <typename A, typename B> struct greater : bool_<(A::value > B::value)> {};
template <typename A, typename B> struct less_equal : bool_<(A::value <= B::value)> {};
template <typename A, typename B> struct greater_equal : bool_<(A::value >= B::value)> {};
}}
#endif
