#ifndef RKB_SHIM_BOOST_MPL_BOOL_HPP
#define RKB_SHIM_BOOST_MPL_BOOL_HPP
namespace boost { namespace mpl {
struct integral_c_tag { static const int value = 0; };
template <bool C> struct bool_ {
  static const bool value = C;
  typedef integral_c_tag tag;
  typedef bool_ type;
  typedef bool value_type;
  constexpr operator bool() const { return C; }
};
template <bool C> const bool bool_<C>::value;
typedef bool_<true> true_;
typedef bool_<false> false_;
template <int N> struct int_ {
  static const int value = N;
  typedef integral_c_tag tag;
  typedef int_ type;
  typedef int value_type;
  typedef int_<N + 1> next;
  typedef int_<N - 1> prior;
  constexpr operator int() const { return N; }
};
template <int N> const int int_<N>::value;
template <typename T, T N> struct integral_c {
  static const T value = N;
  typedef integral_c_tag tag;
  typedef integral_c type;
  typedef T value_type;
  constexpr operator T() const { return N; }
};
template <typename T, T N> const T integral_c<T, N>::value;
template <bool C, typename T1, typename T2> struct if_c { typedef T1 type; };
template <typename T1, typename T2> struct if_c<false, T1, T2> { typedef T2 type; };
template <typename C, typename T1, typename T2> struct if_ : if_c<static_cast<bool>(C::value), T1, T2> {};
}}
#endif
