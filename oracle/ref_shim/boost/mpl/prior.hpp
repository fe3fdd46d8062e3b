#ifndef RKB_SHIM_BOOST_MPL_PRIOR_HPP
#define RKB_SHIM_BOOST_MPL_PRIOR_HPP
namespace boost { namespace mpl {
template <typename T> struct prior { typedef typename T::prior type; };
template <typename T> struct next { typedef typename T::next type; };
}}
#endif
