#ifndef RKB_SHIM_BOOST_CONCEPT_REQUIRES_HPP
#define RKB_SHIM_BOOST_CONCEPT_REQUIRES_HPP
namespace boost { namespace rkb_shim_detail {
template <typename F> struct unparen;
template <typename R> struct unparen<void(R)> { typedef R type; };
template <> struct unparen<void()> { typedef void type; };
}}
#define BOOST_CONCEPT_REQUIRES(models, result) \
  typename ::boost::rkb_shim_detail::unparen<void result>::type
#endif
