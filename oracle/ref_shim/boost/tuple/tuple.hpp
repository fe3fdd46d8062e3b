// Boost shim (test infrastructure): boost::tuple / tie / get are the std ones.
#ifndef RKB_SHIM_BOOST_TUPLE_HPP
#define RKB_SHIM_BOOST_TUPLE_HPP
#include <tuple>
#include <utility>
namespace boost {
using std::tuple;
using std::tie;
using std::get;
using std::make_tuple;
namespace tuples {
using std::tuple;
using std::tie;
using std::get;
template <class T> struct length : std::tuple_size<T> {};
template <int N, class T> struct element : std::tuple_element<N, T> {};
}  // namespace tuples
}  // namespace boost
#endif
