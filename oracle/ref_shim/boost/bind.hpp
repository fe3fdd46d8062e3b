// Boost shim for the reference build: core/optimization/line_search.hpp names boost::bind, boost::cref
// and the _1 placeholder inside function templates the proximity path never instantiates.
#ifndef RKB_SHIM_BOOST_BIND_HPP
#define RKB_SHIM_BOOST_BIND_HPP
#include <functional>
namespace boost {
using std::bind;
using std::cref;
using std::ref;
}
using std::placeholders::_1;
using std::placeholders::_2;
#endif
