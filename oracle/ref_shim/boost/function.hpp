#ifndef RKB_SHIM_BOOST_FUNCTION_HPP
#define RKB_SHIM_BOOST_FUNCTION_HPP
#include <functional>
namespace boost { using std::function; }
#endif
