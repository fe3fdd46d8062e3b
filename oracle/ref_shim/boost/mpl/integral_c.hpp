#ifndef RKB_SHIM_BOOST_MPL_INTEGRAL_C_HPP
#define RKB_SHIM_BOOST_MPL_INTEGRAL_C_HPP
#include "bool.hpp"
#endif
