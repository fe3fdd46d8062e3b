#ifndef RKB_SHIM_BOOST_WEAK_PTR_HPP
#define RKB_SHIM_BOOST_WEAK_PTR_HPP
#include "shared_ptr.hpp"
#endif
