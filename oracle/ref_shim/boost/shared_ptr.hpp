#ifndef RKB_SHIM_BOOST_SHARED_PTR_HPP
#define RKB_SHIM_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
using std::shared_ptr; using std::weak_ptr; using std::enable_shared_from_this;
using std::static_pointer_cast; using std::dynamic_pointer_cast; using std::const_pointer_cast;
using std::make_shared;
}
#endif
