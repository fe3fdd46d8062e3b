#ifndef RKB_SHIM_BOOST_NONCOPYABLE_HPP
#define RKB_SHIM_BOOST_NONCOPYABLE_HPP
namespace boost {
class noncopyable {
protected:
  noncopyable() = default;
  ~noncopyable() = default;
  noncopyable(const noncopyable&) = delete;
  noncopyable& operator=(const noncopyable&) = delete;
};
}
#endif
