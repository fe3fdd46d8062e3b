#ifndef RKB_SHIM_BOOST_STATIC_ASSERT_HPP
#define RKB_SHIM_BOOST_STATIC_ASSERT_HPP
#define BOOST_STATIC_ASSERT(...) static_assert(__VA_ARGS__, #__VA_ARGS__)
#define BOOST_STATIC_ASSERT_MSG(cond, msg) static_assert(cond, msg)
#endif
