// Boost shim (test infrastructure): property maps are only named, never used, by the code paths instantiated here.
#ifndef RKB_SHIM_BOOST_PROPERTY_MAP_HPP
#define RKB_SHIM_BOOST_PROPERTY_MAP_HPP
namespace boost {
template <class PMap>
struct property_traits {
  typedef typename PMap::value_type value_type;
  typedef typename PMap::key_type key_type;
  typedef typename PMap::reference reference;
};
}  // namespace boost
#endif
