#ifndef RKB_SHIM_BOOST_RANDOM_HPP
#define RKB_SHIM_BOOST_RANDOM_HPP
// Boost.Random names used by ReaK's samplers, over <random> (test infrastructure: no sampler is run by the checks)
#include <random>
namespace boost {
namespace random {
using std::mt19937;
using std::normal_distribution;
using std::lognormal_distribution;
using std::random_device;
template <class Engine, class Real = double>
struct uniform_01 {
  Engine eng;
  explicit uniform_01(Engine e) : eng(e) {}
  Real operator()() { return std::generate_canonical<Real, 53>(eng); }
};
template <class Engine, class Dist>
struct variate_generator {
  Engine eng;
  Dist dist;
  variate_generator(Engine e, Dist d) : eng(e), dist(d) {}
  typename Dist::result_type operator()() { return dist(eng); }
};
}
using random::mt19937;
using random::normal_distribution;
using random::lognormal_distribution;
using random::random_device;
using random::uniform_01;
using random::variate_generator;
}
#endif
