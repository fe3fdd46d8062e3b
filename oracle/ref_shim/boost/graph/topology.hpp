// Boost shim (test infrastructure): ctrl/path_planning/topological_search.hpp includes <boost/graph/topology.hpp> but the
// code instantiated here (min_dist_linear_search) only needs std::numeric_limits and boost::tie from what it drags in.
#ifndef RKB_SHIM_BOOST_GRAPH_TOPOLOGY_HPP
#define RKB_SHIM_BOOST_GRAPH_TOPOLOGY_HPP
#include <limits>
#include <tuple>
namespace boost { using std::tie; }
#endif
