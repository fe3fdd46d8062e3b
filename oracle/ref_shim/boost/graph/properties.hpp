// Boost shim (test infrastructure): named by ctrl/path_planning/topological_search.hpp, nothing of it is instantiated here.
#ifndef RKB_SHIM_BOOST_GRAPH_PROPERTIES_HPP
#define RKB_SHIM_BOOST_GRAPH_PROPERTIES_HPP
#endif
