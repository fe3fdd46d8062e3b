// Boost shim (test infrastructure): just enough of boost::graph_traits for ReaK's node generators
// (ctrl/graph_alg/node_generators.hpp), which only name the vertex descriptor type and the null vertex.
#ifndef RKB_SHIM_BOOST_GRAPH_CONCEPTS_HPP
#define RKB_SHIM_BOOST_GRAPH_CONCEPTS_HPP
namespace boost {
template <class Graph>
struct graph_traits {
  typedef typename Graph::vertex_descriptor vertex_descriptor;
  typedef typename Graph::edge_descriptor edge_descriptor;
  static vertex_descriptor null_vertex() { return Graph::null_vertex(); }
};
// direction category: a graph type says `static const bool is_directed = ...;`
template <class Graph> struct is_directed_graph { static const bool value = Graph::is_directed; typedef is_directed_graph type; };
template <class Graph> struct is_undirected_graph { static const bool value = !Graph::is_directed; typedef is_undirected_graph type; };
}  // namespace boost
#endif
