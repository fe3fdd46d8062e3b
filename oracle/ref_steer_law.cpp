// TEST INFRASTRUCTURE — two pieces of reference code that sit behind headers g++ 13 rejects, compiled all the same:
//   (1) IHAQR_topology::get_bounded_input (examples/misc/IHAQR_topology.hpp:304-327), what oracle/steer_law.h restates;
//   (2) ctrl::detail::runge_kutta4_integrate_impl (ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp:50-97), RK4 with an
//       input trajectory, what kto_rk4_inputs / rkb_rollout_rk4_inputs restate — run over the live kte_nl_system.
// For (1) the UNMODIFIED examples/misc/IHAQR_topology.hpp is instantiated on ReaK's own hyperbox_topology and a
// stand-in system type (only its typedefs are read: get_bounded_input touches m_input_space, m_input_rate_space and
// m_time_step), and the protected member is reached through a derived class.
//
// everything the two integrator headers include, first and unharmed
#include <ReaK/core/base/named_object.hpp>
#include <ReaK/ctrl/ctrl_sys/state_space_sys_concept.hpp>
#include <ReaK/ctrl/topologies/metric_space_concept.hpp>
#include <ReaK/ctrl/topologies/temporal_space_concept.hpp>
#include <ReaK/ctrl/interpolation/spatial_trajectory_concept.hpp>
#include <ReaK/core/integrators/integration_exceptions.hpp>
#include <ReaK/core/lin_alg/vect_alg.hpp>
#include <ReaK/core/lin_alg/arithmetic_tuple.hpp>
// The two headers IHAQR_topology.hpp includes next have a *_factory::load that hands its iarchive to named_object::save
// (runge_kutta4_integrator_sys.hpp:267, dormand_prince45_integrator_sys.hpp:385); g++ rejects that statement even though
// the factories are never instantiated.  They are read with the parameter type spelled as the output archive, which makes
// that one statement well-formed; the integrate_impl function templates above the factories are not touched by the spelling.
#define iarchive oarchive
#include <ReaK/ctrl/sys_integrators/dormand_prince45_integrator_sys.hpp>
#include <ReaK/ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp>
#undef iarchive
#include <ReaK/examples/misc/IHAQR_topology.hpp>
#include <ReaK/examples/misc/MEAQR_topology.hpp>
#include <ReaK/ctrl/topologies/hyperball_topology.hpp>
#include <ReaK/ctrl/ctrl_sys/kte_nl_system.hpp>

#include <cmath>
#include <cstddef>
#include <stdint.h>

// ---- an input trajectory given by its values at every half step (what rkb_rollout_rk4_inputs takes) --------------------
namespace {
struct node_traj {
  struct point_type { ReaK::vect_n<double> pt; double time; };   // (the integrator reads .pt, like a temporal_point)
  typedef long long const_waypoint_descriptor;
  const double* nodes;
  int nu;
  long long n_nodes;
  double half_dt;
  std::pair<long long, point_type> at(long long j) const {
    if (j < 0) j = 0;
    if (j >= n_nodes) j = n_nodes - 1;
    point_type p;
    p.pt = ReaK::vect_n<double>(nu);
    for (int k = 0; k < nu; ++k) p.pt[k] = nodes[j * nu + k];
    p.time = double(j) * half_dt;
    return std::make_pair(j, p);
  }
  std::pair<long long, point_type> get_waypoint_at_time(double t) const { return at(llround(t / half_dt)); }
  std::pair<long long, point_type> move_time_diff_from(const std::pair<long long, point_type>& wp, double dt) const {
    return at(wp.first + llround(dt / half_dt));
  }
};
}  // namespace
namespace ReaK { namespace pp {
template <> struct spatial_trajectory_traits<node_traj> {   // the two typedefs runge_kutta4_integrate_impl asks for
  typedef node_traj::point_type point_type;
  typedef node_traj::const_waypoint_descriptor const_waypoint_descriptor;
};
}}

namespace {
using namespace ReaK;

struct law_system : public named_object {   // the typedefs ss_system_traits / linear_ss_system_traits read
  typedef vect_n<double> point_type;
  typedef vect_n<double> point_difference_type;
  typedef vect_n<double> point_derivative_type;
  typedef double time_type;
  typedef double time_difference_type;
  typedef vect_n<double> input_type;
  typedef vect_n<double> output_type;
  typedef mat<double, mat_structure::rectangular> matrixA_type;
  typedef mat<double, mat_structure::rectangular> matrixB_type;
  typedef mat<double, mat_structure::rectangular> matrixC_type;
  typedef mat<double, mat_structure::rectangular> matrixD_type;
  BOOST_STATIC_CONSTANT(std::size_t, dimensions = 0);
  BOOST_STATIC_CONSTANT(std::size_t, input_dimensions = 0);
  BOOST_STATIC_CONSTANT(std::size_t, output_dimensions = 0);
  // the dynamics are the live kte_nl_system's (set for the steering loop below; get_bounded_input never asks)
  const ctrl::kte_nl_system* live;
  law_system() : live(NULL) {}
  template <typename Space>
  vect_n<double> get_state_derivative(const Space&, const vect_n<double>& x, const vect_n<double>& u, double t) const {
    return live->get_state_derivative(*live, x, u, t);
  }
  // compute_linearization_data / compute_IHAQR_data are compiled with move_position_toward_impl but never run here: the
  // points arrive with their linearisation and gains filled in
  template <typename Space>
  void get_linear_blocks(matrixA_type&, matrixB_type&, matrixC_type&, matrixD_type&, const Space&, double, const vect_n<double>&,
                         const vect_n<double>&) const {
    throw std::logic_error("law_system: not linearisable");
  }
  virtual void RK_CALL save(serialization::oarchive& A, unsigned int) const { named_object::save(A, named_object::getStaticObjectType()->TypeVersion()); }
  virtual void RK_CALL load(serialization::iarchive& A, unsigned int) { named_object::load(A, named_object::getStaticObjectType()->TypeVersion()); }
  RK_RTTI_MAKE_CONCRETE_1BASE(law_system, 0xC23FFF01, 1, "rkb_law_system", named_object)
};
struct law_sampler : public named_object {
  virtual void RK_CALL save(serialization::oarchive& A, unsigned int) const { named_object::save(A, named_object::getStaticObjectType()->TypeVersion()); }
  virtual void RK_CALL load(serialization::iarchive& A, unsigned int) { named_object::load(A, named_object::getStaticObjectType()->TypeVersion()); }
  RK_RTTI_MAKE_CONCRETE_1BASE(law_sampler, 0xC23FFF02, 1, "rkb_law_sampler", named_object)
};
typedef pp::hyperball_topology<vect_n<double> > law_space;   // vector space with the Euclidean distance() member the loop asks for
typedef pp::IHAQR_topology<law_space, law_system, law_sampler> law_topology;

struct law_access : public law_topology {
  law_access(const vect_n<double>& lo, const vect_n<double>& hi, const vect_n<double>& bw, double T,
             const shared_ptr<law_system>& sys = shared_ptr<law_system>(), double horizon = 10.0, double threshold = 1.0)
      : law_topology("law", sys, law_space(), lo, hi, bw, mat<double, mat_structure::diagonal>(),
                     mat<double, mat_structure::diagonal>(), T, horizon, threshold) {}
  law_topology::point_type move(const law_topology::point_type& a, double fraction, const law_topology::point_type& b) const {
    return this->move_position_toward_impl(a, fraction, b, false);
  }
  vect_n<double> bounded(const vect_n<double>& u_prev, const vect_n<double>& u_bias, const vect_n<double>& u_corr) const {
    return this->get_bounded_input(u_prev, u_bias, u_corr);
  }
};
}  // namespace

// u_out = IHAQR_topology::get_bounded_input(u_prev, u_bias, u_correction) (IHAQR_topology.hpp:304-327) for `count` triples;
// input box [lo, hi], rate box [-bandwidth, bandwidth] (the constructor's aInputBandwidth, :440), time step T.
extern "C" int rkref_ihaqr_bounded_input(int nu, const double* lo, const double* hi, const double* bandwidth, double T, int count,
                                         const double* u_prev, const double* u_bias, const double* u_corr, double* u_out) {
  try {
    vect_n<double> vlo(nu), vhi(nu), vbw(nu), a(nu), b(nu), c(nu);
    for (int k = 0; k < nu; ++k) { vlo[k] = lo[k]; vhi[k] = hi[k]; vbw[k] = bandwidth[k]; }
    law_access topo(vlo, vhi, vbw, T);
    for (int i = 0; i < count; ++i) {
      for (int k = 0; k < nu; ++k) { a[k] = u_prev[i * nu + k]; b[k] = u_bias[i * nu + k]; c[k] = u_corr[i * nu + k]; }
      const vect_n<double> r = topo.bounded(a, b, c);
      for (int k = 0; k < nu; ++k) u_out[i * nu + k] = r[k];
    }
    return 0;
  } catch (...) { return -1; }
}

// ctrl::detail::runge_kutta4_integrate_impl ITSELF (runge_kutta4_integrator_sys.hpp:50-97) over the live kte_nl_system
// `sys` (rkref_kte_nl_system of a handle of ref_lib.cpp) with the input read from the node trajectory: n_steps steps of dt
// from x0[i]; u_nodes [N][2 n_steps + 1][nu].  status: bit 0 singular mass matrix, bit 1 non-finite (as the batch API).
extern "C" int rkref_rk4_inputs_concept(const void* sys, int nx, int nu, std::size_t N, const double* x0, const double* u_nodes, double dt,
                                        int n_steps, double* xout, int32_t* status) {
  const ctrl::kte_nl_system& S = *static_cast<const ctrl::kte_nl_system*>(sys);
  const law_space space;
  const long long J = 2LL * n_steps + 1;
  for (std::size_t i = 0; i < N; ++i) {
    vect_n<double> a(nx), b(nx);
    for (int k = 0; k < nx; ++k) a[k] = b[k] = x0[i * nx + k];
    int32_t st = 0;
    if (n_steps > 0) {
      node_traj traj;
      traj.nodes = u_nodes + i * J * nu; traj.nu = nu; traj.n_nodes = J; traj.half_dt = 0.5 * dt;
      try {
        // the loop is time-driven (:72-73): an end time half a step short of n_steps dt makes it run exactly n_steps steps
        ctrl::detail::runge_kutta4_integrate_impl(space, S, a, b, traj, 0.0, (double(n_steps) - 0.5) * dt, dt);
      } catch (singularity_error&) { st |= 1; }
    }
    for (int k = 0; k < nx; ++k) { xout[i * nx + k] = b[k]; if (!std::isfinite(b[k])) st |= 2; }
    if (status) status[i] = st;
  }
  return 0;
}

// IHAQR_topology::move_position_toward_impl ITSELF (IHAQR_topology.hpp:337-381, with_collision_check = false) over the live
// kte_nl_system `sys`: per tuple the loop `while (t < horizon && |x - goal| > threshold) { u = get_bounded_input(u_prev,
// u_bias, -K (x - goal)); runge_kutta4_integrate_impl over [t, t + T] with step T * 1e-2; accept }`.  The two points arrive
// with their payloads filled in (a: u_prev0 as lin_data->u with a zero bias; b: u_bias as lin_data->u with a zero bias, the
// gain K [nu][nx]) so that neither the linearisation nor the ARE solver runs.  fraction = 1: the goal is
// m_space.move_position_toward(a.x, 1.0, b.x).  x_out: the state the loop ended on.
extern "C" int rkref_ihaqr_move_toward(const void* sys, int nx, int nu, const double* lo, const double* hi, const double* bandwidth, double T,
                                       double horizon, double threshold, std::size_t N, const double* x0, const double* goal,
                                       const double* u_prev0, const double* u_bias, const double* gain, double* x_out) {
  try {
    vect_n<double> vlo(nu), vhi(nu), vbw(nu);
    for (int k = 0; k < nu; ++k) { vlo[k] = lo[k]; vhi[k] = hi[k]; vbw[k] = bandwidth[k]; }
    shared_ptr<law_system> S(new law_system());
    S->live = static_cast<const ctrl::kte_nl_system*>(sys);
    law_access topo(vlo, vhi, vbw, T, S, horizon, threshold);
    typedef law_topology::point_type point;
    for (std::size_t i = 0; i < N; ++i) {
      vect_n<double> xa(nx), xb(nx), ua(nu), ub(nu), zero(nu);
      for (int k = 0; k < nx; ++k) { xa[k] = x0[i * nx + k]; xb[k] = goal[i * nx + k]; }
      for (int k = 0; k < nu; ++k) { ua[k] = u_prev0[i * nu + k]; ub[k] = u_bias[i * nu + k]; zero[k] = 0.0; }
      point a(xa), b(xb);
      a.lin_data = shared_ptr<point::linearization_payload>(new point::linearization_payload());
      b.lin_data = shared_ptr<point::linearization_payload>(new point::linearization_payload());
      a.IHAQR_data = shared_ptr<point::IHAQR_payload>(new point::IHAQR_payload());
      b.IHAQR_data = shared_ptr<point::IHAQR_payload>(new point::IHAQR_payload());
      a.lin_data->u = ua; a.IHAQR_data->u_bias = zero;
      b.lin_data->u = ub; b.IHAQR_data->u_bias = zero;
      b.IHAQR_data->K = mat<double, mat_structure::rectangular>(nu, nx);
      for (int r = 0; r < nu; ++r)
        for (int c = 0; c < nx; ++c) b.IHAQR_data->K(r, c) = gain[(i * nu + r) * nx + c];
      const point res = topo.move(a, 1.0, b);
      for (int k = 0; k < nx; ++k) x_out[i * nx + k] = res.x[k];
    }
    return 0;
  } catch (std::exception& e) { return -1; }
}

// MEAQR_topology::steer_with_constant_control ITSELF (examples/misc/MEAQR_topology.hpp:503-561, no steer record; with
// `is_free` given, with_collision_check = true and the virtual is_free_impl answers through it) over the live kte_nl_system: `while (t < time_limit && |x - goal| > threshold) { u = t < T ? u0 - K eta -
// K H^-1 (x - goal) : get_bounded_input(u_prev, u0 - K eta, -K H^-1 (x - goal)); runge_kutta4_integrate_impl over [t, t + T]
// with step T * 1e-1; accept }`.  Called with H = I (its Cholesky factor is I: the back-substitution returns x - goal) and
// eta = 0, so that u0 is the batch call's u_bias and K its gain.  Out: the end state, the last input, the time reached.
namespace {
typedef pp::MEAQR_topology<law_space, law_system, law_sampler> meaqr_topology;
typedef int (*is_free_fn)(const double* x, int nx, void* ctx);
struct meaqr_access : public meaqr_topology {
  is_free_fn free_fn;   // the collision environment: MEAQR_topology::is_free_impl (:416) is virtual, "always free" in the base class
  void* free_ctx;
  explicit meaqr_access(const shared_ptr<law_topology>& s) : meaqr_topology("meaqr", s), free_fn(NULL), free_ctx(NULL) {}
  virtual bool is_free_impl(const vect_n<double>& a) const {
    if (!free_fn) return true;
    std::vector<double> x(a.size());
    for (std::size_t k = 0; k < a.size(); ++k) x[k] = a[k];
    return free_fn(&x[0], (int)x.size(), free_ctx) != 0;
  }
  bool steer(const mat<double, mat_structure::square>& H, const mat<double, mat_structure::rectangular>& K, const vect_n<double>& eta,
             const vect_n<double>& u0, vect_n<double>& u_prev, vect_n<double>& x, const vect_n<double>& goal, double& t, double limit) const {
    return this->steer_with_constant_control(H, K, eta, u0, u_prev, x, goal, t, limit, free_fn != NULL, NULL);
  }
};
}  // namespace
extern "C" int rkref_meaqr_steer(const void* sys, int nx, int nu, const double* lo, const double* hi, const double* bandwidth, double T,
                                 double time_limit, double threshold, std::size_t N, const double* x0, const double* goal,
                                 const double* u_prev0, const double* u_bias, const double* gain, double* x_out, double* u_out, double* t_out,
                                 is_free_fn is_free, void* ctx, int32_t* collided) {
  try {
    vect_n<double> vlo(nu), vhi(nu), vbw(nu);
    for (int k = 0; k < nu; ++k) { vlo[k] = lo[k]; vhi[k] = hi[k]; vbw[k] = bandwidth[k]; }
    shared_ptr<law_system> S(new law_system());
    S->live = static_cast<const ctrl::kte_nl_system*>(sys);
    shared_ptr<law_topology> ih(new law_access(vlo, vhi, vbw, T, S, 1e30, threshold));
    meaqr_access topo(ih);
    topo.free_fn = is_free;   // NULL: with_collision_check = false; else the loop asks it about every x_next (:550)
    topo.free_ctx = ctx;
    mat<double, mat_structure::square> H(nx);
    for (int r = 0; r < nx; ++r)
      for (int c = 0; c < nx; ++c) H(r, c) = r == c ? 1.0 : 0.0;
    vect_n<double> eta(nx);
    for (int k = 0; k < nx; ++k) eta[k] = 0.0;
    for (std::size_t i = 0; i < N; ++i) {
      vect_n<double> x(nx), g(nx), up(nu), u0(nu);
      for (int k = 0; k < nx; ++k) { x[k] = x0[i * nx + k]; g[k] = goal[i * nx + k]; }
      for (int k = 0; k < nu; ++k) { up[k] = u_prev0[i * nu + k]; u0[k] = u_bias[i * nu + k]; }
      mat<double, mat_structure::rectangular> K(nu, nx);
      for (int r = 0; r < nu; ++r)
        for (int c = 0; c < nx; ++c) K(r, c) = gain[(i * nu + r) * nx + c];
      double t = 0.0;
      const bool was_free = topo.steer(H, K, eta, u0, up, x, g, t, time_limit);
      if (collided) collided[i] = was_free ? 0 : 1;
      for (int k = 0; k < nx; ++k) x_out[i * nx + k] = x[k];
      for (int k = 0; k < nu; ++k) u_out[i * nu + k] = up[k];
      t_out[i] = t;
    }
    return 0;
  } catch (std::exception& e) { return -1; }
}
