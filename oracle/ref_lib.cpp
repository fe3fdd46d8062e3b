// ref_lib.cpp — TEST INFRASTRUCTURE.  Wraps the UNMODIFIED ReaK sources (read where they lie,
// under /root/reference/src) behind a small C interface so the parity tests and the
// cpu_baseline leg of bench.py can run the reference itself.  Nothing in the product path
// links or loads this file.  Built by oracle/Makefile into oracle/_ref/libreak_ref.so.
//
// The model is assembled from the same flat descriptor (include/reak_b200.h) the CUDA
// library consumes, using the reference's own classes exactly the way
// examples/robot_airship/old/CRS_A465_models.cpp:260-822 and ctrl/mbd_kte/test_bm.cpp:46-72 do,
// evaluated with ctrl/ctrl_sys/kte_nl_system.hpp and integrated with
// core/integrators/fixed_step_integrators.hpp (runge_kutta4_integrator<double>).

#include <ReaK/core/base/defs.hpp>
#include <ReaK/core/lin_alg/mat_alg.hpp>
#include <ReaK/core/lin_alg/vect_alg.hpp>
#include <ReaK/core/lin_alg/mat_num_exceptions.hpp>
#include <ReaK/core/kinetostatics/kinetostatics.hpp>
#include <ReaK/core/kinetostatics/motion_jacobians.hpp>
#include <ReaK/core/integrators/fixed_step_integrators.hpp>

#include <ReaK/ctrl/mbd_kte/kte_map_chain.hpp>
#include <ReaK/ctrl/mbd_kte/revolute_joint.hpp>
#include <ReaK/ctrl/mbd_kte/prismatic_joint.hpp>
#include <ReaK/ctrl/mbd_kte/free_joints.hpp>
#include <ReaK/ctrl/mbd_kte/rigid_link.hpp>
#include <ReaK/ctrl/mbd_kte/inertia.hpp>
#include <ReaK/ctrl/mbd_kte/spring.hpp>
#include <ReaK/ctrl/mbd_kte/damper.hpp>
#include <ReaK/ctrl/mbd_kte/torsion_spring.hpp>
#include <ReaK/ctrl/mbd_kte/torsion_damper.hpp>
#include <ReaK/ctrl/mbd_kte/driving_actuator.hpp>
#include <ReaK/ctrl/mbd_kte/jacobian_joint_map.hpp>
#include <ReaK/ctrl/mbd_kte/mass_matrix_calculator.hpp>
#include <ReaK/ctrl/ctrl_sys/kte_nl_system.hpp>
#include <ReaK/geometry/shapes/plane.hpp>
#include <ReaK/geometry/shapes/sphere.hpp>
#include <ReaK/geometry/shapes/capped_cylinder.hpp>
#include <ReaK/geometry/shapes/cylinder.hpp>
#include <ReaK/geometry/shapes/box.hpp>
#include <ReaK/geometry/shapes/circle.hpp>
#include <ReaK/geometry/shapes/capped_rectangle.hpp>
#include <ReaK/geometry/shapes/rectangle.hpp>
#include <ReaK/geometry/proximity/proxy_query_model.hpp>
#include <ReaK/ctrl/graph_alg/node_generators.hpp>
#include <ReaK/ctrl/kte_models/manip_dynamics_model.hpp>
#include <ReaK/core/serialization/xml_archiver.hpp>
#include <ReaK/ctrl/path_planning/topological_search.hpp>

#include "../include/reak_b200.h"
// libreak_b200.so is only needed by rkref_bridge_gpu_check: keep its symbols weak so that this
// checker library still loads (RTLD_NOW) where the product library is not loaded.
#pragma weak rkb_chain_create
#pragma weak rkb_chain_destroy
#pragma weak rkb_chain_input_dim
#pragma weak rkb_chain_state_dim
#pragma weak rkb_chain_is_serial
#pragma weak rkb_eval
#pragma weak rkb_gen_forces
#pragma weak rkb_mass_matrix
#pragma weak rkb_frames
#pragma weak rkb_chain_frame_count
#pragma weak rkb_twist_shaping
#pragma weak rkb_twist_shaping_rows
#pragma weak rkb_twist_shaping_mcm
#pragma weak rkb_steer_batch
#pragma weak rkb_last_kernel_ms
#pragma weak rkb_last_cuda_error
#pragma weak rkb_rollout_rk4
#pragma weak rkb_rollout
#pragma weak rkb_steer_feedback
#pragma weak rkb_strerror
#pragma weak rkb_proxy_create
#pragma weak rkb_proxy_destroy
#pragma weak rkb_min_distance
#pragma weak rkb_is_free
#pragma weak rkb_nearest
#include "../include/reak_b200/reak_bridge.hpp"
#include "steer_law.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

using namespace ReaK;

namespace {

struct ref_model {
  std::vector<shared_ptr<frame_3D<double> > > f3;
  std::vector<shared_ptr<frame_2D<double> > > f2;
  std::vector<shared_ptr<gen_coord<double> > > coords;
  std::vector<shared_ptr<jacobian_gen_3D<double> > > jac3;
  std::vector<shared_ptr<jacobian_gen_2D<double> > > jac2;
  std::vector<shared_ptr<kte::kte_map> > elems;
  shared_ptr<kte::kte_map_chain> chain;
  shared_ptr<kte::mass_matrix_calc> mcalc;
  std::vector<shared_ptr<kte::inertia_gen> > in_gen;   // in the order they were handed to mcalc
  std::vector<shared_ptr<kte::inertia_2D> > in_2d;
  std::vector<shared_ptr<kte::inertia_3D> > in_3d;
  ctrl::kte_nl_system sys;
  int n, nu;
  // free_joint_3D coordinate frames (kte_nl_system::dofs_3D) and their Jacobian holders
  std::vector<shared_ptr<gen_coord<double> > > aux_coords;  // RKB_COORD_GEN: gen_coords that are not system states
  std::vector<shared_ptr<frame_3D<double> > > fcoord;
  std::vector<shared_ptr<jacobian_3D_3D<double> > > jac33;
  std::vector<shared_ptr<frame_2D<double> > > fcoord2;   // free_joint_2D coordinate frames (kte_nl_system::dofs_2D)
  std::vector<shared_ptr<jacobian_2D_2D<double> > > jac22;
  int nx, na;  // state / acceleration dimensions: 2 n + 13 n_free, n + 6 n_free
};

struct ref_handle {
  rkb_chain_desc desc;
  std::vector<rkb_element> elements;
  ref_model* proto;
};

ref_model* build_model(const rkb_chain_desc& d) {
  ref_model* m = new ref_model();
  m->n = d.n_coords;
  m->nu = d.n_inputs;
  const bool is3 = (d.dim == 3);
  for (int i = 0; i < d.n_frames; ++i) {
    if (is3) m->f3.push_back(shared_ptr<frame_3D<double> >(new frame_3D<double>()));
    else     m->f2.push_back(shared_ptr<frame_2D<double> >(new frame_2D<double>()));
  }
  int n_aux = 0;
  for (int e = 0; e < d.n_elements; ++e) if (d.elements[e].kind == RKB_COORD_GEN) ++n_aux;
  std::vector<shared_ptr<gen_coord<double> > > all_coords;  // system coordinates, then the auxiliary gen_coords
  for (int i = 0; i < d.n_coords; ++i) {
    m->coords.push_back(shared_ptr<gen_coord<double> >(new gen_coord<double>()));
    if (is3) m->jac3.push_back(shared_ptr<jacobian_gen_3D<double> >(new jacobian_gen_3D<double>()));
    else     m->jac2.push_back(shared_ptr<jacobian_gen_2D<double> >(new jacobian_gen_2D<double>()));
  }
  for (int e = 0; e < d.n_elements; ++e)
    if (d.elements[e].kind == RKB_FREE_3D) {
      if (!is3 || d.elements[e].coord != (int)m->fcoord.size()) { delete m; return NULL; }
      m->fcoord.push_back(shared_ptr<frame_3D<double> >(new frame_3D<double>()));
      m->jac33.push_back(shared_ptr<jacobian_3D_3D<double> >(new jacobian_3D_3D<double>()));
    }
  for (int e = 0; e < d.n_elements; ++e)
    if (d.elements[e].kind == RKB_FREE_2D) {
      if (is3 || d.elements[e].coord != (int)m->fcoord2.size()) { delete m; return NULL; }
      m->fcoord2.push_back(shared_ptr<frame_2D<double> >(new frame_2D<double>()));
      m->jac22.push_back(shared_ptr<jacobian_2D_2D<double> >(new jacobian_2D_2D<double>()));
    }
  m->nx = 2 * m->n + 13 * (int)m->fcoord.size() + 7 * (int)m->fcoord2.size();
  m->na = m->n + 6 * (int)m->fcoord.size() + 3 * (int)m->fcoord2.size();
  all_coords = m->coords;
  for (int i = 0; i < n_aux; ++i) all_coords.push_back(shared_ptr<gen_coord<double> >(new gen_coord<double>()));
  m->aux_coords.assign(all_coords.begin() + d.n_coords, all_coords.end());
  const rkb_base_frame& b = d.base;
  if (is3) {
    frame_3D<double>& B = *m->f3[d.base_frame];
    B.Position = vect<double,3>(b.position[0], b.position[1], b.position[2]);
    B.Quat = quaternion<double>(vect<double,4>(b.quat[0], b.quat[1], b.quat[2], b.quat[3]));
    B.Velocity = vect<double,3>(b.velocity[0], b.velocity[1], b.velocity[2]);
    B.AngVelocity = vect<double,3>(b.ang_velocity[0], b.ang_velocity[1], b.ang_velocity[2]);
    B.Acceleration = vect<double,3>(b.acceleration[0], b.acceleration[1], b.acceleration[2]);
    B.AngAcceleration = vect<double,3>(b.ang_acceleration[0], b.ang_acceleration[1], b.ang_acceleration[2]);
  } else {
    frame_2D<double>& B = *m->f2[d.base_frame];
    B.Position = vect<double,2>(b.position[0], b.position[1]);
    B.Rotation = rot_mat_2D<double>(b.quat[0]);
    B.Velocity = vect<double,2>(b.velocity[0], b.velocity[1]);
    B.AngVelocity = b.ang_velocity[0];
    B.Acceleration = vect<double,2>(b.acceleration[0], b.acceleration[1]);
    B.AngAcceleration = b.ang_acceleration[0];
  }

  m->chain = shared_ptr<kte::kte_map_chain>(new kte::kte_map_chain("chain"));
  m->mcalc = shared_ptr<kte::mass_matrix_calc>(new kte::mass_matrix_calc("mcalc"));
  std::vector<shared_ptr<kte::reacting_kte_gen> > joint_of_elem(d.n_elements);
  std::vector<shared_ptr<kte::inertia_gen> > gen_inertias;
  std::vector<shared_ptr<kte::driving_actuator_gen> > actuators(d.n_inputs);

  // Pass 1: joints (an actuator precedes the joint it drives in the CRS chain order,
  // CRS_A465_models.cpp:748-788, and needs the joint object at construction).
  for (int e = 0; e < d.n_elements; ++e) {
    const rkb_element& E = d.elements[e];
    const std::string nm = "e" + std::to_string(e);
    switch (E.kind) {
      case RKB_REVOLUTE_3D:
        joint_of_elem[e] = shared_ptr<kte::reacting_kte_gen>(new kte::revolute_joint_3D(nm, m->coords[E.coord],
            vect<double,3>(E.p[0], E.p[1], E.p[2]), m->f3[E.frame_a], m->f3[E.frame_b], m->jac3[E.coord]));
        break;
      case RKB_PRISMATIC_3D:
        joint_of_elem[e] = shared_ptr<kte::reacting_kte_gen>(new kte::prismatic_joint_3D(nm, m->coords[E.coord],
            vect<double,3>(E.p[0], E.p[1], E.p[2]), m->f3[E.frame_a], m->f3[E.frame_b], m->jac3[E.coord]));
        break;
      case RKB_REVOLUTE_2D:
        joint_of_elem[e] = shared_ptr<kte::reacting_kte_gen>(new kte::revolute_joint_2D(nm, m->coords[E.coord],
            m->f2[E.frame_a], m->f2[E.frame_b], m->jac2[E.coord]));
        break;
      case RKB_PRISMATIC_2D:
        joint_of_elem[e] = shared_ptr<kte::reacting_kte_gen>(new kte::prismatic_joint_2D(nm, m->coords[E.coord],
            vect<double,2>(E.p[0], E.p[1]), m->f2[E.frame_a], m->f2[E.frame_b], m->jac2[E.coord]));
        break;
      default: break;
    }
  }

  for (int e = 0; e < d.n_elements; ++e) {
    const rkb_element& E = d.elements[e];
    const std::string nm = "e" + std::to_string(e);
    shared_ptr<kte::kte_map> k;
    switch (E.kind) {
      case RKB_REVOLUTE_3D: case RKB_PRISMATIC_3D: case RKB_REVOLUTE_2D: case RKB_PRISMATIC_2D:
        k = joint_of_elem[e]; break;
      case RKB_COORD_GEN: {  // not a KTE: a free-standing gen_coord with the values it keeps
        gen_coord<double>& g = *all_coords.at(E.coord);
        g.q = E.p[0]; g.q_dot = E.p[1]; g.q_ddot = E.p[2];
        m->elems.push_back(shared_ptr<kte::kte_map>());
        continue; }
      case RKB_RIGID_LINK_GEN:
        k = shared_ptr<kte::kte_map>(new kte::rigid_link_gen(nm, all_coords.at(E.coord), all_coords.at(E.aux), E.p[0])); break;
      case RKB_SPRING_GEN:
        k = shared_ptr<kte::kte_map>(new kte::spring_gen(nm, all_coords.at(E.coord), all_coords.at(E.aux), E.p[0], E.p[1], E.p[2])); break;
      case RKB_DAMPER_GEN:
        k = shared_ptr<kte::kte_map>(new kte::damper_gen(nm, all_coords.at(E.coord), all_coords.at(E.aux), E.p[0])); break;
      case RKB_FREE_2D:
        k = shared_ptr<kte::kte_map>(new kte::free_joint_2D(nm, m->fcoord2[E.coord], m->f2[E.frame_a], m->f2[E.frame_b], m->jac22[E.coord]));
        break;
      case RKB_FREE_3D:
        k = shared_ptr<kte::kte_map>(new kte::free_joint_3D(nm, m->fcoord[E.coord], m->f3[E.frame_a], m->f3[E.frame_b], m->jac33[E.coord]));
        break;
      case RKB_RIGID_LINK_3D: {
        pose_3D<double> off(weak_ptr<pose_3D<double> >(), vect<double,3>(E.p[0], E.p[1], E.p[2]),
                            quaternion<double>(vect<double,4>(E.p[3], E.p[4], E.p[5], E.p[6])));
        k = shared_ptr<kte::kte_map>(new kte::rigid_link_3D(nm, m->f3[E.frame_a], m->f3[E.frame_b], off));
        break; }
      case RKB_INERTIA_3D: {
        shared_ptr<kte::joint_dependent_frame_3D> dep(new kte::joint_dependent_frame_3D(m->f3[E.frame_a]));
        for (int c = 0; c < d.n_coords; ++c)
          if ((E.upstream >> c) & 1u) dep->add_joint(m->coords[c], m->jac3[c]);
        for (std::size_t c = 0; c < m->fcoord.size(); ++c)
          if ((E.upstream >> (32 + c)) & 1u) dep->add_joint(m->fcoord[c], m->jac33[c]);
        shared_ptr<kte::inertia_3D> in(new kte::inertia_3D(nm, dep, E.p[0],
            mat<double,mat_structure::symmetric>(E.p[1], E.p[2], E.p[3], E.p[4], E.p[5], E.p[6])));
        *m->mcalc << in; m->in_3d.push_back(in); k = in; break; }
      case RKB_INERTIA_GEN: {
        shared_ptr<kte::joint_dependent_gen_coord> dep(new kte::joint_dependent_gen_coord(m->coords[E.coord]));
        dep->add_joint(m->coords[E.coord], shared_ptr<jacobian_gen_gen<double> >(new jacobian_gen_gen<double>(1.0, 0.0)));
        shared_ptr<kte::inertia_gen> in(new kte::inertia_gen(nm, dep, E.p[0]));
        gen_inertias.push_back(in); k = in; break; }
      case RKB_ACTUATOR_GEN: {
        if (E.frame_b < 0 || E.frame_b >= d.n_elements || !joint_of_elem[E.frame_b]) { delete m; return NULL; }
        shared_ptr<kte::driving_actuator_gen> a(new kte::driving_actuator_gen(nm, m->coords[E.coord], joint_of_elem[E.frame_b]));
        actuators.at(E.aux) = a; k = a; break; }
      case RKB_TORSION_SPRING_3D:
        k = shared_ptr<kte::kte_map>(new kte::torsion_spring_3D(nm, m->f3[E.frame_a], m->f3[E.frame_b], E.p[0], E.p[1])); break;
      case RKB_TORSION_DAMPER_3D:
        k = shared_ptr<kte::kte_map>(new kte::torsion_damper_3D(nm, m->f3[E.frame_a], m->f3[E.frame_b], E.p[0])); break;
      case RKB_SPRING_3D:
        k = shared_ptr<kte::kte_map>(new kte::spring_3D(nm, m->f3[E.frame_a], m->f3[E.frame_b], E.p[0], E.p[1], E.p[2])); break;
      case RKB_DAMPER_3D:
        k = shared_ptr<kte::kte_map>(new kte::damper_3D(nm, m->f3[E.frame_a], m->f3[E.frame_b], E.p[0])); break;
      case RKB_RIGID_LINK_2D: {
        pose_2D<double> off(weak_ptr<pose_2D<double> >(), vect<double,2>(E.p[0], E.p[1]), rot_mat_2D<double>(E.p[2]));
        k = shared_ptr<kte::kte_map>(new kte::rigid_link_2D(nm, m->f2[E.frame_a], m->f2[E.frame_b], off));
        break; }
      case RKB_INERTIA_2D: {
        kte::jacobian_joint_map_2D jm;
        for (int c = 0; c < d.n_coords; ++c)
          if ((E.upstream >> c) & 1u) jm[m->coords[c]] = m->jac2[c];
        shared_ptr<kte::joint_dependent_frame_2D> dep(new kte::joint_dependent_frame_2D(m->f2[E.frame_a], jm));
        for (std::size_t c = 0; c < m->fcoord2.size(); ++c)
          if ((E.upstream >> (32 + c)) & 1u) dep->add_joint(m->fcoord2[c], m->jac22[c]);
        shared_ptr<kte::inertia_2D> in(new kte::inertia_2D(nm, dep, E.p[0], E.p[1]));
        *m->mcalc << in; m->in_2d.push_back(in); k = in; break; }
      case RKB_TORSION_SPRING_2D:
        k = shared_ptr<kte::kte_map>(new kte::torsion_spring_2D(nm, m->f2[E.frame_a], m->f2[E.frame_b], E.p[0], E.p[1])); break;
      case RKB_TORSION_DAMPER_2D:
        k = shared_ptr<kte::kte_map>(new kte::torsion_damper_2D(nm, m->f2[E.frame_a], m->f2[E.frame_b], E.p[0])); break;
      case RKB_SPRING_2D:
        k = shared_ptr<kte::kte_map>(new kte::spring_2D(nm, m->f2[E.frame_a], m->f2[E.frame_b], E.p[0], E.p[1], E.p[2])); break;
      case RKB_DAMPER_2D:
        k = shared_ptr<kte::kte_map>(new kte::damper_2D(nm, m->f2[E.frame_a], m->f2[E.frame_b], E.p[0])); break;
      default:
        delete m; return NULL;
    }
    m->elems.push_back(k);
    *m->chain << k;
  }
  // CRS_A465_models.cpp:791-822: link inertias, then motor inertias, then the coordinates.
  for (std::size_t i = 0; i < gen_inertias.size(); ++i) *m->mcalc << gen_inertias[i];
  m->in_gen = gen_inertias;
  for (int c = 0; c < d.n_coords; ++c) *m->mcalc << m->coords[c];
  for (std::size_t c = 0; c < m->fcoord2.size(); ++c) *m->mcalc << m->fcoord2[c];
  for (std::size_t c = 0; c < m->fcoord.size(); ++c) *m->mcalc << m->fcoord[c];

  m->sys.dofs_gen = m->coords;
  m->sys.dofs_3D = m->fcoord;
  m->sys.dofs_2D = m->fcoord2;
  for (int i = 0; i < d.n_inputs; ++i) {
    if (!actuators[i]) { delete m; return NULL; }
    m->sys.inputs.push_back(actuators[i]);
  }
  m->sys.chain = m->chain;
  m->sys.mass_calc = m->mcalc;
  return m;
}

// Same wrapper as num_int_dtnl_sys::rate_function_impl (num_int_dtnl_system.hpp:85-99).
class rate_fn : public state_rate_function<double> {
 public:
  const ctrl::kte_nl_system* sys;
  vect_n<double> u;
  rate_fn(const ctrl::kte_nl_system* s, const vect_n<double>& aU) : sys(s), u(aU) {}
  virtual void RK_CALL computeStateRate(double t, const vect_n<double>& x, vect_n<double>& xd) {
    xd = sys->get_state_derivative(*sys, x, u, t);
  }
};

// The same wrapper with the input read from a sampled trajectory: node round(t / (dt / 2)).  runge_kutta4_integrator
// asks for the rate at t, t + dt/2 (twice) and t + dt (fixed_step_integrators.hpp:273-291) — the instants at which
// ctrl::detail::runge_kutta4_integrate_impl reads its input trajectory (runge_kutta4_integrator_sys.hpp:66-91).
class rate_fn_nodes : public state_rate_function<double> {
 public:
  const ctrl::kte_nl_system* sys;
  const double* nodes;
  int nu;
  long long n_nodes;
  double half_dt;
  rate_fn_nodes(const ctrl::kte_nl_system* s, const double* aNodes, int aNu, long long aCount, double aHalf)
      : sys(s), nodes(aNodes), nu(aNu), n_nodes(aCount), half_dt(aHalf) {}
  virtual void RK_CALL computeStateRate(double t, const vect_n<double>& x, vect_n<double>& xd) {
    long long j = llround(t / half_dt);
    if (j < 0) j = 0;
    if (j >= n_nodes) j = n_nodes - 1;
    vect_n<double> u(nu);
    for (int k = 0; k < nu; ++k) u[k] = nodes[j * nu + k];
    xd = sys->get_state_derivative(*sys, x, u, t);
  }
};

// scheme: enum rkb_scheme (1 euler, 2 midpoint, 4 runge_kutta4, 5 runge_kutta5)
integrator<double>* make_integrator(int scheme, const vect_n<double>& x, double dt, const shared_ptr<state_rate_function<double> >& fn) {
  switch (scheme) {
    case RKB_SCHEME_EULER: return new euler_integrator<double>("euler", x, 0.0, dt, fn);
    case RKB_SCHEME_MIDPOINT: return new midpoint_integrator<double>("midpoint", x, 0.0, dt, fn);
    case RKB_SCHEME_RK5: return new runge_kutta5_integrator<double>("rk5", x, 0.0, dt, fn);
    default: return new runge_kutta4_integrator<double>("rk4", x, 0.0, dt, fn);
  }
}

void rk4_range(ref_model* m, std::size_t i0, std::size_t i1, const double* x0, const double* u,
               double dt, int n_steps, double* xout, int32_t* status, int scheme = RKB_SCHEME_RK4) {
  const int nx = m->nx, nu = m->nu;
  for (std::size_t i = i0; i < i1; ++i) {
    vect_n<double> x(nx), uu(nu);
    for (int k = 0; k < nx; ++k) x[k] = x0[i * nx + k];
    for (int k = 0; k < nu; ++k) uu[k] = u[i * nu + k];
    int32_t st = 0;
    if (n_steps > 0) {
      shared_ptr<state_rate_function<double> > fn(new rate_fn(&m->sys, uu));
      integrator<double>* integ = make_integrator(scheme, x, dt, fn);
      try {
        // The loops of fixed_step_integrators.hpp (:76, :191, :275, :365) are time-driven; an end time
        // half a step short of n_steps*dt makes them run exactly n_steps steps.  (runge_kutta5 moves
        // its clock back and forth within a step, :392-397, but ends each step at t + dt.)
        integ->integrate((double(n_steps) - 0.5) * dt);
      } catch (singularity_error&) { st |= RKB_STATUS_SINGULAR; }
      int k = 0;
      for (std::vector<double>::const_iterator it = integ->getStateBegin(); it != integ->getStateEnd(); ++it, ++k)
        xout[i * nx + k] = *it;
      delete integ;
    } else {
      for (int k = 0; k < nx; ++k) xout[i * nx + k] = x[k];
    }
    for (int k = 0; k < nx; ++k) if (!std::isfinite(xout[i * nx + k])) st |= RKB_STATUS_NONFINITE;
    if (status) status[i] = st;
  }
}

}  // namespace

extern "C" {

// the live kte_nl_system of a handle, for the units that drive it through other reference code (ref_steer_law.cpp)
const void* rkref_kte_nl_system(void* hv) { return &static_cast<ref_handle*>(hv)->proto->sys; }

// RK4 with an input trajectory sampled at every half step ([N][2 n_steps + 1][nu]): the reference's own
// runge_kutta4_integrator<double> over its kte_nl_system, the rate function reading the node that belongs to the time it is asked at.
int rkref_rk4_inputs(void* hv, std::size_t N, const double* x0, const double* u_nodes, double dt, int n_steps, double* xout, int32_t* status) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  const long long J = 2LL * n_steps + 1;
  for (std::size_t i = 0; i < N; ++i) {
    vect_n<double> x(nx);
    for (int k = 0; k < nx; ++k) x[k] = x0[i * nx + k];
    int32_t st = 0;
    if (n_steps > 0) {
      shared_ptr<state_rate_function<double> > fn(new rate_fn_nodes(&m->sys, u_nodes + i * J * nu, nu, J, 0.5 * dt));
      runge_kutta4_integrator<double> integ("rk4", x, 0.0, dt, fn);
      try {
        integ.integrate((double(n_steps) - 0.5) * dt);
      } catch (singularity_error&) { st |= RKB_STATUS_SINGULAR; }
      int k = 0;
      for (std::vector<double>::const_iterator it = integ.getStateBegin(); it != integ.getStateEnd(); ++it, ++k) xout[i * nx + k] = *it;
    } else {
      for (int k = 0; k < nx; ++k) xout[i * nx + k] = x[k];
    }
    for (int k = 0; k < nx; ++k) if (!std::isfinite(xout[i * nx + k])) st |= RKB_STATUS_NONFINITE;
    if (status) status[i] = st;
  }
  return 0;
}

// a29: the legacy manipulator model, kte::manipulator_dynamics_model::computeStateRate (ctrl/kte_models/
// manip_dynamics_model.cpp:152-218; the same code as ctrl/mbd_kte/manipulator_model.cpp:292-355, whose translation unit
// g++ 13 rejects).  Its state is BLOCKED — all positions, then all velocities — and it does not zero q_ddot itself
// (the coordinates keep the 0 they were created with).  The model is assembled over the very objects of this handle:
// coordinates, inertias (in mass_calc order), inputs and the chain.
int rkref_manip_state_rate(void* hv, std::size_t N, const double* x_blocked, const double* u, double* xdot_blocked, int32_t* status) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int n = m->n, nx = 2 * n, nu = m->nu;
  kte::manipulator_dynamics_model mdl("legacy");
  mdl.setModel(m->chain);
  for (int c = 0; c < n; ++c) mdl << m->coords[c];
  for (std::size_t i = 0; i < m->in_3d.size(); ++i) mdl << m->in_3d[i];
  for (std::size_t i = 0; i < m->in_2d.size(); ++i) mdl << m->in_2d[i];
  for (std::size_t i = 0; i < m->in_gen.size(); ++i) mdl << m->in_gen[i];
  for (std::size_t i = 0; i < m->sys.inputs.size(); ++i) mdl << m->sys.inputs[i];
  if ((int)mdl.getJointStatesCount() != nx || (int)mdl.getInputsCount() != nu) return -1;
  for (int c = 0; c < n; ++c) m->coords[c]->q_ddot = 0.0;
  vect_n<double> x(nx), xd(nx), uu(nu);
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) x[k] = x_blocked[i * nx + k];
    for (int k = 0; k < nu; ++k) uu[k] = u[i * nu + k];
    int32_t st = 0;
    try {
      mdl.setInput(uu);
      mdl.computeStateRate(0.0, x, xd);
      for (int k = 0; k < nx; ++k) xdot_blocked[i * nx + k] = xd[k];
    } catch (singularity_error&) {
      st |= RKB_STATUS_SINGULAR;
      for (int k = 0; k < nx; ++k) xdot_blocked[i * nx + k] = 0.0;
    }
    if (status) status[i] = st;
  }
  return 0;
}

// f4: the live kte_nl_system of this handle written by the reference's own XML archiver (core/serialization/
// xml_archiver.cpp) — the `.rkx` format of examples/robot_airship/build_P3R3R_model.cpp:78.  what: 0 = the whole
// kte_nl_system (dofs, inputs, chain, mass_calc), 1 = the kte_map_chain alone.
int rkref_save_rkx(void* hv, const char* path, int what) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  try {
    serialization::xml_oarchive out(path);
    if (what == 1) {
      out << m->chain;
    } else {
      shared_ptr<ctrl::kte_nl_system> sys(new ctrl::kte_nl_system(m->sys));
      out << sys;
    }
    return 0;
  } catch (std::exception&) {
    return -1;
  }
}

// f4: nearest neighbours by the reference's own linear scan — ReaK::pp::min_dist_linear_search
// (ctrl/path_planning/topological_search.hpp:91-112 for k == 1 without a radius, :238-270 otherwise) over vect_n points
// with the metric of the vect_n topologies, norm_2(difference) (core/lin_alg/vect_alg.hpp:2314-2333).
namespace {
struct nn_distance {  // distance from the query to vertex number v
  const std::vector<vect_n<double> >* pts;
  const vect_n<double>* q;
  double operator()(std::size_t v) const { return norm_2((*pts)[v] - *q); }
};
}
int rkref_nearest(std::size_t V, const double* vertices, std::size_t Q, const double* queries, int dim, int k, double radius,
                  int32_t* index, double* distance, int32_t* count) {
  std::vector<vect_n<double> > pts(V, vect_n<double>(dim));
  std::vector<std::size_t> ids(V);
  for (std::size_t i = 0; i < V; ++i) { ids[i] = i; for (int c = 0; c < dim; ++c) pts[i][c] = vertices[i * dim + c]; }
  vect_n<double> q(dim);
  for (std::size_t i = 0; i < Q; ++i) {
    for (int c = 0; c < dim; ++c) q[c] = queries[i * dim + c];
    nn_distance dist; dist.pts = &pts; dist.q = &q;
    for (int r = 0; r < k; ++r) { index[i * k + r] = -1; if (distance) distance[i * k + r] = std::numeric_limits<double>::infinity(); }
    int found = 0;
    if (k == 1 && radius == std::numeric_limits<double>::infinity()) {
      std::vector<std::size_t>::iterator it = pp::min_dist_linear_search<double>(ids.begin(), ids.end(), dist);
      if (it != ids.end()) { index[i] = int32_t(*it); if (distance) distance[i] = dist(*it); found = 1; }
    } else {
      std::vector<std::size_t> out(k);
      std::vector<std::size_t>::iterator last = pp::min_dist_linear_search<double>(ids.begin(), ids.end(), out.begin(), dist, std::size_t(k), radius);
      for (std::vector<std::size_t>::iterator o = out.begin(); o != last; ++o, ++found) {
        index[i * k + found] = int32_t(*o);
        if (distance) distance[i * k + found] = dist(*o);
      }
    }
    if (count) count[i] = found;
  }
  return 0;
}

// GPU drop-in check of ReaK::pp::batched_neighbor_search (reak_bridge.hpp): a batch of query points against a range of
// vertex ids with a position map, both forms, compared in C++ with ReaK::pp::min_dist_linear_search run point by point.
// Needs libreak_b200.so loaded first (RTLD_GLOBAL) and a GPU.  Returns the number of mismatches, or -1 (msg).
namespace {
struct nn_position_map {
  typedef vect_n<double> value_type;
  typedef std::size_t key_type;
  typedef const vect_n<double>& reference;
  const std::vector<vect_n<double> >* pts;
};
inline const vect_n<double>& get(const nn_position_map& m, std::size_t v) { return (*m.pts)[v]; }
struct nn_space {};
}
int rkref_nn_bridge_check(std::size_t V, const double* vertices, std::size_t Q, const double* queries, int dim, int k, double radius,
                          char* msg, int msg_len) {
  try {
    if (!rkb_nearest) throw std::runtime_error("libreak_b200.so is not loaded (load it with RTLD_GLOBAL first)");
    std::vector<vect_n<double> > pts(V, vect_n<double>(dim)), qs(Q, vect_n<double>(dim));
    std::vector<std::size_t> ids(V);
    for (std::size_t i = 0; i < V; ++i) { ids[i] = i; for (int c = 0; c < dim; ++c) pts[i][c] = vertices[i * dim + c]; }
    for (std::size_t i = 0; i < Q; ++i) for (int c = 0; c < dim; ++c) qs[i][c] = queries[i * dim + c];
    nn_position_map pm; pm.pts = &pts;
    pp::batched_neighbor_search nn(0);
    int bad = 0;
    std::vector<std::vector<std::size_t>::iterator> one = nn(qs, ids.begin(), ids.end(), nn_space(), pm);
    std::vector<std::vector<std::size_t> > many;
    nn(qs, ids.begin(), ids.end(), many, nn_space(), pm, std::size_t(k), radius);
    for (std::size_t i = 0; i < Q; ++i) {
      nn_distance dist; dist.pts = &pts; dist.q = &qs[i];
      std::vector<std::size_t>::iterator it = pp::min_dist_linear_search<double>(ids.begin(), ids.end(), dist);
      if (it != one[i]) ++bad;
      std::vector<std::size_t> out(k);
      std::vector<std::size_t>::iterator last = pp::min_dist_linear_search<double>(ids.begin(), ids.end(), out.begin(), dist, std::size_t(k), radius);
      out.erase(last, out.end());
      if (out != many[i]) ++bad;
    }
    return bad;
  } catch (std::exception& e) {
    if (msg && msg_len > 0) { std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
    return -1;
  }
}

void* rkref_create(const rkb_chain_desc* d) {
  if (!d || !d->elements) return NULL;
  ref_handle* h = new ref_handle();
  h->desc = *d;
  h->elements.assign(d->elements, d->elements + d->n_elements);
  h->desc.elements = h->elements.data();
  try { h->proto = build_model(h->desc); } catch (...) { h->proto = NULL; }
  if (!h->proto) { delete h; return NULL; }
  return h;
}

void rkref_destroy(void* hv) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  if (!h) return;
  delete h->proto;
  delete h;
}

int rkref_eval(void* hv, std::size_t N, const double* x, const double* u, double* xdot, int32_t* status) {
  ref_model* m = static_cast<ref_handle*>(hv)->proto;
  const int nx = m->nx, nu = m->nu;
  vect_n<double> p(nx), uu(nu);
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    for (int k = 0; k < nu; ++k) uu[k] = u[i * nu + k];
    int32_t st = 0;
    try {
      vect_n<double> pd = m->sys.get_state_derivative(m->sys, p, uu, 0.0);
      for (int k = 0; k < nx; ++k) xdot[i * nx + k] = pd[k];
    } catch (singularity_error&) {
      st |= RKB_STATUS_SINGULAR;
      for (int k = 0; k < nx; ++k) xdot[i * nx + k] = std::nan("");
    }
    if (status) status[i] = st;
  }
  return 0;
}

// gen_coord::f after doMotion/clearForce/doForce with q_ddot = 0 (kte_nl_system.hpp:240-253).
int rkref_gen_forces(void* hv, std::size_t N, const double* x, const double* u, double* f) {
  ref_model* m = static_cast<ref_handle*>(hv)->proto;
  const int nx = m->nx, nu = m->nu;
  vect_n<double> p(nx), uu(nu);
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    for (int k = 0; k < nu; ++k) uu[k] = u[i * nu + k];
    m->sys.apply_states_and_inputs(p, uu);
    m->chain->doMotion();
    m->chain->clearForce();
    m->chain->doForce();
    for (int k = 0; k < m->n; ++k) f[i * m->na + k] = m->coords[k]->f;
    for (std::size_t c = 0; c < m->fcoord.size(); ++c)
      for (int k = 0; k < 3; ++k) {
        f[i * m->na + m->n + 6 * c + k] = m->fcoord[c]->Force[k];
        f[i * m->na + m->n + 6 * c + 3 + k] = m->fcoord[c]->Torque[k];
      }
    for (std::size_t c = 0; c < m->fcoord2.size(); ++c) {
      f[i * m->na + m->n + 3 * c] = m->fcoord2[c]->Force[0];
      f[i * m->na + m->n + 3 * c + 1] = m->fcoord2[c]->Force[1];
      f[i * m->na + m->n + 3 * c + 2] = m->fcoord2[c]->Torque;
    }
  }
  return 0;
}

// M (and Mdot) row-major n x n per sample (mass_matrix_calculator.cpp:80-98).
int rkref_mass(void* hv, std::size_t N, const double* x, double* M, double* Mdot) {
  ref_model* m = static_cast<ref_handle*>(hv)->proto;
  const int nx = m->nx, nu = m->nu, n = m->na;
  vect_n<double> p(nx), uu(nu, 0.0);
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    m->sys.apply_states_and_inputs(p, uu);
    m->chain->doMotion();
    mat<double,mat_structure::symmetric> Ms(n);
    mat<double,mat_structure::square> Md(n);
    if (Mdot) m->mcalc->getMassMatrixAndDerivative(Ms, Md);
    else      m->mcalc->getMassMatrix(Ms);
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) {
        M[(i * n + r) * n + c] = Ms(r, c);
        if (Mdot) Mdot[(i * n + r) * n + c] = Md(r, c);
      }
  }
  return 0;
}

// World-frame kinematics of every frame after doMotion, for debugging the restatement:
// per frame 19 doubles: Position3, Quat4, Velocity3, AngVelocity3, Acceleration3, AngAcceleration3
// (2D: Position2, (cos,sin), Velocity2, AngVelocity, Acceleration2, AngAcceleration, rest 0).
int rkref_frames(void* hv, const double* x, const double* u, double* out) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  vect_n<double> p(nx), uu(nu);
  for (int k = 0; k < nx; ++k) p[k] = x[k];
  for (int k = 0; k < nu; ++k) uu[k] = u[k];
  m->sys.apply_states_and_inputs(p, uu);
  m->chain->doMotion();
  m->chain->clearForce();
  m->chain->doForce();
  for (int i = 0; i < h->desc.n_frames; ++i) {
    double* o = out + 25 * i;
    for (int k = 0; k < 25; ++k) o[k] = 0.0;
    if (h->desc.dim == 3) {
      const frame_3D<double>& F = *m->f3[i];
      for (int k = 0; k < 3; ++k) { o[k] = F.Position[k]; o[7 + k] = F.Velocity[k]; o[10 + k] = F.AngVelocity[k];
                                    o[13 + k] = F.Acceleration[k]; o[16 + k] = F.AngAcceleration[k];
                                    o[19 + k] = F.Force[k]; o[22 + k] = F.Torque[k]; }
      for (int k = 0; k < 4; ++k) o[3 + k] = F.Quat[k];
    } else {
      const frame_2D<double>& F = *m->f2[i];
      for (int k = 0; k < 2; ++k) { o[k] = F.Position[k]; o[3 + k] = F.Rotation[k]; o[7 + k] = F.Velocity[k];
                                    o[13 + k] = F.Acceleration[k]; o[19 + k] = F.Force[k]; }
      o[10] = F.AngVelocity; o[16] = F.AngAcceleration; o[22] = F.Torque;
    }
  }
  return 0;
}

// ---- planar models (proxy_query_pair_2D, proxy_query_model.cpp:73-212) -----------------------------------------
namespace {
struct planar_pair {
  std::vector<shared_ptr<geom::shape_2D> > shapes;
  shared_ptr<geom::proxy_query_model_2D> mdl[2];
  shared_ptr<geom::proxy_query_pair_2D> pair;
};
// rkb_shape of a planar shape: position[0..1], quat[0..1] = (cos, sin) of its own rotation
bool build_planar_pair(ref_model* m, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2, planar_pair& P) {
  P.mdl[0] = shared_ptr<geom::proxy_query_model_2D>(new geom::proxy_query_model_2D("model1"));
  P.mdl[1] = shared_ptr<geom::proxy_query_model_2D>(new geom::proxy_query_model_2D("model2"));
  for (int k = 0; k < n1 + n2; ++k) {
    const rkb_shape& s = k < n1 ? m1[k] : m2[k - n1];
    shared_ptr<pose_2D<double> > anchor;
    if (s.anchor >= 0) anchor = m->f2[s.anchor];
    const pose_2D<double> pose(weak_ptr<pose_2D<double> >(), vect<double,2>(s.position[0], s.position[1]),
                               rot_mat_2D<double>(vect<double,2>(s.quat[0], s.quat[1])));
    shared_ptr<geom::shape_2D> sh;
    switch (s.kind) {
      case RKB_SHAPE_CIRCLE: sh = shared_ptr<geom::shape_2D>(new geom::circle("ci", anchor, pose, s.dims[0])); break;
      case RKB_SHAPE_CRECT: sh = shared_ptr<geom::shape_2D>(new geom::capped_rectangle("cr", anchor, pose, vect<double,2>(s.dims[0], s.dims[1]))); break;
      case RKB_SHAPE_RECTANGLE: sh = shared_ptr<geom::shape_2D>(new geom::rectangle("re", anchor, pose, vect<double,2>(s.dims[0], s.dims[1]))); break;
      default: return false;
    }
    P.shapes.push_back(sh);
    P.mdl[k < n1 ? 0 : 1]->addShape(sh);
  }
  P.pair = shared_ptr<geom::proxy_query_pair_2D>(new geom::proxy_query_pair_2D("pair", P.mdl[0], P.mdl[1]));
  return true;
}
int min_distance_2d(ref_model* m, std::size_t N, const double* x, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2,
                    double* dist, int32_t* finder, double* points) {
  planar_pair P;
  if (!build_planar_pair(m, m1, n1, m2, n2, P)) return -1;
  const int nx = m->nx, nu = m->nu;
  vect_n<double> p(nx), uu(nu);
  for (int k = 0; k < nu; ++k) uu[k] = 0.0;
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    m->sys.apply_states_and_inputs(p, uu);
    m->chain->doMotion();
    shared_ptr<geom::proximity_finder_2D> f = P.pair->findMinimumDistance();
    if (!f) {
      dist[i] = std::numeric_limits<double>::infinity();
      if (finder) finder[i] = -1;
      if (points) for (int k = 0; k < 6; ++k) points[6 * i + k] = 0.0;
      continue;
    }
    const geom::proximity_record_2D r = f->getLastResult();
    dist[i] = r.mDistance;
    if (points) {
      points[6 * i + 0] = r.mPoint1[0]; points[6 * i + 1] = r.mPoint1[1]; points[6 * i + 2] = 0.0;
      points[6 * i + 3] = r.mPoint2[0]; points[6 * i + 4] = r.mPoint2[1]; points[6 * i + 5] = 0.0;
    }
    if (finder) {  // every pair of planar shapes has a finder: index a * n2 + b, found through the finder's two shapes
      finder[i] = -2;
      const geom::shape_2D* s1 = f->getShape1().get();
      const geom::shape_2D* s2 = f->getShape2().get();
      for (int a = 0; a < n1 && finder[i] < 0; ++a)
        for (int b = 0; b < n2; ++b) {
          const geom::shape_2D* sa = P.shapes[a].get();
          const geom::shape_2D* sb = P.shapes[n1 + b].get();
          if ((sa == s1 && sb == s2) || (sa == s2 && sb == s1)) { finder[i] = (int32_t)(a * n2 + b); break; }
        }
    }
  }
  return n1 * n2;
}
int collision_points_2d(ref_model* m, std::size_t N, const double* x, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2,
                        int max_records, int32_t* count, int32_t* finder, double* records) {
  planar_pair P;
  if (!build_planar_pair(m, m1, n1, m2, n2, P)) return -1;
  const int nx = m->nx, nu = m->nu;
  vect_n<double> p(nx), uu(nu);
  for (int k = 0; k < nu; ++k) uu[k] = 0.0;
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    m->sys.apply_states_and_inputs(p, uu);
    m->chain->doMotion();
    std::vector<geom::proximity_record_2D> out;
    P.pair->gatherCollisionPoints(out);
    count[i] = int32_t(out.size());
    for (int r = 0; r < max_records; ++r) {
      double* o = records + (i * max_records + r) * 7;
      if (r < int(out.size())) {
        o[0] = out[r].mDistance;
        o[1] = out[r].mPoint1[0]; o[2] = out[r].mPoint1[1]; o[3] = 0.0;
        o[4] = out[r].mPoint2[0]; o[5] = out[r].mPoint2[1]; o[6] = 0.0;
      } else {
        o[0] = std::numeric_limits<double>::infinity();
        for (int k = 1; k < 7; ++k) o[k] = 0.0;
      }
      if (finder) finder[i * max_records + r] = -1;
    }
  }
  return 0;
}
}  // namespace

// proxy_query_pair_3D::findMinimumDistance of the live reference (geometry/proximity/proxy_query_model.cpp)
// for two proximity models given as rkb_shape lists, the shapes of either riding on frames of this model
// (anchor = frame id) or fixed in the world (-1), after kte_map_chain::doMotion at each state x[i].
// finder[i] = index of the returned finder in createProxFinderList order, found through its two shapes.
int rkref_min_distance(void* hv, std::size_t N, const double* x, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2,
                       double* dist, int32_t* finder, double* points) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  if (h->desc.dim == 2) return min_distance_2d(m, N, x, m1, n1, m2, n2, dist, finder, points);
  if (h->desc.dim != 3) return -1;
  const int nx = m->nx, nu = m->nu;
  std::vector<shared_ptr<geom::shape_3D> > shapes;
  shared_ptr<geom::proxy_query_model_3D> mdl[2];
  mdl[0] = shared_ptr<geom::proxy_query_model_3D>(new geom::proxy_query_model_3D("model1"));
  mdl[1] = shared_ptr<geom::proxy_query_model_3D>(new geom::proxy_query_model_3D("model2"));
  for (int k = 0; k < n1 + n2; ++k) {
    const rkb_shape& s = k < n1 ? m1[k] : m2[k - n1];
    shared_ptr<pose_3D<double> > anchor;
    if (s.anchor >= 0) anchor = m->f3[s.anchor];
    const pose_3D<double> pose(weak_ptr<pose_3D<double> >(), vect<double,3>(s.position[0], s.position[1], s.position[2]),
                               quaternion<double>(vect<double,4>(s.quat[0], s.quat[1], s.quat[2], s.quat[3])));
    shared_ptr<geom::shape_3D> sh;
    switch (s.kind) {
      case RKB_SHAPE_PLANE: sh = shared_ptr<geom::shape_3D>(new geom::plane("p", anchor, pose, vect<double,2>(s.dims[0], s.dims[1]))); break;
      case RKB_SHAPE_SPHERE: sh = shared_ptr<geom::shape_3D>(new geom::sphere("s", anchor, pose, s.dims[0])); break;
      case RKB_SHAPE_CCYLINDER: sh = shared_ptr<geom::shape_3D>(new geom::capped_cylinder("cc", anchor, pose, s.dims[0], s.dims[1])); break;
      case RKB_SHAPE_CYLINDER: sh = shared_ptr<geom::shape_3D>(new geom::cylinder("cy", anchor, pose, s.dims[0], s.dims[1])); break;
      case RKB_SHAPE_BOX: sh = shared_ptr<geom::shape_3D>(new geom::box("b", anchor, pose, vect<double,3>(s.dims[0], s.dims[1], s.dims[2]))); break;
      default: return -1;
    }
    shapes.push_back(sh);
    mdl[k < n1 ? 0 : 1]->addShape(sh);
  }
  geom::proxy_query_pair_3D pair("pair", mdl[0], mdl[1]);
  // finder order, restated only to turn the returned finder into an index: one finder per pair the
  // reference creates one for (proxy_query_model.cpp:212-384)
  std::vector<std::pair<int, int> > order;
  for (int a = 0; a < n1; ++a)
    for (int b = 0; b < n2; ++b) {
      const int ka = m1[a].kind, kb = m2[b].kind, lo = std::min(ka, kb), hi = std::max(ka, kb);
      const bool has = lo == RKB_SHAPE_PLANE || lo == RKB_SHAPE_SPHERE ||
                       (lo == RKB_SHAPE_CCYLINDER && (hi == RKB_SHAPE_CCYLINDER || hi == RKB_SHAPE_BOX));
      if (has) order.push_back(std::make_pair(a, b));
    }
  vect_n<double> p(nx), uu(nu);
  for (int k = 0; k < nu; ++k) uu[k] = 0.0;
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    m->sys.apply_states_and_inputs(p, uu);
    m->chain->doMotion();
    shared_ptr<geom::proximity_finder_3D> f = pair.findMinimumDistance();
    if (!f) {
      dist[i] = std::numeric_limits<double>::infinity();
      if (finder) finder[i] = -1;
      if (points) for (int k = 0; k < 6; ++k) points[6 * i + k] = 0.0;
      continue;
    }
    const geom::proximity_record_3D r = f->getLastResult();
    dist[i] = r.mDistance;
    if (points) for (int k = 0; k < 3; ++k) { points[6 * i + k] = r.mPoint1[k]; points[6 * i + 3 + k] = r.mPoint2[k]; }
    if (finder) {
      finder[i] = -2;
      const geom::shape_3D* s1 = f->getShape1().get();
      const geom::shape_3D* s2 = f->getShape2().get();
      for (std::size_t q = 0; q < order.size(); ++q) {
        const geom::shape_3D* a = shapes[order[q].first].get();
        const geom::shape_3D* b = shapes[n1 + order[q].second].get();
        if ((a == s1 && b == s2) || (a == s2 && b == s1)) { finder[i] = (int32_t)q; break; }
      }
    }
  }
  return (int)order.size();
}

// proxy_query_pair_3D::gatherCollisionPoints of the live reference (proxy_query_model.cpp:402-421) after doMotion at
// each state: count[i] records, the first max_records of them in records[i][r][7] (mDistance, mPoint1, mPoint2) with the
// finder index (createProxFinderList order, recovered through the record's position in a full scan) in finder[i][r].
int rkref_collision_points(void* hv, std::size_t N, const double* x, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2,
                           int max_records, int32_t* count, int32_t* finder, double* records) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  if (h->desc.dim == 2) return collision_points_2d(m, N, x, m1, n1, m2, n2, max_records, count, finder, records);
  if (h->desc.dim != 3) return -1;
  const int nx = m->nx, nu = m->nu;
  std::vector<shared_ptr<geom::shape_3D> > shapes;
  shared_ptr<geom::proxy_query_model_3D> mdl[2];
  mdl[0] = shared_ptr<geom::proxy_query_model_3D>(new geom::proxy_query_model_3D("model1"));
  mdl[1] = shared_ptr<geom::proxy_query_model_3D>(new geom::proxy_query_model_3D("model2"));
  for (int k = 0; k < n1 + n2; ++k) {
    const rkb_shape& s = k < n1 ? m1[k] : m2[k - n1];
    shared_ptr<pose_3D<double> > anchor;
    if (s.anchor >= 0) anchor = m->f3[s.anchor];
    const pose_3D<double> pose(weak_ptr<pose_3D<double> >(), vect<double,3>(s.position[0], s.position[1], s.position[2]),
                               quaternion<double>(vect<double,4>(s.quat[0], s.quat[1], s.quat[2], s.quat[3])));
    shared_ptr<geom::shape_3D> sh;
    switch (s.kind) {
      case RKB_SHAPE_PLANE: sh = shared_ptr<geom::shape_3D>(new geom::plane("p", anchor, pose, vect<double,2>(s.dims[0], s.dims[1]))); break;
      case RKB_SHAPE_SPHERE: sh = shared_ptr<geom::shape_3D>(new geom::sphere("s", anchor, pose, s.dims[0])); break;
      case RKB_SHAPE_CCYLINDER: sh = shared_ptr<geom::shape_3D>(new geom::capped_cylinder("cc", anchor, pose, s.dims[0], s.dims[1])); break;
      case RKB_SHAPE_CYLINDER: sh = shared_ptr<geom::shape_3D>(new geom::cylinder("cy", anchor, pose, s.dims[0], s.dims[1])); break;
      case RKB_SHAPE_BOX: sh = shared_ptr<geom::shape_3D>(new geom::box("b", anchor, pose, vect<double,3>(s.dims[0], s.dims[1], s.dims[2]))); break;
      default: return -1;
    }
    shapes.push_back(sh);
    mdl[k < n1 ? 0 : 1]->addShape(sh);
  }
  geom::proxy_query_pair_3D pair("pair", mdl[0], mdl[1]);
  vect_n<double> p(nx), uu(nu);
  for (int k = 0; k < nu; ++k) uu[k] = 0.0;
  for (std::size_t i = 0; i < N; ++i) {
    for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
    m->sys.apply_states_and_inputs(p, uu);
    m->chain->doMotion();
    std::vector<geom::proximity_record_3D> out;
    pair.gatherCollisionPoints(out);
    count[i] = int32_t(out.size());
    for (int r = 0; r < max_records; ++r) {
      double* o = records + (i * max_records + r) * 7;
      if (r < int(out.size())) {
        o[0] = out[r].mDistance;
        for (int k = 0; k < 3; ++k) { o[1 + k] = out[r].mPoint1[k]; o[4 + k] = out[r].mPoint2[k]; }
      } else {
        o[0] = std::numeric_limits<double>::infinity();
        for (int k = 1; k < 7; ++k) o[k] = 0.0;
      }
      if (finder) finder[i * max_records + r] = -1;  // (the records do not say which finder made them)
    }
  }
  return 0;
}

// reak_bridge.hpp's compile_proxy_model on live geom:: shapes riding on this model's frames: builds the shapes
// from `in` (as rkref_min_distance does), compiles the system and the model through the bridge and hands the
// shape list it derives back, so that a test can check rkb_shape -> ReaK shapes -> rkb_shape is the identity
// (anchors included: the bridge numbers frames its own way, the test compares through the frame order).
int rkref_bridge_proxy(void* hv, const rkb_shape* in, int n, rkb_shape* out, int* anchor_frame_of_desc, char* err, int err_len) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  if (h->desc.dim == 2) {  // planar models: the 2D overload of compile_proxy_model
    try {
      planar_pair P;
      if (!build_planar_pair(m, in, n, in, 0, P)) throw std::runtime_error("not a planar shape");
      reak_b200::chain_builder b = reak_b200::compile_kte_system(m->sys);
      std::vector<rkb_shape> shapes = reak_b200::compile_proxy_model(*P.mdl[0], b);
      if ((int)shapes.size() != n) throw std::runtime_error("shape count differs");
      for (int k = 0; k < n; ++k) {
        out[k] = shapes[k];
        anchor_frame_of_desc[k] = -1;
        if (shapes[k].anchor >= 0)
          for (std::size_t f = 0; f < m->f2.size(); ++f) {
            std::map<const void*, int>::const_iterator it = b.frame_ids.find(static_cast<const void*>(m->f2[f].get()));
            if (it != b.frame_ids.end() && it->second == shapes[k].anchor) anchor_frame_of_desc[k] = (int)f;
          }
      }
      return n;
    } catch (std::exception& e) {
      if (err && err_len > 0) { std::strncpy(err, e.what(), err_len - 1); err[err_len - 1] = 0; }
      return -1;
    }
  }
  try {
    geom::proxy_query_model_3D mdl("m");
    for (int k = 0; k < n; ++k) {
      const rkb_shape& s = in[k];
      shared_ptr<pose_3D<double> > anchor;
      if (s.anchor >= 0) anchor = m->f3[s.anchor];
      const pose_3D<double> pose(weak_ptr<pose_3D<double> >(), vect<double,3>(s.position[0], s.position[1], s.position[2]),
                                 quaternion<double>(vect<double,4>(s.quat[0], s.quat[1], s.quat[2], s.quat[3])));
      switch (s.kind) {
        case RKB_SHAPE_PLANE: mdl.addShape(shared_ptr<geom::shape_3D>(new geom::plane("p", anchor, pose, vect<double,2>(s.dims[0], s.dims[1])))); break;
        case RKB_SHAPE_SPHERE: mdl.addShape(shared_ptr<geom::shape_3D>(new geom::sphere("s", anchor, pose, s.dims[0]))); break;
        case RKB_SHAPE_CCYLINDER: mdl.addShape(shared_ptr<geom::shape_3D>(new geom::capped_cylinder("cc", anchor, pose, s.dims[0], s.dims[1]))); break;
        case RKB_SHAPE_CYLINDER: mdl.addShape(shared_ptr<geom::shape_3D>(new geom::cylinder("cy", anchor, pose, s.dims[0], s.dims[1]))); break;
        default: mdl.addShape(shared_ptr<geom::shape_3D>(new geom::box("b", anchor, pose, vect<double,3>(s.dims[0], s.dims[1], s.dims[2])))); break;
      }
    }
    reak_b200::chain_builder b = reak_b200::compile_kte_system(m->sys);
    std::vector<rkb_shape> shapes = reak_b200::compile_proxy_model(mdl, b);
    if ((int)shapes.size() != n) throw std::runtime_error("shape count differs");
    for (int k = 0; k < n; ++k) {
      out[k] = shapes[k];
      // translate the bridge's frame id back to this descriptor's numbering through the frame object
      anchor_frame_of_desc[k] = -1;
      if (shapes[k].anchor >= 0)
        for (std::size_t f = 0; f < m->f3.size(); ++f) {
          std::map<const void*, int>::const_iterator it = b.frame_ids.find(static_cast<const void*>(m->f3[f].get()));
          if (it != b.frame_ids.end() && it->second == shapes[k].anchor) anchor_frame_of_desc[k] = (int)f;
        }
    }
    return n;
  } catch (std::exception& e) {
    if (err && err_len > 0) { std::strncpy(err, e.what(), err_len - 1); err[err_len - 1] = 0; }
    return -1;
  }
}

// Runs include/reak_b200/reak_bridge.hpp on the LIVE ReaK objects of this model (built by
// build_model above from the descriptor) and hands the descriptor it derives back, so that a test
// can check that descriptor -> ReaK objects -> descriptor is the identity.  Returns the number of
// elements, or -1 (and the message in `err`) when the bridge rejects the chain.
int rkref_bridge_desc(void* hv, rkb_chain_desc* out, rkb_element* elems, int max_elems, char* err, int err_len) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  try {
    reak_b200::chain_builder b = reak_b200::compile_kte_system(h->proto->sys);
    rkb_chain_desc d = b.desc();
    if (d.n_elements > max_elems) return -2;
    for (int i = 0; i < d.n_elements; ++i) elems[i] = b.element(i);
    *out = d;
    out->elements = elems;
    return d.n_elements;
  } catch (std::exception& e) {
    if (err && err_len > 0) { std::strncpy(err, e.what(), err_len - 1); err[err_len - 1] = 0; }
    return -1;
  }
}

// f4: a `.rkx` file read back by the reference's own xml_iarchive into a live kte_nl_system, then flattened by the
// bridge: what a ReaK process that loads the model file and hands it to the batched path ends up with.  The product's
// own reader (rkb_rkx_read, no ReaK code) must produce the same descriptor.  Returns the element count or -1 (err).
int rkref_load_rkx_desc(const char* path, rkb_chain_desc* out, rkb_element* elems, int max_elems, char* err, int err_len) {
  try {
    // the archive creates objects through the type repository: make sure the classes of this path are registered
    {
      ctrl::kte_nl_system s0; kte::kte_map_chain c0; kte::mass_matrix_calc m0;
      kte::revolute_joint_3D a1; kte::revolute_joint_2D a2; kte::prismatic_joint_3D a3; kte::prismatic_joint_2D a4; kte::free_joint_3D a5; kte::free_joint_2D a6;
      kte::rigid_link_3D b1; kte::rigid_link_2D b2; kte::rigid_link_gen b3; kte::inertia_3D c1; kte::inertia_2D c2; kte::inertia_gen c3;
      kte::driving_actuator_gen d1; kte::torsion_spring_3D e1; kte::torsion_spring_2D e2; kte::torsion_damper_3D e3; kte::torsion_damper_2D e4;
      kte::spring_3D f1; kte::spring_2D f2; kte::spring_gen f3; kte::damper_3D g1; kte::damper_2D g2; kte::damper_gen g3;
      kte::joint_dependent_frame_3D h1; kte::joint_dependent_frame_2D h2; kte::joint_dependent_gen_coord h3;
      shared_ptr<rtti::so_type> keep[] = {s0.getObjectType(), c0.getObjectType(), m0.getObjectType(), a1.getObjectType(), a2.getObjectType(),
        a3.getObjectType(), a4.getObjectType(), a5.getObjectType(), a6.getObjectType(), jacobian_2D_2D<double>().getObjectType(), b1.getObjectType(), b2.getObjectType(), b3.getObjectType(), c1.getObjectType(),
        c2.getObjectType(), c3.getObjectType(), d1.getObjectType(), e1.getObjectType(), e2.getObjectType(), e3.getObjectType(), e4.getObjectType(),
        f1.getObjectType(), f2.getObjectType(), f3.getObjectType(), g1.getObjectType(), g2.getObjectType(), g3.getObjectType(),
        h1.getObjectType(), h2.getObjectType(), h3.getObjectType(), gen_coord<double>().getObjectType(), frame_3D<double>().getObjectType(),
        frame_2D<double>().getObjectType(), jacobian_gen_3D<double>().getObjectType(), jacobian_gen_2D<double>().getObjectType(),
        jacobian_gen_gen<double>().getObjectType(), jacobian_3D_3D<double>().getObjectType()};
      (void)keep;
    }
    shared_ptr<ctrl::kte_nl_system> sys;
    {
      serialization::xml_iarchive in(path);
      in >> sys;
    }
    if (!sys) throw std::runtime_error("the archive did not hold a kte_nl_system");
    reak_b200::chain_builder b = reak_b200::compile_kte_system(*sys);
    rkb_chain_desc d = b.desc();
    if (d.n_elements > max_elems) return -2;
    for (int i = 0; i < d.n_elements; ++i) elems[i] = b.element(i);
    *out = d;
    out->elements = elems;
    return d.n_elements;
  } catch (std::exception& e) {
    if (err && err_len > 0) { std::strncpy(err, e.what(), err_len - 1); err[err_len - 1] = 0; }
    return -1;
  }
}

// GPU drop-in check, all in C++ with ReaK's own types: a ReaK::ctrl::kte_batch_system is built
// from the live kte_nl_system through the bridge and compared, sample by sample, with that
// kte_nl_system (state derivative) and with runge_kutta4_integrator<double> driven by it (one
// step through get_next_state, n_steps through the batched call).  err[0..2] receive the max
// relative errors |a-b| / max(1,|b|).  Needs libreak_b200.so loaded first (RTLD_GLOBAL) and a GPU.
// Returns 0, or -1 with the exception text in `msg`.
int rkref_bridge_gpu_check(void* hv, std::size_t N, const double* x, const double* u, double dt, int n_steps,
                           double* err, char* msg, int msg_len) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  try {
    if (!rkb_chain_create) throw std::runtime_error("libreak_b200.so is not loaded (load it with RTLD_GLOBAL first)");
    ctrl::kte_batch_system bs(m->sys, 0, dt);
    err[0] = err[1] = err[2] = 0.0;
    std::vector<double> ref_n(N * nx), got_n(N * nx);
    std::vector<int32_t> st(N);
    rk4_range(m, 0, N, x, u, dt, n_steps, ref_n.data(), NULL);
    bs.batch().get_next_states(N, x, u, n_steps, dt, got_n.data(), st.data());
    for (std::size_t i = 0; i < N * nx; ++i)
      err[2] = std::max(err[2], std::fabs(got_n[i] - ref_n[i]) / std::max(1.0, std::fabs(ref_n[i])));
    std::vector<double> ref_1(nx);
    for (std::size_t i = 0; i < N && i < 16; ++i) {
      vect_n<double> p(nx), uu(nu);
      for (int k = 0; k < nx; ++k) p[k] = x[i * nx + k];
      for (int k = 0; k < nu; ++k) uu[k] = u[i * nu + k];
      vect_n<double> a = bs.get_state_derivative(m->sys, p, uu, 0.0);
      vect_n<double> b = m->sys.get_state_derivative(m->sys, p, uu, 0.0);
      for (int k = 0; k < nx; ++k) err[0] = std::max(err[0], std::fabs(a[k] - b[k]) / std::max(1.0, std::fabs(b[k])));
      vect_n<double> nxt = bs.get_next_state(m->sys, p, uu, 0.0);
      rk4_range(m, i, i + 1, x, u, dt, 1, ref_1.data() - i * nx, NULL);
      for (int k = 0; k < nx; ++k) err[1] = std::max(err[1], std::fabs(nxt[k] - ref_1[k]) / std::max(1.0, std::fabs(ref_1[k])));
    }
    // dimension errors surface as the reference's std::range_error
    bool threw = false;
    try { bs.get_state_derivative(m->sys, vect_n<double>(nx + 1), vect_n<double>(nu), 0.0); } catch (std::range_error&) { threw = true; }
    if (!threw) throw std::runtime_error("size mismatch did not raise std::range_error");
    return 0;
  } catch (std::exception& e) {
    if (msg && msg_len > 0) { std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
    return -1;
  }
}

// mass_matrix_calc::get_TMT_TdMT of the live model at state x (one sample): Tcm, Tcm_dot rows x n row-major,
// Mcm rows x rows.  Pass NULL pointers to query the row count.
int rkref_tmt(void* hv, const double* x, double* Tcm, double* Mcm, double* Tcm_dot) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  vect_n<double> p(nx), u(nu, 0.0);
  for (int k = 0; k < nx; ++k) p[k] = x[k];
  m->sys.apply_states_and_inputs(p, u);
  m->chain->doMotion();
  mat<double, mat_structure::rectangular> T, Td;
  mat<double, mat_structure::symmetric> Mc;
  m->mcalc->get_TMT_TdMT(T, Mc, Td);
  const int rows = int(T.get_row_count()), n = int(T.get_col_count());
  if (!Tcm || !Mcm || !Tcm_dot) return rows;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < n; ++c) { Tcm[r * n + c] = T(r, c); Tcm_dot[r * n + c] = Td(r, c); }
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < rows; ++c) Mcm[r * rows + c] = Mc(r, c);
  return rows;
}

// GPU drop-in check of ReaK::pp::kte_steer_space (reak_bridge.hpp): P pairs steered in one batched call
// and again one by one; every claim is re-derived with the reference on the CPU — all n_controls
// candidates of every pair are integrated with runge_kutta4_integrator, the winner must be the arg-min
// of the distance to the target, the returned point its end state, the steer record its state after
// every control interval.  err[0] point, err[1] record, err[2] single-vs-batched.  Returns 0 / -1 (msg).
int rkref_steer_space_check(void* hv, std::size_t P, const double* a, const double* b, double fraction, const double* u_lo,
                            const double* u_hi, int n_controls, int n_intervals, int steps, double dt, double* err, char* msg,
                            int msg_len) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  try {
    if (!rkb_chain_create) throw std::runtime_error("libreak_b200.so is not loaded (load it with RTLD_GLOBAL first)");
    vect_n<double> lo(nu), hi(nu);
    for (int k = 0; k < nu; ++k) { lo[k] = u_lo[k]; hi[k] = u_hi[k]; }
    pp::kte_steer_space space(m->sys, lo, hi, n_controls, n_intervals, steps, dt, 1234ull);
    std::vector<vect_n<double> > A(P, vect_n<double>(nx)), B(P, vect_n<double>(nx)), res;
    for (std::size_t i = 0; i < P; ++i)
      for (int k = 0; k < nx; ++k) { A[i][k] = a[i * nx + k]; B[i][k] = b[i * nx + k]; }
    std::vector<pp::kte_steer_space::steer_record_type> rec;
    space.steer_positions_toward(A, fraction, B, res, &rec);
    const std::vector<double> U = space.last_controls();
    const std::vector<int32_t> best = space.last_choice();
    err[0] = err[1] = err[2] = 0.0;
    std::vector<double> x0(n_controls * nx), xe(n_controls * nx), xw(nx);
    for (std::size_t i = 0; i < P; ++i) {
      for (int r = 0; r < n_controls; ++r)
        for (int k = 0; k < nx; ++k) x0[r * nx + k] = A[i][k];
      rk4_range(m, 0, n_controls, x0.data(), U.data() + i * n_controls * nu, dt, n_intervals * steps, xe.data(), NULL);
      int arg = 0;
      double dmin = 1e300;
      for (int r = 0; r < n_controls; ++r) {
        double d2 = 0.0;
        for (int k = 0; k < nx; ++k) {
          const double g = A[i][k] + fraction * (B[i][k] - A[i][k]);
          d2 += (xe[r * nx + k] - g) * (xe[r * nx + k] - g);
        }
        if (std::sqrt(d2) < dmin) { dmin = std::sqrt(d2); arg = r; }
      }
      if (arg != best[i]) throw std::runtime_error("the chosen control is not the arg-min of the reference rollouts");
      for (int k = 0; k < nx; ++k)
        err[0] = std::max(err[0], std::fabs(res[i][k] - xe[arg * nx + k]) / std::max(1.0, std::fabs(xe[arg * nx + k])));
      if ((int)rec[i].size() != n_intervals + 1) throw std::runtime_error("steer record has the wrong length");
      for (int k = 0; k < nx; ++k) if (rec[i][0][k] != A[i][k]) throw std::runtime_error("steer record does not start at the start point");
      for (int j = 1; j <= n_intervals; ++j) {
        rk4_range(m, 0, 1, &A[i][0], U.data() + (i * n_controls + arg) * nu, dt, j * steps, xw.data(), NULL);
        for (int k = 0; k < nx; ++k) err[1] = std::max(err[1], std::fabs(rec[i][j][k] - xw[k]) / std::max(1.0, std::fabs(xw[k])));
      }
    }
    // the concept's single-pair expression draws the same controls after the same reseed
    space.reseed(1234ull);
    std::pair<vect_n<double>, pp::kte_steer_space::steer_record_type> one = space.steer_position_toward(A[0], fraction, B[0]);
    for (int k = 0; k < nx; ++k) err[2] = std::max(err[2], std::fabs(one.first[k] - res[0][k]));
    if (one.second.size() != rec[0].size()) throw std::runtime_error("single and batched steer records differ in length");
    if (std::fabs(space.distance(A[0], B[0]) - norm_2(A[0] - B[0])) > 0.0) throw std::runtime_error("distance is not the Euclidean one");
    return 0;
  } catch (std::exception& e) {
    if (msg && msg_len > 0) { std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
    return -1;
  }
}

}  // extern "C" (reopened below)

namespace {

// The smallest graph the reference's node generators can walk: vertices carry `position`, edges what
// mg_edge_data / optimal_mg_edge carry (ctrl/path_planning/any_motion_graphs.hpp): a steer record and a weight.
struct tiny_graph {
  typedef std::size_t vertex_descriptor;
  typedef std::pair<std::size_t, std::size_t> edge_descriptor;
  static const bool is_directed = false;
  static vertex_descriptor null_vertex() { return std::size_t(-1); }
  struct vertex_bundled { vect_n<double> position; };
  struct edge_bundled { std::vector<vect_n<double> > steer_record; double weight; edge_bundled() : weight(0.0) {} };
  std::vector<vertex_bundled> v;
  vertex_bundled& operator[](vertex_descriptor u) { return v[u]; }
  const vertex_bundled& operator[](vertex_descriptor u) const { return v[u]; }
};

// planning_visitor's steering, one candidate at a time, written with the reference's own pieces: the
// is_steerable_space dispatch of planning_visitors.hpp:251-296 (both overloads, so that the trait decides), the
// boost::tie / get(distance_metric, space.get_super_space()) expressions of :258-260 and the acceptance test of
// :349-360 plus is_position_free (:242).  (planning_visitors.hpp itself needs Boost.Graph, Boost.Any, Boost.Range
// and Boost.Random through motion_planner_base.hpp / any_motion_graphs.hpp and cannot be compiled here.)
template <typename Space>
struct one_at_a_time_visitor {
  typedef typename pp::topology_traits<Space>::point_type point_type;
  const Space* space;
  double tol;
  int* branch;  // [0] calls that took the steerable branch, [1] the move_position_toward branch (the node puller copies its visitor)
  one_at_a_time_visitor(const Space& s, double t, int* counters) : space(&s), tol(t), branch(counters) {}

  template <typename SwitchFreeSpace>
  typename boost::enable_if<pp::is_steerable_space<SwitchFreeSpace>, double>::type dispatched_steer_towards_position(
      const SwitchFreeSpace& sp, const point_type& p_src, const point_type& p_dest, point_type& p_result, double fraction,
      tiny_graph::edge_bundled& ep_result) const {
    ++branch[0];
    boost::tie(p_result, ep_result.steer_record) = sp.steer_position_toward(p_src, fraction, p_dest);
    ep_result.weight = get(pp::distance_metric, sp.get_super_space())(p_src, p_result, sp.get_super_space());
    return ep_result.weight;
  }
  template <typename SwitchFreeSpace>
  typename boost::disable_if<pp::is_steerable_space<SwitchFreeSpace>, double>::type dispatched_steer_towards_position(
      const SwitchFreeSpace& sp, const point_type& p_src, const point_type& p_dest, point_type& p_result, double fraction,
      tiny_graph::edge_bundled& ep_result) const {
    ++branch[1];
    p_result = sp.move_position_toward(p_src, fraction, p_dest);
    ep_result.weight = get(pp::distance_metric, sp.get_super_space())(p_src, p_result, sp.get_super_space());
    return ep_result.weight;
  }
  boost::tuple<point_type, bool, tiny_graph::edge_bundled> steer_towards_position(const point_type& p, std::size_t u, tiny_graph& g) const {
    boost::tuple<point_type, bool, tiny_graph::edge_bundled> result;
    const double traveled = dispatched_steer_towards_position(*space, g[u].position, p, boost::get<0>(result), 1.0, boost::get<2>(result));
    const double best_case = get(pp::distance_metric, space->get_super_space())(g[u].position, p, space->get_super_space());
    boost::get<1>(result) = (!std::isinf(traveled)) && (traveled < 2.0 * best_case) && (traveled > tol * best_case) &&
                            space->is_free(boost::get<0>(result));
    return result;
  }
};

shared_ptr<geom::proxy_query_model_3D> shapes_to_model(ref_model* m, const rkb_shape* in, int n, const char* name) {
  shared_ptr<geom::proxy_query_model_3D> mdl(new geom::proxy_query_model_3D(name));
  for (int k = 0; k < n; ++k) {
    const rkb_shape& s = in[k];
    shared_ptr<pose_3D<double> > anchor;
    if (s.anchor >= 0) anchor = m->f3[s.anchor];
    const pose_3D<double> pose(weak_ptr<pose_3D<double> >(), vect<double,3>(s.position[0], s.position[1], s.position[2]),
                               quaternion<double>(vect<double,4>(s.quat[0], s.quat[1], s.quat[2], s.quat[3])));
    shared_ptr<geom::shape_3D> sh;
    switch (s.kind) {
      case RKB_SHAPE_PLANE: sh = shared_ptr<geom::shape_3D>(new geom::plane("p", anchor, pose, vect<double,2>(s.dims[0], s.dims[1]))); break;
      case RKB_SHAPE_SPHERE: sh = shared_ptr<geom::shape_3D>(new geom::sphere("s", anchor, pose, s.dims[0])); break;
      case RKB_SHAPE_CCYLINDER: sh = shared_ptr<geom::shape_3D>(new geom::capped_cylinder("cc", anchor, pose, s.dims[0], s.dims[1])); break;
      case RKB_SHAPE_CYLINDER: sh = shared_ptr<geom::shape_3D>(new geom::cylinder("cy", anchor, pose, s.dims[0], s.dims[1])); break;
      case RKB_SHAPE_BOX: sh = shared_ptr<geom::shape_3D>(new geom::box("b", anchor, pose, vect<double,3>(s.dims[0], s.dims[1], s.dims[2]))); break;
      default: throw std::runtime_error("unknown shape kind");
    }
    mdl->addShape(sh);
  }
  return mdl;
}

}  // namespace

extern "C" {

// SURVEY 8(f) rank 1, planner side.  ReaK::pp::kte_steer_space (reak_bridge.hpp) under the reference's own planner
// plumbing, as far as it compiles without Boost.Graph:
//   * the traits the planners dispatch on (is_steerable_space / is_metric_space / is_point_distribution) and the
//     tagged get() overloads they enable (metric_space_concept.hpp:288-293, default_random_sampler.hpp:82-88);
//   * the valid expression of SteerableSpaceConcept (steerable_space_concept.hpp:60-85);
//   * ReaK::graph::detail::rrg_node_puller<Graph>::expand_to_nearest — the REAL one, ctrl/graph_alg/node_generators.hpp:59-75 —
//     run twice over the same K candidate vertices: with a visitor that steers one candidate at a time (the
//     reference's way) and with batched_steer_visitor, which steered all K in one GPU call beforehand.  Same seed,
//     so both must pull the same vertex, the same point, the same steer record and the same edge weight.
// x_nodes: K vertex positions, target: the sample, shapes (n1 + n2 > 0): a collision environment for is_free.
// out: [0] vertex pulled one-at-a-time, [1] vertex pulled batched, [2] steer-branch count, [3] move-branch count,
//      [4] candidates the batched visitor prepared;  free_flags: is_free of the K nodes (space.is_free, one by one)
//      and of the same K nodes through are_free (K more);  p_new: the pulled point (batched run);  sample: one random_point().
int rkref_planner_dispatch_check(void* hv, int K, const double* x_nodes, const double* target, const double* u_lo, const double* u_hi,
                                 const double* x_lo, const double* x_hi, int n_controls, int n_intervals, int steps, double dt,
                                 double tol, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2,
                                 int* out, int* free_flags, double* p_new, double* sample, double* err, char* msg, int msg_len) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  try {
    if (!rkb_chain_create) throw std::runtime_error("libreak_b200.so is not loaded (load it with RTLD_GLOBAL first)");
    typedef pp::kte_steer_space space_t;
    static_assert(pp::is_steerable_space<space_t>::value, "kte_steer_space must be steerable for the planners");
    static_assert(pp::is_metric_space<space_t>::value && pp::is_point_distribution<space_t>::value, "metric / sampling traits");
    static_assert(!pp::is_steerable_space<tiny_graph>::value, "the primary template stays false");
    vect_n<double> lo(nu), hi(nu), xl(nx), xh(nx), tgt(nx);
    for (int k = 0; k < nu; ++k) { lo[k] = u_lo[k]; hi[k] = u_hi[k]; }
    for (int k = 0; k < nx; ++k) { xl[k] = x_lo[k]; xh[k] = x_hi[k]; tgt[k] = target[k]; }
    tiny_graph g;
    g.v.resize(K);
    std::vector<std::size_t> Nc;
    for (int i = 0; i < K; ++i) {
      g.v[i].position = vect_n<double>(nx);
      for (int k = 0; k < nx; ++k) g.v[i].position[k] = x_nodes[i * nx + k];
      Nc.push_back(i);
    }
    shared_ptr<geom::proxy_query_model_3D> mdl1, mdl2;
    if (n1 + n2 > 0) { mdl1 = shapes_to_model(m, m1, n1, "model1"); mdl2 = shapes_to_model(m, m2, n2, "model2"); }
    // ---- run 1: one candidate at a time -----------------------------------------------------------------
    space_t sp1(m->sys, lo, hi, n_controls, n_intervals, steps, dt, 4321ull);
    sp1.set_state_bounds(xl, xh);
    if (mdl1) sp1.add_proxy_pair(*mdl1, *mdl2);
    {  // SteerableSpaceConcept's valid expression, verbatim
      vect_n<double> p1 = g.v[0].position, p2 = tgt;
      pp::steerable_space_traits<space_t>::steer_record_type st_rec;
      double d = 0.5;
      boost::tie(p1, st_rec) = sp1.steer_position_toward(p1, d, p2);
      if ((int)st_rec.size() != n_intervals + 1) throw std::runtime_error("steer record has the wrong length");
      sp1.reseed(4321ull);
    }
    int branches[2] = {0, 0};
    one_at_a_time_visitor<space_t> vis1(sp1, tol, branches);
    vect_n<double> p_a = tgt;
    boost::tuple<std::size_t, bool, tiny_graph::edge_bundled> r1 =
        graph::detail::rrg_node_puller<tiny_graph>::expand_to_nearest(p_a, Nc, g, vis1);
    // ---- run 2: all candidates in one batched call, then the same node puller -----------------------------
    space_t sp2(m->sys, lo, hi, n_controls, n_intervals, steps, dt, 4321ull);
    sp2.set_state_bounds(xl, xh);
    if (mdl1) sp2.add_proxy_pair(*mdl1, *mdl2);
    pp::batched_steer_visitor<tiny_graph> vis2(sp2, tol);
    vis2.prepare(tgt, Nc, g);
    vect_n<double> p_b = tgt;
    boost::tuple<std::size_t, bool, tiny_graph::edge_bundled> r2 =
        graph::detail::rrg_node_puller<tiny_graph>::expand_to_nearest(p_b, Nc, g, vis2);
    out[0] = boost::get<1>(r1) ? (int)boost::get<0>(r1) : -1;
    out[1] = boost::get<1>(r2) ? (int)boost::get<0>(r2) : -1;
    out[2] = branches[0]; out[3] = branches[1]; out[4] = (int)vis2.prepared();
    err[0] = err[1] = err[2] = 0.0;
    if (out[0] != out[1]) throw std::runtime_error("batched and one-at-a-time node pulling chose different vertices");
    for (int k = 0; k < nx; ++k) { err[0] = std::max(err[0], std::fabs(p_a[k] - p_b[k])); p_new[k] = p_b[k]; }
    if (boost::get<1>(r1)) {
      const tiny_graph::edge_bundled &e1 = boost::get<2>(r1), &e2 = boost::get<2>(r2);
      if (e1.steer_record.size() != e2.steer_record.size()) throw std::runtime_error("steer records differ in length");
      for (std::size_t j = 0; j < e1.steer_record.size(); ++j)
        for (int k = 0; k < nx; ++k) err[1] = std::max(err[1], std::fabs(e1.steer_record[j][k] - e2.steer_record[j][k]));
      err[2] = std::fabs(e1.weight - e2.weight);
      // the weight is the reference's default_distance_metric on the super-space, i.e. space.distance
      if (std::fabs(e2.weight - sp2.distance(g.v[out[1]].position, p_b)) > 0.0) throw std::runtime_error("edge weight is not the metric's distance");
    }
    // ---- is_free, one by one and batched; random_point through the reference's default_random_sampler --------
    std::vector<vect_n<double> > pts;
    for (int i = 0; i < K; ++i) { pts.push_back(g.v[i].position); free_flags[i] = sp2.is_free(g.v[i].position) ? 1 : 0; }
    std::vector<char> fr;
    sp2.are_free(pts, fr);
    for (int i = 0; i < K; ++i) free_flags[K + i] = fr[i];
    vect_n<double> rp = get(pp::random_sampler, sp2)(sp2);
    for (int k = 0; k < nx; ++k) {
      sample[k] = rp[k];
      if (!(rp[k] >= xl[k] && rp[k] <= xh[k])) throw std::runtime_error("random_point left the state box");
    }
    return 0;
  } catch (std::exception& e) {
    if (msg && msg_len > 0) { std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
    return -1;
  }
}

// n_workers > 1: the samples are block-partitioned over forked worker processes.  Threads do
// not scale here: every rk_dynamic_ptr_cast in the reference bumps the atomic reference count
// of shared static type descriptors, so threads serialise on those cache lines (measured:
// 8 threads = 1.04x, 8 processes = 3.7x on the 8-vCPU build container).  A forked child owns a
// copy-on-write image of the whole model, which is the "one model instance per worker" the
// non-re-entrant KTE objects need.  Results come back through a shared anonymous mapping.
// Returns wall seconds of the integration (< 0 on failure).
double rkref_integrate(void* hv, std::size_t N, const double* x0, const double* u, int scheme, double dt, int n_steps,
                       double* xout, int32_t* status, int n_workers);

// The steering loop with the reference's own dynamics and runge_kutta4_integrator inside; the feedback
// law is the restatement of oracle/steer_law.h, itself pinned against IHAQR_topology / MEAQR_topology's own members (ref_steer_law.cpp).
int rkref_steer_feedback(void* hv, std::size_t N, const double* x0, const double* goal, const double* u_bias, const double* gain,
                         double* u_prev, double T, double dt, int substeps, int max_intervals, double proximity, int saturate_first,
                         const double* lo, const double* hi, const double* dlo, const double* dhi,
                         double* x_out, int32_t* n_done, double* traj, int32_t* status) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  if (nu > STEER_MAX_INPUTS) return -1;
  for (std::size_t i = 0; i < N; ++i) {
    std::vector<double> x(x0 + i * nx, x0 + (i + 1) * nx), xn(nx), u(nu ? nu : 1), up(nu ? nu : 1);
    for (int j = 0; j < nu; ++j) up[j] = u_prev[i * nu + j];
    int k = 0, st = 0;
    while (k < max_intervals) {
      int32_t s1 = 0;
      if (!steer_next_input(nx, nu, T, proximity, (!saturate_first && k == 0), lo, hi, dlo, dhi, &x[0], goal + i * nx,
                            u_bias + i * nu, gain + i * std::size_t(nu) * nx, &up[0], &u[0]))
        break;
      rk4_range(m, 0, 1, &x[0], &u[0], dt, substeps, &xn[0], &s1);
      st |= s1;
      x = xn;
      for (int j = 0; j < nu; ++j) up[j] = u[j];
      if (traj) for (int j = 0; j < nx; ++j) traj[(i * std::size_t(max_intervals) + k) * nx + j] = x[j];
      ++k;
    }
    for (int j = 0; j < nx; ++j) x_out[i * nx + j] = x[j];
    for (int j = 0; j < nu; ++j) u_prev[i * nu + j] = up[j];
    if (n_done) n_done[i] = k;
    if (status) status[i] = st;
  }
  return 0;
}
// rkref_steer_feedback with with_collision_check = true: after every interval the live proxy pairs are queried at
// x_next (kte_map_chain::doMotion + proxy_query_pair_3D::findMinimumDistance, the body of is_free_impl,
// MEAQR_topology.hpp:921-940); a state that is not free is not accepted and ends the loop
// (MEAQR_topology.hpp:550-559).  Pair p is (m1s[p][0..n1s[p]), m2s[p][0..n2s[p])).
int rkref_steer_feedback_checked(void* hv, std::size_t N, const double* x0, const double* goal, const double* u_bias, const double* gain,
                                 double* u_prev, double T, double dt, int substeps, int max_intervals, double proximity, int saturate_first,
                                 const double* lo, const double* hi, const double* dlo, const double* dhi,
                                 const rkb_shape* const* m1s, const int* n1s, const rkb_shape* const* m2s, const int* n2s, int n_pairs,
                                 double* x_out, int32_t* n_done, int32_t* collided, double* traj, int32_t* status) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx, nu = m->nu;
  if (nu > STEER_MAX_INPUTS || h->desc.dim != 3) return -1;
  std::vector<shared_ptr<geom::proxy_query_pair_3D> > pairs;
  for (int p = 0; p < n_pairs; ++p) {
    shared_ptr<geom::proxy_query_model_3D> mdl[2];
    for (int w = 0; w < 2; ++w) {
      mdl[w] = shared_ptr<geom::proxy_query_model_3D>(new geom::proxy_query_model_3D("m"));
      const rkb_shape* sh = w == 0 ? m1s[p] : m2s[p];
      const int n = w == 0 ? n1s[p] : n2s[p];
      for (int k = 0; k < n; ++k) {
        const rkb_shape& s = sh[k];
        shared_ptr<pose_3D<double> > anchor;
        if (s.anchor >= 0) anchor = m->f3[s.anchor];
        const pose_3D<double> pose(weak_ptr<pose_3D<double> >(), vect<double,3>(s.position[0], s.position[1], s.position[2]),
                                   quaternion<double>(vect<double,4>(s.quat[0], s.quat[1], s.quat[2], s.quat[3])));
        switch (s.kind) {
          case RKB_SHAPE_PLANE: mdl[w]->addShape(shared_ptr<geom::shape_3D>(new geom::plane("p", anchor, pose, vect<double,2>(s.dims[0], s.dims[1])))); break;
          case RKB_SHAPE_SPHERE: mdl[w]->addShape(shared_ptr<geom::shape_3D>(new geom::sphere("s", anchor, pose, s.dims[0]))); break;
          case RKB_SHAPE_CCYLINDER: mdl[w]->addShape(shared_ptr<geom::shape_3D>(new geom::capped_cylinder("cc", anchor, pose, s.dims[0], s.dims[1]))); break;
          case RKB_SHAPE_CYLINDER: mdl[w]->addShape(shared_ptr<geom::shape_3D>(new geom::cylinder("cy", anchor, pose, s.dims[0], s.dims[1]))); break;
          case RKB_SHAPE_BOX: mdl[w]->addShape(shared_ptr<geom::shape_3D>(new geom::box("b", anchor, pose, vect<double,3>(s.dims[0], s.dims[1], s.dims[2])))); break;
          default: return -1;
        }
      }
    }
    pairs.push_back(shared_ptr<geom::proxy_query_pair_3D>(new geom::proxy_query_pair_3D("pair", mdl[0], mdl[1])));
  }
  vect_n<double> pv(nx), uv(nu);
  for (int j = 0; j < nu; ++j) uv[j] = 0.0;
  for (std::size_t i = 0; i < N; ++i) {
    std::vector<double> x(x0 + i * nx, x0 + (i + 1) * nx), xn(nx), u(nu ? nu : 1), up(nu ? nu : 1);
    for (int j = 0; j < nu; ++j) up[j] = u_prev[i * nu + j];
    int k = 0, st = 0, hit = 0;
    while (k < max_intervals) {
      int32_t s1 = 0;
      if (!steer_next_input(nx, nu, T, proximity, (!saturate_first && k == 0), lo, hi, dlo, dhi, &x[0], goal + i * nx,
                            u_bias + i * nu, gain + i * std::size_t(nu) * nx, &up[0], &u[0]))
        break;
      rk4_range(m, 0, 1, &x[0], &u[0], dt, substeps, &xn[0], &s1);
      st |= s1;
      // is_free_impl(x_next)
      for (int j = 0; j < nx; ++j) pv[j] = xn[j];
      m->sys.apply_states_and_inputs(pv, uv);
      m->chain->doMotion();
      bool is_free = true;
      for (std::size_t p = 0; p < pairs.size() && is_free; ++p) {
        shared_ptr<geom::proximity_finder_3D> f = pairs[p]->findMinimumDistance();
        if (f && f->getLastResult().mDistance < 0.0) is_free = false;
      }
      if (!is_free) { hit = 1; break; }
      x = xn;
      for (int j = 0; j < nu; ++j) up[j] = u[j];
      if (traj) for (int j = 0; j < nx; ++j) traj[(i * std::size_t(max_intervals) + k) * nx + j] = x[j];
      ++k;
    }
    for (int j = 0; j < nx; ++j) x_out[i * nx + j] = x[j];
    for (int j = 0; j < nu; ++j) u_prev[i * nu + j] = up[j];
    if (n_done) n_done[i] = k;
    if (collided) collided[i] = hit;
    if (status) status[i] = st;
  }
  return 0;
}
double rkref_rk4(void* hv, std::size_t N, const double* x0, const double* u, double dt, int n_steps,
                 double* xout, int32_t* status, int n_workers) {
  return rkref_integrate(hv, N, x0, u, RKB_SCHEME_RK4, dt, n_steps, xout, status, n_workers);
}

double rkref_integrate(void* hv, std::size_t N, const double* x0, const double* u, int scheme, double dt, int n_steps,
                       double* xout, int32_t* status, int n_workers) {
  ref_handle* h = static_cast<ref_handle*>(hv);
  ref_model* m = h->proto;
  const int nx = m->nx;
  if (n_workers < 1) n_workers = 1;
  if (std::size_t(n_workers) > N && N > 0) n_workers = int(N);
  const auto t0 = std::chrono::steady_clock::now();
  if (n_workers == 1 || N == 0) {
    rk4_range(m, 0, N, x0, u, dt, n_steps, xout, status, scheme);
  } else {
    const std::size_t bytes_x = N * nx * sizeof(double), bytes_s = N * sizeof(int32_t);
    void* shm = mmap(NULL, bytes_x + bytes_s, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shm == MAP_FAILED) return -1.0;
    double* sx = static_cast<double*>(shm);
    int32_t* ss = reinterpret_cast<int32_t*>(static_cast<char*>(shm) + bytes_x);
    std::vector<pid_t> pids;
    bool ok = true;
    for (int t = 0; t < n_workers; ++t) {
      std::size_t i0 = N * t / n_workers, i1 = N * (t + 1) / n_workers;
      pid_t pid = fork();
      if (pid == 0) {
        rk4_range(m, i0, i1, x0, u, dt, n_steps, sx, ss, scheme);
        _exit(0);
      }
      if (pid < 0) { ok = false; break; }
      pids.push_back(pid);
    }
    for (std::size_t i = 0; i < pids.size(); ++i) {
      int wst = 0;
      if (waitpid(pids[i], &wst, 0) < 0 || !WIFEXITED(wst) || WEXITSTATUS(wst) != 0) ok = false;
    }
    if (ok) {
      std::memcpy(xout, sx, bytes_x);
      if (status) std::memcpy(status, ss, bytes_s);
    }
    munmap(shm, bytes_x + bytes_s);
    if (!ok) return -1.0;
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
