// kte_batch_propagator.hpp — C++ host side of the B200 batched KTE-chain propagator.
//
// Header-only, C++11, no Boost, no CUDA headers: it only drives the C-ABI of reak_b200.h
// (link with -lreak_b200).  Two pieces:
//
//   reak_b200::chain_builder        assembles the flat rkb_chain_desc with the vocabulary of
//                                   ReaK::kte (frames, generalized coordinates, joints, links,
//                                   inertias, springs, dampers, actuators), in kte_map_chain order.
//   reak_b200::kte_batch_propagator the batched counterpart of ReaK::ctrl::kte_nl_system
//                                   (ctrl/ctrl_sys/kte_nl_system.hpp:68-423) wrapped in
//                                   num_int_dtnl_sys<..., runge_kutta4_integrator>
//                                   (ctrl/ctrl_sys/num_int_dtnl_system.hpp:166-180).
//
// The single-sample members keep the reference's concept signatures — SSSystemConcept
// (ctrl/ctrl_sys/state_space_sys_concept.hpp:111-137: get_state_derivative(space, p, u, t),
// get_output, get_state_dimensions, get_input_dimensions, get_output_dimensions) and
// DiscreteSSSConcept (ctrl/ctrl_sys/discrete_sss_concept.hpp:40-141: get_time_step,
// get_next_state(space, p, u, t)) — and raise what the reference raises: std::range_error on a
// size mismatch (kte_nl_system.hpp:181-187), singularity_error where linsolve_Cholesky throws it
// (core/lin_alg/mat_cholesky.hpp:80-82), impossible_integration for a zero step
// (core/integrators/fixed_step_integrators.hpp:258-266).  The batched members are additive.
// reak_bridge.hpp builds a chain_builder from live ReaK objects.
#ifndef REAK_B200_KTE_BATCH_PROPAGATOR_HPP
#define REAK_B200_KTE_BATCH_PROPAGATOR_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <map>
#include <vector>

#include "../reak_b200.h"

namespace reak_b200 {

/// Counterpart of ReaK::singularity_error (core/lin_alg/mat_num_exceptions.hpp:45-56).
class singularity_error : public std::runtime_error {
 public:
  explicit singularity_error(const std::string& what) : std::runtime_error("Singularity Error: " + what) {}
};
/// Counterpart of ReaK::impossible_integration (core/integrators/integration_exceptions.hpp:38).
class impossible_integration : public std::runtime_error {
 public:
  explicit impossible_integration(const std::string& what) : std::runtime_error(what) {}
};
/// Any other failure of the C-ABI (CUDA error, unsupported chain, ...).
class propagator_error : public std::runtime_error {
 public:
  int code;
  propagator_error(int aCode, const std::string& where)
      : std::runtime_error(where + ": " + rkb_strerror(aCode) + (aCode == RKB_ERR_CUDA ? std::string(" [") + rkb_last_cuda_error() + "]" : std::string())),
        code(aCode) {}
};

/// Builds an rkb_chain_desc.  Frames and coordinates are handed out as integer ids; elements are
/// appended in kte_map_chain order (the order `chain << element` would give).
class chain_builder {
 public:
  explicit chain_builder(int dim = 3) : mDim(dim), mFrames(0), mCoords(0), mInputs(0), mBaseFrame(-1), mAux(0) {
    mBase = rkb_base_frame();
    mBase.quat[0] = dim == 3 ? 1.0 : 0.0;
  }

  int add_frame() { return mFrames++; }
  /// frame object of the host model (e.g. a ReaK frame_3D) -> frame id; filled by reak_bridge.hpp so that
  /// proximity shapes can be anchored to the frames they ride on
  std::map<const void*, int> frame_ids;
  int add_coord() { return mCoords++; }  ///< state slot j <-> kte_nl_system::dofs_gen[j]

  /// The un-driven root frame (robot_base in CRS_A465_models.cpp:298-301); gravity is an upward Acceleration.
  void set_base(int frame, const rkb_base_frame& kinematics) { mBaseFrame = frame; mBase = kinematics; }
  rkb_base_frame& base() { return mBase; }

  int revolute_joint_3D(int coord, double ax, double ay, double az, int base, int end) {
    return push(RKB_REVOLUTE_3D, base, end, coord, 0, 0, vec(ax, ay, az));
  }
  /// free_joint_3D (free_joints.cpp:119-208): `free_index` = position of its coordinate frame in kte_nl_system::dofs_3D;
  /// adds 13 states and 6 accelerations after the generalized coordinates' (kte_nl_system.hpp:145-147)
  int free_joint_3D(int free_index, int base, int end) { return push(RKB_FREE_3D, base, end, free_index, 0, 0, std::vector<double>()); }
  /// free_joint_2D (free_joints.cpp:33-117): 7 states and 3 accelerations (kte_nl_system.hpp:194-204, 282-291)
  int free_joint_2D(int free_index, int base, int end) { return push(RKB_FREE_2D, base, end, free_index, 0, 0, std::vector<double>()); }
  int prismatic_joint_3D(int coord, double ax, double ay, double az, int base, int end) {
    return push(RKB_PRISMATIC_3D, base, end, coord, 0, 0, vec(ax, ay, az));
  }
  int revolute_joint_2D(int coord, int base, int end) { return push(RKB_REVOLUTE_2D, base, end, coord, 0, 0, std::vector<double>()); }
  int prismatic_joint_2D(int coord, double ax, double ay, int base, int end) {
    return push(RKB_PRISMATIC_2D, base, end, coord, 0, 0, vec(ax, ay));
  }
  /// pose offset: position + quaternion (w, x, y, z)
  int rigid_link_3D(int base, int end, const double p[3], const double q[4]) {
    std::vector<double> v(7);
    for (int i = 0; i < 3; ++i) v[i] = p[i];
    for (int i = 0; i < 4; ++i) v[3 + i] = q[i];
    return push(RKB_RIGID_LINK_3D, base, end, -1, 0, 0, v);
  }
  int rigid_link_2D(int base, int end, double px, double py, double angle) { return push(RKB_RIGID_LINK_2D, base, end, -1, 0, 0, vec(px, py, angle)); }
  /// tensor = (Ixx, Ixy, Ixz, Iyy, Iyz, Izz); upstream bit c <=> coordinate c is in mUpStreamJoints
  int inertia_3D(int frame, double mass, const double tensor[6], std::uint64_t upstream) {
    std::vector<double> v(7);
    v[0] = mass;
    for (int i = 0; i < 6; ++i) v[1 + i] = tensor[i];
    return push(RKB_INERTIA_3D, frame, -1, -1, 0, upstream, v);
  }
  int inertia_2D(int frame, double mass, double moment, std::uint64_t upstream) { return push(RKB_INERTIA_2D, frame, -1, -1, 0, upstream, vec(mass, moment)); }
  int inertia_gen(int coord, double mass) { return push(RKB_INERTIA_GEN, -1, -1, coord, 0, std::uint64_t(1) << coord, vec(mass)); }
  /// driving_actuator_gen on `coord`, reacting on joint element `joint_element`; returns the element index.
  /// `input_index` is its position in kte_nl_system::inputs.
  int driving_actuator_gen(int coord, int joint_element, int input_index) {
    if (input_index + 1 > mInputs) mInputs = input_index + 1;
    return push(RKB_ACTUATOR_GEN, -1, joint_element, coord, input_index, 0, std::vector<double>());
  }
  int torsion_spring(int a1, int a2, double stiffness, double saturation = 0.0) {
    return push(mDim == 3 ? RKB_TORSION_SPRING_3D : RKB_TORSION_SPRING_2D, a1, a2, -1, 0, 0, vec(stiffness, saturation));
  }
  int torsion_damper(int a1, int a2, double damping) {
    return push(mDim == 3 ? RKB_TORSION_DAMPER_3D : RKB_TORSION_DAMPER_2D, a1, a2, -1, 0, 0, vec(damping));
  }
  int spring(int a1, int a2, double rest_length, double stiffness, double saturation = 0.0) {
    return push(mDim == 3 ? RKB_SPRING_3D : RKB_SPRING_2D, a1, a2, -1, 0, 0, vec(rest_length, stiffness, saturation));
  }
  int damper(int a1, int a2, double damping) { return push(mDim == 3 ? RKB_DAMPER_3D : RKB_DAMPER_2D, a1, a2, -1, 0, 0, vec(damping)); }

  /// Elements on generalized coordinates alone (rigid_link.cpp:30-75, spring.cpp:32-96, damper.cpp:32-68).  Their anchors
  /// are system coordinates or auxiliary ones: add_aux_coord() declares a gen_coord that is not a system state (a fixed
  /// anchor, or the end of a rigid_link_gen) holding (q, q_dot, q_ddot), and returns its index.  Call it after every
  /// add_coord() and before the first element that uses it.
  int add_aux_coord(double q = 0.0, double q_dot = 0.0, double q_ddot = 0.0) {
    const int idx = mCoords + mAux++;
    push(RKB_COORD_GEN, -1, -1, idx, 0, 0, vec(q, q_dot, q_ddot));
    return idx;
  }
  int rigid_link_gen(int base_coord, int end_aux_coord, double offset) { return push(RKB_RIGID_LINK_GEN, -1, -1, base_coord, end_aux_coord, 0, vec(offset)); }
  int spring_gen(int c1, int c2, double rest_length, double stiffness, double saturation = 0.0) {
    return push(RKB_SPRING_GEN, -1, -1, c1, c2, 0, vec(rest_length, stiffness, saturation));
  }
  int damper_gen(int c1, int c2, double damping) { return push(RKB_DAMPER_GEN, -1, -1, c1, c2, 0, vec(damping)); }

  /// Actuators may be appended before the joint they drive (CRS order: actuator, rotor, joint, ...).
  void set_actuator_joint(int actuator_element, int joint_element) { mElements.at(actuator_element).frame_b = joint_element; }

  std::size_t element_count() const { return mElements.size(); }
  const rkb_element& element(std::size_t i) const { return mElements.at(i); }

  /// The descriptor points into this builder: keep the builder alive until rkb_chain_create returned.
  rkb_chain_desc desc() const {
    rkb_chain_desc d;
    d.dim = mDim;
    d.n_elements = static_cast<int32_t>(mElements.size());
    d.n_frames = mFrames;
    d.n_coords = mCoords;
    d.n_inputs = mInputs;
    d.base_frame = mBaseFrame;
    d.base = mBase;
    d.elements = mElements.empty() ? NULL : &mElements[0];
    return d;
  }

 private:
  static std::vector<double> vec(double a) { return std::vector<double>(1, a); }
  static std::vector<double> vec(double a, double b) { std::vector<double> v(2); v[0] = a; v[1] = b; return v; }
  static std::vector<double> vec(double a, double b, double c) { std::vector<double> v(3); v[0] = a; v[1] = b; v[2] = c; return v; }
  int push(int kind, int fa, int fb, int coord, int aux, std::uint64_t upstream, const std::vector<double>& p) {
    rkb_element e = rkb_element();
    e.kind = kind; e.frame_a = fa; e.frame_b = fb; e.coord = coord; e.aux = aux; e.upstream = upstream;
    for (std::size_t i = 0; i < p.size() && i < 12; ++i) e.p[i] = p[i];
    mElements.push_back(e);
    return static_cast<int>(mElements.size()) - 1;
  }
  int mDim, mFrames, mCoords, mInputs, mBaseFrame, mAux;
  rkb_base_frame mBase;
  std::vector<rkb_element> mElements;
};

class kte_batch_propagator {
 public:
  typedef std::vector<double> point_type;
  typedef std::vector<double> point_difference_type;
  typedef std::vector<double> point_derivative_type;
  typedef double time_type;
  typedef double time_difference_type;
  typedef std::vector<double> input_type;
  typedef std::vector<double> output_type;

  static const std::size_t dimensions = 0;  // run-time sized, as in kte_nl_system.hpp:94-96
  static const std::size_t input_dimensions = 0;
  static const std::size_t output_dimensions = 0;

  explicit kte_batch_propagator(const rkb_chain_desc& desc, int device = 0, double time_step = 1e-3)
      : mChain(NULL), mDevice(device), mDt(time_step) {
    int rc = rkb_chain_create(&desc, &mChain);
    if (rc != RKB_OK) throw propagator_error(rc, "rkb_chain_create");
  }
  explicit kte_batch_propagator(const chain_builder& builder, int device = 0, double time_step = 1e-3)
      : mChain(NULL), mDevice(device), mDt(time_step) {
    rkb_chain_desc d = builder.desc();
    int rc = rkb_chain_create(&d, &mChain);
    if (rc != RKB_OK) throw propagator_error(rc, "rkb_chain_create");
  }
  /// The kte_nl_system stored in a ReaK XML archive (`.rkx`, core/serialization/xml_archiver.cpp; rkb_rkx_load).
  struct from_rkx_t {};
  kte_batch_propagator(from_rkx_t, const std::string& path, int device = 0, double time_step = 1e-3)
      : mChain(NULL), mDevice(device), mDt(time_step) {
    char err[256] = {0};
    int rc = rkb_rkx_load(path.c_str(), 0u, &mChain, err, sizeof err);
    if (rc != RKB_OK) throw propagator_error(rc, std::string("rkb_rkx_load: ") + err);
  }
  ~kte_batch_propagator() { rkb_chain_destroy(mChain); }

  // ---- SSSystemConcept / DiscreteSSSConcept ----------------------------------------------------
  std::size_t get_state_dimensions() const { return static_cast<std::size_t>(rkb_chain_state_dim(mChain)); }
  std::size_t get_input_dimensions() const { return static_cast<std::size_t>(rkb_chain_input_dim(mChain)); }
  std::size_t get_output_dimensions() const { return 0; }
  time_difference_type get_time_step() const { return mDt; }
  void set_time_step(time_difference_type dt) { mDt = dt; }
  bool is_serial() const { return rkb_chain_is_serial(mChain) == 1; }
  /// Compile the kernels for this chain's own structure at run time (NVRTC; a few seconds, cached per structure).
  void specialize() { check(rkb_chain_specialize(mChain, mDevice), "rkb_chain_specialize"); }
  bool is_specialized() const { return rkb_chain_is_specialized(mChain) == 1; }

  /// kte_nl_system::get_state_derivative (kte_nl_system.hpp:238-346) for one state.
  template <typename StateSpaceType, typename Point, typename Input>
  point_derivative_type get_state_derivative(const StateSpaceType&, const Point& p, const Input& u, const time_type& = 0) const {
    check_sizes(p.size(), u.size());
    std::vector<double> x(p.size()), uu(u.size() ? u.size() : 1), xd(p.size());
    for (std::size_t i = 0; i < p.size(); ++i) x[i] = p[i];
    for (std::size_t i = 0; i < u.size(); ++i) uu[i] = u[i];
    int32_t st = 0;
    check(rkb_eval(mChain, mDevice, 1, &x[0], &uu[0], &xd[0], &st, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL), "rkb_eval");
    if (st & RKB_STATUS_SINGULAR) throw singularity_error("A");
    return xd;
  }
  /// num_int_dtnl_sys::get_next_state (num_int_dtnl_system.hpp:166-180): one RK4 step of get_time_step(), input held.
  template <typename StateSpaceType, typename Point, typename Input>
  point_type get_next_state(const StateSpaceType&, const Point& p, const Input& u, const time_type& = 0) const {
    check_sizes(p.size(), u.size());
    std::vector<double> x(p.size()), uu(u.size() ? u.size() : 1), xn(p.size());
    for (std::size_t i = 0; i < p.size(); ++i) x[i] = p[i];
    for (std::size_t i = 0; i < u.size(); ++i) uu[i] = u[i];
    int32_t st = 0;
    int rc = rkb_rollout_rk4(mChain, mDevice, 1, &x[0], &uu[0], mDt, 1, &xn[0], &st, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL);
    if (rc == RKB_ERR_INTEGRATION) throw impossible_integration("Integration is impossible: the time step is zero");
    check(rc, "rkb_rollout_rk4");
    if (st & RKB_STATUS_SINGULAR) throw singularity_error("A");
    return xn;
  }
  /// kte_nl_system::get_output with no system_output registered (kte_nl_system.hpp:356-372).
  template <typename StateSpaceType, typename Point, typename Input>
  output_type get_output(const StateSpaceType&, const Point&, const Input&, const time_type& = 0) const { return output_type(); }

  // ---- batched API (caller-owned buffers; flags = RKB_MEM_* | RKB_LAYOUT_*) ---------------------
  void get_state_derivatives(std::size_t n, const double* x, const double* u, double* xdot, int32_t* status = NULL,
                             unsigned flags = 0, void* stream = NULL) const {
    check(rkb_eval(mChain, mDevice, n, x, u, xdot, status, flags, stream), "rkb_eval");
  }
  void get_next_states(std::size_t n, const double* x, const double* u, int n_steps, double dt, double* x_out,
                       int32_t* status = NULL, unsigned flags = 0, void* stream = NULL) const {
    int rc = rkb_rollout_rk4(mChain, mDevice, n, x, u, dt, n_steps, x_out, status, flags, stream);
    if (rc == RKB_ERR_INTEGRATION) throw impossible_integration("Integration is impossible: zero step or negative step count");
    check(rc, "rkb_rollout_rk4");
  }
  /// Any fixed-step scheme of fixed_step_integrators.hpp (scheme = RKB_SCHEME_*) over n_intervals control
  /// intervals with the input constant within each — what a planner gets by calling get_next_state once
  /// per interval.  u: [n][n_intervals][n_inputs]; x_traj (nullable): [n][n_intervals][2 dof] (AoS).
  void rollout(std::size_t n, const double* x, const double* u, int scheme, int n_intervals, int steps_per_interval, double dt,
               double* x_out, double* x_traj = NULL, int32_t* status = NULL, unsigned flags = 0, void* stream = NULL) const {
    rkb_rollout_opts o = rkb_rollout_opts();
    o.scheme = scheme; o.n_intervals = n_intervals; o.steps_per_interval = steps_per_interval; o.dt = dt;
    int rc = rkb_rollout(mChain, mDevice, n, x, u, &o, x_out, x_traj, status, flags, stream);
    if (rc == RKB_ERR_INTEGRATION) throw impossible_integration("Integration is impossible: zero step, negative step count or no interval");
    check(rc, "rkb_rollout");
  }
  /// Host (AoS) buffers sharded over several GPUs of this box from one process: contiguous blocks of
  /// samples, one copy/compute pipeline per device, no inter-GPU communication.
  void get_next_states_multi(const std::vector<int>& devices, std::size_t n, const double* x, const double* u, int n_steps,
                             double dt, double* x_out, int32_t* status = NULL) const {
    int rc = rkb_rollout_rk4_multi(mChain, static_cast<int>(devices.size()), devices.empty() ? NULL : &devices[0], n, x, u, dt,
                                   n_steps, x_out, status);
    if (rc == RKB_ERR_INTEGRATION) throw impossible_integration("Integration is impossible: zero step or negative step count");
    check(rc, "rkb_rollout_rk4_multi");
  }
  void get_gen_forces(std::size_t n, const double* x, const double* u, double* f, unsigned flags = 0, void* stream = NULL) const {
    check(rkb_gen_forces(mChain, mDevice, n, x, u, f, flags, stream), "rkb_gen_forces");
  }
  void get_mass_matrices(std::size_t n, const double* x, double* M, double* Mdot = NULL, unsigned flags = 0, void* stream = NULL) const {
    check(rkb_mass_matrix(mChain, mDevice, n, x, M, Mdot, flags, stream), "rkb_mass_matrix");
  }
  /// Every frame after doMotion / clearForce / doForce: [n][frame_count()][RKB_FRAME_DOUBLES], layout in reak_b200.h.
  int frame_count() const { return rkb_chain_frame_count(mChain); }
  void get_frames(std::size_t n, const double* x, const double* u, double* frames, unsigned flags = 0, void* stream = NULL) const {
    check(rkb_frames(mChain, mDevice, n, x, u, frames, flags, stream), "rkb_frames");
  }
  /// proxy_query_pair_3D (geometry/proximity/proxy_query_model.hpp): two shape lists, the shapes anchored to frames
  /// of this chain or fixed in the world.  The caller owns the handle: rkb_proxy_destroy.
  rkb_proxy* make_proxy_pair(const std::vector<rkb_shape>& model1, const std::vector<rkb_shape>& model2) const {
    rkb_proxy* p = NULL;
    check(rkb_proxy_create(mChain, model1.empty() ? NULL : &model1[0], static_cast<int>(model1.size()),
                           model2.empty() ? NULL : &model2[0], static_cast<int>(model2.size()), &p), "rkb_proxy_create");
    return p;
  }
  /// proxy_query_pair_3D::findMinimumDistance at every state (proxy_query_model.cpp:388-412): mDistance [n], finder
  /// index [n] (nullable), mPoint1 / mPoint2 [n][6] (nullable).
  void get_min_distances(const rkb_proxy* pair, std::size_t n, const double* x, double* distance, int32_t* finder = NULL,
                         double* points = NULL, unsigned flags = 0, void* stream = NULL) const {
    check(rkb_min_distance(mChain, pair, mDevice, n, x, distance, finder, points, flags, stream), "rkb_min_distance");
  }
  /// Compiles the query of this chain and pair as straight-line CUDA now (rkb_proxy_specialize: NVRTC, a few seconds, cubin
  /// cached on disk) instead of waiting for the background compilation that starts once 4096 states have been queried.
  /// False when libnvrtc.so.12 is not installed (the interpreter kernel keeps serving the pair).
  bool specialize_proxy_pair(rkb_proxy* pair) const {
    const int rc = rkb_proxy_specialize(pair, mDevice);
    if (rc == RKB_ERR_UNSUPPORTED) return false;
    check(rc, "rkb_proxy_specialize");
    return true;
  }
  /// The steering kernel with the collision test of `pairs` built in (rkb_steer_checked_specialize): the steering loops with
  /// with_collision_check = true (MEAQR_topology.hpp:550-559) become one launch.  False for interpreter chains / no NVRTC.
  bool specialize_checked_steering(const std::vector<const rkb_proxy*>& pairs) const {
    const int rc = rkb_steer_checked_specialize(mChain, mDevice, pairs.empty() ? NULL : &pairs[0], static_cast<int>(pairs.size()));
    if (rc == RKB_ERR_UNSUPPORTED) return false;
    check(rc, "rkb_steer_checked_specialize");
    return true;
  }
  /// manip_dk_proxy_env_impl::is_free for every state: is_free[i] = 1 unless some pair reports a negative distance
  void get_is_free(const std::vector<const rkb_proxy*>& pairs, std::size_t n, const double* x, int32_t* is_free,
                   unsigned flags = 0, void* stream = NULL) const {
    check(rkb_is_free(mChain, mDevice, n, x, pairs.empty() ? NULL : &pairs[0], static_cast<int>(pairs.size()), is_free, flags, stream),
          "rkb_is_free");
  }
  /// manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99) for one state
  bool is_free(const std::vector<const rkb_proxy*>& pairs, const point_type& p) const {
    if (p.size() != get_state_dimensions()) throw std::range_error("State vector dimension mismatch!");
    if (pairs.empty()) return true;
    int32_t ok = 1;
    get_is_free(pairs, 1, &p[0], &ok);
    return ok != 0;
  }
  /// mass_matrix_calc::get_TMT_TdMT (mass_matrix_calculator.cpp:100-287): Tcm and (nullable) Tcm_dot, [n][rows][dof];
  /// twist_shaping_rows() and twist_shaping_mcm(Mcm) give the row count and the constant rows x rows Mcm.
  int twist_shaping_rows() const { return rkb_twist_shaping_rows(mChain); }
  void twist_shaping_mcm(double* Mcm) const { check(rkb_twist_shaping_mcm(mChain, Mcm), "rkb_twist_shaping_mcm"); }
  void get_twist_shaping(std::size_t n, const double* x, double* Tcm, double* Tcm_dot = NULL, unsigned flags = 0, void* stream = NULL) const {
    check(rkb_twist_shaping(mChain, mDevice, n, x, Tcm, Tcm_dot, flags, stream), "rkb_twist_shaping");
  }
  void steer_batch(std::size_t n_pairs, std::size_t n_rollouts, const double* x0, const double* goal, const double* u, int n_steps,
                   double dt, int32_t* best_idx, double* best_x, double* best_cost = NULL, int32_t* status = NULL,
                   unsigned flags = 0, void* stream = NULL) const {
    int rc = rkb_steer_batch(mChain, mDevice, n_pairs, n_rollouts, x0, goal, u, dt, n_steps, best_idx, best_x, best_cost, status, flags, stream);
    if (rc == RKB_ERR_INTEGRATION) throw impossible_integration("Integration is impossible: zero step or negative step count");
    check(rc, "rkb_steer_batch");
  }
  /// The whole loop of steer_with_constant_control (examples/misc/MEAQR_topology.hpp:503-561) /
  /// IHAQR_topology::move_position_toward_impl (examples/misc/IHAQR_topology.hpp:349-378) for n tuples:
  /// state feedback, get_bounded_input, one RK4 control interval, goal-proximity stop, steer record.
  void steer_feedback(std::size_t n, const double* x0, const double* x_goal, const double* u_bias, const double* gain, double* u_prev,
                      const rkb_steer_opts& opts, double* x_out, int32_t* n_done, double* x_traj = NULL, int32_t* status = NULL,
                      unsigned flags = 0, void* stream = NULL) const {
    int rc = rkb_steer_feedback(mChain, mDevice, n, x0, x_goal, u_bias, gain, u_prev, &opts, x_out, n_done, x_traj, status, flags, stream);
    if (rc == RKB_ERR_INTEGRATION) throw impossible_integration("Integration is impossible: zero step, no substep or negative time limit");
    check(rc, "rkb_steer_feedback");
  }
  double last_kernel_ms() const { return rkb_last_kernel_ms(mChain); }
  rkb_chain* handle() const { return mChain; }
  /// rkb_chain_set_option (RKB_OPT_*): execution choices that do not change results beyond rounding
  void set_option(int option, long long value) { check(rkb_chain_set_option(mChain, option, value), "rkb_chain_set_option"); }
  long long get_option(int option) const { return rkb_chain_get_option(mChain, option); }
  /// A caller's std::vector is pageable memory: the library then pays the driver's synchronous staged copies (about
  /// half the end-to-end rate).  pin() page-locks the vector's storage once (rkb_host_pin) so that every later call on
  /// it copies by DMA, overlapped with the kernels; unpin() before the vector is resized or destroyed.
  template <typename T> static void pin(std::vector<T>& v) { if (!v.empty()) check(rkb_host_pin(&v[0], v.size() * sizeof(T)), "rkb_host_pin"); }
  template <typename T> static void unpin(std::vector<T>& v) { if (!v.empty()) check(rkb_host_unpin(&v[0]), "rkb_host_unpin"); }

 private:
  kte_batch_propagator(const kte_batch_propagator&);
  kte_batch_propagator& operator=(const kte_batch_propagator&);
  void check_sizes(std::size_t np, std::size_t nu) const {
    if (np != get_state_dimensions()) throw std::range_error("State vector dimension mismatch!");
    if (nu != get_input_dimensions()) throw std::range_error("Input vector dimension mismatch!");
  }
  static void check(int rc, const char* where) {
    if (rc != RKB_OK) throw propagator_error(rc, where);
  }
  rkb_chain* mChain;
  int mDevice;
  double mDt;
};

}  // namespace reak_b200

#endif
