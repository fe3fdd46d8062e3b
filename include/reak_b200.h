/*
 * reak_b200.h — C-ABI of the B200 batched KTE-chain propagator.
 *
 * This is the drop-in boundary for ONE path of ReaK: evaluating and integrating a serial
 * Kinetostatic-Transmission-Element chain over a batch of independent states.  Every entry
 * point names the reference interface it replaces (paths relative to the ReaK source tree):
 *
 *   rkb_chain_create      <- model assembly: kte_map_chain::operator<<  (ctrl/mbd_kte/kte_map_chain.hpp:98-102),
 *                            mass_matrix_calc::operator<<               (ctrl/mbd_kte/mass_matrix_calculator.cpp:30-78),
 *                            kte_nl_system public members               (ctrl/ctrl_sys/kte_nl_system.hpp:70-78)
 *   rkb_eval              <- kte_nl_system::get_state_derivative        (ctrl/ctrl_sys/kte_nl_system.hpp:238-346);
 *                            with RKB_LAYOUT_BLOCKED: manipulator_dynamics_model::computeStateRate
 *                            (ctrl/mbd_kte/manipulator_model.cpp:292-355), the state_rate_function of the legacy models
 *   rkb_rollout_rk4       <- runge_kutta4_integrator<double>::integrate (core/integrators/fixed_step_integrators.hpp:256-293)
 *                            driven through num_int_dtnl_sys::get_next_state (ctrl/ctrl_sys/num_int_dtnl_system.hpp:166-180)
 *   rkb_rollout_rk4_inputs <- ctrl::detail::runge_kutta4_integrate_impl (ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp:50-97):
 *                            RK4 with an input trajectory read at t, t + dt/2, t + dt
 *   rkb_rollout           <- the same with euler / midpoint / runge_kutta4 / runge_kutta5 integrators
 *                            (fixed_step_integrators.hpp:60-399) over a sequence of control intervals
 *   rkb_mass_matrix       <- mass_matrix_calc::getMassMatrix / getMassMatrixAndDerivative
 *                                                                       (ctrl/mbd_kte/mass_matrix_calculator.cpp:80-98)
 *   rkb_twist_shaping     <- mass_matrix_calc::get_TMT_TdMT             (ctrl/mbd_kte/mass_matrix_calculator.cpp:100-287)
 *   rkb_frames            <- kte_map_chain::doMotion / doForce as seen on the frames (frame_3D / frame_2D members)
 *   rkb_steer_feedback_checked <- the same loops with with_collision_check = true (is_free_impl after every interval)
 *   rkb_is_free           <- manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99)
 *   rkb_min_distance      <- proxy_query_pair_3D::findMinimumDistance (geometry/proximity/proxy_query_model.cpp:388-412)
 *                            with the pair finders of geometry/proximity/prox_*_*.cpp: the is_free test of
 *                            ctrl/topologies/manip_free_workspace.hpp:77-99 on every propagated state
 *   rkb_gen_forces        <- kte_map_chain::doMotion/clearForce/doForce (ctrl/mbd_kte/kte_map_chain.hpp:71-89); returns gen_coord::f
 *   rkb_steer_batch       <- the inner loop of steer_with_constant_control (examples/misc/MEAQR_topology.hpp:503-561),
 *                            many (start, goal, control) tuples per call
 *   rkb_steer_feedback    <- that loop whole — state feedback, get_bounded_input (examples/misc/IHAQR_topology.hpp:304-327),
 *                            one RK4 control interval, goal-proximity stop, steer record — for many tuples per call
 *
 * Conventions
 *   - All arithmetic is IEEE double, like the reference.
 *   - State per sample: 2n doubles interleaved (q0, qd0, q1, qd1, ...) — kte_nl_system.hpp:189-193 — followed, for a
 *     chain with a free_joint_3D, by the 13 states of its coordinate frame: Position (3), Quat (w,x,y,z; normalised when
 *     the state is applied, the integrators advance the raw vector), Velocity (3), AngVelocity (3) —
 *     kte_nl_system.hpp:145-147, 205-219.  State derivative: Velocity, QuatDot, 6 accelerations (:293-308).  Generalised
 *     forces, M, Mdot and the columns of the twist-shaping matrix then have n + 6 entries / rows / columns (the six of
 *     the free joint last: Force, Torque / jacobian_3D_3D columns, mass_matrix_calculator.cpp:232-276).  A planar chain
 *     with a free_joint_2D carries 7 states — Position (2), Rotation (cos, sin; normalised when applied), Velocity (2),
 *     AngVelocity — and 3 accelerations instead (kte_nl_system.hpp:194-204, 282-291; jacobian_2D_2D columns).  At most one
 *     free joint per chain; interpreter kernels; not with RKB_LAYOUT_BLOCKED.  A rotor (inertia_gen) on coordinate 0 next to
 *     a free joint is rejected: the reference itself dereferences a null pointer there (mass_matrix_calculator.cpp:226-233).
 *     Input per sample: one double per driving_actuator_gen, in input-index order.
 *   - Buffers are caller-owned.  RKB_MEM_DEVICE pointers must be valid on `device`;
 *     RKB_MEM_HOST pointers are staged through device memory by the library
 *     (pinned host memory makes the copies asynchronous up to the final sync).
 *   - RKB_LAYOUT_AOS is the reference layout [N][dim]; RKB_LAYOUT_SOA is [dim][N];
 *     RKB_LAYOUT_BLOCKED orders a state as (q..., qd...) like manipulator_dynamics_model::computeStateRate.
 *   - No exception crosses this boundary.  Return 0 = ok, negative = error (rkb_strerror).
 *     Per-sample `status` words mirror the reference's exceptions:
 *       bit 0 (RKB_STATUS_SINGULAR)  Cholesky pivot < 1e-8, where linsolve_Cholesky throws
 *                                    singularity_error (core/lin_alg/mat_cholesky.hpp:80-82, 546)
 *       bit 1 (RKB_STATUS_NONFINITE) a non-finite state derivative was produced
 *   - There is no CPU fallback: every compute entry point fails with RKB_ERR_CUDA when no
 *     usable sm_100 device is present.
 */
#ifndef REAK_B200_H
#define REAK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RKB_VERSION 120

#if defined(__GNUC__)
#define RKB_API __attribute__((visibility("default")))
#else
#define RKB_API
#endif

/* ---- element kinds (one per in-scope kte_map subclass) ------------------------------------ */
enum rkb_kind {
  RKB_REVOLUTE_3D       = 1,  /* revolute_joint_3D   (revolute_joint.cpp:121-213)  p[0..2] = axis                   */
  RKB_PRISMATIC_3D      = 2,  /* prismatic_joint_3D  (prismatic_joint.cpp:129-222) p[0..2] = axis                   */
  RKB_FREE_3D           = 3,  /* free_joint_3D       (free_joints.cpp:123-208)     coord = index of its coordinate frame in dofs_3D */
  RKB_RIGID_LINK_3D     = 4,  /* rigid_link_3D       (rigid_link.cpp:152-185)      p[0..2] = offset, p[3..6] = quat (w,x,y,z) */
  RKB_INERTIA_3D        = 5,  /* inertia_3D          (inertia.cpp:111-121)         p[0] = mass, p[1..6] = Ixx Ixy Ixz Iyy Iyz Izz */
  RKB_INERTIA_GEN       = 6,  /* inertia_gen         (inertia.cpp:47-53)           p[0] = mass (jacobian_gen_gen(1,0))          */
  RKB_ACTUATOR_GEN      = 7,  /* driving_actuator_gen (driving_actuator.cpp:31-38) aux = input index, frame_b = joint element index */
  RKB_TORSION_SPRING_3D = 8,  /* torsion_spring_3D   (torsion_spring.cpp:106-129)  p[0] = stiffness, p[1] = saturation */
  RKB_TORSION_DAMPER_3D = 9,  /* torsion_damper_3D   (torsion_damper.cpp:93-104)   p[0] = damping                    */
  RKB_SPRING_3D         = 10, /* spring_3D           (spring.cpp:178-207)          p[0] = rest length, p[1] = stiffness, p[2] = saturation */
  RKB_DAMPER_3D         = 11, /* damper_3D           (damper.cpp:136-149)          p[0] = damping                    */
  /* elements on generalized coordinates alone.  Their anchors are system coordinates (index < n_coords) or AUXILIARY
   * coordinates — gen_coord objects that are not system states: fixed anchors, or the end of a rigid_link_gen — declared by
   * RKB_COORD_GEN records with indices n_coords, n_coords + 1, ... (n_coords + auxiliaries <= RKB_MAX_COORDS). */
  RKB_RIGID_LINK_GEN    = 12, /* rigid_link_gen      (rigid_link.cpp:34-75)        coord = base, aux = end (auxiliary), p[0] = offset */
  RKB_SPRING_GEN        = 13, /* spring_gen          (spring.cpp:50-84)            coord = anchor 1, aux = anchor 2, p as RKB_SPRING_3D */
  RKB_DAMPER_GEN        = 14, /* damper_gen          (damper.cpp:48-57)            coord = anchor 1, aux = anchor 2, p[0] = damping  */
  RKB_COORD_GEN         = 15, /* an auxiliary gen_coord (gen_coord.hpp:44-178): coord = its index, p[0..2] = q, q_dot, q_ddot it holds */
  RKB_REVOLUTE_2D       = 17, /* revolute_joint_2D   (revolute_joint.cpp:32-116)                                     */
  RKB_PRISMATIC_2D      = 18, /* prismatic_joint_2D  (prismatic_joint.cpp:33-123)  p[0..1] = axis                    */
  RKB_FREE_2D           = 19, /* free_joint_2D       (free_joints.cpp:33-117)      coord = index of its coordinate frame in dofs_2D */
  RKB_RIGID_LINK_2D     = 20, /* rigid_link_2D       (rigid_link.cpp:87-139)       p[0..1] = offset, p[2] = angle    */
  RKB_INERTIA_2D        = 21, /* inertia_2D          (inertia.cpp:77-86)           p[0] = mass, p[1] = moment of inertia */
  RKB_TORSION_SPRING_2D = 24, /* torsion_spring_2D   (torsion_spring.cpp:50-71)    p[0] = stiffness, p[1] = saturation */
  RKB_TORSION_DAMPER_2D = 25, /* torsion_damper_2D   (torsion_damper.cpp:49-58)    p[0] = damping                    */
  RKB_SPRING_2D         = 26, /* spring_2D           (spring.cpp:116-143)          as RKB_SPRING_3D                  */
  RKB_DAMPER_2D         = 27  /* damper_2D           (damper.cpp:88-102)           p[0] = damping                    */
};

#define RKB_MAX_COORDS 16 /* largest number of generalized coordinates one chain may carry */

/* One KTE of the chain, in kte_map_chain order (doMotion order; doForce runs it reversed). */
typedef struct rkb_element {
  int32_t  kind;      /* enum rkb_kind */
  int32_t  frame_a;   /* joint/link: base frame; spring/damper: anchor 1; inertia_2D/3D: CoM frame; else -1 */
  int32_t  frame_b;   /* joint/link: end frame;  spring/damper: anchor 2; actuator: index of the reacting joint element; else -1 */
  int32_t  coord;     /* generalized coordinate (joints, inertia_gen, actuator); else -1 */
  int32_t  aux;       /* actuator: input index; _gen link / spring / damper: second coordinate; else 0 */
  int32_t  reserved;
  uint64_t upstream;  /* inertias: bit c set <=> coordinate c is in mUpStreamJoints (jacobian_joint_map.hpp:252-331);
                         bit 32 + i set <=> the coordinate frame of free joint i is in mUpStream3DJoints */
  double   p[12];     /* parameters, see enum rkb_kind */
} rkb_element;

/* Kinematics of the un-driven root frame (robot_base in CRS_A465_models.cpp:298-301). */
typedef struct rkb_base_frame {
  double position[3];
  double quat[4];          /* (w,x,y,z); 2D chains: rotation angle in quat[0] */
  double velocity[3];
  double ang_velocity[3];  /* 2D chains: scalar in [0] */
  double acceleration[3];  /* gravity enters as an upward base acceleration */
  double ang_acceleration[3];
} rkb_base_frame;

typedef struct rkb_chain_desc {
  int32_t dim;          /* 2 or 3: all frames of a chain are frame_2D or frame_3D */
  int32_t n_elements;
  int32_t n_frames;     /* frame ids are 0..n_frames-1; base_frame is never written by doMotion */
  int32_t n_coords;     /* = kte_nl_system::dofs_gen.size(); state index j <-> coordinate j   */
  int32_t n_inputs;     /* = kte_nl_system::get_input_dimensions()                            */
  int32_t base_frame;
  rkb_base_frame base;
  const rkb_element* elements;
} rkb_chain_desc;

typedef struct rkb_chain rkb_chain; /* opaque */

/* ---- flags ---------------------------------------------------------------------------------- */
#define RKB_MEM_HOST    0u
#define RKB_MEM_DEVICE  1u
#define RKB_LAYOUT_AOS  0u
#define RKB_LAYOUT_SOA  2u
/* State vectors as manipulator_dynamics_model keeps them (ctrl/mbd_kte/manipulator_model.cpp:292-355):
 * all positions, then all velocities — (q0 .. qn-1, qd0 .. qdn-1) — instead of kte_nl_system's
 * interleaved pairs.  Applies to every state-shaped buffer of a call (x, xdot, x_out, x_traj, goal,
 * best_x); combines with AOS / SOA.  The gain columns of rkb_steer_feedback follow the same order. */
#define RKB_LAYOUT_BLOCKED 4u

#define RKB_STATUS_SINGULAR  1
#define RKB_STATUS_NONFINITE 2

/* ---- error codes ---------------------------------------------------------------------------- */
#define RKB_OK                 0
#define RKB_ERR_INVALID       -1  /* null pointer / bad argument / malformed descriptor        */
#define RKB_ERR_UNSUPPORTED   -2  /* chain topology or element outside the compiled path       */
#define RKB_ERR_DIMENSION     -3  /* std::range_error of kte_nl_system::apply_states_and_inputs */
#define RKB_ERR_CUDA          -4  /* CUDA runtime failure or no usable device                  */
#define RKB_ERR_NOMEM         -5
#define RKB_ERR_INTEGRATION   -6  /* impossible_integration (dt == 0, n_steps < 0)             */

RKB_API int         rkb_version(void);
/* 16 hex digits: a hash over the sources this binary was built from (profiles and benches name the build they measured). */
RKB_API const char* rkb_build_id(void);
RKB_API const char* rkb_strerror(int code);
/* Text of the last CUDA error seen by the calling thread ("" if none). */
RKB_API const char* rkb_last_cuda_error(void);

/* Validate `desc`, lower it to the device program and return a handle.  The descriptor is
 * copied; the caller may free it afterwards. */
RKB_API int  rkb_chain_create(const rkb_chain_desc* desc, rkb_chain** out);
/* ... with a say in the lowering (0 = rkb_chain_create).  The results of every entry point are the same up to
 * rounding whatever the flags; they exist so that the kernel families can be compared with each other. */
#define RKB_CREATE_INTERPRETER 1u  /* run on the interpreter kernels even if the chain has the serial form        */
#define RKB_CREATE_GENERAL     2u  /* serial kernels without structural specialisation (axis / link / tensor shape) */
RKB_API int  rkb_chain_create_ex(const rkb_chain_desc* desc, unsigned create_flags, rkb_chain** out);

/* Execution options of a handle (none changes results beyond rounding; all have measured defaults):
 *   RKB_OPT_SPLIT_MAX_SAMPLES  RK4 rollouts, control sequences and steering loops of at most this many samples run with
 *                              one sample on a PAIR of warps (forces | mass matrix + solve) instead of one thread per
 *                              sample: ~1.5x faster while the batch cannot fill the GPU.  0 = never, -1 = default (8192
 *                              for chains of 4 or more coordinates, 0 for shorter ones, whose evaluation is too brief).
 *   RKB_OPT_FUSED_STEER        1 (default): rkb_steer_feedback on a serial chain is one launch; 0: one control-law and
 *                              one rollout launch per interval (what interpreter chains always do)
 *   RKB_OPT_FUSED_SEQUENCE     1 (default): rkb_rollout with the RK4 scheme on a serial chain is one launch; 0: one per interval
 *   RKB_OPT_HOST_PIPELINE      1 (default): large RKB_MEM_HOST rollouts are cut into chunks whose copies overlap the kernels
 *   RKB_OPT_AUTO_SPECIALIZE    1 (default): a serial chain whose structure the shipped kernels do not fully match gets kernels
 *                              compiled for exactly its structure on the first call of >= 4096 samples — from the disk cache
 *                              ($RKB_CACHE_DIR, else $XDG_CACHE_HOME/reak_b200, else ~/.cache/reak_b200; RKB_CACHE_DIR=off
 *                              disables it) or by NVRTC on a background thread; calls made meanwhile run on the shipped
 *                              kernels, so results may change at rounding level once the new kernels take over.  0: only on
 *                              rkb_chain_specialize. */
enum rkb_option { RKB_OPT_SPLIT_MAX_SAMPLES = 1, RKB_OPT_FUSED_STEER = 2, RKB_OPT_FUSED_SEQUENCE = 3, RKB_OPT_HOST_PIPELINE = 4,
                  RKB_OPT_AUTO_SPECIALIZE = 5 };
RKB_API int       rkb_chain_set_option(rkb_chain* chain, int option, long long value);
RKB_API long long rkb_chain_get_option(const rkb_chain* chain, int option);
RKB_API void rkb_chain_destroy(rkb_chain* chain);

RKB_API int  rkb_chain_state_dim(const rkb_chain* chain);  /* kte_nl_system::get_state_dimensions, kte_nl_system.hpp:145-147 */
RKB_API int  rkb_chain_input_dim(const rkb_chain* chain);  /* kte_nl_system::get_input_dimensions, kte_nl_system.hpp:153-158 */
RKB_API int  rkb_chain_dof(const rkb_chain* chain);
/* 1 when the chain was lowered to the register-resident serial-chain kernels, 0 when it runs on
 * the interpreter kernels (any element order, 2D frames, two-anchor springs/dampers). */
RKB_API int  rkb_chain_is_serial(const rkb_chain* chain);
/* Structure found in the descriptor (axis-aligned joints, axis-aligned unrotated links, diagonal
 * inertia tensors; 8 bits per stage) and the part of it the selected kernels are specialised on. */
RKB_API unsigned long long rkb_chain_shape(const rkb_chain* chain);
RKB_API unsigned long long rkb_chain_kernel_shape(const rkb_chain* chain);

/* Run-time specialisation.  The library ships the general serial kernels and those specialised for the shapes of
 * the reference's own models; rkb_chain_specialize compiles the same kernel source for exactly this chain's
 * structure (NVRTC: a few seconds, cached per structure for the life of the process) and routes every later
 * launch of this handle to the result — about twice the speed of the general code for a chain with
 * axis-aligned joints that is not among the shipped shapes.  Results are those of the general code up to rounding.
 * RKB_ERR_UNSUPPORTED: the chain runs on the interpreter kernels, or libnvrtc.so.12 is not installed. */
RKB_API int rkb_chain_specialize(rkb_chain* chain, int device);
RKB_API int rkb_chain_is_specialized(const rkb_chain* chain);

/* xdot[i] = get_state_derivative(x[i], u[i]).  x: N x 2n, u: N x n_inputs, xdot: N x 2n,
 * status: N (nullable).  `stream` is a cudaStream_t (NULL = default stream). */
RKB_API int rkb_eval(rkb_chain* chain, int device, size_t n_samples,
             const double* x, const double* u, double* xdot, int32_t* status,
             unsigned flags, void* stream);

/* x_out[i] = state after exactly n_steps RK4 steps of size dt from x0[i] with u[i] held
 * constant (zero-order hold, num_int_dtnl_system.hpp:166-180).  The step COUNT is explicit:
 * the reference's time-driven loop would take one step more or less depending on rounding. */
RKB_API int rkb_rollout_rk4(rkb_chain* chain, int device, size_t n_samples,
                    const double* x0, const double* u, double dt, int n_steps,
                    double* x_out, int32_t* status, unsigned flags, void* stream);

/* ---- general fixed-step rollout ----------------------------------------------------------------
 * The fixed-step schemes of core/integrators/fixed_step_integrators.hpp, driven the way a planner
 * drives them: n_intervals control intervals, the input constant within an interval
 * (num_int_dtnl_sys::get_next_state once per interval, ctrl/ctrl_sys/num_int_dtnl_system.hpp:166-180;
 * steer_with_constant_control's inner loop, examples/misc/MEAQR_topology.hpp:539-547), and
 * steps_per_interval integrator steps of size dt per interval. */
enum rkb_scheme {
  RKB_SCHEME_EULER    = 1,  /* euler_integrator<T>::integrate          fixed_step_integrators.hpp:64-84   */
  RKB_SCHEME_MIDPOINT = 2,  /* midpoint_integrator<T>::integrate       fixed_step_integrators.hpp:177-202 */
  RKB_SCHEME_RK4      = 4,  /* runge_kutta4_integrator<T>::integrate   fixed_step_integrators.hpp:257-293 */
  RKB_SCHEME_RK5      = 5   /* runge_kutta5_integrator<T>::integrate   fixed_step_integrators.hpp:351-399 */
};

typedef struct rkb_rollout_opts {
  int32_t scheme;              /* enum rkb_scheme */
  int32_t n_intervals;         /* >= 1 */
  int32_t steps_per_interval;  /* >= 0 */
  int32_t reserved;            /* must be 0 */
  double  dt;                  /* integrator step, != 0 (negative integrates backwards) */
} rkb_rollout_opts;

/* x0: N x 2n.  u: one input vector per sample and interval — AOS [N][n_intervals][n_inputs],
 * SOA [n_intervals][n_inputs][N].  x_out: N x 2n, the state after the last interval.
 * x_traj (nullable): the state at the end of every interval — AOS [N][n_intervals][2n],
 * SOA [n_intervals][2n][N].  status: N (nullable), bits OR-ed over the whole rollout. */
RKB_API int rkb_rollout(rkb_chain* chain, int device, size_t n_samples,
                        const double* x0, const double* u, const rkb_rollout_opts* opts,
                        double* x_out, double* x_traj, int32_t* status, unsigned flags, void* stream);

/* RK4 with an input TRAJECTORY instead of a held input — ctrl::detail::runge_kutta4_integrate_impl
 * (ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp:50-97) reads u_traj at t for the first evaluation of a step, at
 * t + dt/2 for the second and third and at t + dt for the fourth.  u_nodes holds that trajectory sampled at every half
 * step: 2 n_steps + 1 nodes per sample, AOS [N][2 n_steps + 1][n_inputs], SOA [2 n_steps + 1][n_inputs][N].  With all
 * nodes equal this is rkb_rollout_rk4. */
RKB_API int rkb_rollout_rk4_inputs(rkb_chain* chain, int device, size_t n_samples,
                                   const double* x0, const double* u_nodes, double dt, int n_steps,
                                   double* x_out, int32_t* status, unsigned flags, void* stream);

/* One rank's share of a sample-sharded rollout WITH its all-gather (one process per GPU): the rank
 * integrates its n_samples states and the kernel stores every end state to row `row_offset + i` of EACH of the n_dest
 * (<= 8) destination buffers — its own copy of the gathered batch and the other ranks' copies, mapped into this process
 * as peer memory (CUDA IPC handles, or torch.distributed's symmetric memory).  The stores travel over NVLink while the
 * other CTAs integrate; nothing is left to gather afterwards — the ranks only need a barrier before they read.
 * x_out_dest[d]: [n_total][2n] AoS; status_dest (nullable, entries nullable): [n_total].  RKB_MEM_DEVICE | RKB_LAYOUT_AOS
 * (| RKB_LAYOUT_BLOCKED) only; serial chains only (RKB_ERR_UNSUPPORTED otherwise).  Returns without synchronising. */
RKB_API int rkb_rollout_rk4_scatter(rkb_chain* chain, int device, size_t n_samples, const double* x0, const double* u,
                                    double dt, int n_steps, int n_dest, double* const* x_out_dest, int32_t* const* status_dest,
                                    size_t row_offset, unsigned flags, void* stream);

/* rkb_rollout_rk4 on HOST buffers (AoS), sharded over `n_devices` GPUs of this box from one
 * process: contiguous blocks of samples, one copy/compute pipeline per device, no inter-GPU
 * communication (the reference has no counterpart; per sample the semantics are unchanged). */
RKB_API int rkb_rollout_rk4_multi(rkb_chain* chain, int n_devices, const int* devices, size_t n_samples,
                                  const double* x0, const double* u, double dt, int n_steps,
                                  double* x_out, int32_t* status);

/* Generalised force gen_coord::f after doMotion/clearForce/doForce with q_ddot = 0
 * (tau - h(q,qd) in ReaK's convention).  f: N x n. */
RKB_API int rkb_gen_forces(rkb_chain* chain, int device, size_t n_samples,
                   const double* x, const double* u, double* f,
                   unsigned flags, void* stream);

/* M[i] (n x n, full symmetric storage, row-major) and, if Mdot != NULL, its time
 * derivative at state x[i].  AOS: [N][n][n]; SOA: [n*n][N]. */
RKB_API int rkb_mass_matrix(rkb_chain* chain, int device, size_t n_samples,
                    const double* x, double* M, double* Mdot,
                    unsigned flags, void* stream);

/* Every frame of the chain after doMotion / clearForce / doForce (ctrl/mbd_kte/kte_map_chain.hpp:71-89) at
 * state x[i], input u[i], q_ddot = 0 — the direct kinematics (poses, velocities, accelerations) a collision
 * check or a measurement model reads, and the wrenches the elements left on the frames.  25 doubles per frame:
 *   frame_3D  Position 0-2, Quat (w,x,y,z) 3-6, Velocity 7-9, AngVelocity 10-12, Acceleration 13-15,
 *             AngAcceleration 16-18, Force 19-21, Torque 22-24   (conventions of frame_3D.hpp:69-78)
 *   frame_2D  Position 0-1, Rotation (cos, sin) 3-4, Velocity 7-8, AngVelocity 10, Acceleration 13-14,
 *             AngAcceleration 16, Force 19-20, Torque 22; the other slots are zero
 * frames: AOS [N][n_frames][25], SOA [n_frames * 25][N].  Interpreter kernels (the serial kernels work in
 * link-local coordinates and never build these frames). */
#define RKB_FRAME_DOUBLES 25
RKB_API int rkb_chain_frame_count(const rkb_chain* chain);
/* Samples one full wave of the RK4 rollout kernel holds on `device` (SMs x resident CTAs x 128; 0 for interpreter
 * chains): cut batches that are processed piecewise at multiples of it, every launch ends with a partial wave. */
RKB_API long long rkb_chain_wave_samples(rkb_chain* chain, int device);
RKB_API int rkb_frames(rkb_chain* chain, int device, size_t n_samples,
                       const double* x, const double* u, double* frames, unsigned flags, void* stream);

/* ---- proximity queries on the propagated states ---------------------------------------
 * A proximity model is a list of primitive shapes (geometry/shapes/{plane,sphere,capped_cylinder,cylinder,box}.hpp),
 * each riding on a frame of the chain (geometry_3D::mAnchor) or fixed in the world, with its own pose relative to
 * that anchor (geometry_3D::mPose).  rkb_proxy pairs two models the way proxy_query_pair_3D does
 * (geometry/proximity/proxy_query_model.hpp): one finder per (shape of model 1, shape of model 2) the reference has
 * a finder for, in createProxFinderList order (proxy_query_model.cpp:212-384; planar chains with planar shapes:
 * proxy_query_pair_2D, :73-160). */
enum rkb_shape_kind {
  RKB_SHAPE_PLANE     = 1,   /* dims: x, y extents (only the culling test looks at them)  plane.hpp      */
  RKB_SHAPE_SPHERE    = 2,   /* dims: radius                                              sphere.hpp     */
  RKB_SHAPE_CCYLINDER = 3,   /* dims: length, radius; axis = local z                      capped_cylinder.hpp */
  RKB_SHAPE_CYLINDER  = 4,   /* dims: length, radius; axis = local z                      cylinder.hpp   */
  RKB_SHAPE_BOX       = 5,   /* dims: x, y, z extents                                     box.hpp        */
  /* planar shapes, for planar chains (proxy_query_pair_2D, proxy_query_model.cpp:73-212): position[0..1] and
   * quat[0..1] = (cos, sin) of the shape's own rotation (pose_2D: Position, rot_mat_2D), the rest ignored; every pair of
   * planar shapes has a finder (prox_circle_circle / circle_crect / circle_rectangle / crect_crect / crect_rectangle /
   * rectangle_rectangle.cpp).  Points come back as (x, y, 0). */
  RKB_SHAPE_CIRCLE    = 6,   /* dims: radius                                              circle.hpp     */
  RKB_SHAPE_CRECT     = 7,   /* dims: length along x, width (= diameter of the round caps)  capped_rectangle.hpp */
  RKB_SHAPE_RECTANGLE = 8    /* dims: x, y extents                                        rectangle.hpp  */
};
typedef struct rkb_shape {
  int32_t kind;          /* rkb_shape_kind */
  int32_t anchor;        /* frame id of the chain descriptor the shape rides on, -1 = world */
  double  position[3];   /* mPose.Position, in the anchor's coordinates */
  double  quat[4];       /* mPose.Quat (w,x,y,z) */
  double  dims[3];
} rkb_shape;
#define RKB_PROXY_MAX_SHAPES 16   /* per model */
typedef struct rkb_proxy rkb_proxy;

RKB_API int  rkb_proxy_create(const rkb_chain* chain, const rkb_shape* model1, int n1,
                              const rkb_shape* model2, int n2, rkb_proxy** out);
RKB_API void rkb_proxy_destroy(rkb_proxy* proxy);
/* number of finders; finder k pairs shape *i1 of model 1 with shape *i2 of model 2 */
/* Run-time specialisation of a pair's query (rkb_min_distance, rkb_is_free, rkb_steer_feedback_checked).  The library
 * ships an interpreter kernel that reads the chain and the shape list at run time; rkb_proxy_specialize writes the query of
 * THIS chain and pair as straight-line CUDA (forward kinematics with the chain's constants, world-fixed shapes as literals,
 * every finder of createProxFinderList with its argument order resolved), compiles it with NVRTC (a few seconds; cubins
 * are cached on disk like those of rkb_chain_specialize) and routes the pair's launches to it: ~2x the interpreter.
 * Results agree with the interpreter kernel to rounding.  By default (RKB_PROXY_OPT_AUTO_SPECIALIZE = 1, and the chain's
 * RKB_OPT_AUTO_SPECIALIZE on) the same happens on a background thread once 4096 states have been queried (in one call or
 * over many small ones); queries made meanwhile run on the interpreter.
 * RKB_ERR_UNSUPPORTED: libnvrtc.so.12 is not installed. */
enum rkb_proxy_option { RKB_PROXY_OPT_AUTO_SPECIALIZE = 1, RKB_PROXY_OPT_MIN_BLOCKS = 2 /* CTAs per SM compiled for, 1..8 */ };
RKB_API int  rkb_proxy_set_option(rkb_proxy* proxy, int option, long long value);
RKB_API int  rkb_proxy_specialize(rkb_proxy* proxy, int device);
RKB_API int  rkb_proxy_is_specialized(const rkb_proxy* proxy);
/* test hook: the CUDA source rkb_proxy_specialize compiles (NUL-terminated); returns its size including the NUL,
 * out may be NULL to ask for the size */
RKB_API int  rkb_proxy_source(const rkb_proxy* proxy, char* out, size_t size);

RKB_API int  rkb_proxy_finder_count(const rkb_proxy* proxy);
RKB_API int  rkb_proxy_finder(const rkb_proxy* proxy, int k, int* i1, int* i2);

/* diagnostic: copies the lowered device program (an opaque blob: normalised poses, bounding radii) into out;
 * returns its size in bytes, or RKB_ERR_INVALID when `size` is too small.  tests/host_build uses it to run the
 * device source of the finders on the host against the compiled reference. */
RKB_API int  rkb_proxy_program(const rkb_proxy* proxy, void* out, size_t size);
/* Test hook: the interpreter program the descriptor was lowered to (an internal structure of this build, rkb_types.h:
 * GenericProgram), for tests/host_build/generic_host.cpp which runs the device source of the interpreter on the host. */
RKB_API int  rkb_chain_program(const rkb_chain* chain, void* out, size_t size);

/* proxy_query_pair_3D::findMinimumDistance (proxy_query_model.cpp:388-412) after the chain's doMotion at state
 * x[i], including its bounding-sphere culling, for every sample:
 *   distance[i]  mLastResult.mDistance of the finder it returns (+inf when the pair has no finder);
 *                the state is free in the sense of manip_dk_proxy_env_impl::is_free
 *                (ctrl/topologies/manip_free_workspace.hpp:77-99) iff distance[i] >= 0 for every proxy pair
 *   finder[i]    index of that finder (-1 when there is none), nullable
 *   points[i]    mPoint1, mPoint2 of its record (6 doubles, world coordinates), nullable
 * x as in rkb_eval (only the positions matter).  flags: RKB_LAYOUT_* as everywhere. */
RKB_API int rkb_min_distance(rkb_chain* chain, const rkb_proxy* proxy, int device, size_t n_samples,
                             const double* x, double* distance, int32_t* finder, double* points,
                             unsigned flags, void* stream);


/* proxy_query_pair_3D::gatherCollisionPoints (geometry/proximity/proxy_query_model.cpp:402-421) at the pose of every state:
 * each finder of the pair (createProxFinderList order) whose bounding spheres overlap is evaluated and every one that
 * reports a NEGATIVE distance contributes a record.
 *   count   [N]                    colliding finders of the state (can exceed max_records: only the first are stored)
 *   finder  [N][max_records]       nullable: finder index of each stored record, -1 beyond count
 *   records [N][max_records][7]    mDistance, mPoint1 (3), mPoint2 (3) in world coordinates; distance +inf beyond count
 * 3D chains, RKB_LAYOUT_AOS. */
RKB_API int rkb_collision_points(rkb_chain* chain, const rkb_proxy* proxy, int device, size_t n_samples, const double* x, int max_records,
                                 int32_t* count, int32_t* finder, double* records, unsigned flags, void* stream);

/* manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99) and the is_free_impl of the
 * steering topologies (examples/misc/MEAQR_topology.hpp:921-940) for every state: is_free[i] = 1 unless some proxy
 * pair's findMinimumDistance reports a negative distance at x[i] (a pair without finders never objects), else 0.
 * pairs: n_pairs >= 1 handles made for `chain`.  One proximity launch per pair and one combining launch. */
RKB_API int rkb_is_free(rkb_chain* chain, int device, size_t n_samples, const double* x,
                        const rkb_proxy* const* pairs, int n_pairs, int32_t* is_free,
                        unsigned flags, void* stream);

/* Twist-shaping matrix Tcm and its time derivative Tcm_dot of mass_matrix_calc::get_TMT_TdMT
 * (ctrl/mbd_kte/mass_matrix_calculator.cpp:100-287) at state x[i].  Rows: one per inertia_gen, then three per
 * inertia_2D (v2, w), then six per inertia_3D (v3, w3: jacobian_gen_3D::get_jac_relative_to,
 * core/kinetostatics/motion_jacobians.hpp:238-279), each group in chain order; columns: the coordinates.
 * M = Tcm^T Mcm Tcm and Mdot = Tcm_dot^T Mcm Tcm + transpose.  Tcm, Tcm_dot (nullable): AOS [N][rows][n],
 * SOA [rows * n][N].  Mcm is constant: rkb_twist_shaping_mcm writes it (rows x rows, row-major, host memory).
 * Always evaluated by the interpreter kernels (the serial kernels never form Tcm). */
RKB_API int rkb_twist_shaping_rows(const rkb_chain* chain);
RKB_API int rkb_twist_shaping_mcm(const rkb_chain* chain, double* Mcm);
RKB_API int rkb_twist_shaping(rkb_chain* chain, int device, size_t n_samples,
                              const double* x, double* Tcm, double* Tcm_dot, unsigned flags, void* stream);

/* Direct-kinematics Jacobian of ONE frame — an end effector — and its time derivative at state x[i]: what
 * manip_kin_mdl_jac_calculator::getJacobianMatrixAndDerivative (ctrl/mbd_kte/manipulator_model_helper.hpp:342-...) stacks
 * for the dependent frames of a manipulator, i.e. jacobian_gen_3D / _2D::get_jac_relative_to(frame) of every upstream
 * joint (core/kinetostatics/motion_jacobians.hpp:238-279, 139-147).  `frame`: frame id of the descriptor; `upstream`: bit
 * c set <=> coordinate c moves the frame (the joint_dependent_frame's mUpStreamJoints); other columns are zero.
 * J, Jdot (nullable): 3D chains [N][6][n] — rows v (3) then w (3), in the frame's own coordinates — 2D chains [N][3][n];
 * SOA [rows * n][N].  Interpreter kernels. */
RKB_API int rkb_frame_jacobian(rkb_chain* chain, int device, size_t n_samples, const double* x, int frame, uint64_t upstream,
                               double* J, double* Jdot, unsigned flags, void* stream);


/* ---- model files -----------------------------------------------------------------------------------------
 * Read a ctrl::kte_nl_system (ctrl/ctrl_sys/kte_nl_system.hpp:376-395: dofs_gen, dofs_3D, inputs, chain, mass_calc) from
 * a ReaK XML archive — the `.rkx` files ReaK::serialization::xml_oarchive writes (core/serialization/xml_archiver.cpp,
 * examples/robot_airship/build_P3R3R_model.cpp:78) — without any ReaK code: classes are recognised by their RTTI
 * numbers, shared objects by their object_ID.  The descriptor is the one include/reak_b200/reak_bridge.hpp derives from
 * the same system loaded by ReaK's own xml_iarchive (same element order, frame and coordinate numbering).  Values are
 * what the file holds (the archiver prints 6 significant digits).
 *   rkb_rkx_read  fills *desc and elements[0 .. n) (desc->elements = elements); returns the element count n >= 0.
 *                 elements == NULL: count and header only.
 *   rkb_rkx_load  reads and lowers in one go (create_flags as rkb_chain_create_ex).
 * Errors: RKB_ERR_INVALID (file cannot be opened), RKB_ERR_UNSUPPORTED (not a kte_nl_system archive, or an element
 * outside the compiled set), RKB_ERR_NOMEM (max_elements too small); `err` receives a message. */
RKB_API int rkb_rkx_read(const char* path, rkb_chain_desc* desc, rkb_element* elements, int max_elements, char* err, size_t err_len);
RKB_API int rkb_rkx_load(const char* path, unsigned create_flags, rkb_chain** out, char* err, size_t err_len);

/* ---- nearest neighbours -----------------------------------------------------------------------------------
 * The k vertices nearest to each query point, the search the planners run before every steer:
 * ReaK::pp::linear_neighbor_search / dvp_tree (ctrl/path_planning/topological_search.hpp:91-112, 238-270, 586-596;
 * metric_space_search.hpp, dvp_tree_detail.hpp — a vantage-point tree is an exact search, it returns what the linear scan
 * returns) under distance(a, b) = norm_2(difference(b, a)) of vect_n points (core/lin_alg/vect_alg.hpp:2314-2333).
 *   vertices [n_vertices][dim], queries [n_queries][dim]   row-major (the points of the motion graph / the samples)
 *   radius   only vertices with distance < radius qualify (+inf: no limit), as min_dist_linear_search's `radius`
 *   index    [n_queries][k]  vertex indices by ascending distance; -1 where fewer than k qualify
 *   distance [n_queries][k]  nullable; +inf where index is -1        count [n_queries] nullable: qualifying neighbours (<= k)
 * Distances are bit-identical to the reference's arithmetic (separately rounded multiply / add in index order, sqrt);
 * among equal distances the lowest vertex index comes first (the element min_dist_linear_search meets first).
 * flags: RKB_MEM_HOST or RKB_MEM_DEVICE (all buffers alike; device: asynchronous on `stream`). */
#define RKB_NEAREST_MAX_K   16
#define RKB_NEAREST_MAX_DIM 48
RKB_API int rkb_nearest(int device, size_t n_vertices, const double* vertices, size_t n_queries, const double* queries, int dim, int k,
                        double radius, int32_t* index, double* distance, int32_t* count, unsigned flags, void* stream);
RKB_API const char* rkb_nearest_last_error(void);

/* Linearisation of the dynamics about (x[i], u[i]) — what a linear-quadratic steering asks its system for
 * (get_linear_blocks in examples/misc/IHAQR_topology.hpp:240-258, MEAQR_topology.hpp): A[i] = d xdot / d x
 * (2n x 2n, row-major) and B[i] = d xdot / d u (2n x n_inputs), by central differences of get_state_derivative with the
 * step h = eps max(1, |component|) per state / input component (eps <= 0: 1e-6); the divisor is the difference of the
 * perturbed values as formed in floating point.  A or B may be NULL.  status (nullable): the bits of every
 * evaluation involved, OR-ed.  AOS, interleaved states only (RKB_LAYOUT_SOA / _BLOCKED: RKB_ERR_UNSUPPORTED). */
RKB_API int rkb_linearize(rkb_chain* chain, int device, size_t n_samples, const double* x, const double* u, double eps,
                          double* A, double* B, int32_t* status, unsigned flags, void* stream);

/* For each of n_pairs (start, goal) pairs roll out n_rollouts constant controls for n_steps
 * RK4 steps and keep the rollout whose end state is closest to the goal (Euclidean norm over
 * the 2n state components).  x0, goal: P x 2n; u: P x R x n_inputs (AOS) or [n_inputs][P*R] (SOA);
 * best_idx: P; best_x: P x 2n; best_cost: P (nullable); status: P x R (nullable). */
RKB_API int rkb_steer_batch(rkb_chain* chain, int device, size_t n_pairs, size_t n_rollouts,
                    const double* x0, const double* goal, const double* u,
                    double dt, int n_steps,
                    int32_t* best_idx, double* best_x, double* best_cost, int32_t* status,
                    unsigned flags, void* stream);

/* ---- closed-loop steering ------------------------------------------------------------------------
 * The loop of steer_with_constant_control (examples/misc/MEAQR_topology.hpp:503-561) and of
 * IHAQR_topology::move_position_toward_impl (examples/misc/IHAQR_topology.hpp:349-378) for N independent
 * (start, goal, gains) tuples at once, collision checking off:
 *
 *   while (k < max_intervals && ||x - x_goal||_2 > goal_proximity) {
 *     correction = -gain (x - x_goal)                       // -K H^-1 (x - x_goal) resp. -K (x - x_goal)
 *     u = (k == 0 && !saturate_first) ? u_bias + correction // MEAQR_topology.hpp:521-522
 *                                     : get_bounded_input(u_prev, u_bias, correction)  // IHAQR_topology.hpp:304-327
 *     x = `substeps` RK4 steps of `dt` with u held          // runge_kutta4_integrate_impl, one control interval
 *     u_prev = u;  ++k;  record x                           // the steer record
 *   }
 *
 * get_bounded_input clamps u_bias to the input box, bisects the correction (10 halvings) until the sum
 * is inside the box, and limits the rate (u - u_prev) / time_step to the rate box; a NULL box is unbounded. */
typedef struct rkb_steer_opts {
  double  time_step;        /* control interval T of the rate limit (m_time_step), > 0                      */
  double  dt;               /* RK4 step inside an interval (MEAQR: T / 10, IHAQR: T / 100), != 0             */
  double  goal_proximity;   /* m_goal_proximity_threshold                                                   */
  int32_t substeps;         /* RK4 steps per interval, >= 1                                                 */
  int32_t max_intervals;    /* the time limit, in intervals, >= 0                                           */
  int32_t saturate_first;   /* 0: the first interval is not saturated (MEAQR); 1: every interval is (IHAQR) */
  int32_t reserved;         /* must be 0                                                                    */
  const double* u_lower;    /* n_inputs doubles in HOST memory, or NULL; lower < upper component-wise      */
  const double* u_upper;
  const double* du_lower;   /* input-rate box, likewise                                                     */
  const double* du_upper;
} rkb_steer_opts;

/* x0, x_goal: N x 2n.  u_bias: N x n_inputs.  gain: N x n_inputs x 2n (row-major per sample).
 * u_prev: N x n_inputs, in: the input applied before the steer, out: the last input applied.
 * x_out: N x 2n.  n_done: N, intervals performed.  x_traj (nullable): N x max_intervals x 2n, slot j of
 * sample i holds the state after interval j for j < n_done[i]; later slots are unspecified.
 * AOS only (RKB_LAYOUT_SOA -> RKB_ERR_UNSUPPORTED); RKB_MEM_HOST or RKB_MEM_DEVICE. */
RKB_API int rkb_steer_feedback(rkb_chain* chain, int device, size_t n_samples,
                               const double* x0, const double* x_goal, const double* u_bias, const double* gain,
                               double* u_prev, const rkb_steer_opts* opts,
                               double* x_out, int32_t* n_done, double* x_traj, int32_t* status,
                               unsigned flags, void* stream);

/* rkb_steer_feedback with with_collision_check = true: after every control interval the state it ended on is
 * tested like is_free_impl (examples/misc/MEAQR_topology.hpp:921-940: no proxy pair reports a negative
 * findMinimumDistance); a state that is not free is NOT accepted — the loop of that sample stops on the last free
 * state, u_prev keeps the last accepted input, no waypoint is recorded (MEAQR_topology.hpp:550-559,
 * IHAQR_topology.hpp:369-375) — and collided[i] = 1 (was_collision_free == false), else 0.
 * pairs: n_pairs >= 1 handles of rkb_proxy_create made for `chain`.  Runs interval by interval (control law,
 * rollout into a scratch state, one proximity launch per pair, acceptance). */
RKB_API int rkb_steer_feedback_checked(rkb_chain* chain, int device, size_t n_samples,
                                       const double* x0, const double* x_goal, const double* u_bias, const double* gain,
                                       double* u_prev, const rkb_steer_opts* opts,
                                       const rkb_proxy* const* pairs, int n_pairs,
                                       double* x_out, int32_t* n_done, int32_t* collided, double* x_traj, int32_t* status,
                                       unsigned flags, void* stream);

/* Checked steering in ONE launch.  For a serial chain the library can compile its steering kernel with the collision
 * test of a given set of proxy pairs built in (the generated query of rkb_proxy_specialize inside the loop: integrate an
 * interval, test the state it ends on, accept or stop — no state round trips between four kernels per interval).
 * rkb_steer_checked_specialize does it now (NVRTC, a few seconds, cubin cached on disk); by default it happens in the
 * background once 4096 tuples have been steered with these pairs (in one call or over many; while every pair and the chain
 * have their AUTO_SPECIALIZE option on), calls made meanwhile run interval by interval.  Results: those of the interval-by-
 * interval path up to rounding.  RKB_ERR_UNSUPPORTED: interpreter chain, RKB_OPT_FUSED_STEER off, or no libnvrtc.so.12. */
RKB_API int rkb_steer_checked_specialize(rkb_chain* chain, int device, const rkb_proxy* const* pairs, int n_pairs);
RKB_API int rkb_steer_checked_is_specialized(rkb_chain* chain, const rkb_proxy* const* pairs, int n_pairs);
/* test hook: the CUDA source rkb_steer_checked_specialize compiles, its last line naming the kernel instantiated from it
 * ("// kernel: <name expression>"); returns its size including the NUL, out may be NULL to ask for the size */
RKB_API int rkb_steer_checked_source(rkb_chain* chain, const rkb_proxy* const* pairs, int n_pairs, char* out, size_t size);

/* Device-side timing of the last compute launch issued through `chain` on the calling
 * thread, in milliseconds (CUDA events on the launch stream); < 0 if none. Blocks until
 * that launch has finished. */
RKB_API double rkb_last_kernel_ms(rkb_chain* chain);
/* Number of kernels launched through this handle since creation. */
RKB_API uint64_t rkb_launch_count(const rkb_chain* chain);

/* Host buffers.  RKB_MEM_HOST pointers may be pageable (a std::vector's storage): the library then pays the
 * driver's synchronous staged copy.  rkb_host_pin page-locks a caller-owned range once (cudaHostRegister,
 * portable across devices) so that every later call on it copies by DMA, overlapped with the kernels;
 * rkb_host_unpin undoes it (before the memory is freed).  Pinning twice / unpinning unpinned memory is not an error.
 * Threading: a handle serialises its calls (internal mutex); RKB_MEM_DEVICE calls return without synchronising and
 * use per-handle scratch buffers, so a handle may have work in flight on ONE stream at a time — make a second
 * handle (rkb_chain_create is cheap) for a second concurrent stream. */
RKB_API int rkb_host_pin(void* ptr, size_t bytes);
RKB_API int rkb_host_unpin(void* ptr);

/* Instrumentation: best-of-runs throughput of a pure DFMA loop on `device` for about `seconds`
 * (TFLOP/s, 2 flops per DFMA) — the FP64 roofline denominator bench.py reports against. */
RKB_API int rkb_measure_fp64_peak(int device, double seconds, double* tflops_out, double* sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* REAK_B200_H */
