// rkb_types.h — internal structures shared by the host lowering (rkb_api.cu) and the kernels.
//
// A descriptor (include/reak_b200.h) is lowered once, at rkb_chain_create, into one or both of
//   * SerialParams  — the canonical serial-chain form consumed by the register-resident
//                     kernels in kte_serial.cuh (joint -> link -> inertia stages), passed BY VALUE
//                     as a __grid_constant__ kernel parameter so that every chain constant is a
//                     constant-bank operand of the FP64 instructions, and
//   * GenericProgram — the element list as is, consumed by the interpreter kernels in
//                     kte_generic.cuh (any element order, 2D chains, two-anchor springs/dampers).
#ifndef RKB_TYPES_H
#define RKB_TYPES_H

#include <stdint.h>
#include "../../include/reak_b200.h"

#define RKB_SERIAL_MAX_DOF 8

// stage flags (uniform across the grid, so branching on them never diverges)
#define RKB_ST_PRISMATIC 1u  // joint is prismatic_joint_3D instead of revolute_joint_3D
#define RKB_ST_LINKROT   2u  // rigid link carries a rotation (Ro != I)
#define RKB_ST_SPRING    4u  // torsion_spring_3D across the joint (anchors: joint base, joint end)
#define RKB_ST_DAMPER    8u  // torsion_damper_3D across the joint
#define RKB_ST_INERTIA   16u // an inertia_3D sits on the link end frame
#define RKB_ST_LINK      32u // a rigid_link_3D follows the joint (else end frame == joint end)
#define RKB_ST_SPRING_2D 64u // the spring is a torsion_spring_2D: plain wrapped angle, no axis_angle dead zone

// kernel-template feature mask: which optional code is compiled in
#define RKB_FL_PRISMATIC 1
#define RKB_FL_LINKROT   2
#define RKB_FL_SPRINGS   4
#define RKB_FL_ALL       7

// Structural promises, 8 bits per stage (stage k at bits [8k, 8k+8)): see kte_serial.cuh
//   bits 0-2  axis   0 = general (revolute with any axis, or prismatic), 1/2/3 = revolute about +-e_x/e_y/e_z,
//                    5/6/7 = prismatic along +-e_x/e_y/e_z
//   bits 3-4  link   0 = general (none / any offset / rotated), 1/2/3 = offset along e_x/e_y/e_z, no rotation
//   bit  5    inertia 0 = general (none / full tensor), 1 = present with a diagonal tensor
//   bits 6-7  sign of an axis-aligned joint axis: 0 = read at run time, 1 = +e_D, 3 = -e_D
#define RKB_SHAPE_STAGE(ax, lk, in) ((unsigned long long)((ax) | ((lk) << 3) | ((in) << 5)))
#define RKB_SHAPE_STAGE_SIGNED(ax, lk, in, sign) (RKB_SHAPE_STAGE(ax, lk, in) | ((unsigned long long)(sign) << 6))
#define RKB_SHAPE_AT(code, k) ((unsigned long long)(code) << (8 * (k)))

struct SerialStage {
  double ax[3];    // joint axis as stored (mAxis): used for angular/linear velocity terms and force projection
  double an[3];    // normalised axis (axis_angle ctor, rotations_3D.hpp:1962-1974): used for the rotation
  double aa[6];    // an an^T: xx, yy, zz, xy, xz, yz
  double po[3];    // rigid link offset position
  double Ro[9];    // rigid link offset rotation, row-major, v_base = Ro v_end
  double m;        // inertia_3D mass
  double I[6];     // inertia tensor xx, xy, xz, yy, yz, zz
  double rotor;    // sum of inertia_gen masses on this coordinate (diagonal of M)
  double mc;       // composite mass: sum of the inertia_3D masses of this and all later stages
  double mcpo[3];  // mc * po
  double ks, sat;  // torsion spring stiffness / saturation
  double cd;       // torsion damper coefficient
  uint32_t flags;
  int32_t  coord;  // state slot of this joint's coordinate (q at 2*coord, qd at 2*coord+1)
  int32_t  input;  // input index of the driving_actuator_gen on this joint, -1 if none
  int32_t  pad;
};

struct SerialParams {
  int32_t n;            // number of stages == number of coordinates
  int32_t n_inputs;
  double  w0[3];        // base frame AngVelocity (local)
  double  al0[3];       // base frame AngAcceleration (local)
  double  a0[3];        // base frame Acceleration rotated into base-local coordinates
  SerialStage st[RKB_SERIAL_MAX_DOF];
};

// strided view of a batch buffer: element k of sample i lives at p[i * si + k * sk].
// State buffers hold (q, q_dot) of coordinate c of an n-coordinate chain at elements 2c, 2c + 1
// (interleaved, kte_nl_system.hpp:189-193) or, when `blocked`, at elements c, n + c
// (manipulator_dynamics_model::computeStateRate, ctrl/mbd_kte/manipulator_model.cpp:292-355).
struct BatchView {
  double*  p;
  long long si, sk;
  int blocked;
};
struct ConstBatchView {
  const double* p;
  long long si, sk;
  int blocked;
};
#if defined(__CUDACC__)
__host__ __device__
#endif
inline long long rkb_state_q(int blocked, int n, int c) { (void)n; return blocked ? c : 2 * c; }
#if defined(__CUDACC__)
__host__ __device__
#endif
inline long long rkb_state_qd(int blocked, int n, int c) { return blocked ? n + c : 2 * c + 1; }

// Explicit one-step scheme in "start of step plus weighted increments" form: after evaluation s
// (k_s = dt f(X_s)) the next evaluation point — or, for s = stages - 1, the end of the step — is
// w + sum_{j <= s} c[s][j] k_j.  Filled by the host from the formulas of
// core/integrators/fixed_step_integrators.hpp (euler :64-84, midpoint :177-202, runge_kutta4 :257-293,
// runge_kutta5 :351-399).
#define RKB_RK_MAX_STAGES 6
struct RkTable {
  int32_t stages;
  int32_t pad;
  double  c[RKB_RK_MAX_STAGES][RKB_RK_MAX_STAGES];
};

// One control interval: n_steps integrator steps with the input held constant.  A sequence of
// intervals is a sequence of launches (the state makes one round trip through HBM per interval,
// 2 x 16 n bytes per sample against >= 4 evaluations of the chain).
struct RolloutArgs {
  ConstBatchView x0, u;
  BatchView      xout;     // may alias x0 (each thread reads its own row before it writes it)
  BatchView      traj;     // nullable (p == 0): a second copy of the end state (the interval's slot of x_traj)
  int32_t*       status;   // nullable
  long long      n_samples;
  long long      x0_div;   // sample i starts from row i / x0_div of x0 (steer batch: rollouts per pair; else 1)
  double         dt;
  int32_t        n_steps;
  int32_t        status_or;  // != 0: OR the status bits into status[i] instead of overwriting it
  const int32_t* active;     // nullable: samples with active[i] == 0 are left untouched (closed-loop steering)
  long long      u_node_stride;  // interpreter kernels, RK4: != 0: the input at every half step, node j at u.p[.. + j * u_node_stride]
};

// Rollout whose results go to several destination buffers at once (this GPU's own and, over NVLink, the peers'
// copies of the gathered batch): the all-gather of a sample-sharded job done by the stores of the kernel itself.
#define RKB_MAX_DEST 8
struct RolloutScatterArgs {
  ConstBatchView x0, u;
  double*   xout[RKB_MAX_DEST];    // AoS [n_total][2n] each; sample i of this launch is row row_offset + i
  int32_t*  status[RKB_MAX_DEST];  // [n_total] each (entries may be null)
  long long n_samples, row_offset;
  double    dt;
  int32_t   n_steps, n_dest, blocked, pad;
};

// A control sequence in one launch (serial kernels, RK4): n_intervals intervals of n_steps steps
struct RolloutSeqArgs {
  ConstBatchView x0, u;    // input k of interval j of sample i: u.p[i * u.si + j * u_sj + k * u.sk]
  BatchView      xout;
  BatchView      traj;     // nullable; slot j of sample i at traj.p + i * traj.si + j * traj_sj
  int32_t*       status;   // nullable
  long long      n_samples, u_sj, traj_sj;
  double         dt;
  int32_t        n_steps, n_intervals;
  int32_t        half_step_nodes;  // != 0: ONE interval, u holds the input at every half step (2 n_steps + 1 nodes, stride u_sj)
  int32_t        pad;
};

// The whole steering loop in one launch (serial kernels), all buffers device-resident AoS
struct SteerArgs {
  const double* x0;      // [N][nx]
  const double* goal;    // [N][nx]
  const double* u_bias;  // [N][nu]
  const double* gain;    // [N][nu][nx]
  double*       u_prev;  // [N][nu] in / out
  double*       xout;    // [N][nx]
  double*       traj;    // nullable, [N][max_intervals][nx]
  int32_t*      n_done;  // [N]
  int32_t*      status;  // nullable
  long long     n_samples;
  int32_t       nu, max_intervals, substeps, saturate_first, have_u_box, have_du_box, blocked, pad;
  double        time_step, dt, proximity;
  double        u_lo[RKB_MAX_COORDS], u_hi[RKB_MAX_COORDS], du_lo[RKB_MAX_COORDS], du_hi[RKB_MAX_COORDS];
  int32_t*      collided;  // [N], kernels generated with a collision test only (else unused): 1 when the loop stopped on a state that was not free
};

// One pass of the steering loop head (rkb_steer.cu), all buffers device-resident AoS
struct SteerLawArgs {
  const double* x0;      // [N][nx] start states (read at interval 0)
  double*       x;       // [N][nx] working state, advanced in place by the rollout kernel
  const double* goal;    // [N][nx]
  const double* u_bias;  // [N][nu]
  const double* gain;    // [N][nu][nx]
  double*       u_prev;  // [N][nu] in: previous input; out: input of the next interval
  int32_t*      n_done;  // [N] intervals performed so far
  int32_t*      active;  // [N] 1 when the next interval is to be integrated
  long long     n_samples;
  int32_t       nx, nu, interval, saturate_first, have_u_box, have_du_box;
  double        time_step, proximity;
  double        u_lo[RKB_MAX_COORDS], u_hi[RKB_MAX_COORDS], du_lo[RKB_MAX_COORDS], du_hi[RKB_MAX_COORDS];
  double*       u_next;  // nullable: where the interval's input goes when acceptance is decided later (collision check)
};

// acceptance of a control interval after its collision test (MEAQR_topology.hpp:550-559)
struct SteerCommitArgs {
  double*       x;        // [N][nx] accepted state
  const double* x_next;   // [N][nx] state the interval ended on
  double*       u_prev;   // [N][nu]
  const double* u_next;   // [N][nu] input the interval was integrated with
  double*       traj;     // nullable: steer record, [N][J][nx]
  const double* dist;     // [n_pairs][N] minimum distance of every proxy pair at x_next
  int32_t*      n_done;
  int32_t*      active;
  int32_t*      collided; // [N] set when the loop ended on a state that was not free
  long long     n_samples;
  int32_t       nx, nu, interval, max_intervals, n_pairs, pad;
};

struct EvalArgs {
  ConstBatchView x, u;
  BatchView      out;      // xdot (2n), f (n) or M (n*n) depending on the kernel
  BatchView      out2;     // Mdot (n*n) for the mass kernel, else unused
  int32_t*       status;   // nullable
  long long      n_samples;
};

// ---- proximity models (kte_proximity.cuh) ------------------------------------------------------
#define RKB_PROX_MAX_SHAPES 16   // per model
struct ProxShape {
  int32_t kind;      // rkb_shape_kind
  int32_t anchor;    // chain frame id, -1 = world
  double  pos[3], quat[4];
  double  dims[3];
  double  brad;      // shape_3D::getBoundingRadius
  double  rot[9];    // world-fixed shapes (anchor < 0): the rotation table of kte_proximity.cuh, formed once on the host
};
struct ProxProgram {
  int32_t   n1, n2;
  // Frames the kernel has to keep after the forward kinematics has moved on: those a shape rides on and those an
  // element reads that is not the one right after their writer.  All other frames only ever live in registers.
  int32_t   n_slots;
  int8_t    slot_of[40];                 // per chain frame (RKB_GEN_MAX_FRAMES): storage slot, -1 = never stored
  int32_t   pad;
  ProxShape s[2 * RKB_PROX_MAX_SHAPES];  // model 1 then model 2
};

// ---- generic interpreter program ------------------------------------------------------------
#define RKB_GEN_MAX_FRAMES 40
#define RKB_GEN_MAX_ELEMENTS 96

// free_joint_3D (free_joints.cpp:123-208): at most RKB_GEN_MAX_FREE per chain, each adding 13 states (position,
// quaternion, velocity, angular velocity of the coordinate frame) and 6 accelerations after the generalized coordinates'
// (kte_nl_system.hpp:145-147, 205-219, 293-308).  Interpreter kernels only.
#define RKB_GEN_MAX_FREE 1
#define RKB_GEN_MAX_ACC (RKB_MAX_COORDS + 6 * RKB_GEN_MAX_FREE)
#define RKB_GEN_MAX_STATE (2 * RKB_MAX_COORDS + 13 * RKB_GEN_MAX_FREE)
#define RKB_GEN_FREE_BIT 16  // GenericElement::upstream: bit c = coordinate c, bit 16 + i = free joint i

struct GenericElement {
  int32_t kind, fa, fb, coord, aux, pad;
  uint32_t upstream;
  int32_t  row;      // first row of this inertia in the twist-shaping matrix (rkb_twist_shaping)
  double   p[12];    // as rkb_element::p; rigid_link_3D: p[3..6] normalised quaternion
};

struct GenericProgram {
  int32_t dim, n_elements, n_frames, n_coords, n_inputs, base_frame;
  int32_t n_free, free_elem[RKB_GEN_MAX_FREE];  // free joints and the element index of each
  int32_t free_states, free_acc;                // per free joint: 13 / 6 (free_joint_3D), 7 / 3 (free_joint_2D)
  int32_t n_aux;  // auxiliary gen_coords (RKB_COORD_GEN), indexed n_coords .. n_coords + n_aux - 1 like the coordinates
  double  aux_q[RKB_MAX_COORDS], aux_qd[RKB_MAX_COORDS];  // the values they hold unless a rigid_link_gen writes them
  double  base[19];  // p3 q4 v3 w3 a3 al3 (2D: p2, -, cos, sin, -, -, v2, -, w, -, -, a2, -, al)
  int32_t jelem[RKB_MAX_COORDS];  // element index of the joint that owns each coordinate
  GenericElement el[RKB_GEN_MAX_ELEMENTS];
};

#endif
