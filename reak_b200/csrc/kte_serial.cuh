// kte_serial.cuh — register-resident FP64 kernels for canonical serial KTE chains (sm_100a).
//
// One thread integrates one sample.  All per-sample quantities live in registers (plus a
// per-thread shared-memory column for the few arrays that must survive between sweeps); every
// chain constant is read from the __grid_constant__ SerialParams kernel parameter, i.e. it is
// a constant-bank operand of the DFMA/DMUL that uses it and costs no register.  The stage
// loops are fully unrolled at compile time (template<int N>).
//
// What is computed is exactly what ReaK's kte_nl_system::get_state_derivative computes
// (ctrl/ctrl_sys/kte_nl_system.hpp:238-346) for a chain of
//   [driving_actuator_gen] [inertia_gen] joint [torsion_spring_3D] [torsion_damper_3D]
//   [rigid_link_3D] [inertia_3D]
// stages — but formulated in link-local coordinates so that no global frame is ever built:
//
//  sweep 1 (doMotion, kte_map_chain.hpp:71-76):  w, al (local angular velocity/acceleration)
//      and a (linear acceleration rotated into the local frame) are pushed through each joint
//      (revolute_joint.cpp:121-152 / prismatic_joint.cpp:129-161) and link
//      (rigid_link.cpp:156 -> frame_3D.hpp:236-251); the d'Alembert wrench of the stage's
//      inertia_3D (inertia.cpp:111-121) is formed on the spot and parked in shared memory.
//  sweep 2 (doForce, kte_map_chain.hpp:78-83, reverse order): wrenches are shifted back through
//      links (rigid_link.cpp:170-177) and joints (revolute_joint.cpp:172-184,
//      prismatic_joint.cpp:181-193) INCLUDING the reference's removal of the axial component
//      and the actuator reaction (driving_actuator.cpp:31-38, revolute_joint.cpp:210-213),
//      which makes gen_coord::f differ from the textbook tau - h whenever joint axes are not
//      orthogonal; torsion springs/dampers across a joint (torsion_spring.cpp:106-129,
//      torsion_damper.cpp:93-104) reduce to scalars along the joint axis.
//  sweep 3 (mass_matrix_calc::getMassMatrix, mass_matrix_calculator.cpp:80-87,100-287):
//      M = Tcm^T Mcm Tcm, whose twist-shaping columns the reference re-derives from relative
//      frames per (joint, inertia) pair (jacobian_gen_3D::get_jac_relative_to,
//      motion_jacobians.hpp:238-251), is regrouped by composite inertias: one inward pass carries
//      the composite (mass, first moment, tensor) of everything beyond a joint, and each column
//      C_k S_k is carried inward and projected on the axes below; rotor inertias (inertia_gen)
//      land on the diagonal.  Same matrix, O(n^2) small-vector work instead of O(n^2) frame algebra.
//  solve  (linsolve_Cholesky, core/lin_alg/mat_cholesky.hpp:63-84,160-179) with the reference's
//      "pivot < 1e-8 before the square root" singularity test reported in the status word.
//  RK4    (runge_kutta4_integrator<T>::integrate, core/integrators/fixed_step_integrators.hpp:256-293),
//      constant input over the rollout (num_int_dtnl_sys::get_next_state, num_int_dtnl_system.hpp:166-180).
//
// Structural specialisation.  The third template argument SHAPE carries 8 bits per stage
// (rkb_types.h, RKB_SHAPE_*) that promise structure the lowering found in the descriptor: a
// revolute axis that is +-e_x/e_y/e_z, a link offset along one coordinate axis without rotation,
// a diagonal inertia tensor.  A promised stage uses the 4-multiply planar rotation, the
// 2-multiply axis cross products and the 3-multiply tensor product instead of the dense 3x3
// forms; SHAPE = 0 is the fully general code.  Results are the same up to rounding.
#ifndef RKB_KTE_SERIAL_CUH
#define RKB_KTE_SERIAL_CUH

#include <cuda_runtime.h>
#include <math.h>
#include "rkb_types.h"

namespace rkb {

#define RKB_DEV __device__ __forceinline__

typedef unsigned long long shape_t;

struct vec3 { double c[3]; };

RKB_DEV vec3 mk(double x, double y, double z) { vec3 r; r.c[0] = x; r.c[1] = y; r.c[2] = z; return r; }
RKB_DEV vec3 operator+(vec3 a, vec3 b) { return mk(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
RKB_DEV vec3 operator-(vec3 a, vec3 b) { return mk(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]); }
RKB_DEV vec3 operator*(double s, vec3 a) { return mk(s * a.c[0], s * a.c[1], s * a.c[2]); }
RKB_DEV double dot(vec3 a, vec3 b) { return a.c[0] * b.c[0] + a.c[1] * b.c[1] + a.c[2] * b.c[2]; }
RKB_DEV vec3 cross(vec3 a, vec3 b) {
  return mk(a.c[1] * b.c[2] - a.c[2] * b.c[1], a.c[2] * b.c[0] - a.c[0] * b.c[2], a.c[0] * b.c[1] - a.c[1] * b.c[0]);
}
RKB_DEV vec3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }

// 3x3 rotation held row-major: v_parent = R v_child
struct mat3 { double m[9]; };
RKB_DEV vec3 mul(const mat3& R, vec3 v) {
  return mk(R.m[0] * v.c[0] + R.m[1] * v.c[1] + R.m[2] * v.c[2],
            R.m[3] * v.c[0] + R.m[4] * v.c[1] + R.m[5] * v.c[2],
            R.m[6] * v.c[0] + R.m[7] * v.c[1] + R.m[8] * v.c[2]);
}
RKB_DEV vec3 tmul(const mat3& R, vec3 v) {  // R^T v
  return mk(R.m[0] * v.c[0] + R.m[3] * v.c[1] + R.m[6] * v.c[2],
            R.m[1] * v.c[0] + R.m[4] * v.c[1] + R.m[7] * v.c[2],
            R.m[2] * v.c[0] + R.m[5] * v.c[1] + R.m[8] * v.c[2]);
}
RKB_DEV vec3 mulc(const double* R, vec3 v) {  // constant-bank matrix (row-major) times v
  return mk(R[0] * v.c[0] + R[1] * v.c[1] + R[2] * v.c[2], R[3] * v.c[0] + R[4] * v.c[1] + R[5] * v.c[2],
            R[6] * v.c[0] + R[7] * v.c[1] + R[8] * v.c[2]);
}
RKB_DEV vec3 tmulc(const double* R, vec3 v) {
  return mk(R[0] * v.c[0] + R[3] * v.c[1] + R[6] * v.c[2], R[1] * v.c[0] + R[4] * v.c[1] + R[7] * v.c[2],
            R[2] * v.c[0] + R[5] * v.c[1] + R[8] * v.c[2]);
}
RKB_DEV vec3 symmul(const double* I, vec3 v) {  // I = xx xy xz yy yz zz
  return mk(I[0] * v.c[0] + I[1] * v.c[1] + I[2] * v.c[2], I[1] * v.c[0] + I[3] * v.c[1] + I[4] * v.c[2],
            I[2] * v.c[0] + I[4] * v.c[1] + I[5] * v.c[2]);
}
RKB_DEV vec3 diagmul(const double* I, vec3 v) { return mk(I[0] * v.c[0], I[3] * v.c[1], I[5] * v.c[2]); }

// axis_angle(q, axis).getRotMat(), rotations_3D.hpp:2159-2178, from cos/sin of the joint angle
RKB_DEV mat3 rodrigues(const SerialStage& S, double c, double s) {
  const double omc = 1.0 - c;
  const double t12 = omc * S.aa[3], t13 = omc * S.aa[4], t23 = omc * S.aa[5];
  const double t01 = s * S.an[0], t02 = s * S.an[1], t03 = s * S.an[2];
  mat3 R;
  R.m[0] = c + omc * S.aa[0]; R.m[1] = t12 - t03;         R.m[2] = t13 + t02;
  R.m[3] = t12 + t03;         R.m[4] = c + omc * S.aa[1]; R.m[5] = t23 - t01;
  R.m[6] = t13 - t02;         R.m[7] = t23 + t01;         R.m[8] = c + omc * S.aa[2];
  return R;
}

// Rotation by (c, s) about coordinate axis D (0 = x, 1 = y, 2 = z); D1, D2 follow D cyclically.
// R_D v and R_D^T v touch only the two off-axis components: 4 multiplies instead of 9.
template <int D>
RKB_DEV vec3 rot_axis(double c, double s, vec3 v) {
  constexpr int D1 = (D + 1) % 3, D2 = (D + 2) % 3;
  vec3 r;
  r.c[D] = v.c[D];
  r.c[D1] = c * v.c[D1] - s * v.c[D2];
  r.c[D2] = s * v.c[D1] + c * v.c[D2];
  return r;
}
template <int D>
RKB_DEV vec3 rotT_axis(double c, double s, vec3 v) {
  constexpr int D1 = (D + 1) % 3, D2 = (D + 2) % 3;
  vec3 r;
  r.c[D] = v.c[D];
  r.c[D1] = c * v.c[D1] + s * v.c[D2];
  r.c[D2] = c * v.c[D2] - s * v.c[D1];
  return r;
}

// Angle of axis_angle(conj(Q_base) * Q_end) times its axis, expressed as a signed scalar along the
// normalised joint axis (rotations_3D.hpp:1985-2010): the angle wrapped to (-pi, pi], zero inside
// the reference's |sin(q/2)| <= 1e-7 dead zone.
RKB_DEV double wrapped_angle(double q) {
  const double two_pi = 6.283185307179586476925286766559;
  double r = q - two_pi * rint(q * 0.15915494309189533576888376337251);
  if (fabs(r) <= 2.0e-7) r = 0.0;
  return r;
}

// torsion spring (+ saturation) and damper across a revolute joint, as a signed scalar along the axis.
// torsion_spring_3D goes through axis_angle (dead zone, torsion_spring.cpp:106-129); torsion_spring_2D takes
// the angle of the relative rotation as it is (atan2, torsion_spring.cpp:50-71).
RKB_DEV double spring_scalar(const SerialStage& S, double q) {
  double r;
  if (S.flags & RKB_ST_SPRING_2D) {
    const double two_pi = 6.283185307179586476925286766559;
    r = q - two_pi * rint(q * 0.15915494309189533576888376337251);
  } else {
    r = wrapped_angle(q);
  }
  double mag = S.ks * fabs(r);  // stiffness * angle_diff.angle(), angle >= 0
  if (S.sat > 0.0 && fabs(mag) > S.sat) mag = (mag > 0.0) ? S.sat : -S.sat;
  return r < 0.0 ? -mag : mag;
}

// ---- symmetric 3x3 tensors held as xx xy xz yy yz zz ---------------------------------------------
__host__ __device__ constexpr int sym_idx(int i, int j) {
  return (i <= j) ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : (j == 0 ? i : (j == 1 ? 2 + i : 5));
}
// ---- sine and cosine of the joint angles ---------------------------------------------------------
// The CUDA library's sincos() spends more issue slots on loading its polynomial coefficients
// (two UMOV per 64-bit immediate), on F2I/I2F and on its slow-path branch than on arithmetic, and the
// branch keeps the compiler from interleaving the six independent evaluations: ncu attributed 29 % of
// the rollout kernel's stall samples to it.  This version has the same structure — three-term
// Cody-Waite reduction by pi/2 (exact for |q| < 2^30), degree-13/12 polynomials on [-pi/4, pi/4],
// error below 1.6 ulp — but rounds with the 1.5*2^52 trick instead of F2I/I2F and has no branch,
// so the six evaluations of a stage interleave.  Arguments outside the reduction's range (and NaN)
// are left to sincos() by the caller.
// sin(r) = r + r z (S1 + z (S2 + ... z S6)),  cos(r) = 1 - z/2 + z^2 (C1 + z (C2 + ... z C6)),  z = r^2, |r| <= pi/4.
// The coefficients are literals on purpose: FP64 instructions take constants only through uniform
// registers, of which there are 63; a table would be hoisted out of the rollout loop into 32 of them
// and push chain constants out (R2UR.FILL traffic), whereas literals are re-materialised by UMOVs on
// the uniform datapath, which costs no issue slot of the vector pipes.
#define RKB_S1 -1.66666666666666324348e-01
#define RKB_S2 8.33333333332248946124e-03
#define RKB_S3 -1.98412698298579493134e-04
#define RKB_S4 2.75573137070700676789e-06
#define RKB_S5 -2.50507602534068634195e-08
#define RKB_S6 1.58969099521155010221e-10
#define RKB_C1 4.16666666666666019037e-02
#define RKB_C2 -1.38888888888741095749e-03
#define RKB_C3 2.48015872894767294178e-05
#define RKB_C4 -2.75573143513906633035e-07
#define RKB_C5 2.08757232129817482790e-09
#define RKB_C6 -1.13596475577881948265e-11
#define RKB_2_OVER_PI 6.36619772367581382433e-01
#define RKB_PIO2_HI 1.570796326794896557999e+00   // pi/2 split in three doubles
#define RKB_PIO2_MID 6.123233995736766035869e-17
#define RKB_PIO2_LO -1.497384904859169832944e-33

#define RKB_SINCOS_MAX 1.0e9  // |q| below this: k = rint(q 2/pi) fits 31 bits and the reduction is accurate

RKB_DEV void sincos_reduced(double q, double& sn, double& cs) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
  const double t = fma(q, RKB_2_OVER_PI, magic);
  const int k = __double2loint(t);          // low mantissa word = the integer, two's complement
  const double kd = t - magic;
  double r = fma(-kd, RKB_PIO2_HI, q);
  r = fma(-kd, RKB_PIO2_MID, r);
  r = fma(-kd, RKB_PIO2_LO, r);
  const double z = r * r;
  double ps = fma(z, RKB_S6, RKB_S5);
  double pc = fma(z, RKB_C6, RKB_C5);
  ps = fma(z, ps, RKB_S4); pc = fma(z, pc, RKB_C4);
  ps = fma(z, ps, RKB_S3); pc = fma(z, pc, RKB_C3);
  ps = fma(z, ps, RKB_S2); pc = fma(z, pc, RKB_C2);
  ps = fma(z, ps, RKB_S1); pc = fma(z, pc, RKB_C1);
  const double s0 = fma(r * z, ps, r);
  const double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));
  // quadrant k mod 4: (sin, cos) = (s0, c0), (c0, -s0), (-s0, -c0), (-c0, s0)
  double s = (k & 1) ? c0 : s0;
  double c = (k & 1) ? s0 : c0;
  if (k & 2) s = -s;
  if ((k + 1) & 2) c = -c;
  sn = s; cs = c;
}

// ---- forward-mode scalars for the mass-matrix sweep ------------------------------------------------
// mass_sweep<..., T> below runs on T = double (M only) or T = dual (value and time derivative: M and
// Mdot = d/dt M along q_dot, what mass_matrix_calc::getMassMatrixAndDerivative assembles from Tcm_dot,
// mass_matrix_calculator.cpp:89-98).  Chain constants stay plain doubles, so a constant times a dual
// costs two multiplies, not three.
struct dual { double v, d; };
RKB_DEV dual mkd(double v, double d) { dual r; r.v = v; r.d = d; return r; }
RKB_DEV dual operator+(dual a, dual b) { return mkd(a.v + b.v, a.d + b.d); }
RKB_DEV dual operator-(dual a, dual b) { return mkd(a.v - b.v, a.d - b.d); }
RKB_DEV dual operator-(dual a) { return mkd(-a.v, -a.d); }
RKB_DEV dual operator*(dual a, dual b) { return mkd(a.v * b.v, a.v * b.d + a.d * b.v); }
RKB_DEV dual operator*(double a, dual b) { return mkd(a * b.v, a * b.d); }
RKB_DEV dual operator*(dual a, double b) { return mkd(a.v * b, a.d * b); }
RKB_DEV dual operator+(dual a, double b) { return mkd(a.v + b, a.d); }
RKB_DEV dual operator+(double a, dual b) { return mkd(a + b.v, b.d); }
RKB_DEV dual operator-(dual a, double b) { return mkd(a.v - b, a.d); }
RKB_DEV dual operator-(double a, dual b) { return mkd(a - b.v, -b.d); }
RKB_DEV dual& operator+=(dual& a, dual b) { a = a + b; return a; }
RKB_DEV dual& operator-=(dual& a, dual b) { a = a - b; return a; }
RKB_DEV dual& operator+=(dual& a, double b) { a.v += b; return a; }
RKB_DEV double value_of(double a) { return a; }
RKB_DEV double value_of(dual a) { return a.v; }
RKB_DEV double deriv_of(double) { return 0.0; }
RKB_DEV double deriv_of(dual a) { return a.d; }
template <class T> RKB_DEV T zero_of();
template <> RKB_DEV double zero_of<double>() { return 0.0; }
template <> RKB_DEV dual zero_of<dual>() { return mkd(0.0, 0.0); }

template <class T> struct vec3t { T c[3]; };
template <class T> RKB_DEV vec3t<T> mkt(T x, T y, T z) { vec3t<T> r; r.c[0] = x; r.c[1] = y; r.c[2] = z; return r; }
template <class T> RKB_DEV vec3t<T> operator+(vec3t<T> a, vec3t<T> b) { return mkt<T>(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
template <class T> RKB_DEV vec3t<T> cross_cv(vec3 a, vec3t<T> b) {  // constant x variable
  return mkt<T>(a.c[1] * b.c[2] - a.c[2] * b.c[1], a.c[2] * b.c[0] - a.c[0] * b.c[2], a.c[0] * b.c[1] - a.c[1] * b.c[0]);
}
template <class T> RKB_DEV vec3t<T> cross_vc(vec3t<T> a, vec3 b) {  // variable x constant
  return mkt<T>(a.c[1] * b.c[2] - a.c[2] * b.c[1], a.c[2] * b.c[0] - a.c[0] * b.c[2], a.c[0] * b.c[1] - a.c[1] * b.c[0]);
}
template <class T> RKB_DEV vec3t<T> cross_vv(vec3t<T> a, vec3t<T> b) {
  return mkt<T>(a.c[1] * b.c[2] - a.c[2] * b.c[1], a.c[2] * b.c[0] - a.c[0] * b.c[2], a.c[0] * b.c[1] - a.c[1] * b.c[0]);
}
template <class T> RKB_DEV T dot_cv(vec3 a, vec3t<T> b) { return a.c[0] * b.c[0] + a.c[1] * b.c[1] + a.c[2] * b.c[2]; }
template <class T> RKB_DEV T dot_vv(vec3t<T> a, vec3t<T> b) { return a.c[0] * b.c[0] + a.c[1] * b.c[1] + a.c[2] * b.c[2]; }
template <class T> struct mat3t { T m[9]; };
template <class T> RKB_DEV vec3t<T> mul(const mat3t<T>& R, vec3t<T> v) {
  return mkt<T>(R.m[0] * v.c[0] + R.m[1] * v.c[1] + R.m[2] * v.c[2], R.m[3] * v.c[0] + R.m[4] * v.c[1] + R.m[5] * v.c[2],
                R.m[6] * v.c[0] + R.m[7] * v.c[1] + R.m[8] * v.c[2]);
}
template <class T> RKB_DEV vec3t<T> mulc_t(const double* R, vec3t<T> v) {
  return mkt<T>(R[0] * v.c[0] + R[1] * v.c[1] + R[2] * v.c[2], R[3] * v.c[0] + R[4] * v.c[1] + R[5] * v.c[2],
                R[6] * v.c[0] + R[7] * v.c[1] + R[8] * v.c[2]);
}
template <class T> RKB_DEV vec3t<T> symmul_vc(const T (&I)[6], vec3 v) {  // variable tensor times constant vector
  return mkt<T>(I[0] * v.c[0] + I[1] * v.c[1] + I[2] * v.c[2], I[1] * v.c[0] + I[3] * v.c[1] + I[4] * v.c[2],
                I[2] * v.c[0] + I[4] * v.c[1] + I[5] * v.c[2]);
}
template <class T> RKB_DEV mat3t<T> rodrigues_t(const SerialStage& S, T c, T s) {
  const T omc = 1.0 - c;
  const T t12 = omc * S.aa[3], t13 = omc * S.aa[4], t23 = omc * S.aa[5];
  const T t01 = s * S.an[0], t02 = s * S.an[1], t03 = s * S.an[2];
  mat3t<T> R;
  R.m[0] = c + omc * S.aa[0]; R.m[1] = t12 - t03;         R.m[2] = t13 + t02;
  R.m[3] = t12 + t03;         R.m[4] = c + omc * S.aa[1]; R.m[5] = t23 - t01;
  R.m[6] = t13 - t02;         R.m[7] = t23 + t01;         R.m[8] = c + omc * S.aa[2];
  return R;
}
template <int D, class T>
RKB_DEV vec3t<T> rot_axis_t(T c, T s, vec3t<T> v) {
  constexpr int D1 = (D + 1) % 3, D2 = (D + 2) % 3;
  vec3t<T> r;
  r.c[D] = v.c[D];
  r.c[D1] = c * v.c[D1] - s * v.c[D2];
  r.c[D2] = s * v.c[D1] + c * v.c[D2];
  return r;
}
// I <- R I R^T for a rotation by (c, s) about coordinate axis D: 18 multiplies
template <int D, class T>
RKB_DEV void sym_rotate_axis_t(T c, T s, T (&I)[6]) {
  constexpr int p = (D + 1) % 3, q = (D + 2) % 3;
  const T ipr = I[sym_idx(p, D)], iqr = I[sym_idx(q, D)];
  I[sym_idx(p, D)] = c * ipr - s * iqr;
  I[sym_idx(q, D)] = s * ipr + c * iqr;
  const T ipp = I[sym_idx(p, p)], ipq = I[sym_idx(p, q)], iqq = I[sym_idx(q, q)];
  const T app = c * ipp - s * ipq, apq = c * ipq - s * iqq;  // A = R I (2x2 block)
  const T aqp = s * ipp + c * ipq, aqq = s * ipq + c * iqq;
  I[sym_idx(p, p)] = c * app - s * apq;
  I[sym_idx(p, q)] = s * app + c * apq;
  I[sym_idx(q, q)] = s * aqp + c * aqq;
}
// I <- R I R^T for a general rotation
template <class T>
RKB_DEV void sym_rotate_t(const mat3t<T>& R, T (&I)[6]) {
  T A[9];  // A = R I
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      A[3 * i + j] = R.m[3 * i] * I[sym_idx(0, j)] + R.m[3 * i + 1] * I[sym_idx(1, j)] + R.m[3 * i + 2] * I[sym_idx(2, j)];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      I[sym_idx(i, j)] = A[3 * i] * R.m[3 * j] + A[3 * i + 1] * R.m[3 * j + 1] + A[3 * i + 2] * R.m[3 * j + 2];
}
template <class T>
RKB_DEV void sym_rotate_c(const double* R, T (&I)[6]) {  // constant rotation (row-major)
  T A[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      A[3 * i + j] = R[3 * i] * I[sym_idx(0, j)] + R[3 * i + 1] * I[sym_idx(1, j)] + R[3 * i + 2] * I[sym_idx(2, j)];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      I[sym_idx(i, j)] = A[3 * i] * R[3 * j] + A[3 * i + 1] * R[3 * j + 1] + A[3 * i + 2] * R[3 * j + 2];
}
// Parallel-axis move of a composite (mass mc, first moment h, tensor I): every mass point goes from
// r to r + p, p constant, mcp = mc * p.  With w = h + mcp / 2:  I += 2 (w.p) 1 - w p^T - p w^T,  h += mcp.
template <class T>
RKB_DEV void sym_shift_c(vec3 p, vec3 mcp, vec3t<T>& h, T (&I)[6]) {
  const vec3t<T> w = mkt<T>(h.c[0] + 0.5 * mcp.c[0], h.c[1] + 0.5 * mcp.c[1], h.c[2] + 0.5 * mcp.c[2]);
  const T d = 2.0 * dot_cv(p, w);
  I[0] += d - 2.0 * (w.c[0] * p.c[0]);
  I[3] += d - 2.0 * (w.c[1] * p.c[1]);
  I[5] += d - 2.0 * (w.c[2] * p.c[2]);
  I[1] -= w.c[0] * p.c[1] + w.c[1] * p.c[0];
  I[2] -= w.c[0] * p.c[2] + w.c[2] * p.c[0];
  I[4] -= w.c[1] * p.c[2] + w.c[2] * p.c[1];
  h.c[0] += mcp.c[0]; h.c[1] += mcp.c[1]; h.c[2] += mcp.c[2];
}
// ... and by a variable offset r = q * axis of a prismatic joint (mc constant)
template <class T>
RKB_DEV void sym_shift_v(vec3t<T> p, double mc, vec3t<T>& h, T (&I)[6]) {
  const vec3t<T> mcp = mkt<T>(mc * p.c[0], mc * p.c[1], mc * p.c[2]);
  const vec3t<T> w = mkt<T>(h.c[0] + 0.5 * mcp.c[0], h.c[1] + 0.5 * mcp.c[1], h.c[2] + 0.5 * mcp.c[2]);
  const T d = 2.0 * dot_vv(w, p);
  I[0] += d - 2.0 * (w.c[0] * p.c[0]);
  I[3] += d - 2.0 * (w.c[1] * p.c[1]);
  I[5] += d - 2.0 * (w.c[2] * p.c[2]);
  I[1] -= w.c[0] * p.c[1] + p.c[0] * w.c[1];
  I[2] -= w.c[0] * p.c[2] + p.c[0] * w.c[2];
  I[4] -= w.c[1] * p.c[2] + p.c[1] * w.c[2];
  h = h + mcp;
}

template <int N>
struct SerialState {
  double q[N], qd[N], u[N];
};

__host__ __device__ constexpr int shape_axraw(shape_t s, int k) { return (int)((s >> (8 * k)) & 7u); }      // 0 general, 1..3 revolute about x,y,z, 5..7 prismatic along x,y,z
__host__ __device__ constexpr int shape_ax(shape_t s, int k) { return shape_axraw(s, k) <= 3 ? shape_axraw(s, k) : 0; }      // revolute about e_D: D + 1, else 0
__host__ __device__ constexpr int shape_px(shape_t s, int k) { return shape_axraw(s, k) >= 5 ? shape_axraw(s, k) - 4 : 0; }  // prismatic along e_D: D + 1, else 0
__host__ __device__ constexpr int shape_lk(shape_t s, int k) { return (int)((s >> (8 * k + 3)) & 3u); }     // 0 general, 1..3 offset along x,y,z, no rotation
__host__ __device__ constexpr int shape_in(shape_t s, int k) { return (int)((s >> (8 * k + 5)) & 1u); }     // 0 general, 1 diagonal tensor present
__host__ __device__ constexpr int shape_sign(shape_t s, int k) { return (int)((s >> (8 * k + 6)) & 3u); }   // 0 read the sign at run time, 1 axis = +e_D, 3 axis = -e_D
// sign of an axis-aligned joint axis: a literal when the shape promises it (the multiplications by it fold away)
#define RKB_AXIS_SIGN(SH, k, runtime) (shape_sign(SH, k) == 1 ? 1.0 : (shape_sign(SH, k) == 3 ? -1.0 : (runtime)))

// cos and sin of every revolute joint angle (1, 0 for a prismatic joint); the sign of an axis-aligned
// joint's axis is folded into the sine, which is all the specialised rotations need.
template <int N, int FL, shape_t SHAPE>
RKB_DEV void serial_trig_raw(const SerialParams& P, const SerialState<N>& X, double (&cs)[N], double (&sn)[N]) {
  bool in_range = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const SerialStage& S = P.st[k];
    constexpr shape_t SH = SHAPE;
    const int AX = shape_ax(SH, k);
    const bool prismatic = AX == 0 && (FL & RKB_FL_PRISMATIC) && (S.flags & RKB_ST_PRISMATIC);
    if (!prismatic) {
      sincos_reduced(X.q[k], sn[k], cs[k]);
      in_range = in_range && (fabs(X.q[k]) < RKB_SINCOS_MAX);
    } else { sn[k] = 0.0; cs[k] = 1.0; }
  }
  if (!in_range) {  // a huge or non-finite angle somewhere in this sample: library path, never taken in practice
#pragma unroll
    for (int k = 0; k < N; ++k) {  // (static indices: a rolled loop would push q, sn, cs into local memory)
      const bool prismatic = (FL & RKB_FL_PRISMATIC) && (P.st[k].flags & RKB_ST_PRISMATIC);
      if (!prismatic) sincos(X.q[k], &sn[k], &cs[k]);
    }
  }
}
template <int N, int FL, shape_t SHAPE>
RKB_DEV void serial_trig_fold(const SerialParams& P, double (&sn)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    constexpr shape_t SH = SHAPE;
    const int AX = shape_ax(SH, k);
    if (AX != 0) sn[k] *= RKB_AXIS_SIGN(SH, k, P.st[k].ax[AX - 1]);  // axis = +-e_D: fold the sign into the sine
  }
}
template <int N, int FL, shape_t SHAPE>
RKB_DEV void serial_trig(const SerialParams& P, const SerialState<N>& X, double (&cs)[N], double (&sn)[N]) {
  serial_trig_raw<N, FL, SHAPE>(P, X, cs, sn);
  serial_trig_fold<N, FL, SHAPE>(P, sn);
}

// sin and cos of w + d from those of w for a small d (|d| < 2^-5): sin d and cos d - 1 from their Taylor
// polynomials (truncation below 3e-17 relative), then the angle-sum formulas written as corrections to
// the base values.  14 FP64 instructions and nothing else, against 21 plus the quadrant logic of a full
// evaluation; the RK4 stages 2-4 of a step differ from its start by dt q_dot / 2 or dt q_dot.
#define RKB_SMALL_ANGLE 0.03125
RKB_DEV void sincos_shift(double d, double sw, double cw, double& sn, double& cs) {
  const double z = d * d;
  double ps = fma(z, -1.98412698412698412698e-04, 8.33333333333333333333e-03);
  ps = fma(z, ps, -1.66666666666666666667e-01);
  const double sd = fma(d * z, ps, d);                      // sin d
  double pc = fma(z, -1.38888888888888888889e-03, 4.16666666666666666667e-02);
  pc = fma(z, pc, -0.5);
  const double cm = z * pc;                                 // cos d - 1
  sn = sw + fma(cw, sd, sw * cm);
  cs = cw + fma(-sw, sd, cw * cm);
}

// ---- sweep 3: mass matrix by composite inertias, inward -------------------------------------------
// M = Tcm^T Mcm Tcm (mass_matrix_calculator.cpp:80-87) regrouped: with C_k the composite inertia
// (mass mc, first moment h, tensor I about the frame origin) of all inertias at or beyond stage k,
// column k of M is S_j^T X_{j<-k} (C_k S_k) for j <= k.  The composite walks inward once
// (rotate at the joint, parallel-axis shift at the link); each new column (f, n) = C_k S_k is then
// carried inward through the joints below it and projected on their axes.  Rotor inertias
// (inertia_gen) land on the diagonal.
// T = double gives M; T = dual carries d/dt alongside (cs, sn and the prismatic displacements have
// derivatives -sin q q_dot, cos q q_dot and q_dot), which yields Mdot in the same pass.
// Mp[i*(i+1)/2 + j] = M(j,i), j <= i, in stage order.
template <int N, int FL, shape_t SHAPE, class T>
RKB_DEV void mass_sweep(const SerialParams& P, const T (&cs)[N], const T (&sn)[N], const T (&qp)[N], T (&Mp)[N * (N + 1) / 2]) {
  vec3t<T> h = mkt<T>(zero_of<T>(), zero_of<T>(), zero_of<T>());
  T I[6];  // xx xy xz yy yz zz
#pragma unroll
  for (int d = 0; d < 6; ++d) I[d] = zero_of<T>();
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    const SerialStage& S = P.st[k];
    constexpr shape_t SH = SHAPE;
    const int AX = shape_ax(SH, k), LK = shape_lk(SH, k), IN = shape_in(SH, k);
    // [1] the stage's own inertia_3D sits at the link end frame, centre of mass on its origin
    if (IN == 1) { I[0] += S.I[0]; I[3] += S.I[3]; I[5] += S.I[5]; }
    else if (S.flags & RKB_ST_INERTIA) {
#pragma unroll
      for (int d = 0; d < 6; ++d) I[d] += S.I[d];
    }
    // [2] link end frame -> joint end frame
    if (LK != 0) {
      // offset L e_D: w = h + mc po / 2, I += 2 (w.po) 1 - w po^T - po w^T, h += mc po
      const int D = LK - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
      const double L = S.po[D];
      const T dd = L * (2.0 * h.c[D] + S.mcpo[D]);
      I[sym_idx(D1, D1)] += dd;
      I[sym_idx(D2, D2)] += dd;
      I[sym_idx(D1, D)] -= L * h.c[D1];
      I[sym_idx(D2, D)] -= L * h.c[D2];
      h.c[D] += S.mcpo[D];
    } else if (S.flags & RKB_ST_LINK) {
      if ((FL & RKB_FL_LINKROT) && (S.flags & RKB_ST_LINKROT)) {
        h = mulc_t<T>(S.Ro, h);
        sym_rotate_c<T>(S.Ro, I);
      }
      sym_shift_c<T>(ld3(S.po), ld3(S.mcpo), h, I);
    }
    // [3] column k at the joint end frame, then inward
    vec3t<T> f, n;
    const bool prismatic_k = AX == 0 && (FL & RKB_FL_PRISMATIC) && (S.flags & RKB_ST_PRISMATIC);
    if (AX != 0) {
      // revolute about sg e_D: n = I a, f = a x h
      const int D = AX - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
      const double sg = RKB_AXIS_SIGN(SH, k, S.ax[D]);
      n.c[0] = sg * I[sym_idx(0, D)]; n.c[1] = sg * I[sym_idx(1, D)]; n.c[2] = sg * I[sym_idx(2, D)];
      f.c[D] = zero_of<T>(); f.c[D1] = (-sg) * h.c[D2]; f.c[D2] = sg * h.c[D1];
      Mp[k * (k + 1) / 2 + k] = I[sym_idx(D, D)] + S.rotor;
    } else if (shape_px(SH, k) != 0) {
      // prismatic along sg e_D: f = mc sg e_D, n = h x (sg e_D), M(k,k) = mc + rotor
      const int D = shape_px(SH, k) - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
      const double sg = RKB_AXIS_SIGN(SH, k, S.ax[D]);
      f.c[D] = zero_of<T>() + sg * S.mc; f.c[D1] = zero_of<T>(); f.c[D2] = zero_of<T>();
      n.c[D] = zero_of<T>(); n.c[D1] = sg * h.c[D2]; n.c[D2] = (-sg) * h.c[D1];
      Mp[k * (k + 1) / 2 + k] = zero_of<T>() + (S.mc + S.rotor);
    } else if (!prismatic_k) {
      const vec3 ax = ld3(S.ax);
      n = symmul_vc<T>(I, ax);
      f = cross_cv<T>(ax, h);
      Mp[k * (k + 1) / 2 + k] = dot_cv<T>(ax, n) + S.rotor;
    } else {
      const vec3 ax = ld3(S.ax);
      f = mkt<T>(zero_of<T>() + S.mc * ax.c[0], zero_of<T>() + S.mc * ax.c[1], zero_of<T>() + S.mc * ax.c[2]);
      n = cross_vc<T>(h, ax);
      Mp[k * (k + 1) / 2 + k] = dot_cv<T>(ax, f) + S.rotor;
    }
#pragma unroll
    for (int j = k; j >= 1; --j) {
      // joint j: end frame -> base frame (= link end frame of stage j-1)
      const SerialStage& Sj = P.st[j];
      const int AXj = shape_ax(SH, j);
      if (AXj != 0) {
        const int D = AXj - 1;
        if (D == 0) { f = rot_axis_t<0, T>(cs[j], sn[j], f); n = rot_axis_t<0, T>(cs[j], sn[j], n); }
        else if (D == 1) { f = rot_axis_t<1, T>(cs[j], sn[j], f); n = rot_axis_t<1, T>(cs[j], sn[j], n); }
        else { f = rot_axis_t<2, T>(cs[j], sn[j], f); n = rot_axis_t<2, T>(cs[j], sn[j], n); }
      } else if (shape_px(SH, j) != 0) {
        // n += (L e_D) x f, L = sg q
        const int D = shape_px(SH, j) - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const T L = RKB_AXIS_SIGN(SH, j, Sj.ax[D]) * qp[j];
        n.c[D1] -= L * f.c[D2];
        n.c[D2] += L * f.c[D1];
      } else if (!((FL & RKB_FL_PRISMATIC) && (Sj.flags & RKB_ST_PRISMATIC))) {
        const mat3t<T> R = rodrigues_t<T>(Sj, cs[j], sn[j]);
        f = mul(R, f); n = mul(R, n);
      } else {
        const vec3 ax = ld3(Sj.ax);
        n = n + cross_vv<T>(mkt<T>(qp[j] * ax.c[0], qp[j] * ax.c[1], qp[j] * ax.c[2]), f);
      }
      // link j-1: end frame -> joint end frame, then project on joint j-1
      const SerialStage& Si = P.st[j - 1];
      const int AXi = shape_ax(SH, j - 1), LKi = shape_lk(SH, j - 1);
      if (LKi != 0) {
        const int D = LKi - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const double L = Si.po[D];
        n.c[D1] -= L * f.c[D2];
        n.c[D2] += L * f.c[D1];
      } else if (Si.flags & RKB_ST_LINK) {
        if ((FL & RKB_FL_LINKROT) && (Si.flags & RKB_ST_LINKROT)) { f = mulc_t<T>(Si.Ro, f); n = mulc_t<T>(Si.Ro, n); }
        n = n + cross_cv<T>(ld3(Si.po), f);
      }
      T mjk;
      if (AXi != 0) mjk = RKB_AXIS_SIGN(SH, j - 1, Si.ax[AXi - 1]) * n.c[AXi - 1];
      else if (shape_px(SH, j - 1) != 0) mjk = RKB_AXIS_SIGN(SH, j - 1, Si.ax[shape_px(SH, j - 1) - 1]) * f.c[shape_px(SH, j - 1) - 1];
      else if (!((FL & RKB_FL_PRISMATIC) && (Si.flags & RKB_ST_PRISMATIC))) mjk = dot_cv<T>(ld3(Si.ax), n);
      else mjk = dot_cv<T>(ld3(Si.ax), f);
      Mp[k * (k + 1) / 2 + (j - 1)] = mjk;
    }
    // [4] composite: joint end frame -> joint base frame
    if (k > 0) {
      if (AX != 0) {
        const int D = AX - 1;
        if (D == 0) { h = rot_axis_t<0, T>(cs[k], sn[k], h); sym_rotate_axis_t<0, T>(cs[k], sn[k], I); }
        else if (D == 1) { h = rot_axis_t<1, T>(cs[k], sn[k], h); sym_rotate_axis_t<1, T>(cs[k], sn[k], I); }
        else { h = rot_axis_t<2, T>(cs[k], sn[k], h); sym_rotate_axis_t<2, T>(cs[k], sn[k], I); }
      } else if (shape_px(SH, k) != 0) {
        // parallel-axis move by L e_D, L = sg q (the link-offset formulas with a variable offset)
        const int D = shape_px(SH, k) - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const T L = RKB_AXIS_SIGN(SH, k, S.ax[D]) * qp[k];
        const T mcL = S.mc * L;
        const T dd = L * (2.0 * h.c[D] + mcL);
        I[sym_idx(D1, D1)] += dd;
        I[sym_idx(D2, D2)] += dd;
        I[sym_idx(D1, D)] -= L * h.c[D1];
        I[sym_idx(D2, D)] -= L * h.c[D2];
        h.c[D] += mcL;
      } else if (!prismatic_k) {
        const mat3t<T> R = rodrigues_t<T>(S, cs[k], sn[k]);
        h = mul(R, h);
        sym_rotate_t<T>(R, I);
      } else {
        const vec3 ax = ld3(S.ax);
        sym_shift_v<T>(mkt<T>(qp[k] * ax.c[0], qp[k] * ax.c[1], qp[k] * ax.c[2]), S.mc, h, I);
      }
    }
  }
}

// The evaluation proper.  (`sm`, this thread's shared-memory column with stride SMS, is no longer
// used by the sweeps; the rollout kernel keeps its RK4 state there.)
// Returns f (generalised forces, gen_coord::f) and, if WANT_M, the packed upper triangle
// Mp[i*(i+1)/2 + j] = M(j,i), j <= i, in stage order.
// With HAVE_TRIG the caller has filled cs / sn (sign-folded, see serial_trig).
template <int N, int FL, shape_t SHAPE, int SMS, bool WANT_F, bool WANT_M, bool HAVE_TRIG = false>
RKB_DEV void serial_sweeps(const SerialParams& P, const SerialState<N>& X, double (&cs)[N], double (&sn)[N],
                           double (&f)[N], double (&Mp)[N * (N + 1) / 2], double* sm) {
  if (!HAVE_TRIG) serial_trig<N, FL, SHAPE>(P, X, cs, sn);
  // ---- sweep 1: kinematics outward, inertia wrenches parked --------------------------------
  if (WANT_F) {
    vec3 w = ld3(P.w0), al = ld3(P.al0), a = ld3(P.a0);
    vec3 F = mk(0, 0, 0), T = mk(0, 0, 0);  // running wrench of sweep 2
    // d'Alembert wrenches of stages 0..N-2, held until sweep 2 picks them up.  (They used to be parked
    // in shared memory; the compiler forwarded every store to its load and kept the values in
    // registers anyway, so the stores were 30 dead STS per evaluation.  It now decides itself which of
    // them to keep and which to spill.)
    double park[6 * (N > 1 ? N - 1 : 1)];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const SerialStage& S = P.st[k];
      constexpr shape_t SH = SHAPE;
      const int AX = shape_ax(SH, k), LK = shape_lk(SH, k), IN = shape_in(SH, k);
      if (AX != 0) {
        // revolute about +-e_D (revolute_joint.cpp:121-152)
        const int D = AX - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const double g = RKB_AXIS_SIGN(SH, k, S.ax[D]) * X.qd[k];  // q_dot * axis has the single component g
        vec3 wt, alt, at;
        if (D == 0) { wt = rotT_axis<0>(cs[k], sn[k], w); alt = rotT_axis<0>(cs[k], sn[k], al); at = rotT_axis<0>(cs[k], sn[k], a); }
        else if (D == 1) { wt = rotT_axis<1>(cs[k], sn[k], w); alt = rotT_axis<1>(cs[k], sn[k], al); at = rotT_axis<1>(cs[k], sn[k], a); }
        else { wt = rotT_axis<2>(cs[k], sn[k], w); alt = rotT_axis<2>(cs[k], sn[k], al); at = rotT_axis<2>(cs[k], sn[k], a); }
        // al = R^T al + wt x (g e_D): (v x e_D)[D1] = v[D2], (v x e_D)[D2] = -v[D1]
        alt.c[D1] += wt.c[D2] * g;
        alt.c[D2] -= wt.c[D1] * g;
        wt.c[D] += g;
        w = wt; al = alt; a = at;
      } else if (shape_px(SH, k) != 0) {
        // prismatic along sg e_D (prismatic_joint.cpp:129-161): r = L e_D, r_dot = Ld e_D;
        // a += w x (w x r) + 2 w x r_dot + al x r, with (v x e_D)[D1] = v[D2], (v x e_D)[D2] = -v[D1]
        const int D = shape_px(SH, k) - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const double sg = RKB_AXIS_SIGN(SH, k, S.ax[D]);
        const double L = sg * X.q[k], Ld2 = 2.0 * (sg * X.qd[k]);
        a.c[D] -= L * (w.c[D1] * w.c[D1] + w.c[D2] * w.c[D2]);
        a.c[D1] += L * (w.c[D] * w.c[D1] + al.c[D2]) + Ld2 * w.c[D2];
        a.c[D2] += L * (w.c[D] * w.c[D2] - al.c[D1]) - Ld2 * w.c[D1];
      } else {
        const vec3 ax = ld3(S.ax);
        const bool prismatic = (FL & RKB_FL_PRISMATIC) && (S.flags & RKB_ST_PRISMATIC);
        if (!prismatic) {
          const mat3 R2 = rodrigues(S, cs[k], sn[k]);
          const vec3 wt = tmul(R2, w);
          const vec3 qda = X.qd[k] * ax;
          al = tmul(R2, al) + cross(wt, qda);  // + q_ddot * axis, zero in this pass (kte_nl_system.hpp:192)
          w = wt + qda;
          a = tmul(R2, a);
        } else {
          const vec3 r = X.q[k] * ax, rd = X.qd[k] * ax;
          a = a + cross(w, cross(w, r)) + 2.0 * cross(w, rd) + cross(al, r);
        }
      }
      if (LK != 0) {
        // link offset L e_D, no rotation: a += w x (w x po) + al x po = L (w_D w - |w|^2 e_D + al x e_D)
        const int D = LK - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const double L = S.po[D];
        a.c[D] -= L * (w.c[D1] * w.c[D1] + w.c[D2] * w.c[D2]);
        a.c[D1] += L * (w.c[D] * w.c[D1] + al.c[D2]);
        a.c[D2] += L * (w.c[D] * w.c[D2] - al.c[D1]);
      } else if (S.flags & RKB_ST_LINK) {
        const vec3 po = ld3(S.po);
        a = a + cross(w, cross(w, po)) + cross(al, po);
        if ((FL & RKB_FL_LINKROT) && (S.flags & RKB_ST_LINKROT)) {
          a = tmulc(S.Ro, a); w = tmulc(S.Ro, w); al = tmulc(S.Ro, al);
        }
      }
      // inertia_3D::doForce: F -= m * (R^T a_global) ; T -= I al + w x (I w)
      vec3 Fk = mk(0, 0, 0), Tk = mk(0, 0, 0);
      if (IN == 1) {
        Fk = (-S.m) * a;
        const vec3 Iw = diagmul(S.I, w);
        Tk = mk(0, 0, 0) - (diagmul(S.I, al) + cross(w, Iw));
      } else if (S.flags & RKB_ST_INERTIA) {
        Fk = (-S.m) * a;
        const vec3 Iw = symmul(S.I, w);
        Tk = mk(0, 0, 0) - (symmul(S.I, al) + cross(w, Iw));
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (k == N - 1) { F.c[d] = Fk.c[d]; T.c[d] = Tk.c[d]; }  // the outermost wrench is consumed first: keep it in registers
        else { park[6 * k + d] = Fk.c[d]; park[6 * k + 3 + d] = Tk.c[d]; }
      }
    }
    // ---- sweep 2: wrenches inward -----------------------------------------------------------
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      const SerialStage& S = P.st[k];
      constexpr shape_t SH = SHAPE;
      const int AX = shape_ax(SH, k), LK = shape_lk(SH, k);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (k < N - 1) { F.c[d] += park[6 * k + d]; T.c[d] += park[6 * k + 3 + d]; }
      }
      if (LK != 0) {
        // T += po x F = L (e_D x F): [D1] -= L F[D2], [D2] += L F[D1]   (rigid_link.cpp:170-177, Ro = I)
        const int D = LK - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const double L = S.po[D];
        T.c[D1] -= L * F.c[D2];
        T.c[D2] += L * F.c[D1];
      } else if (S.flags & RKB_ST_LINK) {
        if ((FL & RKB_FL_LINKROT) && (S.flags & RKB_ST_LINKROT)) { F = mulc(S.Ro, F); T = mulc(S.Ro, T); }
        T = T + cross(ld3(S.po), F);
      }
      if (AX != 0) {
        // revolute about +-e_D (revolute_joint.cpp:172-184 + actuator reaction :210-213)
        const int D = AX - 1;
        const double sg = RKB_AXIS_SIGN(SH, k, S.ax[D]);
        double tsd = 0.0;  // spring + damper torque along e_D
        if (FL & RKB_FL_SPRINGS) {
          if (S.flags & RKB_ST_SPRING) tsd = sg * spring_scalar(S, X.q[k]);
          if (S.flags & RKB_ST_DAMPER) tsd += sg * (S.cd * X.qd[k]);
        }
        f[k] = sg * (T.c[D] - tsd) + X.u[k];
        T.c[D] = tsd - sg * X.u[k];  // axial part removed, reaction and spring/damper on the base side added
        if (D == 0) { F = rot_axis<0>(cs[k], sn[k], F); T = rot_axis<0>(cs[k], sn[k], T); }
        else if (D == 1) { F = rot_axis<1>(cs[k], sn[k], F); T = rot_axis<1>(cs[k], sn[k], T); }
        else { F = rot_axis<2>(cs[k], sn[k], F); T = rot_axis<2>(cs[k], sn[k], T); }
      } else if (shape_px(SH, k) != 0) {
        // prismatic along sg e_D (prismatic_joint.cpp:181-193 + reaction :219-222): f = F.a + u,
        // T += (q a) x F, F -= (F.a) a + u a  ->  F_D = -sg u
        const int D = shape_px(SH, k) - 1, D1 = (D + 1) % 3, D2 = (D + 2) % 3;
        const double sg = RKB_AXIS_SIGN(SH, k, S.ax[D]);
        const double L = sg * X.q[k];
        f[k] = sg * F.c[D] + X.u[k];
        T.c[D1] -= L * F.c[D2];
        T.c[D2] += L * F.c[D1];
        F.c[D] = -(sg * X.u[k]);
      } else {
        const vec3 ax = ld3(S.ax);
        const bool prismatic = (FL & RKB_FL_PRISMATIC) && (S.flags & RKB_ST_PRISMATIC);
        if (!prismatic) {
          // torsion spring / damper between the joint's base and end frames act along the axis
          vec3 tsd = mk(0, 0, 0);
          if (FL & RKB_FL_SPRINGS) {
            if (S.flags & RKB_ST_SPRING) tsd = spring_scalar(S, X.q[k]) * ld3(S.an);
            if (S.flags & RKB_ST_DAMPER) tsd = tsd + (S.cd * X.qd[k]) * ax;
            T = T - tsd;
          }
          const double ta = dot(T, ax);
          f[k] = ta + X.u[k];
          const mat3 R = rodrigues(S, cs[k], sn[k]);
          F = mul(R, F);
          T = mul(R, T - ta * ax) - X.u[k] * ax;  // actuator reaction on the joint base
          if (FL & RKB_FL_SPRINGS) T = T + tsd;
        } else {
          const double fa = dot(F, ax);
          f[k] = fa + X.u[k];
          T = T + cross(X.q[k] * ax, F);
          F = F - fa * ax - X.u[k] * ax;
        }
      }
    }
  }
  // ---- sweep 3: mass matrix -----------------------------------------------------------------------
  if (WANT_M) mass_sweep<N, FL, SHAPE, double>(P, cs, sn, X.q, Mp);
}

// 1/d for a pivot d >= 1e-8 (anything else has already raised the singular status): the 2^-23
// hardware seed (MUFU.RCP64H) and two Newton steps, relative error ~2^-52, no slow path.
RKB_DEV double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// M q_ddot = f on the packed matrix; b is overwritten with the solution.  The factorisation is the
// square-root-free L D L^T of core/lin_alg/mat_cholesky.hpp:134-157 (same pivots as the Cholesky
// of :63-84: D(i) = L_chol(i,i)^2), so the reference's "pivot < 1e-8 -> singularity_error" test
// (:80-82) is applied to D(i) unchanged, and only one reciprocal per pivot is needed — no sqrt.
// Row i keeps W(i,k) = L(i,k) D(k) in a scratch row while L(i,k) overwrites the matrix.
template <int N>
RKB_DEV int ldl_factor_packed(double (&Mp)[N * (N + 1) / 2], double (&inv)[N]) {
  int st = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double W[N];
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double s = Mp[i * (i + 1) / 2 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-W[k], Mp[j * (j + 1) / 2 + k], s);
      W[j] = s;                              // L(i,j) D(j)
      Mp[i * (i + 1) / 2 + j] = s * inv[j];  // L(i,j)
    }
    double d = Mp[i * (i + 1) / 2 + i];
#pragma unroll
    for (int k = 0; k < i; ++k) d = fma(-W[k], Mp[i * (i + 1) / 2 + k], d);
    if (!(d >= 1.0e-8)) st = RKB_STATUS_SINGULAR;
    inv[i] = fast_rcp(d);
  }
  return st;
}
// ... and the two substitutions with the factors: b <- (L D L^T)^-1 b
template <int N>
RKB_DEV void ldl_apply_packed(const double (&Mp)[N * (N + 1) / 2], const double (&inv)[N], double (&b)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {  // L y = b
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-Mp[i * (i + 1) / 2 + k], b[k], s);
    b[i] = s;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {  // D L^T x = y
    double s = b[i] * inv[i];
#pragma unroll
    for (int k = N - 1; k > i; --k) s = fma(-Mp[k * (k + 1) / 2 + i], b[k], s);
    b[i] = s;
  }
}
template <int N>
RKB_DEV int cholesky_solve_packed(double (&Mp)[N * (N + 1) / 2], double (&b)[N]) {
  double inv[N];
  const int st = ldl_factor_packed<N>(Mp, inv);
  ldl_apply_packed<N>(Mp, inv, b);
  return st;
}

// q_ddot = M^-1 f for the state in X; returns the status bits.
template <int N, int FL, shape_t SHAPE, int SMS>
RKB_DEV int serial_accel(const SerialParams& P, const SerialState<N>& X, double (&qdd)[N], double* sm) {
  double cs[N], sn[N], Mp[N * (N + 1) / 2];
  serial_sweeps<N, FL, SHAPE, SMS, true, true>(P, X, cs, sn, qdd, Mp, sm);
  return cholesky_solve_packed<N>(Mp, qdd);
}
// ... with cos / sin supplied by the caller
template <int N, int FL, shape_t SHAPE, int SMS>
RKB_DEV int serial_accel_trig(const SerialParams& P, const SerialState<N>& X, double (&cs)[N], double (&sn)[N], double (&qdd)[N], double* sm) {
  double Mp[N * (N + 1) / 2];
  serial_sweeps<N, FL, SHAPE, SMS, true, true, true>(P, X, cs, sn, qdd, Mp, sm);
  return cholesky_solve_packed<N>(Mp, qdd);
}

// ---- one sample on two warps (small batches) -----------------------------------------------------------
// A lone warp per SM sub-partition — the situation of a batch of a few hundred to a few thousand samples — issues
// one FP64 instruction every 2.5 cycles and spends ~2600 cycles on the ~750 FP64 instructions of a 6-coordinate
// evaluation, while the longest chain of dependent instructions in it is only ~62 (500 cycles): the warp is bound by
// its own issue slot, not by latency (tools/small_batch_microbench.cu, tools/flop_count.py: critical_path).  The
// evaluation has two independent halves of almost equal size — sweeps 1-2 (the generalised forces) and sweep 3 plus
// the factorisation (the mass matrix) — so a PAIR of warps on two sub-partitions takes one each: the "force" warp
// hands f over through shared memory, the "mass" warp solves and hands q_ddot back, both advance their own copy of
// the RK4 state.  Two named-barrier hand-offs per evaluation (~75 cycles each, same microbenchmark).
struct DuoCtx {
  double* xf;   // this sample's exchange column for f     (N doubles, stride 32)
  double* xq;   // ...                          for q_ddot
  int role;     // 0: forces, 1: mass matrix + solve (warp-uniform)
  int bar_f, bar_q;  // named barrier ids of this pair
};
RKB_DEV void named_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
RKB_DEV void named_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int N, int FL, shape_t SHAPE, int SMS>
RKB_DEV int duo_accel_trig(const SerialParams& P, const SerialState<N>& X, double (&cs)[N], double (&sn)[N], double (&qdd)[N], double* sm,
                           const DuoCtx& duo) {
  double Mp[N * (N + 1) / 2];
  int st = 0;
  if (duo.role == 0) {
    serial_sweeps<N, FL, SHAPE, SMS, true, false, true>(P, X, cs, sn, qdd, Mp, sm);
#pragma unroll
    for (int k = 0; k < N; ++k) duo.xf[k * 32] = qdd[k];
    named_arrive(duo.bar_f);   // f is there
    named_sync(duo.bar_q);     // wait for q_ddot
#pragma unroll
    for (int k = 0; k < N; ++k) qdd[k] = duo.xq[k * 32];
  } else {
    double inv[N];
    serial_sweeps<N, FL, SHAPE, SMS, false, true, true>(P, X, cs, sn, qdd, Mp, sm);
    st = ldl_factor_packed<N>(Mp, inv);
    named_sync(duo.bar_f);
#pragma unroll
    for (int k = 0; k < N; ++k) qdd[k] = duo.xf[k * 32];
    ldl_apply_packed<N>(Mp, inv, qdd);
#pragma unroll
    for (int k = 0; k < N; ++k) duo.xq[k * 32] = qdd[k];
    named_arrive(duo.bar_q);
  }
  return st;
}

template <int N>
RKB_DEV void load_state(const SerialParams& P, const ConstBatchView& x, const ConstBatchView& u, long long i, SerialState<N>& X) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int c = P.st[k].coord;
    X.q[k] = x.p[i * x.si + rkb_state_q(x.blocked, N, c) * x.sk];
    X.qd[k] = x.p[i * x.si + rkb_state_qd(x.blocked, N, c) * x.sk];
    const int in = P.st[k].input;
    X.u[k] = (in >= 0) ? u.p[i * u.si + in * u.sk] : 0.0;
  }
}

#ifndef RKB_BLOCK
#define RKB_BLOCK 128
#endif
// per-thread shared-memory doubles: for the rollout the
// state at the start of the step (w), k1 + 2 k2, and cos / sin of the joint angles at the start of the step.  k3 is not stored: the stage-4 update starts
// from x = w + k3, so k3 is recovered as x - w (exact up to one rounding of x, i.e. ~1e-16 |x|).
#define RKB_SMEM_EVAL(n) 1
#define RKB_SMEM_ROLLOUT(n) (6 * (n))
// Resident CTAs per SM the compiler must leave room for.  The kernels are bound by register-file
// reads and the FP64 pipe, not by latency (tools/issue_model.py), so occupancy only matters through
// the spills a tighter register budget causes: up to 6 coordinates the specialised code is equally
// fast at 4 CTAs (128 registers) and 3 CTAs (168) per SM and keeps 4; from 7 coordinates on 3 CTAs
// are 5 % faster (7-DOF rollout 3.24 vs 3.41 ms).  The general code needs the full 255 registers and
// loses 10 % when squeezed.
#ifdef RKB_FORCE_MINBLOCKS
#define RKB_MINBLOCKS(shape, n, smem_doubles_per_thread) (RKB_FORCE_MINBLOCKS)
#endif
#ifndef RKB_MINBLOCKS
#define RKB_FITS(blocks, smem_doubles_per_thread) ((blocks) * ((smem_doubles_per_thread) * RKB_BLOCK * 8 + 1024) <= 227 * 1024)
#define RKB_MINBLOCKS(shape, n, smem_doubles_per_thread) \
  ((shape) == 0 ? 1 : (((n) <= 6 && RKB_FITS(4, smem_doubles_per_thread)) ? 4 : (RKB_FITS(3, smem_doubles_per_thread) ? 3 : 1)))
#endif

// ---- coalesced result write-back ------------------------------------------------------------------
// A CTA's results form one contiguous tile [rows][DIM] of the caller's AoS buffer.  A thread that
// wrote its own DIM doubles straight to global memory would touch 32 different sectors per store
// instruction (8 useful bytes each).  Instead every thread parks its row in shared memory (row
// stride DIM + 1 doubles: odd, so the 64-bit accesses of a warp never collide on a bank), and the
// CTA then streams the tile out with consecutive threads writing consecutive 128-bit words.
// SoA buffers need none of this: element k of consecutive samples is already contiguous.
template <int DIM>
RKB_DEV bool tile_is_aos(const BatchView& o) { return o.sk == 1 && o.si == DIM; }

template <int DIM>
RKB_DEV void tile_write_back(double* smem, const BatchView& o, long long tile_first, long long n_samples) {
  constexpr int STRIDE = DIM + 1;
  const long long rows_ll = n_samples - tile_first;
  const int rows = rows_ll < RKB_BLOCK ? (int)rows_ll : RKB_BLOCK;
  const int total = rows * DIM;
  double* g = o.p + tile_first * DIM;
  if ((DIM % 2 == 0) && ((reinterpret_cast<unsigned long long>(g) & 15ull) == 0)) {
    for (int e = 2 * threadIdx.x; e < total; e += 2 * RKB_BLOCK) {
      const int row = e / DIM, col = e - row * DIM;
      double2 v;
      v.x = smem[row * STRIDE + col];
      v.y = smem[row * STRIDE + col + 1];
      *reinterpret_cast<double2*>(g + e) = v;  // STG.128, consecutive lanes -> consecutive 16 bytes
    }
  } else {
    for (int e = threadIdx.x; e < total; e += RKB_BLOCK) {
      const int row = e / DIM, col = e - row * DIM;
      g[e] = smem[row * STRIDE + col];
    }
  }
}

// (The mirror image for the INPUTS — cooperative 128-bit tile loads into padded shared memory,
// then each thread picks up its row — was built and measured: 0.218 ms instead of 0.164 ms per 2^21
// 6-DOF evaluations.  The two extra barriers and the shared-memory round trip cost more than the
// strided LDG.64 they replace, whose sectors are all consumed out of L1 anyway.  Inputs are
// therefore read directly, once per thread.)
#define RKB_SMEM_TILE(dim) ((dim) + 1)
#define RKB_MAX2(a, b) ((a) > (b) ? (a) : (b))
#define RKB_SMEM_EVAL_K(n) RKB_MAX2(RKB_SMEM_EVAL(n), RKB_SMEM_TILE(2 * (n)))
#define RKB_SMEM_FORCES_K(n) RKB_MAX2(RKB_SMEM_EVAL(n), RKB_SMEM_TILE(n))
#define RKB_SMEM_MASS_K(n) RKB_MAX2(RKB_SMEM_EVAL(n), RKB_SMEM_TILE((n) * (n)))
#define RKB_SMEM_ROLLOUT_K(n) RKB_MAX2(RKB_SMEM_ROLLOUT(n), RKB_SMEM_TILE(2 * (n)))

// ---- kernels ---------------------------------------------------------------------------------
// xdot = get_state_derivative(x, u)
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_BLOCK, RKB_MINBLOCKS(SHAPE, N, RKB_SMEM_EVAL_K(N))) serial_eval_kernel(const __grid_constant__ SerialParams P, const EvalArgs A) {
  extern __shared__ double smem[];
  const long long tile_first = (long long)blockIdx.x * RKB_BLOCK;
  const long long i = tile_first + threadIdx.x;
  const bool active = i < A.n_samples;
  const bool staged = tile_is_aos<2 * N>(A.out);
  double xd[2 * N];
  int st = 0;
  if (active) {
    SerialState<N> X;
    load_state<N>(P, A.x, A.u, i, X);
    double qdd[N];
    st = serial_accel<N, FL, SHAPE, RKB_BLOCK>(P, X, qdd, smem + threadIdx.x);
    bool finite = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      xd[2 * k] = X.qd[k];
      xd[2 * k + 1] = qdd[k];
      finite = finite && isfinite(qdd[k]) && isfinite(X.qd[k]);
    }
    if (!finite) st |= RKB_STATUS_NONFINITE;
    if (A.status) A.status[i] = st;
    if (!staged) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int c = P.st[k].coord;
        A.out.p[i * A.out.si + rkb_state_q(A.out.blocked, N, c) * A.out.sk] = xd[2 * k];
        A.out.p[i * A.out.si + rkb_state_qd(A.out.blocked, N, c) * A.out.sk] = xd[2 * k + 1];
      }
    }
  }
  if (staged) {
    __syncthreads();  // every thread is done with its wrench column
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int c = P.st[k].coord;
        smem[threadIdx.x * (2 * N + 1) + rkb_state_q(A.out.blocked, N, c)] = xd[2 * k];
        smem[threadIdx.x * (2 * N + 1) + rkb_state_qd(A.out.blocked, N, c)] = xd[2 * k + 1];
      }
    }
    __syncthreads();
    tile_write_back<2 * N>(smem, A.out, tile_first, A.n_samples);
  }
}

// f = gen_coord::f after doMotion / clearForce / doForce with q_ddot = 0
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_BLOCK) serial_forces_kernel(const __grid_constant__ SerialParams P, const EvalArgs A) {
  extern __shared__ double smem[];
  const long long tile_first = (long long)blockIdx.x * RKB_BLOCK;
  const long long i = tile_first + threadIdx.x;
  const bool active = i < A.n_samples;
  const bool staged = tile_is_aos<N>(A.out);
  double f[N];
  if (active) {
    SerialState<N> X;
    load_state<N>(P, A.x, A.u, i, X);
    double cs[N], sn[N], Mp[N * (N + 1) / 2];
    serial_sweeps<N, FL, SHAPE, RKB_BLOCK, true, false>(P, X, cs, sn, f, Mp, smem + threadIdx.x);
    if (!staged) {
#pragma unroll
      for (int k = 0; k < N; ++k) A.out.p[i * A.out.si + P.st[k].coord * A.out.sk] = f[k];
    }
  }
  if (staged) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) smem[threadIdx.x * (N + 1) + P.st[k].coord] = f[k];
    }
    __syncthreads();
    tile_write_back<N>(smem, A.out, tile_first, A.n_samples);
  }
}

// M = getMassMatrix (full symmetric n x n, row-major per sample in AoS) and, with WANT_DOT, its time
// derivative Mdot = getMassMatrixAndDerivative's second result (mass_matrix_calculator.cpp:89-98) from
// the same sweep run on forward-mode scalars.
template <int N, int DIMS>
RKB_DEV void store_sym(const double (&Mp)[N * (N + 1) / 2], const SerialParams& P, const BatchView& o, long long i, bool staged, bool active,
                       double* smem, long long tile_first, long long n_samples) {
  if (active && !staged) {
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        const int ca = P.st[a].coord, cb = P.st[b].coord;
        const double v = Mp[a * (a + 1) / 2 + b];
        o.p[i * o.si + (long long)(ca * N + cb) * o.sk] = v;
        o.p[i * o.si + (long long)(cb * N + ca) * o.sk] = v;
      }
  }
  if (staged) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          const int ca = P.st[a].coord, cb = P.st[b].coord;
          const double v = Mp[a * (a + 1) / 2 + b];
          smem[threadIdx.x * (N * N + 1) + ca * N + cb] = v;
          smem[threadIdx.x * (N * N + 1) + cb * N + ca] = v;
        }
    }
    __syncthreads();
    tile_write_back<N * N>(smem, o, tile_first, n_samples);
  }
}

template <int N, int FL, shape_t SHAPE, bool WANT_DOT>
__global__ void __launch_bounds__(RKB_BLOCK) serial_mass_kernel(const __grid_constant__ SerialParams P, const EvalArgs A) {
  extern __shared__ double smem[];
  const long long tile_first = (long long)blockIdx.x * RKB_BLOCK;
  const long long i = tile_first + threadIdx.x;
  const bool active = i < A.n_samples;
  double Mp[N * (N + 1) / 2], Mdp[WANT_DOT ? N * (N + 1) / 2 : 1];
  if (active) {
    SerialState<N> X;
    ConstBatchView nou = A.x;
    load_state<N>(P, A.x, nou, i, X);
    double cs[N], sn[N], f[N];
    if (!WANT_DOT) {
      serial_sweeps<N, FL, SHAPE, RKB_BLOCK, false, true>(P, X, cs, sn, f, Mp, smem + threadIdx.x);
    } else {
      serial_trig<N, FL, SHAPE>(P, X, cs, sn);
      dual dcs[N], dsn[N], dq[N], Md[N * (N + 1) / 2];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        constexpr shape_t SH = SHAPE;
        const int AX = shape_ax(SH, k);
        const bool prismatic = AX == 0 && (FL & RKB_FL_PRISMATIC) && (P.st[k].flags & RKB_ST_PRISMATIC);
        // d/dt cos q = -sin q q_dot, d/dt sin q = cos q q_dot; with the axis sign sg = +-1 folded in,
        // sn = sg sin q: d/dt sn = sg cos q q_dot and sin q = sg sn
        const double sg = (AX != 0) ? RKB_AXIS_SIGN(SH, k, P.st[k].ax[AX - 1]) : 1.0;
        if (!prismatic) {
          dcs[k] = mkd(cs[k], -(sg * sn[k]) * X.qd[k]);
          dsn[k] = mkd(sn[k], (sg * cs[k]) * X.qd[k]);
        } else { dcs[k] = mkd(1.0, 0.0); dsn[k] = mkd(0.0, 0.0); }
        dq[k] = mkd(X.q[k], X.qd[k]);
      }
      mass_sweep<N, FL, SHAPE, dual>(P, dcs, dsn, dq, Md);
#pragma unroll
      for (int e = 0; e < N * (N + 1) / 2; ++e) { Mp[e] = Md[e].v; Mdp[e] = Md[e].d; }
    }
  }
  store_sym<N, N * N>(Mp, P, A.out, i, tile_is_aos<N * N>(A.out), active, smem, tile_first, A.n_samples);
  if (WANT_DOT) {
    double (&Mdq)[N * (N + 1) / 2] = reinterpret_cast<double (&)[N * (N + 1) / 2]>(Mdp);
    store_sym<N, N * N>(Mdq, P, A.out2, i, tile_is_aos<N * N>(A.out2), active, smem, tile_first, A.n_samples);
  }
}

template <int N>
RKB_DEV void store_state(const SerialParams& P, const BatchView& o, long long off, const SerialState<N>& X) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int c = P.st[k].coord;
    o.p[off + rkb_state_q(o.blocked, N, c) * o.sk] = X.q[k];
    o.p[off + rkb_state_qd(o.blocked, N, c) * o.sk] = X.qd[k];
  }
}

// n_steps RK4 steps of size dt on the state in X with the inputs X.u held; returns the status bits.
// sm is this thread's shared-memory column: w (2N), k1 + 2 k2 (2N), cos / sin at the start of the step (2N).
// Inputs that vary within a step (ctrl::detail::runge_kutta4_integrate_impl, ctrl/sys_integrators/
// runge_kutta4_integrator_sys.hpp:50-97): the input trajectory sampled at every half step; evaluation 1 of step s reads
// node 2 s, evaluations 2 and 3 node 2 s + 1, evaluation 4 node 2 s + 2.
struct NodeInputs {
  const double* p;    // this sample's node 0
  long long sj, sk;   // input k of node j at p[j * sj + k * sk]
};

template <int N, int FL, shape_t SHAPE, int SMS, bool DUO = false>
RKB_DEV int rk4_steps(const SerialParams& P, SerialState<N>& X, double dt, int n_steps, double* sm, const DuoCtx* duo = nullptr,
                      const NodeInputs* nodes = nullptr) {
  double* sw = sm;                  // state at the start of the step (w)
  double* sa = sw + 2 * N * SMS;    // k1 + 2 k2
  double* sb = sa + 2 * N * SMS;    // cos, sin of the joint angles at the start of the step
  const double sixth = 1.0 / 6.0;
  int st = 0;
  const int total = 4 * n_steps;
#pragma unroll 1
  for (int it = 0; it < total; ++it) {
    const int stage = it & 3;
    double qdd[N], cs[N], sn[N];
    if (nodes) {
      const long long node = 2 * (long long)(it >> 2) + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int in = P.st[k].input;
        X.u[k] = (in >= 0) ? nodes->p[node * nodes->sj + in * nodes->sk] : 0.0;
      }
    }
    {
      // Stage 1 evaluates sin / cos in full and keeps them; stages 2-4 sit at w + d with d = dt q_dot / 2
      // or dt q_dot, and get theirs by sincos_shift unless some |d| is not small (or not finite).
      double d[N];
      bool full = stage == 0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        d[k] = X.q[k] - sw[(2 * k) * SMS];  // (stage 1 reads the previous step's w: unused, `full` is already set)
        const bool prismatic = shape_ax(SHAPE, k) == 0 && (FL & RKB_FL_PRISMATIC) && (P.st[k].flags & RKB_ST_PRISMATIC);
        if (!prismatic) full = full || !(fabs(d[k]) < RKB_SMALL_ANGLE);
      }
      if (full) {
        serial_trig_raw<N, FL, SHAPE>(P, X, cs, sn);
        if (stage == 0) {
#pragma unroll
          for (int k = 0; k < N; ++k) { sb[(2 * k) * SMS] = cs[k]; sb[(2 * k + 1) * SMS] = sn[k]; }
        }
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const bool prismatic = shape_ax(SHAPE, k) == 0 && (FL & RKB_FL_PRISMATIC) && (P.st[k].flags & RKB_ST_PRISMATIC);
          if (!prismatic) sincos_shift(d[k], sb[(2 * k + 1) * SMS], sb[(2 * k) * SMS], sn[k], cs[k]);
          else { sn[k] = 0.0; cs[k] = 1.0; }
        }
      }
      serial_trig_fold<N, FL, SHAPE>(P, sn);
    }
    if (DUO) st |= duo_accel_trig<N, FL, SHAPE, SMS>(P, X, cs, sn, qdd, sm, *duo);
    else st |= serial_accel_trig<N, FL, SHAPE, SMS>(P, X, cs, sn, qdd, sm);
    // state derivative f = (qd, qdd) interleaved; the four stages of fixed_step_integrators.hpp:277-289
    if (stage == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double k1q = X.qd[k] * dt, k1v = qdd[k] * dt;
        sw[(2 * k) * SMS] = X.q[k]; sw[(2 * k + 1) * SMS] = X.qd[k];
        sa[(2 * k) * SMS] = k1q;    sa[(2 * k + 1) * SMS] = k1v;
        X.q[k] += k1q * 0.5; X.qd[k] += k1v * 0.5;
      }
    } else if (stage == 1) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double k2q = X.qd[k] * dt, k2v = qdd[k] * dt;
        sa[(2 * k) * SMS] += k2q * 2.0; sa[(2 * k + 1) * SMS] += k2v * 2.0;
        X.q[k] = sw[(2 * k) * SMS] + k2q * 0.5; X.qd[k] = sw[(2 * k + 1) * SMS] + k2v * 0.5;
      }
    } else if (stage == 2) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double k3q = X.qd[k] * dt, k3v = qdd[k] * dt;
        X.q[k] = sw[(2 * k) * SMS] + k3q; X.qd[k] = sw[(2 * k + 1) * SMS] + k3v;
      }
    } else {
      // x += (k1 + 2 k2 + k4) / 6 - k3 (2/3), with the division by 6 as a multiplication
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double k4q = X.qd[k] * dt, k4v = qdd[k] * dt;
        const double k3q = X.q[k] - sw[(2 * k) * SMS], k3v = X.qd[k] - sw[(2 * k + 1) * SMS];
        X.q[k] += (sa[(2 * k) * SMS] + k4q) * sixth - k3q * (2.0 / 3.0);
        X.qd[k] += (sa[(2 * k + 1) * SMS] + k4v) * sixth - k3v * (2.0 / 3.0);
      }
    }
  }
  return st;
}

// n_steps of fixed-step RK4 (runge_kutta4_integrator<T>::integrate, fixed_step_integrators.hpp:256-293)
// with the input held constant (num_int_dtnl_sys::get_next_state, num_int_dtnl_system.hpp:166-180).
// Per-thread shared-memory column: w (2N) and k1 + 2 k2 (2N).
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_BLOCK, RKB_MINBLOCKS(SHAPE, N, RKB_SMEM_ROLLOUT(N))) serial_rollout_kernel(const __grid_constant__ SerialParams P, const RolloutArgs A) {
  extern __shared__ double smem[];
  constexpr int SMS = RKB_BLOCK;
  const long long i = (long long)blockIdx.x * RKB_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  if (A.active && !A.active[i]) return;
  double* sm = smem + threadIdx.x;
  SerialState<N> X;
  {
    const long long i0 = A.x0_div > 1 ? i / A.x0_div : i;
    ConstBatchView xv = A.x0;
    xv.p += i0 * xv.si - i * xv.si;  // rows of x0 are shared by x0_div consecutive samples (steer batch)
    load_state<N>(P, xv, A.u, i, X);
  }
  int st = rk4_steps<N, FL, SHAPE, SMS>(P, X, A.dt, A.n_steps, sm);
  bool finite = true;
#pragma unroll
  for (int k = 0; k < N; ++k) finite = finite && isfinite(X.q[k]) && isfinite(X.qd[k]);
  store_state<N>(P, A.xout, i * A.xout.si, X);
  if (A.traj.p) store_state<N>(P, A.traj, i * A.traj.si, X);
  if (!finite) st |= RKB_STATUS_NONFINITE;
  if (A.status) A.status[i] = A.status_or ? (A.status[i] | st) : st;
}

// The sample-sharded job's rollout AND its all-gather in one kernel: every CTA parks its tile of end states in shared
// memory (the RK4 columns are free by then) and streams it, as coalesced 128-bit stores, to its place in EVERY
// destination buffer — this GPU's copy of the gathered batch and, through NVLink peer mappings, the other GPUs'.  The
// 96 bytes per sample and peer are nothing against the ~20 us of arithmetic behind them, so the transfer hides under
// the integration of the other CTAs and no collective runs afterwards (only a barrier before anyone reads).
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_BLOCK, RKB_MINBLOCKS(SHAPE, N, RKB_SMEM_ROLLOUT_K(N))) serial_rollout_scatter_kernel(const __grid_constant__ SerialParams P, const __grid_constant__ RolloutScatterArgs A) {
  extern __shared__ double smem[];
  constexpr int SMS = RKB_BLOCK;
  const long long tile_first = (long long)blockIdx.x * RKB_BLOCK;
  const long long i = tile_first + threadIdx.x;
  const bool active = i < A.n_samples;
  SerialState<N> X;
  int st = 0;
  if (active) {
    load_state<N>(P, A.x0, A.u, i, X);
    st = rk4_steps<N, FL, SHAPE, SMS>(P, X, A.dt, A.n_steps, smem + threadIdx.x);
    bool finite = true;
#pragma unroll
    for (int k = 0; k < N; ++k) finite = finite && isfinite(X.q[k]) && isfinite(X.qd[k]);
    if (!finite) st |= RKB_STATUS_NONFINITE;
  }
  __syncthreads();  // every thread is done with its RK4 columns
  if (active) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const int c = P.st[k].coord;
      smem[threadIdx.x * (2 * N + 1) + rkb_state_q(A.blocked, N, c)] = X.q[k];
      smem[threadIdx.x * (2 * N + 1) + rkb_state_qd(A.blocked, N, c)] = X.qd[k];
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int d = 0; d < A.n_dest; ++d) {
    const BatchView o = {A.xout[d] + A.row_offset * (2 * N), 2 * N, 1, A.blocked};
    tile_write_back<2 * N>(smem, o, tile_first, A.n_samples);
    if (active && A.status[d]) A.status[d][A.row_offset + i] = st;
  }
}

// A piecewise-constant control sequence in one launch: interval j integrates n_steps RK4 steps with the
// j-th input of the sample (num_int_dtnl_sys::get_next_state once per interval) and leaves its end state in
// slot j of the trajectory.  Kept apart from serial_rollout_kernel: the interval loop around the RK4 loop
// costs that kernel 0.8 % through extra spills.
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_BLOCK, RKB_MINBLOCKS(SHAPE, N, RKB_SMEM_ROLLOUT(N))) serial_rollout_seq_kernel(const __grid_constant__ SerialParams P, const RolloutSeqArgs A) {
  extern __shared__ double smem[];
  constexpr int SMS = RKB_BLOCK;
  const long long i = (long long)blockIdx.x * RKB_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  double* sm = smem + threadIdx.x;
  SerialState<N> X;
  load_state<N>(P, A.x0, A.u, i, X);
  int st = 0;
  if (A.half_step_nodes) {  // one interval whose input is sampled at every half step
    const NodeInputs nd = {A.u.p + i * A.u.si, A.u_sj, A.u.sk};
    st |= rk4_steps<N, FL, SHAPE, SMS>(P, X, A.dt, A.n_steps, sm, nullptr, &nd);
    if (A.traj.p) store_state<N>(P, A.traj, i * A.traj.si, X);
  } else {
#pragma unroll 1
    for (int j = 0; j < A.n_intervals; ++j) {
      if (j > 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const int in = P.st[k].input;
          X.u[k] = (in >= 0) ? A.u.p[i * A.u.si + j * A.u_sj + in * A.u.sk] : 0.0;
        }
      }
      st |= rk4_steps<N, FL, SHAPE, SMS>(P, X, A.dt, A.n_steps, sm);
      if (A.traj.p) store_state<N>(P, A.traj, i * A.traj.si + j * A.traj_sj, X);
    }
  }
  bool finite = true;
#pragma unroll
  for (int k = 0; k < N; ++k) finite = finite && isfinite(X.q[k]) && isfinite(X.qd[k]);
  store_state<N>(P, A.xout, i * A.xout.si, X);
  if (!finite) st |= RKB_STATUS_NONFINITE;
  if (A.status) A.status[i] = st;
}

// The closed-loop steering loop of steer_with_constant_control (examples/misc/MEAQR_topology.hpp:503-561) /
// IHAQR_topology::move_position_toward_impl (examples/misc/IHAQR_topology.hpp:349-378) in ONE launch: per
// control interval the goal-proximity test, the state feedback u = u_bias - G (x - x_goal) through
// get_bounded_input (IHAQR_topology.hpp:304-327, see rkb_steer.cu for the same law as a separate pass), then
// `substeps` RK4 steps with that input.  A sample whose loop has ended leaves; its warp carries on.
// CHK: the collision test of `with_collision_check = true` (MEAQR_topology.hpp:550-559: the interval is integrated, the
// state it ends on is tested, and only a free state is accepted — else the loop stops on the last free state).  The
// shipped kernels carry none; rkb_prox_jit.cu generates one per chain and set of proxy pairs (straight-line forward
// kinematics and finders, kte_prox_spec.cuh) and NVRTC compiles this kernel with it: checked steering in one launch.
struct NoSteerCheck {
  static constexpr bool enabled = false;
  static constexpr int fewer_blocks = 0;  // CTAs per SM given up for the test's registers
  template <int N> RKB_DEV static bool is_free(const SerialState<N>&) { return true; }
};
template <int N, int FL, shape_t SHAPE, class CHK = NoSteerCheck>
__global__ void __launch_bounds__(RKB_BLOCK, RKB_MINBLOCKS(SHAPE, N, RKB_SMEM_ROLLOUT(N)) - CHK::fewer_blocks) serial_steer_kernel(const __grid_constant__ SerialParams P, const __grid_constant__ SteerArgs A) {
  extern __shared__ double smem[];
  constexpr int SMS = RKB_BLOCK;
  constexpr int NX = 2 * N;
  const long long i = (long long)blockIdx.x * RKB_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  double* sm = smem + threadIdx.x;
  const int nu = A.nu;
  SerialState<N> X;
  {
    const ConstBatchView xv = {A.x0, NX, 1, A.blocked};
    const ConstBatchView uv = {A.x0, 1, 1, 0};
    load_state<N>(P, xv, uv, i, X);  // (X.u is set by the law below)
  }
  const double T = A.time_step;
  // the input applied last, per stage (stages without a driving actuator carry 0 and are skipped)
  double up[N];
  bool act[N];
#pragma unroll
  for (int s = 0; s < N; ++s) {
    const int in = P.st[s].input;
    act[s] = in >= 0 && in < nu;
    up[s] = act[s] ? A.u_prev[i * nu + in] : 0.0;
  }
  int st = 0, k = 0, collided = 0;
#pragma unroll 1
  for (; k < A.max_intervals; ++k) {
    double up_old[CHK::enabled ? N : 1];
    if (CHK::enabled) {
#pragma unroll
      for (int s = 0; s < N; ++s) up_old[s] = up[s];
    }
    // ---- the law: x - x_goal, distance, correction -G (x - x_goal), get_bounded_input ----------------------
    double dx[NX];
    double d2 = 0.0;
#pragma unroll
    for (int s = 0; s < N; ++s) {
      const int c = P.st[s].coord;
      const double dq = X.q[s] - A.goal[i * NX + rkb_state_q(A.blocked, N, c)];
      const double dd = X.qd[s] - A.goal[i * NX + rkb_state_qd(A.blocked, N, c)];
      dx[2 * s] = dq; dx[2 * s + 1] = dd;
      d2 = fma(dq, dq, d2); d2 = fma(dd, dd, d2);
    }
    if (!(sqrt(d2) > A.proximity)) break;  // MEAQR_topology.hpp:513-514
    double bias[N], corr[N], cur[N];
#pragma unroll
    for (int s = 0; s < N; ++s) {
      bias[s] = corr[s] = cur[s] = 0.0;
      if (!act[s]) continue;
      const int in = P.st[s].input;
      const double* G = A.gain + (i * nu + in) * NX;
      double acc = 0.0;
#pragma unroll
      for (int t = 0; t < N; ++t) {
        const int c = P.st[t].coord;
        acc = fma(G[rkb_state_q(A.blocked, N, c)], dx[2 * t], acc);
        acc = fma(G[rkb_state_qd(A.blocked, N, c)], dx[2 * t + 1], acc);
      }
      corr[s] = -acc;
      bias[s] = A.u_bias[i * nu + in];
    }
    if (k == 0 && !A.saturate_first) {  // MEAQR_topology.hpp:521-522: the first interval is not saturated
#pragma unroll
      for (int s = 0; s < N; ++s) if (act[s]) up[s] = bias[s] + corr[s];
    } else {                            // IHAQR_topology.hpp:304-327
      bool inside = true;
#pragma unroll
      for (int s = 0; s < N; ++s) {
        if (!act[s]) continue;
        const int in = P.st[s].input;
        if (A.have_u_box) { if (bias[s] < A.u_lo[in]) bias[s] = A.u_lo[in]; else if (bias[s] > A.u_hi[in]) bias[s] = A.u_hi[in]; }
        cur[s] = bias[s] + corr[s];
        if (A.have_u_box && ((cur[s] < A.u_lo[in]) || (cur[s] > A.u_hi[in]))) inside = false;
      }
      if (!inside) {
#pragma unroll 1
        for (int j = 0; j < 10; ++j) {
          bool ok = true;
#pragma unroll
          for (int s = 0; s < N; ++s) {
            if (!act[s]) continue;
            const int in = P.st[s].input;
            corr[s] *= 0.5; cur[s] -= corr[s];
            if ((cur[s] < A.u_lo[in]) || (cur[s] > A.u_hi[in])) ok = false;
          }
          if (ok) {
#pragma unroll
            for (int s = 0; s < N; ++s) { bias[s] = cur[s]; cur[s] += corr[s]; }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < N; ++s) {
        if (!act[s]) continue;
        const int in = P.st[s].input;
        double du = ((inside ? cur[s] : bias[s]) - up[s]) * (1.0 / T);
        if (A.have_du_box) { if (du < A.du_lo[in]) du = A.du_lo[in]; else if (du > A.du_hi[in]) du = A.du_hi[in]; }
        up[s] = up[s] + T * du;
      }
    }
#pragma unroll
    for (int s = 0; s < N; ++s) X.u[s] = up[s];
    // ---- one control interval -----------------------------------------------------------------------------
    double q0[CHK::enabled ? N : 1], qd0[CHK::enabled ? N : 1];
    if (CHK::enabled) {
#pragma unroll
      for (int s = 0; s < N; ++s) { q0[s] = X.q[s]; qd0[s] = X.qd[s]; }
    }
    st |= rk4_steps<N, FL, SHAPE, SMS>(P, X, A.dt, A.substeps, sm);
    if (CHK::enabled && !CHK::is_free(X)) {  // rejected: back on the last free state, with the input that led to it
#pragma unroll
      for (int s = 0; s < N; ++s) { X.q[s] = q0[s]; X.qd[s] = qd0[s]; up[s] = up_old[s]; }
      collided = 1;
      break;
    }
    if (A.traj) {
      const BatchView tv = {A.traj + (long long)k * NX, (long long)NX * A.max_intervals, 1, A.blocked};
      store_state<N>(P, tv, i * tv.si, X);
    }
  }
  bool finite = true;
#pragma unroll
  for (int s = 0; s < N; ++s) finite = finite && isfinite(X.q[s]) && isfinite(X.qd[s]);
  const BatchView ov = {A.xout, NX, 1, A.blocked};
  store_state<N>(P, ov, i * NX, X);
#pragma unroll
  for (int s = 0; s < N; ++s) if (act[s]) A.u_prev[i * nu + P.st[s].input] = up[s];  // u_prev = u_current (MEAQR_topology.hpp:553)
  A.n_done[i] = k;
  if (CHK::enabled && A.collided) A.collided[i] = collided;
  if (!finite) st |= RKB_STATUS_NONFINITE;
  if (A.status) A.status[i] = st;
}

// ---- the same three kernels with one sample on a pair of warps (see DuoCtx) ---------------------------------
// CTA = 128 threads = two pairs = 64 samples; warps 2p (forces) and 2p + 1 (mass matrix, solve) of pair p sit on
// different SM sub-partitions.  Lanes beyond the batch integrate a copy of the last sample (they must take part in the
// pair's barriers) and store nothing.  Results are those of the one-thread-per-sample kernels bit for bit: the same
// instruction sequences run, only on two warps.
#define RKB_DUO_BLOCK 128
#define RKB_SMEM_DUO(n) (6 * (n) * RKB_DUO_BLOCK + 2 * 2 * (n) * 32)  // doubles per CTA: RK4 columns + two exchange tiles

template <int N>
RKB_DEV bool duo_setup(double* smem, long long n_samples, DuoCtx& d, long long& i, bool& valid) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = warp >> 1;
  const long long first = ((long long)blockIdx.x * 2 + pair) * 32;
  if (first >= n_samples) return false;  // the whole pair has nothing to do (both of its warps leave)
  double* x = smem + 6 * N * RKB_DUO_BLOCK + pair * (2 * N * 32);
  d.role = warp & 1;
  d.xf = x + lane;
  d.xq = x + N * 32 + lane;
  d.bar_f = 1 + 2 * pair;
  d.bar_q = 2 + 2 * pair;
  i = first + lane;
  valid = i < n_samples;
  if (!valid) i = n_samples - 1;
  return true;
}

template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_DUO_BLOCK, 2) serial_rollout_duo_kernel(const __grid_constant__ SerialParams P, const RolloutArgs A) {
  extern __shared__ double smem[];
  DuoCtx duo;
  long long i;
  bool valid;
  if (!duo_setup<N>(smem, A.n_samples, duo, i, valid)) return;
  double* sm = smem + threadIdx.x;
  SerialState<N> X;
  {
    const long long i0 = A.x0_div > 1 ? i / A.x0_div : i;
    ConstBatchView xv = A.x0;
    xv.p += i0 * xv.si - i * xv.si;
    load_state<N>(P, xv, A.u, i, X);
  }
  int st = rk4_steps<N, FL, SHAPE, RKB_DUO_BLOCK, true>(P, X, A.dt, A.n_steps, sm, &duo);
  if (!valid) return;
  if (duo.role == 0) {
    store_state<N>(P, A.xout, i * A.xout.si, X);
    if (A.traj.p) store_state<N>(P, A.traj, i * A.traj.si, X);
  } else {
    bool finite = true;
#pragma unroll
    for (int k = 0; k < N; ++k) finite = finite && isfinite(X.q[k]) && isfinite(X.qd[k]);
    if (!finite) st |= RKB_STATUS_NONFINITE;
    if (A.status) A.status[i] = A.status_or ? (A.status[i] | st) : st;
  }
}

template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_DUO_BLOCK, 2) serial_rollout_seq_duo_kernel(const __grid_constant__ SerialParams P, const RolloutSeqArgs A) {
  extern __shared__ double smem[];
  DuoCtx duo;
  long long i;
  bool valid;
  if (!duo_setup<N>(smem, A.n_samples, duo, i, valid)) return;
  double* sm = smem + threadIdx.x;
  SerialState<N> X;
  load_state<N>(P, A.x0, A.u, i, X);
  int st = 0;
#pragma unroll 1
  for (int j = 0; j < A.n_intervals; ++j) {
    if (j > 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int in = P.st[k].input;
        X.u[k] = (in >= 0) ? A.u.p[i * A.u.si + j * A.u_sj + in * A.u.sk] : 0.0;
      }
    }
    st |= rk4_steps<N, FL, SHAPE, RKB_DUO_BLOCK, true>(P, X, A.dt, A.n_steps, sm, &duo);
    if (valid && duo.role == 0 && A.traj.p) store_state<N>(P, A.traj, i * A.traj.si + j * A.traj_sj, X);
  }
  if (!valid) return;
  if (duo.role == 0) {
    store_state<N>(P, A.xout, i * A.xout.si, X);
  } else {
    bool finite = true;
#pragma unroll
    for (int k = 0; k < N; ++k) finite = finite && isfinite(X.q[k]) && isfinite(X.qd[k]);
    if (!finite) st |= RKB_STATUS_NONFINITE;
    if (A.status) A.status[i] = st;
  }
}

// The steering loop: samples of a warp stop after different numbers of intervals, but the pair's barriers need every
// lane, so a finished lane keeps riding along with a zero step (its state does not move) until the whole warp is done;
// both warps of the pair hold the same states and therefore agree on when that is.
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_DUO_BLOCK, 2) serial_steer_duo_kernel(const __grid_constant__ SerialParams P, const __grid_constant__ SteerArgs A) {
  extern __shared__ double smem[];
  constexpr int NX = 2 * N;
  DuoCtx duo;
  long long i;
  bool valid;
  if (!duo_setup<N>(smem, A.n_samples, duo, i, valid)) return;
  double* sm = smem + threadIdx.x;
  const int nu = A.nu;
  SerialState<N> X;
  {
    const ConstBatchView xv = {A.x0, NX, 1, A.blocked};
    const ConstBatchView uv = {A.x0, 1, 1, 0};
    load_state<N>(P, xv, uv, i, X);
  }
  const double T = A.time_step;
  double up[N];
  bool act[N];
#pragma unroll
  for (int s = 0; s < N; ++s) {
    const int in = P.st[s].input;
    act[s] = in >= 0 && in < nu;
    up[s] = act[s] ? A.u_prev[i * nu + in] : 0.0;
  }
  int st = 0, n_done = 0;
  bool live = true;
#pragma unroll 1
  for (int k = 0; k < A.max_intervals; ++k) {
    double dx[NX];
    double d2 = 0.0;
#pragma unroll
    for (int s = 0; s < N; ++s) {
      const int c = P.st[s].coord;
      const double dq = X.q[s] - A.goal[i * NX + rkb_state_q(A.blocked, N, c)];
      const double dd = X.qd[s] - A.goal[i * NX + rkb_state_qd(A.blocked, N, c)];
      dx[2 * s] = dq; dx[2 * s + 1] = dd;
      d2 = fma(dq, dq, d2); d2 = fma(dd, dd, d2);
    }
    if (live && !(sqrt(d2) > A.proximity)) live = false;  // MEAQR_topology.hpp:513-514
    if (!__any_sync(0xffffffffu, live)) break;
    double bias[N], corr[N], cur[N], un[N];
#pragma unroll
    for (int s = 0; s < N; ++s) {
      bias[s] = corr[s] = cur[s] = 0.0;
      un[s] = up[s];
      if (!act[s]) continue;
      const int in = P.st[s].input;
      const double* G = A.gain + (i * nu + in) * NX;
      double acc = 0.0;
#pragma unroll
      for (int t = 0; t < N; ++t) {
        const int c = P.st[t].coord;
        acc = fma(G[rkb_state_q(A.blocked, N, c)], dx[2 * t], acc);
        acc = fma(G[rkb_state_qd(A.blocked, N, c)], dx[2 * t + 1], acc);
      }
      corr[s] = -acc;
      bias[s] = A.u_bias[i * nu + in];
    }
    if (k == 0 && !A.saturate_first) {
#pragma unroll
      for (int s = 0; s < N; ++s) if (act[s]) un[s] = bias[s] + corr[s];
    } else {  // IHAQR_topology.hpp:304-327
      bool inside = true;
#pragma unroll
      for (int s = 0; s < N; ++s) {
        if (!act[s]) continue;
        const int in = P.st[s].input;
        if (A.have_u_box) { if (bias[s] < A.u_lo[in]) bias[s] = A.u_lo[in]; else if (bias[s] > A.u_hi[in]) bias[s] = A.u_hi[in]; }
        cur[s] = bias[s] + corr[s];
        if (A.have_u_box && ((cur[s] < A.u_lo[in]) || (cur[s] > A.u_hi[in]))) inside = false;
      }
      if (!inside) {
#pragma unroll 1
        for (int j = 0; j < 10; ++j) {
          bool ok = true;
#pragma unroll
          for (int s = 0; s < N; ++s) {
            if (!act[s]) continue;
            const int in = P.st[s].input;
            corr[s] *= 0.5; cur[s] -= corr[s];
            if ((cur[s] < A.u_lo[in]) || (cur[s] > A.u_hi[in])) ok = false;
          }
          if (ok) {
#pragma unroll
            for (int s = 0; s < N; ++s) { bias[s] = cur[s]; cur[s] += corr[s]; }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < N; ++s) {
        if (!act[s]) continue;
        const int in = P.st[s].input;
        double du = ((inside ? cur[s] : bias[s]) - up[s]) * (1.0 / T);
        if (A.have_du_box) { if (du < A.du_lo[in]) du = A.du_lo[in]; else if (du > A.du_hi[in]) du = A.du_hi[in]; }
        un[s] = up[s] + T * du;
      }
    }
#pragma unroll
    for (int s = 0; s < N; ++s) {
      if (live) up[s] = un[s];
      X.u[s] = up[s];
    }
    const int s_k = rk4_steps<N, FL, SHAPE, RKB_DUO_BLOCK, true>(P, X, live ? A.dt : 0.0, A.substeps, sm, &duo);
    if (live) {
      st |= s_k;
      n_done = k + 1;
      if (valid && duo.role == 0 && A.traj) {
        const BatchView tv = {A.traj + (long long)k * NX, (long long)NX * A.max_intervals, 1, A.blocked};
        store_state<N>(P, tv, i * tv.si, X);
      }
    }
  }
  if (!valid) return;
  if (duo.role == 0) {
    const BatchView ov = {A.xout, NX, 1, A.blocked};
    store_state<N>(P, ov, i * NX, X);
#pragma unroll
    for (int s = 0; s < N; ++s) if (act[s]) A.u_prev[i * nu + P.st[s].input] = up[s];
    A.n_done[i] = n_done;
  } else {
    bool finite = true;
#pragma unroll
    for (int s = 0; s < N; ++s) finite = finite && isfinite(X.q[s]) && isfinite(X.qd[s]);
    if (!finite) st |= RKB_STATUS_NONFINITE;
    if (A.status) A.status[i] = st;
  }
}

// Any explicit one-step scheme given as an RkTable (Euler, midpoint, RK5 — and RK4, which the kernel
// above does faster).  Per-thread shared-memory column: w (2N), then k_0 .. k_{stages-1} (2N each).
template <int N, int FL, shape_t SHAPE>
__global__ void __launch_bounds__(RKB_BLOCK, RKB_MINBLOCKS(SHAPE, N, RKB_SMEM_ROLLOUT(N))) serial_rollout_rk_kernel(const __grid_constant__ SerialParams P, const RolloutArgs A,
                                                                                                       const __grid_constant__ RkTable T) {
  extern __shared__ double smem[];
  constexpr int SMS = RKB_BLOCK;
  const long long i = (long long)blockIdx.x * RKB_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  if (A.active && !A.active[i]) return;
  double* sm = smem + threadIdx.x;
  double* sw = sm;
  double* sk = sw + 2 * N * SMS;
  SerialState<N> X;
  {
    const long long i0 = A.x0_div > 1 ? i / A.x0_div : i;
    ConstBatchView xv = A.x0;
    xv.p += i0 * xv.si - i * xv.si;
    load_state<N>(P, xv, A.u, i, X);
  }
  const double dt = A.dt;
  int st = 0;
#pragma unroll 1
  for (int step = 0; step < A.n_steps; ++step) {
#pragma unroll
    for (int k = 0; k < N; ++k) { sw[(2 * k) * SMS] = X.q[k]; sw[(2 * k + 1) * SMS] = X.qd[k]; }
#pragma unroll 1
    for (int s = 0; s < T.stages; ++s) {
      double qdd[N];
      st |= serial_accel<N, FL, SHAPE, SMS>(P, X, qdd, sm);
      double* ks = sk + s * (2 * N * SMS);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        ks[(2 * k) * SMS] = X.qd[k] * dt; ks[(2 * k + 1) * SMS] = qdd[k] * dt;
        X.q[k] = sw[(2 * k) * SMS]; X.qd[k] = sw[(2 * k + 1) * SMS];
      }
#pragma unroll 1
      for (int j = 0; j <= s; ++j) {
        const double c = T.c[s][j];
        const double* kj = sk + j * (2 * N * SMS);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          X.q[k] = fma(c, kj[(2 * k) * SMS], X.q[k]);
          X.qd[k] = fma(c, kj[(2 * k + 1) * SMS], X.qd[k]);
        }
      }
    }
  }
  bool finite = true;
#pragma unroll
  for (int k = 0; k < N; ++k) finite = finite && isfinite(X.q[k]) && isfinite(X.qd[k]);
  store_state<N>(P, A.xout, i * A.xout.si, X);
  if (A.traj.p) store_state<N>(P, A.traj, i * A.traj.si, X);
  if (!finite) st |= RKB_STATUS_NONFINITE;
  if (A.status) A.status[i] = A.status_or ? (A.status[i] | st) : st;
}

}  // namespace rkb
#endif
