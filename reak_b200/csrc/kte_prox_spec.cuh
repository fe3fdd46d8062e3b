// kte_prox_spec.cuh — what a proximity kernel generated for ONE chain and ONE proxy pair is made of
// (rkb_prox_jit.cu writes the source, NVRTC compiles it; tests/host_build compiles the same source for the host).
//
// The interpreter version (generic_proximity_kernel in kte_generic.cu) reads the element list and the shape list
// at run time: 70 % of the instructions it issues are not FP64 (kind decoding, selects between the two shapes of
// a finder, copies of poses, local-memory frames).  A generated kernel is straight-line code: the forward
// kinematics with the chain's constants as literals (axis-aligned joints and links multiply by one column
// only), world-fixed shapes as literal poses, every finder call with its argument order and dimensions resolved.
// Same arithmetic per operation as motion_pose / prox_shape_pose / prox_compute, minus the products with
// literal zeros; results agree with the interpreter kernel to rounding.
#pragma once

#ifndef RKB_PROX_SPEC_HOST
#include "rkb_types.h"
#define GD __device__ __forceinline__
#define RKB_PROX_SPEC_TABLE __constant__
#ifndef INFINITY
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#endif

namespace {
#include "kte_math.cuh"
#include "kte_proximity.cuh"

// columns of qrot(q) (quaternion::getRotMat, rotations_3D.hpp:986-1000): what mul(qrot(q), c e_k) reads
template <int K>
GD V3 qrot_col(Q4 q) {
  if (K == 0) return v3(1.0 - 2.0 * q.y * q.y - 2.0 * q.z * q.z, 2.0 * q.x * q.y + 2.0 * q.w * q.z, 2.0 * q.x * q.z - 2.0 * q.w * q.y);
  if (K == 1) return v3(2.0 * q.x * q.y - 2.0 * q.w * q.z, 1.0 - 2.0 * q.x * q.x - 2.0 * q.z * q.z, 2.0 * q.w * q.x + 2.0 * q.y * q.z);
  return v3(2.0 * q.w * q.y + 2.0 * q.x * q.z, 2.0 * q.y * q.z - 2.0 * q.w * q.x, 1.0 - 2.0 * q.x * q.x - 2.0 * q.y * q.y);
}
// qmul(a, (c, s e_K)) without the products with zero (rotations_3D.hpp:1093-1098)
template <int K>
GD Q4 qmul_axis(Q4 a, double c, double s) {
  Q4 r;
  if (K == 0) { r.w = c * a.w - s * a.x; r.x = c * a.x + s * a.w; r.y = c * a.y + s * a.z; r.z = c * a.z - s * a.y; }
  else if (K == 1) { r.w = c * a.w - s * a.y; r.x = c * a.x - s * a.z; r.y = c * a.y + s * a.w; r.z = c * a.z + s * a.x; }
  else { r.w = c * a.w - s * a.z; r.x = c * a.x + s * a.y; r.y = c * a.y - s * a.x; r.z = c * a.z + s * a.w; }
  return r;
}
// qrotv(Q, c e_K) without the products with zero (rotations_3D.hpp:1137-1151)
template <int K>
GD V3 qrotv_axis(Q4 Q, double c) {
  const double t0 = Q.w * Q.x, t1 = Q.w * Q.y, t2 = Q.w * Q.z, t3 = -Q.x * Q.x, t4 = Q.x * Q.y, t5 = Q.x * Q.z, t6 = -Q.y * Q.y,
               t7 = Q.y * Q.z, t8 = -Q.z * Q.z;
  if (K == 0) return v3(2.0 * ((t6 + t8) * c) + c, 2.0 * ((t2 + t4) * c), 2.0 * ((t5 - t1) * c));
  if (K == 1) return v3(2.0 * ((t4 - t2) * c), 2.0 * ((t3 + t8) * c) + c, 2.0 * ((t0 + t7) * c));
  return v3(2.0 * ((t1 + t5) * c), 2.0 * ((t7 - t0) * c), 2.0 * ((t3 + t6) * c) + c);
}
GD Q4 q4(double w, double x, double y, double z) { Q4 r; r.w = w; r.x = x; r.y = y; r.z = z; return r; }
GD Pose pose_of(V3 p, Q4 q) { Pose r; r.p = p; r.q = q; return r; }
GD SPose spose_of(V3 p, Q4 q) { SPose r; r.p = p; rot_table(q, r.m); return r; }
GD SPose spose_lit(double px, double py, double pz, double m0, double m1, double m2, double m3, double m4, double m5, double m6, double m7,
                   double m8) {
  SPose r;
  r.p = v3(px, py, pz);
  r.m[0] = m0; r.m[1] = m1; r.m[2] = m2; r.m[3] = m3; r.m[4] = m4; r.m[5] = m5; r.m[6] = m6; r.m[7] = m7; r.m[8] = m8;
  return r;
}
}  // namespace

#ifndef RKB_PROX_SPEC_HOST
// The kernel of a generated source: NC coordinates, NFREE free joints (0 or 1), the search in prox_spec<false> — distance
// and finder index, the buffers of generic_proximity_kernel.  (A query that also wants the two points runs on the
// interpreter kernel: every thread of a warp may hold a different winner, and evaluating its record needs the shapes picked
// by index — measured 0.61 ms there against 0.70 ms (a switch over all finders) and 6 ms (poses in a per-thread array) here.
// prox_spec<true> is what the host build of the CPU suite checks the points of.)
#define RKB_PROX_SPEC_KERNELS(NC, NFREE, MINB)                                                                          \
  extern "C" __global__ void __launch_bounds__(128, MINB) rkb_prox_spec_d(const EvalArgs A) {                           \
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;                                                      \
    if (i >= A.n_samples) return;                                                                                       \
    double q[NC > 0 ? NC : 1];                                                                                          \
    _Pragma("unroll") for (int c = 0; c < NC; ++c) q[c] = A.x.p[i * A.x.si + rkb_state_q(A.x.blocked, NC, c) * A.x.sk]; \
    Pose freec = pose_of(v3(0.0, 0.0, 0.0), q4(1.0, 0.0, 0.0, 0.0));                                                    \
    if (NFREE) {                                                                                                        \
      double s[7];                                                                                                      \
      _Pragma("unroll") for (int k = 0; k < 7; ++k) s[k] = A.x.p[i * A.x.si + (2 * NC + k) * A.x.sk];                   \
      const double nq = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);                                    \
      freec = pose_of(v3(s[0], s[1], s[2]), q4(s[3] / nq, s[4] / nq, s[5] / nq, s[6] / nq));                            \
    }                                                                                                                   \
    ProxRecord R;                                                                                                       \
    const int best = prox_spec<false>(q, freec, R);                                                                     \
    A.out.p[i * A.out.si] = R.d;                                                                                        \
    if (A.status) A.status[i] = best;                                                                                   \
  }
#endif
