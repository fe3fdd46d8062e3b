// Test infrastructure (never shipped, never loaded by the product): lets g++ compile, for the host, the CUDA source
// rkb_proxy_source() returns — the text rkb_proxy_specialize hands to NVRTC — so that the CPU suite can hold the
// generated forward kinematics and finder sequence against the compiled reference without a GPU.
//   g++ -include tests/host_build/prox_spec_host.h -I reak_b200/csrc -x c++ generated.cu
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE 1
#endif
#include <cmath>
#include <cstdint>
#include "rkb_types.h"
#define RKB_PROX_SPEC_HOST 1
#define GD static inline
#define RKB_PROX_SPEC_TABLE static const
using std::fabs;
using std::sqrt;
// q: the chain's coordinates; freec7: position + quaternion state of the free joint (ignored without one)
#define RKB_PROX_SPEC_KERNELS(NC, NFREE, MINB)                                                                              \
  extern "C" int prox_spec_host(const double* q, const double* freec7, int with_points, double* dist, double* pts) {         \
    Pose freec = pose_of(v3(0.0, 0.0, 0.0), q4(1.0, 0.0, 0.0, 0.0));                                                         \
    if (NFREE) {                                                                                                             \
      const double* s = freec7;                                                                                              \
      const double nq = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);                                         \
      freec = pose_of(v3(s[0], s[1], s[2]), q4(s[3] / nq, s[4] / nq, s[5] / nq, s[6] / nq));                                 \
    }                                                                                                                        \
    ProxRecord R;                                                                                                            \
    const int best = with_points ? prox_spec<true>(q, freec, R) : prox_spec<false>(q, freec, R);                             \
    *dist = R.d;                                                                                                             \
    pts[0] = R.p1.x; pts[1] = R.p1.y; pts[2] = R.p1.z; pts[3] = R.p2.x; pts[4] = R.p2.y; pts[5] = R.p2.z;                    \
    return best;                                                                                                             \
  }                                                                                                                          \
  extern "C" int prox_spec_host_dims(int* nc, int* nfree, int* minb) { *nc = NC; *nfree = NFREE; *minb = MINB; return 0; }
