// Test infrastructure (never shipped, never loaded by the product): compiles the DEVICE source
// reak_b200/csrc/kte_proximity.cuh for the host, so that the CPU suite can hold the finders the GPU
// runs against the compiled reference (oracle/_ref) without a GPU.  Built by tests/test_proximity.py
// into a temporary directory with g++.
#include <cmath>
#include <cstdint>
#include "../../include/reak_b200.h"
#include "../../reak_b200/csrc/rkb_types.h"

#define GD static inline
using std::fabs;
using std::sqrt;
#include "../../reak_b200/csrc/kte_math.cuh"   // the helpers kte_generic.cu has ahead of the two headers
#include "../../reak_b200/csrc/kte_proximity.cuh"
#include "../../reak_b200/csrc/kte_proximity2d.cuh"

// program: the ProxProgram rkb_proxy_create lowers (read back through rkb_proxy_program, a test hook of the
// product library); frames: [n_frames][7] world position + quaternion of the chain frames.
extern "C" int prox_host_min_distance(const ProxProgram* P, const double* frames, int n_frames, double* dist, double* pts) {
  Pose fr[RKB_GEN_MAX_FRAMES];  // the kernel's slot array: only the frames the program asks to keep
  for (int f = 0; f < n_frames && f < RKB_GEN_MAX_FRAMES; ++f) {
    if (P->slot_of[f] < 0) continue;
    const double* v = frames + 7 * f;
    Pose& S = fr[P->slot_of[f]];
    S.p = v3(v[0], v[1], v[2]);
    S.q.w = v[3]; S.q.x = v[4]; S.q.y = v[5]; S.q.z = v[6];
  }
  ProxRecord R;
  const int best = prox_min_distance(*P, fr, true, R);
  *dist = R.d;
  pts[0] = R.p1.x; pts[1] = R.p1.y; pts[2] = R.p1.z; pts[3] = R.p2.x; pts[4] = R.p2.y; pts[5] = R.p2.z;
  return best;
}
extern "C" int prox_host_program_size(void) { return (int)sizeof(ProxProgram); }

// proxy_query_pair_3D::gatherCollisionPoints through the same device source: records [max_records][7], finder [max_records]
extern "C" int prox_host_gather(const ProxProgram* P, const double* frames, int n_frames, int max_records, double* records, int32_t* finder) {
  Pose fr[RKB_GEN_MAX_FRAMES];
  for (int f = 0; f < n_frames && f < RKB_GEN_MAX_FRAMES; ++f) {
    if (P->slot_of[f] < 0) continue;
    const double* v = frames + 7 * f;
    Pose& S = fr[P->slot_of[f]];
    S.p = v3(v[0], v[1], v[2]);
    S.q.w = v[3]; S.q.x = v[4]; S.q.y = v[5]; S.q.z = v[6];
  }
  return prox_gather_collisions(*P, fr, max_records, [&](int r, int f, const ProxRecord& R) {
    double* o = records + 7 * r;
    o[0] = R.d; o[1] = R.p1.x; o[2] = R.p1.y; o[3] = R.p1.z; o[4] = R.p2.x; o[5] = R.p2.y; o[6] = R.p2.z;
    finder[r] = f;
  });
}

// ---- planar models: frames [n_frames][7] = (x, y, -, cos, sin, -, -) as rkb_frames / the reference report them ----
static void load_slots2(const ProxProgram* P, const double* frames, int n_frames, Pose2* fr) {
  for (int f = 0; f < n_frames && f < RKB_GEN_MAX_FRAMES; ++f) {
    if (P->slot_of[f] < 0) continue;
    const double* v = frames + 7 * f;
    Pose2& S = fr[P->slot_of[f]];
    S.p = v2(v[0], v[1]);
    S.R.c = v[3]; S.R.s = v[4];
  }
}
extern "C" int prox2d_host_min_distance(const ProxProgram* P, const double* frames, int n_frames, double* dist, double* pts) {
  Pose2 fr[RKB_GEN_MAX_FRAMES];
  load_slots2(P, frames, n_frames, fr);
  ProxRecord2 R;
  const int best = prox_min_distance2(*P, fr, R);
  *dist = R.d;
  pts[0] = R.p1.x; pts[1] = R.p1.y; pts[2] = 0.0; pts[3] = R.p2.x; pts[4] = R.p2.y; pts[5] = 0.0;
  return best;
}
extern "C" int prox2d_host_gather(const ProxProgram* P, const double* frames, int n_frames, int max_records, double* records, int32_t* finder) {
  Pose2 fr[RKB_GEN_MAX_FRAMES];
  load_slots2(P, frames, n_frames, fr);
  return prox_gather_collisions2(*P, fr, max_records, [&](int r, int f, const ProxRecord2& R) {
    double* o = records + 7 * r;
    o[0] = R.d; o[1] = R.p1.x; o[2] = R.p1.y; o[3] = 0.0; o[4] = R.p2.x; o[5] = R.p2.y; o[6] = 0.0;
    finder[r] = f;
  });
}
