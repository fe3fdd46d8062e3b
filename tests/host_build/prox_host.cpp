// Test infrastructure (never shipped, never loaded by the product): compiles the DEVICE source
// reak_b200/csrc/kte_proximity.cuh for the host, so that the CPU suite can hold the finders the GPU
// runs against the compiled reference (oracle/_ref) without a GPU.  Built by tests/test_proximity.py
// into a temporary directory with g++.
#include <cmath>
#include <cstdint>
#include "../../include/reak_b200.h"
#include "../../reak_b200/csrc/rkb_types.h"

#define GD static inline
// the helpers kte_generic.cu defines ahead of including the header
struct V3 { double x, y, z; };
GD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
GD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Q4 { double w, x, y, z; };
GD Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = b.w * a.w - b.x * a.x - b.y * a.y - b.z * a.z;
  r.x = b.w * a.x + b.z * a.y - b.y * a.z + b.x * a.w;
  r.y = b.w * a.y - b.z * a.x + b.x * a.z + b.y * a.w;
  r.z = b.w * a.z + b.y * a.x - b.x * a.y + b.z * a.w;
  return r;
}
GD Q4 qconj(Q4 a) { Q4 r; r.w = a.w; r.x = -a.x; r.y = -a.y; r.z = -a.z; return r; }
using std::fabs;
using std::sqrt;
#include "../../reak_b200/csrc/kte_proximity.cuh"

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
