// rkb_prox_jit.cu — writes the CUDA source of a proximity kernel for ONE chain and ONE proxy pair (host code only).
//
// The source is straight-line code over the prelude kte_prox_spec.cuh: position-level forward kinematics
// (revolute_joint.cpp:121-131, prismatic_joint.cpp:129-140, free_joints.cpp:127, rigid_link.cpp:156 ->
// pose_3D::addBefore) with the chain's constants as hexadecimal literals, the world pose of every shape
// (pose_3D::getGlobalPose, pose_3D.hpp:102-110; world-fixed shapes are literals), and the finders of
// createProxFinderList (proxy_query_model.cpp:212-384) in order, each with the bounding-sphere test of
// findMinimumDistance (proxy_query_model.cpp:388-412) in front of it.  rkb_jit.cu compiles it with NVRTC.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rkb_internal.h"

namespace {

struct Emit {
  std::string s;
  void f(const char* fmt, ...) {  // (any length: a line is never cut short)
    char buf[1024];
    va_list ap, ap2;
    va_start(ap, fmt);
    va_copy(ap2, ap);
    const int n = std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n >= 0 && n < (int)sizeof buf) s += buf;
    else if (n > 0) {
      std::vector<char> big((size_t)n + 1);
      std::vsnprintf(big.data(), big.size(), fmt, ap2);
      s += big.data();
    }
    va_end(ap2);
  }
};

// A chain or shape constant as the kernel reads it.  Doubles whose low word is zero (0, +-1, 0.5, 1.5, ...) fit the
// 32-bit immediate of an FP64 instruction and stay literals, which also lets the compiler fold them; the others go to a
// __constant__ table of the generated source and are read as constant-bank operands — a literal would cost two moves
// into registers wherever it is used (measured on the first generated kernels: a third of all instructions).
std::vector<double>* g_pool = nullptr;  // the table of the source being written (rkb_prox_source is serialised by g_gen_mu)
std::mutex g_gen_mu;
std::string lit(double v) {
  char buf[64];
  if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
  if (v == 1.0) return "1.0";
  if (v == -1.0) return "-1.0";
  unsigned long long bits;
  std::memcpy(&bits, &v, sizeof bits);
  if ((bits & 0xffffffffull) == 0 || !g_pool) {
    std::snprintf(buf, sizeof buf, "%a", v);
    return buf;
  }
  size_t k = 0;
  for (; k < g_pool->size(); ++k)
    if (std::memcmp(&(*g_pool)[k], &v, sizeof v) == 0) break;
  if (k == g_pool->size()) g_pool->push_back(v);
  std::snprintf(buf, sizeof buf, "C[%d]", (int)k);
  return buf;
}
std::string v3lit(const double* p) { return "v3(" + lit(p[0]) + ", " + lit(p[1]) + ", " + lit(p[2]) + ")"; }
std::string q4lit(const double* q) { return "q4(" + lit(q[0]) + ", " + lit(q[1]) + ", " + lit(q[2]) + ", " + lit(q[3]) + ")"; }

// index of the only non-zero component of a 3-vector, -1 when there are none, -2 when there are several
int single_axis(const double* v) {
  int k = -1;
  for (int i = 0; i < 3; ++i)
    if (v[i] != 0.0) k = k == -1 ? i : -2;
  return k;
}

// q * (c, s e_k) or q * (w, x, y, z), as text
std::string qmul_const(const std::string& q, const double* b) {
  if (b[0] == 1.0 && b[1] == 0.0 && b[2] == 0.0 && b[3] == 0.0) return q;
  const int k = single_axis(b + 1);
  char buf[256];
  if (k >= 0) {
    std::snprintf(buf, sizeof buf, "qmul_axis<%d>(%s, %s, %s)", k, q.c_str(), lit(b[0]).c_str(), lit(b[1 + k]).c_str());
    return buf;
  }
  return "qmul(" + q + ", " + q4lit(b) + ")";
}
// p + R(q) v for a literal v, as text
std::string offset_const(const std::string& p, const std::string& q, const double* v) {
  const int k = single_axis(v);
  char buf[256];
  if (k == -1) return p;
  if (k >= 0) {
    std::snprintf(buf, sizeof buf, "%s + %s * qrot_col<%d>(%s)", p.c_str(), lit(v[k]).c_str(), k, q.c_str());
    return buf;
  }
  return p + " + mul(qrot(" + q + "), " + v3lit(v) + ")";
}

void unit_axis_host(const double* a, double* an) {  // axis_angle ctor, rotations_3D.hpp:1962-1974
  const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  if (n > 0.0000001) { an[0] = a[0] / n; an[1] = a[1] / n; an[2] = a[2] / n; }
  else { an[0] = 1.0; an[1] = 0.0; an[2] = 0.0; }
}

// the call of the finder of shapes a (model 1) and b (model 2): the kind listed first takes the first slot
std::string finder_call(const ProxShape& a, const std::string& Pa, const ProxShape& b, const std::string& Pb, const char* pts) {
  const bool swap = b.kind < a.kind;
  const ProxShape& s1 = swap ? b : a;
  const ProxShape& s2 = swap ? a : b;
  const std::string& P1 = swap ? Pb : Pa;
  const std::string& P2 = swap ? Pa : Pb;
  const std::string d1x = lit(s1.dims[0]), d1y = lit(s1.dims[1]), d2x = lit(s2.dims[0]), d2y = lit(s2.dims[1]);
  const std::string d1 = v3lit(s1.dims), d2 = v3lit(s2.dims);
  const std::string t = std::string("<") + pts + ">(";
  if (s1.kind == RKB_SHAPE_PLANE) {
    if (s2.kind == RKB_SHAPE_PLANE) return "prox_plane_plane" + t + P1 + ", " + d1 + ", " + P2 + ", " + d2 + ")";
    if (s2.kind == RKB_SHAPE_SPHERE) return "prox_plane_sphere" + t + P1 + ", " + P2 + ", " + d2x + ")";
    if (s2.kind == RKB_SHAPE_CCYLINDER) return "prox_plane_ccylinder" + t + P1 + ", " + P2 + ", " + d2x + ", " + d2y + ")";
    if (s2.kind == RKB_SHAPE_CYLINDER) return "prox_plane_cylinder" + t + P1 + ", " + P2 + ", " + d2x + ", " + d2y + ")";
    return "prox_plane_box" + t + P1 + ", " + P2 + ", " + d2 + ")";
  }
  if (s1.kind == RKB_SHAPE_SPHERE) {
    if (s2.kind == RKB_SHAPE_SPHERE) return "prox_sphere_sphere" + t + P1 + ", " + d1x + ", " + P2 + ", " + d2x + ")";
    if (s2.kind == RKB_SHAPE_CCYLINDER) return "prox_sphere_ccylinder" + t + P1 + ", " + d1x + ", " + P2 + ", " + d2x + ", " + d2y + ")";
    if (s2.kind == RKB_SHAPE_CYLINDER) return "prox_sphere_cylinder" + t + P1 + ", " + d1x + ", " + P2 + ", " + d2x + ", " + d2y + ")";
    return "prox_sphere_box" + t + P1 + ", " + d1x + ", " + P2 + ", " + d2 + ")";
  }
  if (s2.kind == RKB_SHAPE_CCYLINDER) return "prox_ccylinder_ccylinder" + t + P1 + ", " + d1x + ", " + d1y + ", " + P2 + ", " + d2x + ", " + d2y + ")";
  return "prox_ccylinder_box" + t + P1 + ", " + d1x + ", " + d1y + ", " + P2 + ", " + d2 + ")";
}
bool has_finder(int ka, int kb) {  // proxy_query_model.cpp:212-384
  const int lo = ka < kb ? ka : kb, hi = ka < kb ? kb : ka;
  if (lo == RKB_SHAPE_PLANE || lo == RKB_SHAPE_SPHERE) return true;
  if (lo == RKB_SHAPE_CCYLINDER) return hi == RKB_SHAPE_CCYLINDER || hi == RKB_SHAPE_BOX;
  return false;
}

}  // namespace

namespace {
// `namespace <ns> { tables, prox_spec<PTS>(q, freec, bestR) }` of one chain and pair (inside the caller's anonymous
// namespace); empty when the chain cannot be written as straight-line code.  g_gen_mu held.
std::string prox_body(const GenericProgram& G, const ProxProgram& P, const char* ns) {
  if (G.dim != 3) return std::string();  // planar models run on the interpreter kernel (kte_proximity2d.cuh)
  std::vector<double> pool;
  g_pool = &pool;
  struct Reset { ~Reset() { g_pool = nullptr; } } reset;
  Emit e;
  e.f("template <bool PTS>\nGD int prox_spec(const double* q, const Pose& freec, ProxRecord& bestR) {\n");
  // ---- forward kinematics: one Pose per chain frame, named after the frame
  bool written[RKB_GEN_MAX_FRAMES] = {};
  e.f("  const Pose F%d = pose_of(%s, %s);\n", G.base_frame, v3lit(G.base).c_str(), q4lit(G.base + 3).c_str());
  written[G.base_frame] = true;
  for (int k = 0; k < G.n_elements; ++k) {
    const GenericElement& E = G.el[k];
    if (E.kind != RKB_REVOLUTE_3D && E.kind != RKB_PRISMATIC_3D && E.kind != RKB_RIGID_LINK_3D && E.kind != RKB_FREE_3D) continue;
    if (E.fa < 0 || E.fa >= RKB_GEN_MAX_FRAMES || E.fb < 0 || E.fb >= RKB_GEN_MAX_FRAMES || !written[E.fa] || written[E.fb]) return std::string();
    char Bp[20], Bq[20];
    std::snprintf(Bp, sizeof Bp, "F%d.p", E.fa);
    std::snprintf(Bq, sizeof Bq, "F%d.q", E.fa);
    if (E.kind == RKB_REVOLUTE_3D) {
      double an[3];
      unit_axis_host(E.p, an);
      const int ax = single_axis(an);
      e.f("  double sh%d, ch%d;\n  sincos(0.5 * q[%d], &sh%d, &ch%d);\n", k, k, E.coord, k, k);
      if (ax >= 0 && (an[ax] == 1.0 || an[ax] == -1.0))
        e.f("  const Pose F%d = pose_of(%s, qmul_axis<%d>(%s, ch%d, %ssh%d));\n", E.fb, Bp, ax, Bq, k, an[ax] < 0.0 ? "-" : "", k);
      else
        e.f("  const Pose F%d = pose_of(%s, qmul(%s, q4(ch%d, %s * sh%d, %s * sh%d, %s * sh%d)));\n", E.fb, Bp, Bq, k, lit(an[0]).c_str(), k,
            lit(an[1]).c_str(), k, lit(an[2]).c_str(), k);
    } else if (E.kind == RKB_PRISMATIC_3D) {
      const int ax = single_axis(E.p);
      if (ax >= 0)
        e.f("  const Pose F%d = pose_of(%s + (q[%d] * %s) * qrot_col<%d>(%s), %s);\n", E.fb, Bp, E.coord, lit(E.p[ax]).c_str(), ax, Bq, Bq);
      else
        e.f("  const Pose F%d = pose_of(%s + mul(qrot(%s), q[%d] * %s), %s);\n", E.fb, Bp, Bq, E.coord, v3lit(E.p).c_str(), Bq);
    } else if (E.kind == RKB_FREE_3D) {
      e.f("  const Pose F%d = pose_of(%s + mul(qrot(%s), freec.p), qmul(%s, freec.q));\n", E.fb, Bp, Bq, Bq);
    } else {
      e.f("  const Pose F%d = pose_of(%s, %s);\n", E.fb, offset_const(Bp, Bq, E.p).c_str(), qmul_const(Bq, E.p + 3).c_str());
    }
    written[E.fb] = true;
  }
  // ---- shapes: world-fixed ones are literals, the others ride on their anchor frame
  for (int k = 0; k < P.n1 + P.n2; ++k) {
    const ProxShape& S = P.s[k];
    if (S.anchor < 0) {
      e.f("  const SPose S%d = spose_lit(%s, %s, %s", k, lit(S.pos[0]).c_str(), lit(S.pos[1]).c_str(), lit(S.pos[2]).c_str());
      for (int j = 0; j < 9; ++j) e.f(", %s", lit(S.rot[j]).c_str());
      e.f(");\n");
    } else {
      if (S.anchor >= RKB_GEN_MAX_FRAMES || !written[S.anchor]) return std::string();
      char Fp[20], Fq[20];
      std::snprintf(Fp, sizeof Fp, "F%d.p", S.anchor);
      std::snprintf(Fq, sizeof Fq, "F%d.q", S.anchor);
      const bool zero = S.pos[0] == 0.0 && S.pos[1] == 0.0 && S.pos[2] == 0.0;
      const int ax = single_axis(S.pos);
      char one[160];
      std::snprintf(one, sizeof one, "%s + qrotv_axis<%d>(%s, %s)", Fp, ax < 0 ? 0 : ax, Fq, lit(S.pos[ax < 0 ? 0 : ax]).c_str());
      const std::string p = zero ? std::string(Fp) : ax >= 0 ? std::string(one) : std::string(Fp) + " + qrotv(" + Fq + ", " + v3lit(S.pos) + ")";
      e.f("  const SPose S%d = spose_of(%s, %s);\n", k, p.c_str(), qmul_const(Fq, S.quat).c_str());
    }
  }
  // ---- the search (distances only), then the record of the winner
  e.f("  double min_d = INFINITY;\n  int best = -1;\n");
  int f = 0;
  std::string fa, fb;  // shape indices of every finder, for the record of the winner
  for (int a = 0; a < P.n1; ++a)
    for (int b = 0; b < P.n2; ++b) {
      const ProxShape &Sa = P.s[a], &Sb = P.s[P.n1 + b];
      if (!has_finder(Sa.kind, Sb.kind)) continue;
      char Pa[16], Pb[16];
      std::snprintf(Pa, sizeof Pa, "S%d", a);
      std::snprintf(Pb, sizeof Pb, "S%d", P.n1 + b);
      const std::string call = finder_call(Sa, Pa, Sb, Pb, "false");
      if (f == 0)
        e.f("  { min_d = %s.d; best = 0; }\n", call.c_str());
      else
        e.f("  if (!(norm3(%s.p - %s.p) - %s - %s > min_d)) { const double d = %s.d; if (min_d > d) { min_d = d; best = %d; } }\n", Pb, Pa,
            lit(Sa.brad).c_str(), lit(Sb.brad).c_str(), call.c_str(), f);
      char idx[16];
      std::snprintf(idx, sizeof idx, "%s%d", f ? ", " : "", a);
      fa += idx;
      std::snprintf(idx, sizeof idx, "%s%d", f ? ", " : "", P.n1 + b);
      fb += idx;
      ++f;
    }
  e.f("  bestR.p1 = v3(0.0, 0.0, 0.0); bestR.p2 = v3(0.0, 0.0, 0.0); bestR.d = min_d;\n");
  if (f > 0) {
    // The two points of the winner.  Every thread of a warp may hold a different winner: one finder call on shapes picked
    // by index (poses from a per-thread array, kinds and dimensions from tables) instead of a switch over all of them.
    const int ns = P.n1 + P.n2;
    e.f("  if (PTS) {\n    SPose SP[%d];\n", ns);
    for (int k = 0; k < ns; ++k) e.f("    SP[%d] = S%d;\n", k, k);
    e.f("    const int a = FA[best], b = FB[best];\n");
    e.f("    bestR = prox_compute_kd<true>(KIND[a], v3(DIMS[3 * a], DIMS[3 * a + 1], DIMS[3 * a + 2]), SP[a], KIND[b], "
        "v3(DIMS[3 * b], DIMS[3 * b + 1], DIMS[3 * b + 2]), SP[b]);\n  }\n");
  }
  e.f("  return best;\n}\n}  // namespace %s\n", ns);
  Emit h;
  h.f("// %d elements, %d + %d shapes\nnamespace %s {\n", G.n_elements, P.n1, P.n2, ns);
  h.f("RKB_PROX_SPEC_TABLE double C[%d] = {", (int)(pool.size() ? pool.size() : 1));
  for (size_t k = 0; k < pool.size(); ++k) h.f("%s%a", k ? ", " : "", pool[k]);
  if (pool.empty()) h.f("0.0");
  h.f("};\n");
  if (f > 0) {
    const int ns = P.n1 + P.n2;
    h.f("RKB_PROX_SPEC_TABLE int FA[%d] = {%s};\nRKB_PROX_SPEC_TABLE int FB[%d] = {%s};\n", f, fa.c_str(), f, fb.c_str());
    h.f("RKB_PROX_SPEC_TABLE int KIND[%d] = {", ns);
    for (int k = 0; k < ns; ++k) h.f("%s%d", k ? ", " : "", (int)P.s[k].kind);
    h.f("};\nRKB_PROX_SPEC_TABLE double DIMS[%d] = {", 3 * ns);
    for (int k = 0; k < 3 * ns; ++k) h.f("%s%a", k ? ", " : "", P.s[k / 3].dims[k % 3]);
    h.f("};\n");
  }
  return h.s + e.s;
}
}  // namespace

// The two query kernels (kte_prox_spec.cuh: RKB_PROX_SPEC_KERNELS) of one chain and pair.
// min_blocks: CTAs of 128 threads per SM the kernels are compiled for (register budget 65536 / (128 min_blocks))
std::string rkb_prox_source(const GenericProgram& G, const ProxProgram& P, int min_blocks) {
  std::lock_guard<std::mutex> lock(g_gen_mu);
  const std::string body = prox_body(G, P, "pair0");
  if (body.empty()) return body;
  Emit e;
  e.f("// generated by reak_b200 (rkb_prox_jit.cu)\n#include \"kte_prox_spec.cuh\"\nnamespace {\n");
  e.s += body;
  e.f("using pair0::prox_spec;\n}  // namespace\nRKB_PROX_SPEC_KERNELS(%d, %d, %d)\n", G.n_coords, G.n_free > 0 ? 1 : 0, min_blocks);
  return e.s;
}

// serial_steer_kernel<n, fl, shape, RkbSteerCheck> (kte_serial.cuh) with the collision test of `n_pairs` proxy pairs
// compiled in: the whole checked steering loop in one launch.  coord_of_stage: SerialParams::st[s].coord.  `expr` receives
// the name expression of the kernel.  Serial chains have no free joint.
std::string rkb_steer_checked_source(int n, int fl, unsigned long long shape, const int* coord_of_stage, const GenericProgram& G,
                                     const ProxProgram* const* pairs, int n_pairs, std::string* expr) {
  std::lock_guard<std::mutex> lock(g_gen_mu);
  if (G.n_free > 0 || G.n_coords != n || n_pairs < 1) return std::string();
  Emit e;
  e.f("// generated by reak_b200 (rkb_prox_jit.cu): checked steering, %d proxy pair(s)\n", n_pairs);
  e.f("#include \"kte_serial.cuh\"\n#include \"kte_prox_spec.cuh\"\nnamespace {\n");
  for (int p = 0; p < n_pairs; ++p) {
    char ns[16];
    std::snprintf(ns, sizeof ns, "pair%d", p);
    const std::string body = prox_body(G, *pairs[p], ns);
    if (body.empty()) return body;
    e.s += body;
  }
  // fewer_blocks = 1: the kernel gives up one CTA per SM (4 -> 3 for the 6- and 7-joint arms) so that the test's forward
  // kinematics and finders get their registers.  Measured on the CRS arm against the MD148 lab, 2^18 tuples x 10 intervals
  // (unchecked loop 5.29 ms): inlined at the rollout kernel's occupancy 7.55 ms, as a real call 8.24, inlined with one CTA
  // fewer 6.81, a real call with one CTA fewer 7.44; interval by interval (four launches each) 7.08.  Chains of 7 and 8
  // coordinates already run at 3 CTAs per SM and keep them (7-joint arm: 8.20 ms against 8.41 interval by interval).
  e.f("}  // namespace\nstruct RkbSteerCheck {\n  static constexpr bool enabled = true;\n  static constexpr int fewer_blocks = %d;\n", n <= 6 ? 1 : 0);
  e.f("  template <int N> __device__ __forceinline__ static bool is_free(");
  e.f("const rkb::SerialState<N>& X) {\n    double q[%d];\n", n);
  for (int s = 0; s < n; ++s) {
    if (coord_of_stage[s] < 0 || coord_of_stage[s] >= n) return std::string();
    e.f("    q[%d] = X.q[%d];\n", coord_of_stage[s], s);
  }
  e.f("    const Pose freec = pose_of(v3(0.0, 0.0, 0.0), q4(1.0, 0.0, 0.0, 0.0));\n    ProxRecord R;\n");
  // is_free_impl (MEAQR_topology.hpp:921-940): no pair with a negative minimum distance
  for (int p = 0; p < n_pairs; ++p) e.f("    pair%d::prox_spec<false>(q, freec, R);\n    if (R.d < 0.0) return false;\n", p);
  e.f("    return true;\n  }\n};\n");
  char buf[160];
  std::snprintf(buf, sizeof buf, "rkb::serial_steer_kernel<%d, %d, %lluull, RkbSteerCheck>", n, fl, shape);
  *expr = buf;
  return e.s;
}
