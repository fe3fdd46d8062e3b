// rkb_api.cu — the C-ABI of include/reak_b200.h: descriptor validation and lowering, buffer
// staging, kernel dispatch.  No numerics happen on the host and there is no CPU fallback: every
// compute entry point needs a usable sm_100 device and fails with RKB_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/reak_b200.h"
#include "rkb_internal.h"
#include "build_id.h"

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* where) {
  std::snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", where, cudaGetErrorString(e));
  cudaGetLastError();  // clear the sticky-less error state
  return RKB_ERR_CUDA;
}
#define CU(call)                                          \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; cudaGetLastError(); return RKB_ERR_NOMEM; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DeviceCtx {
  int device = -1;
  bool checked = false;
  GenericProgram* d_prog = nullptr;
  DevBuf in_x, in_u, out_a, out_b, st, scratch_x, scratch_u, scratch_o, scratch_s, in_goal, out_idx, out_cost;
  DevBuf in_bias, in_gain, io_uprev, out_ndone, act;  // closed-loop steering
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-buffer pipeline: copy-in, two alternating compute streams, copy-out
  cudaStream_t s_in = nullptr, s_k[2] = {nullptr, nullptr}, s_out = nullptr;
  cudaEvent_t ev_in[16] = {}, ev_k[16] = {}, ev_join = nullptr;
  bool timed = false;
  bool prepared = false;
};

}  // namespace

struct rkb_chain {
  rkb_chain_desc desc;
  std::vector<rkb_element> elements;
  int n = 0, nu = 0;
  int n_free = 0, nx = 0, na = 0;  // free_joint_3D count; state (2 n + 13 n_free) and acceleration (n + 6 n_free) dimensions
  bool serial_ok = false;
  int serial_fl = 0;
  unsigned long long serial_shape = 0;  // structure found in the descriptor (RKB_SHAPE_*)
  SerialParams sp;
  const SerialKernels* sk = nullptr;
  const JitKernels* jit = nullptr;  // set by rkb_chain_specialize: kernels compiled at run time for this chain's exact shape
  bool generic_ok = false;
  GenericProgram gp;
  std::vector<DeviceCtx*> ctx;  // one per device used
  std::mutex mu;
  uint64_t launches = 0;
  DeviceCtx* last = nullptr;
  // rkb_chain_set_option
  long long split_max = -1;        // largest batch integrated with one sample on a pair of warps; -1: the measured default
  bool fused_steer = true, fused_sequence = true, host_pipeline = true;
  bool auto_specialize = true, auto_done = false;
  unsigned create_flags = 0;
  // checked steering in one launch: kernels generated per set of proxy pairs (rkb_steer_checked_source)
  struct CheckedSteer { std::vector<unsigned long long> key; const SourceKernels* K; bool done; long long seen; };
  std::vector<CheckedSteer> checked;
};

// Below this many samples a batch cannot fill the GPU with one thread per sample (148 SMs x 4 sub-partitions x 32
// lanes = 18944 lone warps) and the pair-of-warps kernels win; measured crossover on B200: profiles/r2_small_batch.md.
constexpr long long kSplitMaxDefault = 8192;
constexpr int kSplitMinCoords = 4;

namespace {

// ---- small host-side algebra used only to lower constants -------------------------------------
void quat_to_rowmajor(const double* q, double* R) {  // quaternion::getRotMat, rotations_3D.hpp:986-1000
  const double t01 = 2.0 * q[0] * q[1], t02 = 2.0 * q[0] * q[2], t03 = 2.0 * q[0] * q[3];
  const double t11 = 2.0 * q[1] * q[1], t12 = 2.0 * q[1] * q[2], t13 = 2.0 * q[1] * q[3];
  const double t22 = 2.0 * q[2] * q[2], t23 = 2.0 * q[2] * q[3], t33 = 2.0 * q[3] * q[3];
  R[0] = 1.0 - t22 - t33; R[1] = t12 - t03; R[2] = t02 + t13;
  R[3] = t12 + t03; R[4] = 1.0 - t11 - t33; R[5] = t23 - t01;
  R[6] = t13 - t02; R[7] = t01 + t23; R[8] = 1.0 - t11 - t22;
}
void unit_quat(const double* in, double* out) {  // explicit quaternion(Vector), rotations_3D.hpp:917-920
  const double n = std::sqrt(in[0] * in[0] + in[1] * in[1] + in[2] * in[2] + in[3] * in[3]);
  for (int i = 0; i < 4; ++i) out[i] = in[i] / n;
}
void tmul3(const double* R, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
bool is_identity_quat(const double* q) { return q[0] == 1.0 && q[1] == 0.0 && q[2] == 0.0 && q[3] == 0.0; }

bool finite_all(const double* p, int n) {
  for (int i = 0; i < n; ++i) if (!std::isfinite(p[i])) return false;
  return true;
}

int validate(const rkb_chain_desc* d) {
  if (!d || (!d->elements && d->n_elements > 0)) return RKB_ERR_INVALID;
  if (d->dim != 2 && d->dim != 3) return RKB_ERR_INVALID;
  if (d->n_elements < 0 || d->n_frames < 1 || d->n_coords < 0 || d->n_inputs < 0) return RKB_ERR_INVALID;
  if (d->n_coords > RKB_MAX_COORDS) return RKB_ERR_UNSUPPORTED;
  if (d->base_frame < 0 || d->base_frame >= d->n_frames) return RKB_ERR_INVALID;
  if (!finite_all(d->base.position, 3) || !finite_all(d->base.quat, 4) || !finite_all(d->base.velocity, 3) ||
      !finite_all(d->base.ang_velocity, 3) || !finite_all(d->base.acceleration, 3) || !finite_all(d->base.ang_acceleration, 3))
    return RKB_ERR_INVALID;
  std::vector<int> coord_joint(d->n_coords, 0), input_used(d->n_inputs, 0), written(d->n_frames, 0);
  int n_free = 0, n_aux = 0;
  for (int e = 0; e < d->n_elements; ++e) if (d->elements[e].kind == RKB_FREE_3D || d->elements[e].kind == RKB_FREE_2D) ++n_free;
  for (int e = 0; e < d->n_elements; ++e) if (d->elements[e].kind == RKB_COORD_GEN) ++n_aux;
  if (d->n_coords + n_aux > RKB_MAX_COORDS) return RKB_ERR_UNSUPPORTED;
  std::vector<int> aux_declared(n_aux, 0), aux_written(n_aux, 0);
  auto any_coord_ok = [&](int c) { return c >= 0 && c < d->n_coords + n_aux && (c < d->n_coords || aux_declared[c - d->n_coords]); };
  if (n_free > RKB_GEN_MAX_FREE) return RKB_ERR_UNSUPPORTED;
  int free_seen = 0;
  for (int e = 0; e < d->n_elements; ++e) {
    const rkb_element& E = d->elements[e];
    if (!finite_all(E.p, 12)) return RKB_ERR_INVALID;
    const bool is3 = E.kind < 16;
    auto frame_ok = [&](int f) { return f >= 0 && f < d->n_frames; };
    auto coord_ok = [&](int c) { return c >= 0 && c < d->n_coords; };
    switch (E.kind) {
      case RKB_REVOLUTE_3D: case RKB_PRISMATIC_3D: case RKB_REVOLUTE_2D: case RKB_PRISMATIC_2D:
        if (!frame_ok(E.frame_a) || !frame_ok(E.frame_b) || !coord_ok(E.coord) || E.frame_a == E.frame_b) return RKB_ERR_INVALID;
        if ((d->dim == 3) != is3) return RKB_ERR_INVALID;
        if (coord_joint[E.coord]++) return RKB_ERR_UNSUPPORTED;  // one joint per coordinate
        if (written[E.frame_b]++ || E.frame_b == d->base_frame) return RKB_ERR_INVALID;
        break;
      case RKB_RIGID_LINK_3D: case RKB_RIGID_LINK_2D:
        if (!frame_ok(E.frame_a) || !frame_ok(E.frame_b) || E.frame_a == E.frame_b) return RKB_ERR_INVALID;
        if ((d->dim == 3) != is3) return RKB_ERR_INVALID;
        if (written[E.frame_b]++ || E.frame_b == d->base_frame) return RKB_ERR_INVALID;
        if (E.kind == RKB_RIGID_LINK_3D) {
          const double n2 = E.p[3] * E.p[3] + E.p[4] * E.p[4] + E.p[5] * E.p[5] + E.p[6] * E.p[6];
          if (!(n2 > 0.0)) return RKB_ERR_INVALID;
        }
        break;
      case RKB_INERTIA_3D: case RKB_INERTIA_2D:
        if (!frame_ok(E.frame_a) || (d->dim == 3) != is3) return RKB_ERR_INVALID;
        {  // bits 0 .. n_coords-1: coordinates; bits 32 .. 32+n_free-1: free-joint frames (mUpStream3DJoints)
          const uint64_t allowed = (d->n_coords >= 32 ? 0xffffffffull : ((1ull << d->n_coords) - 1ull)) | (((1ull << n_free) - 1ull) << 32);
          if (E.upstream & ~allowed) return RKB_ERR_INVALID;
        }
        break;
      case RKB_INERTIA_GEN:
        if (!coord_ok(E.coord)) return RKB_ERR_INVALID;
        if (E.upstream != (1ull << E.coord)) return RKB_ERR_UNSUPPORTED;
        // mass_matrix_calc::get_TMT_TdMT tests mUpStreamJoints.find(mCoords[i]) with i running over the 3D FRAMES and then
        // dereferences mUpStream3DJoints[mFrames3D[i]] (mass_matrix_calculator.cpp:226-233): a rotor on coordinate i < n_free
        // is a null dereference in the reference; there is no behaviour to reproduce
        if (E.coord < n_free) return RKB_ERR_UNSUPPORTED;
        break;
      case RKB_ACTUATOR_GEN: {
        if (!coord_ok(E.coord) || E.aux < 0 || E.aux >= d->n_inputs) return RKB_ERR_INVALID;
        if (E.frame_b < 0 || E.frame_b >= d->n_elements) return RKB_ERR_INVALID;
        const int jk = d->elements[E.frame_b].kind;
        if (jk != RKB_REVOLUTE_3D && jk != RKB_PRISMATIC_3D && jk != RKB_REVOLUTE_2D && jk != RKB_PRISMATIC_2D) return RKB_ERR_INVALID;
        if (input_used[E.aux]++) return RKB_ERR_INVALID;
        break;
      }
      case RKB_TORSION_SPRING_3D: case RKB_TORSION_DAMPER_3D: case RKB_SPRING_3D: case RKB_DAMPER_3D:
      case RKB_TORSION_SPRING_2D: case RKB_TORSION_DAMPER_2D: case RKB_SPRING_2D: case RKB_DAMPER_2D:
        if (!frame_ok(E.frame_a) || !frame_ok(E.frame_b) || (d->dim == 3) != is3) return RKB_ERR_INVALID;
        break;
      case RKB_COORD_GEN:  // declares auxiliary coordinate `coord` (>= n_coords), ahead of the elements that use it
        if (E.coord < d->n_coords || E.coord >= d->n_coords + n_aux || aux_declared[E.coord - d->n_coords]++) return RKB_ERR_INVALID;
        break;
      case RKB_RIGID_LINK_GEN:  // the end is an auxiliary coordinate no other link writes
        if (!any_coord_ok(E.coord) || !any_coord_ok(E.aux) || E.aux < d->n_coords || E.aux == E.coord) return RKB_ERR_INVALID;
        if (aux_written[E.aux - d->n_coords]++) return RKB_ERR_INVALID;
        break;
      case RKB_SPRING_GEN: case RKB_DAMPER_GEN:
        if (!any_coord_ok(E.coord) || !any_coord_ok(E.aux)) return RKB_ERR_INVALID;
        break;
      case RKB_FREE_3D: case RKB_FREE_2D:  // coord = index of the joint's coordinate frame in dofs_3D / dofs_2D, in chain order
        if ((d->dim == 3) != (E.kind == RKB_FREE_3D) || !frame_ok(E.frame_a) || !frame_ok(E.frame_b) || E.frame_a == E.frame_b) return RKB_ERR_INVALID;
        if (E.coord != free_seen++) return RKB_ERR_INVALID;
        if (written[E.frame_b]++ || E.frame_b == d->base_frame) return RKB_ERR_INVALID;
        break;
      default:
        return RKB_ERR_INVALID;
    }
  }
  for (int c = 0; c < d->n_coords; ++c) if (coord_joint[c] != 1) return RKB_ERR_UNSUPPORTED;
  for (int i = 0; i < d->n_inputs; ++i) if (input_used[i] != 1) return RKB_ERR_INVALID;
  // every frame must be reachable: doMotion must write a frame before an element reads it
  std::vector<int> ready(d->n_frames, 0);
  ready[d->base_frame] = 1;
  for (int e = 0; e < d->n_elements; ++e) {
    const rkb_element& E = d->elements[e];
    switch (E.kind) {
      case RKB_REVOLUTE_3D: case RKB_PRISMATIC_3D: case RKB_REVOLUTE_2D: case RKB_PRISMATIC_2D:
      case RKB_RIGID_LINK_3D: case RKB_RIGID_LINK_2D: case RKB_FREE_3D: case RKB_FREE_2D:
        if (!ready[E.frame_a]) return RKB_ERR_UNSUPPORTED;
        ready[E.frame_b] = 1;
        break;
      default: break;
    }
  }
  for (int f = 0; f < d->n_frames; ++f) if (!ready[f]) return RKB_ERR_UNSUPPORTED;
  return RKB_OK;
}

// Try to express the chain as joint/link/inertia stages (see rkb_types.h).  Returns false when
// the chain needs the generic interpreter.
// A planar chain is the 3D chain it describes when drawn in the x-y plane: revolute_joint_2D turns about
// e_z (revolute_joint.cpp:32-116 against :121-213 — the torque left on the base is zero in both, the axial
// part being all there is), prismatic_joint_2D / rigid_link_2D carry planar vectors (prismatic_joint.cpp:33-123,
// rigid_link.cpp:87-139), inertia_2D is a diagonal tensor whose zz entry is the moment of inertia (no
// gyroscopic torque for a rotation about a principal axis, inertia.cpp:77-86 against :111-121), and the base
// frame rotates about e_z.  Two-anchor spring_2D / damper_2D stay with the interpreter.
bool embed_planar(const rkb_chain_desc& d, rkb_chain_desc& out, std::vector<rkb_element>& el) {
  out = d;
  out.dim = 3;
  el.assign(d.elements, d.elements + d.n_elements);
  for (rkb_element& E : el) {
    double p[12];
    std::memcpy(p, E.p, sizeof p);
    std::memset(E.p, 0, sizeof E.p);
    E.reserved = 0;
    switch (E.kind) {
      case RKB_REVOLUTE_2D: E.kind = RKB_REVOLUTE_3D; E.p[2] = 1.0; break;
      case RKB_PRISMATIC_2D: E.kind = RKB_PRISMATIC_3D; E.p[0] = p[0]; E.p[1] = p[1]; break;
      case RKB_RIGID_LINK_2D:
        E.kind = RKB_RIGID_LINK_3D; E.p[0] = p[0]; E.p[1] = p[1];
        if (p[2] == 0.0) { E.p[3] = 1.0; } else { E.p[3] = std::cos(0.5 * p[2]); E.p[6] = std::sin(0.5 * p[2]); }
        break;
      case RKB_INERTIA_2D: E.kind = RKB_INERTIA_3D; E.p[0] = p[0]; E.p[6] = p[1]; break;
      case RKB_TORSION_SPRING_2D: E.kind = RKB_TORSION_SPRING_3D; E.p[0] = p[0]; E.p[1] = p[1]; E.reserved = 1; break;
      case RKB_TORSION_DAMPER_2D: E.kind = RKB_TORSION_DAMPER_3D; E.p[0] = p[0]; break;
      case RKB_INERTIA_GEN: case RKB_ACTUATOR_GEN: std::memcpy(E.p, p, sizeof p); break;
      default: return false;
    }
  }
  out.elements = el.data();
  const double a = d.base.quat[0];
  out.base.position[2] = 0.0; out.base.velocity[2] = 0.0; out.base.acceleration[2] = 0.0;
  out.base.quat[0] = std::cos(0.5 * a); out.base.quat[1] = 0.0; out.base.quat[2] = 0.0; out.base.quat[3] = std::sin(0.5 * a);
  out.base.ang_velocity[2] = d.base.ang_velocity[0]; out.base.ang_velocity[0] = out.base.ang_velocity[1] = 0.0;
  out.base.ang_acceleration[2] = d.base.ang_acceleration[0]; out.base.ang_acceleration[0] = out.base.ang_acceleration[1] = 0.0;
  return true;
}

bool lower_serial(const rkb_chain_desc& d_in, SerialParams& P, int& fl, unsigned long long& shape) {
  rkb_chain_desc d = d_in;
  std::vector<rkb_element> embedded;
  if (d_in.dim == 2 && !embed_planar(d_in, d, embedded)) return false;
  if (d.dim != 3 || d.n_coords < 1 || d.n_coords > RKB_SERIAL_MAX_DOF) return false;
  std::memset(&P, 0, sizeof P);
  P.n = d.n_coords;
  P.n_inputs = d.n_inputs;
  fl = 0;
  // base frame in its own coordinates
  double q0[4], R0[9];
  unit_quat(d.base.quat, q0);
  quat_to_rowmajor(q0, R0);
  double w[3] = {d.base.ang_velocity[0], d.base.ang_velocity[1], d.base.ang_velocity[2]};
  double al[3] = {d.base.ang_acceleration[0], d.base.ang_acceleration[1], d.base.ang_acceleration[2]};
  double a[3];
  tmul3(R0, d.base.acceleration, a);

  std::vector<double> rotor(d.n_coords, 0.0);
  std::vector<int> input_of(d.n_coords, -1);
  int cur = d.base_frame, k = -1;
  int joint_base = -1, joint_end = -1;
  uint64_t mask = 0;
  bool stage_has_link = false, stage_has_inertia = false;
  for (int e = 0; e < d.n_elements; ++e) {
    const rkb_element& E = d.elements[e];
    switch (E.kind) {
      case RKB_ACTUATOR_GEN: {
        const rkb_element& J = d.elements[E.frame_b];
        if (J.coord != E.coord || input_of[E.coord] >= 0) return false;
        input_of[E.coord] = E.aux;
        break;
      }
      case RKB_INERTIA_GEN:
        rotor[E.coord] += E.p[0];
        break;
      case RKB_REVOLUTE_3D: case RKB_PRISMATIC_3D: {
        if (E.frame_a != cur) return false;
        if (++k >= RKB_SERIAL_MAX_DOF) return false;
        SerialStage& S = P.st[k];
        S.coord = E.coord;
        S.input = -1;
        const double nrm = std::sqrt(E.p[0] * E.p[0] + E.p[1] * E.p[1] + E.p[2] * E.p[2]);
        for (int i = 0; i < 3; ++i) S.ax[i] = E.p[i];
        if (nrm > 0.0000001) for (int i = 0; i < 3; ++i) S.an[i] = E.p[i] / nrm;  // rotations_3D.hpp:1962-1974
        else { S.an[0] = 1.0; S.an[1] = 0.0; S.an[2] = 0.0; }
        S.aa[0] = S.an[0] * S.an[0]; S.aa[1] = S.an[1] * S.an[1]; S.aa[2] = S.an[2] * S.an[2];
        S.aa[3] = S.an[0] * S.an[1]; S.aa[4] = S.an[0] * S.an[2]; S.aa[5] = S.an[1] * S.an[2];
        S.Ro[0] = S.Ro[4] = S.Ro[8] = 1.0;
        if (E.kind == RKB_PRISMATIC_3D) { S.flags |= RKB_ST_PRISMATIC; fl |= RKB_FL_PRISMATIC; }
        else if (nrm <= 0.0000001) return false;  // degenerate revolute axis: leave it to the interpreter
        joint_base = E.frame_a; joint_end = E.frame_b; cur = E.frame_b;
        mask |= 1ull << E.coord;
        stage_has_link = stage_has_inertia = false;
        break;
      }
      case RKB_RIGID_LINK_3D: {
        if (E.frame_a != cur) return false;
        double q[4], R[9];
        unit_quat(&E.p[3], q);
        quat_to_rowmajor(q, R);
        if (k < 0) {
          // a link ahead of the first joint only re-bases the root frame (frame_3D.hpp:236-251)
          double t1[3], t2[3], t3[3], acc[3];
          cross3(w, E.p, t1); cross3(w, t1, t2); cross3(al, E.p, t3);
          for (int i = 0; i < 3; ++i) acc[i] = a[i] + t2[i] + t3[i];
          double wn[3], aln[3];
          tmul3(R, acc, a); tmul3(R, w, wn); tmul3(R, al, aln);
          std::memcpy(w, wn, sizeof w); std::memcpy(al, aln, sizeof al);
        } else {
          if (stage_has_link || stage_has_inertia) return false;
          SerialStage& S = P.st[k];
          S.flags |= RKB_ST_LINK;
          for (int i = 0; i < 3; ++i) S.po[i] = E.p[i];
          std::memcpy(S.Ro, R, sizeof R);
          if (!is_identity_quat(q)) { S.flags |= RKB_ST_LINKROT; fl |= RKB_FL_LINKROT; }
          stage_has_link = true;
        }
        cur = E.frame_b;
        break;
      }
      case RKB_INERTIA_3D: {
        if (k < 0) { if (E.upstream != 0) return false; break; }  // rides on the fixed root: no generalised force
        if (E.frame_a != cur || E.upstream != mask) return false;
        SerialStage& S = P.st[k];
        S.flags |= RKB_ST_INERTIA;
        S.m += E.p[0];
        for (int i = 0; i < 6; ++i) S.I[i] += E.p[1 + i];
        stage_has_inertia = true;
        break;
      }
      case RKB_TORSION_SPRING_3D: case RKB_TORSION_DAMPER_3D: {
        if (k < 0 || E.frame_a != joint_base || E.frame_b != joint_end) return false;
        SerialStage& S = P.st[k];
        if (S.flags & RKB_ST_PRISMATIC) break;  // no relative rotation across a prismatic joint: zero torque
        if (E.kind == RKB_TORSION_SPRING_3D) {
          if (S.flags & RKB_ST_SPRING) return false;
          S.flags |= RKB_ST_SPRING; S.ks = E.p[0]; S.sat = E.p[1];
          if (d_in.dim == 2 && E.reserved == 1) S.flags |= RKB_ST_SPRING_2D;  // marked by embed_planar
        } else {
          S.flags |= RKB_ST_DAMPER; S.cd += E.p[0];
        }
        fl |= RKB_FL_SPRINGS;
        break;
      }
      default:
        return false;
    }
  }
  if (k + 1 != d.n_coords) return false;
  for (int s = 0; s <= k; ++s) {
    P.st[s].rotor = rotor[P.st[s].coord];
    P.st[s].input = input_of[P.st[s].coord];
  }
  for (int i = 0; i < 3; ++i) { P.w0[i] = w[i]; P.al0[i] = al[i]; P.a0[i] = a[i]; }
  double mc = 0.0;  // composite masses for the inward mass-matrix pass
  for (int s = k; s >= 0; --s) {
    mc += P.st[s].m;
    P.st[s].mc = mc;
    for (int i = 0; i < 3; ++i) P.st[s].mcpo[i] = mc * P.st[s].po[i];
  }
  // structure the specialised kernels may rely on
  shape = 0;
  for (int s = 0; s <= k; ++s) {
    const SerialStage& S = P.st[s];
    unsigned ax = 0, lk = 0, in = 0, sign = 0;
    for (int dd = 0; dd < 3; ++dd) {  // revolute about / prismatic along +-e_dd
      const int d1 = (dd + 1) % 3, d2 = (dd + 2) % 3;
      if ((S.ax[dd] == 1.0 || S.ax[dd] == -1.0) && S.ax[d1] == 0.0 && S.ax[d2] == 0.0) {
        ax = ((S.flags & RKB_ST_PRISMATIC) ? 4 : 0) + dd + 1;
        sign = S.ax[dd] > 0.0 ? 1u : 3u;
      }
    }
    if (!(S.flags & RKB_ST_LINKROT)) {  // no link at all counts as a zero offset along z
      int nz = 0, which = 2;
      for (int dd = 0; dd < 3; ++dd) if (S.po[dd] != 0.0) { ++nz; which = dd; }
      if (nz <= 1) lk = which + 1;
    }
    if ((S.flags & RKB_ST_INERTIA) && S.I[1] == 0.0 && S.I[2] == 0.0 && S.I[4] == 0.0) in = 1;
    shape |= RKB_SHAPE_AT(RKB_SHAPE_STAGE_SIGNED(ax, lk, in, sign), s);
  }
  return true;
}

bool lower_generic(const rkb_chain_desc& d, GenericProgram& G) {
  if (d.n_elements > RKB_GEN_MAX_ELEMENTS || d.n_frames > RKB_GEN_MAX_FRAMES) return false;
  std::memset(&G, 0, sizeof G);
  G.dim = d.dim; G.n_elements = d.n_elements; G.n_frames = d.n_frames;
  G.n_coords = d.n_coords; G.n_inputs = d.n_inputs; G.base_frame = d.base_frame;
  G.free_states = d.dim == 3 ? 13 : 7; G.free_acc = d.dim == 3 ? 6 : 3;
  if (d.dim == 3) {
    double q[4];
    unit_quat(d.base.quat, q);
    for (int i = 0; i < 3; ++i) {
      G.base[i] = d.base.position[i]; G.base[7 + i] = d.base.velocity[i]; G.base[10 + i] = d.base.ang_velocity[i];
      G.base[13 + i] = d.base.acceleration[i]; G.base[16 + i] = d.base.ang_acceleration[i];
    }
    for (int i = 0; i < 4; ++i) G.base[3 + i] = q[i];
  } else {
    G.base[0] = d.base.position[0]; G.base[1] = d.base.position[1];
    G.base[3] = std::cos(d.base.quat[0]); G.base[4] = std::sin(d.base.quat[0]);  // rot_mat_2D(angle), rotations_2D.hpp:108-112
    G.base[7] = d.base.velocity[0]; G.base[8] = d.base.velocity[1];
    G.base[10] = d.base.ang_velocity[0];
    G.base[13] = d.base.acceleration[0]; G.base[14] = d.base.acceleration[1];
    G.base[16] = d.base.ang_acceleration[0];
  }
  int row_gen = 0, row_2d = 0, row_3d = 0;
  for (int e = 0; e < d.n_elements; ++e) {
    const rkb_element& E = d.elements[e];
    if (E.kind == RKB_INERTIA_GEN) row_2d += 1;
  }
  row_3d = row_2d;
  for (int e = 0; e < d.n_elements; ++e) if (d.elements[e].kind == RKB_INERTIA_2D) row_3d += 3;
  for (int e = 0; e < d.n_elements; ++e) {
    const rkb_element& E = d.elements[e];
    GenericElement& g = G.el[e];
    g.kind = E.kind; g.fa = E.frame_a; g.fb = E.frame_b; g.coord = E.coord; g.aux = E.aux;
    g.upstream = (uint32_t)(E.upstream & 0xffffull) | ((uint32_t)((E.upstream >> 32) & 0xffull) << RKB_GEN_FREE_BIT);
    if (E.kind == RKB_FREE_3D || E.kind == RKB_FREE_2D) { G.free_elem[G.n_free] = e; G.n_free += 1; }
    if (E.kind == RKB_COORD_GEN) { G.aux_q[E.coord - d.n_coords] = E.p[0]; G.aux_qd[E.coord - d.n_coords] = E.p[1]; G.n_aux += 1; }
    std::memcpy(g.p, E.p, sizeof g.p);
    if (E.kind == RKB_RIGID_LINK_3D) unit_quat(&E.p[3], &g.p[3]);
    if (E.kind == RKB_RIGID_LINK_2D) { g.p[3] = std::cos(E.p[2]); g.p[4] = std::sin(E.p[2]); }
    if (E.kind == RKB_INERTIA_GEN) { g.row = row_gen; row_gen += 1; }
    if (E.kind == RKB_INERTIA_2D) { g.row = row_2d; row_2d += 3; }
    if (E.kind == RKB_INERTIA_3D) { g.row = row_3d; row_3d += 6; }
    if (E.kind == RKB_REVOLUTE_3D || E.kind == RKB_PRISMATIC_3D || E.kind == RKB_REVOLUTE_2D || E.kind == RKB_PRISMATIC_2D)
      G.jelem[E.coord] = e;
  }
  return true;
}

// every promise of `have` (8 bits per stage, three fields) is either absent or kept by `chain`
bool shape_compatible(unsigned long long have, unsigned long long chain, int n) {
  for (int k = 0; k < n; ++k) {
    const unsigned h = (unsigned)((have >> (8 * k)) & 0xffu), c = (unsigned)((chain >> (8 * k)) & 0xffu);
    const unsigned hf[4] = {h & 7u, (h >> 3) & 3u, (h >> 5) & 1u, (h >> 6) & 3u}, cf[4] = {c & 7u, (c >> 3) & 3u, (c >> 5) & 1u, (c >> 6) & 3u};
    for (int i = 0; i < 4; ++i) if (hf[i] != 0 && hf[i] != cf[i]) return false;
  }
  return (n >= 8) || (have >> (8 * n)) == 0;
}
int shape_score(unsigned long long have, int n) {
  int sc = 0;
  for (int k = 0; k < n; ++k) {
    const unsigned h = (unsigned)((have >> (8 * k)) & 0xffu);
    sc += ((h & 7u) ? 3 : 0) + (((h >> 3) & 3u) ? 1 : 0) + (((h >> 5) & 1u) ? 1 : 0) + (((h >> 6) & 3u) ? 1 : 0);
  }
  return sc;
}

const SerialKernels* find_serial(int n, int fl, unsigned long long shape, bool no_special) {
  int count = 0;
  const SerialKernels* t = nullptr;
  switch (n) {
    case 1: t = rkb_serial_table_1(&count); break;
    case 2: t = rkb_serial_table_2(&count); break;
    case 3: t = rkb_serial_table_3(&count); break;
    case 4: t = rkb_serial_table_4(&count); break;
    case 5: t = rkb_serial_table_5(&count); break;
    case 6: t = rkb_serial_table_6(&count); break;
    case 7: t = rkb_serial_table_7(&count); break;
    case 8: t = rkb_serial_table_8(&count); break;
    default: return nullptr;
  }
  const SerialKernels* best = nullptr;
  for (int i = 0; i < count; ++i) {
    if ((t[i].fl & fl) != fl || !shape_compatible(t[i].shape, shape, n)) continue;
    if (no_special && t[i].shape != 0) continue;
    if (!best) { best = &t[i]; continue; }
    const int sa = shape_score(t[i].shape, n), sb = shape_score(best->shape, n);
    if (sa > sb || (sa == sb && __builtin_popcount(t[i].fl) < __builtin_popcount(best->fl))) best = &t[i];
  }
  return best;
}

// RAII: switch to `device`, restore on exit
// (cudaGetDevice reports 0 for a thread that never chose a device; restoring that would create a context on
// GPU 0 in every rank of a one-process-per-GPU job, so the previous device is only restored when it differs.)
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev == device) { prev = -1; ok = true; return; }
    ok = cudaSetDevice(device) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int get_ctx(rkb_chain* c, int device, DeviceCtx** out) {
  for (DeviceCtx* x : c->ctx) if (x->device == device) { *out = x; return 0; }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    std::snprintf(g_cuda_err, sizeof g_cuda_err, "device %d is sm_%d%d; this library holds sm_100a code only", device, prop.major, prop.minor);
    return RKB_ERR_CUDA;
  }
  DeviceCtx* x = new (std::nothrow) DeviceCtx();
  if (!x) return RKB_ERR_NOMEM;
  x->device = device;
  if (c->generic_ok) {
    e = cudaMalloc(&x->d_prog, sizeof(GenericProgram));
    if (e != cudaSuccess) { delete x; return cuda_fail(e, "cudaMalloc(program)"); }
    e = cudaMemcpy(x->d_prog, &c->gp, sizeof(GenericProgram), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(x->d_prog); delete x; return cuda_fail(e, "cudaMemcpy(program)"); }
  }
  if (c->sk) {
    e = c->sk->prepare();
    if (e == cudaSuccess && c->jit) e = rkb_jit_prepare(*c->jit);
    if (e != cudaSuccess) { if (x->d_prog) cudaFree(x->d_prog); delete x; return cuda_fail(e, "cudaFuncSetAttribute"); }
  }
  e = cudaEventCreate(&x->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&x->ev1);
  if (e != cudaSuccess) { delete x; return cuda_fail(e, "cudaEventCreate"); }
  c->ctx.push_back(x);
  *out = x;
  return 0;
}

struct Layout {
  bool device, soa, blocked;
};
Layout parse_flags(unsigned flags) {
  return Layout{(flags & RKB_MEM_DEVICE) != 0, (flags & RKB_LAYOUT_SOA) != 0, (flags & RKB_LAYOUT_BLOCKED) != 0};
}

// `blocked` only matters for state buffers ((q, q_dot) interleaved per coordinate, or all q then all q_dot)
ConstBatchView cview(const double* p, long long n, int dim, bool soa, bool blocked = false) {
  return ConstBatchView{p, soa ? 1 : dim, soa ? n : 1, blocked ? 1 : 0};
}
BatchView view(double* p, long long n, int dim, bool soa, bool blocked = false) {
  return BatchView{p, soa ? 1 : dim, soa ? n : 1, blocked ? 1 : 0};
}

// Stage a host input on the device (or pass a device pointer through).
int stage_in(DevBuf& buf, const void* src, size_t bytes, bool on_device, cudaStream_t s, const void** out) {
  if (on_device || bytes == 0) { *out = src; return 0; }
  int rc = buf.ensure(bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, s));
  *out = buf.p;
  return 0;
}
int stage_out(DevBuf& buf, void* dst, size_t bytes, bool on_device, void** out) {
  if (on_device || bytes == 0 || !dst) { *out = dst; return 0; }
  int rc = buf.ensure(bytes);
  if (rc) return rc;
  *out = buf.p;
  return 0;
}
int unstage_out(void* dev, void* dst, size_t bytes, bool on_device, cudaStream_t s) {
  if (on_device || bytes == 0 || !dst) return 0;
  CU(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, s));
  return 0;
}

enum Op { OP_EVAL, OP_FORCES, OP_MASS, OP_TMT, OP_FRAMES };

void maybe_auto_specialize(rkb_chain* c, long long n_samples);  // below, next to the kernel selection

int tmt_rows(const rkb_chain_desc& d) {
  int rows = 0;
  for (int e = 0; e < d.n_elements; ++e) {
    const int k = d.elements[e].kind;
    rows += k == RKB_INERTIA_GEN ? 1 : k == RKB_INERTIA_2D ? 3 : k == RKB_INERTIA_3D ? 6 : 0;
  }
  return rows;
}

int run_eval_like(rkb_chain* c, Op op, int device, size_t N, const double* x, const double* u, double* out, double* out2,
                  int32_t* status, unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (N == 0) return RKB_OK;
  if (!x || !out || (c->nu > 0 && !u && op != OP_MASS && op != OP_TMT)) return RKB_ERR_INVALID;
  if ((op == OP_TMT || op == OP_FRAMES) && !c->generic_ok) return RKB_ERR_UNSUPPORTED;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;  // the legacy blocked order has no place for a free joint's 13 states
  const int n = c->na, nx = c->nx, nu = c->nu;
  const int out_dim = op == OP_EVAL ? nx : op == OP_FORCES ? n : op == OP_MASS ? n * n : op == OP_TMT ? tmt_rows(c->desc) * n
                                                                                           : RKB_FRAME_DOUBLES * c->desc.n_frames;
  if (out_dim == 0) return RKB_OK;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  maybe_auto_specialize(c, (long long)N);
  cudaStream_t s = (cudaStream_t)stream;
  const void *dx = nullptr, *du = nullptr;
  void *dout = nullptr, *dout2 = nullptr, *dst = nullptr;
  if ((rc = stage_in(ctx->in_x, x, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if (op != OP_MASS && op != OP_TMT && nu > 0) { if ((rc = stage_in(ctx->in_u, u, N * nu * sizeof(double), L.device, s, &du))) return rc; }
  if ((rc = stage_out(ctx->out_a, out, N * out_dim * sizeof(double), L.device, &dout))) return rc;
  if ((rc = stage_out(ctx->out_b, out2, N * out_dim * sizeof(double), L.device, &dout2))) return rc;
  if ((rc = stage_out(ctx->st, status, N * sizeof(int32_t), L.device, &dst))) return rc;
  EvalArgs A;
  A.x = cview((const double*)dx, (long long)N, nx, L.soa, L.blocked);
  A.u = cview((const double*)(du ? du : dx), (long long)N, nu > 0 ? nu : 1, L.soa);
  A.out = view((double*)dout, (long long)N, out_dim, L.soa, L.blocked && op == OP_EVAL);
  A.out2 = view((double*)dout2, (long long)N, out_dim, L.soa);
  A.status = (int32_t*)dst;
  A.n_samples = (long long)N;
  CU(cudaEventRecord(ctx->ev0, s));
  cudaError_t e;
  const bool use_serial = c->serial_ok && c->sk && op != OP_TMT && op != OP_FRAMES;
  if (op == OP_TMT) {
    e = rkb_generic_tmt(ctx->d_prog, c->gp, A, s);
  } else if (op == OP_FRAMES) {
    e = rkb_generic_frames(ctx->d_prog, c->gp, A, s);
  } else if (use_serial) {
    if (c->jit)
      e = rkb_jit_launch(*c->jit, op == OP_EVAL ? RKB_JIT_EVAL : op == OP_FORCES ? RKB_JIT_FORCES : (A.out2.p ? RKB_JIT_MASSDOT : RKB_JIT_MASS),
                         c->sp, &A, nullptr, A.n_samples, 0, s);
    else
      e = op == OP_EVAL ? c->sk->eval(c->sp, A, s) : op == OP_FORCES ? c->sk->forces(c->sp, A, s) : c->sk->mass(c->sp, A, s);
  } else if (c->generic_ok) {
    e = op == OP_EVAL ? rkb_generic_eval(ctx->d_prog, c->gp, A, s)
        : op == OP_FORCES ? rkb_generic_forces(ctx->d_prog, c->gp, A, s) : rkb_generic_mass(ctx->d_prog, c->gp, A, s);
  } else {
    return RKB_ERR_UNSUPPORTED;
  }
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  c->launches += 1;
  if ((rc = unstage_out(dout, out, N * out_dim * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dout2, out2, N * out_dim * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dst, status, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

// small batch: one sample on a pair of warps (kte_serial.cuh: DuoCtx).  Not for run-time specialised handles (the
// NVRTC set holds the one-thread-per-sample kernels only).
bool use_split(const rkb_chain* c, long long n_samples) {
  if (!c->serial_ok || !c->sk || c->jit) return false;
  if (c->split_max >= 0) return n_samples <= c->split_max;  // the caller's explicit choice
  // default: a short evaluation does not pay for the two hand-offs (planar 2-link chain: measured 0.9x; from 6
  // coordinates on 1.5x), and beyond ~8192 samples one thread per sample already fills the sub-partitions
  return c->n >= kSplitMinCoords && n_samples <= kSplitMaxDefault;
}

// A serial chain whose structure (axis-aligned joints, links along one axis, diagonal tensors) the shipped kernels only
// partly promise runs on more general code than it needs, at half the speed.  On the first call that is big enough to
// care, kernels for exactly this chain's structure are requested: from the disk cache if a previous process compiled
// them, else from NVRTC on a background thread — the call at hand (and every one until the compilation has finished)
// runs on the shipped kernels.  Called with the handle locked and the device current.
constexpr long long kAutoSpecializeMinSamples = 4096;
void maybe_auto_specialize(rkb_chain* c, long long n_samples) {
  if (!c->auto_specialize || c->auto_done || c->jit || !c->serial_ok || !c->sk) return;
  if (c->sk->shape == c->serial_shape) { c->auto_done = true; return; }  // the shipped kernels already match
  if (n_samples < kAutoSpecializeMinSamples) return;
  const JitKernels* J = nullptr;
  if (rkb_jit_poll(c->n, c->serial_fl, c->serial_shape, &J) != RKB_OK) { c->auto_done = true; return; }  // no NVRTC here: stay as we are
  if (!J) return;  // still compiling
  int cur = -1;
  cudaGetDevice(&cur);
  for (DeviceCtx* x : c->ctx) {
    if (x->device == cur) continue;  // rkb_jit_poll prepared the kernels on the current device
    DeviceGuard g(x->device);
    if (rkb_jit_prepare(*J) != cudaSuccess) { cudaGetLastError(); c->auto_done = true; return; }
  }
  c->jit = J;
  c->auto_done = true;
}

// rollout on device-resident views; used by rkb_rollout_rk4 and rkb_steer_batch
// table == nullptr: the dedicated RK4 kernels (the fast path)
int launch_rollout(rkb_chain* c, DeviceCtx* ctx, const RolloutArgs& A, const RkTable* table, cudaStream_t s) {
  cudaError_t e;
  if (c->serial_ok && c->jit)
    e = table ? rkb_jit_launch(*c->jit, RKB_JIT_ROLLOUT_RK, c->sp, &A, table, A.n_samples,
                               (2 * c->n + 2 * c->n * table->stages) * 128 * (int)sizeof(double), s)
              : rkb_jit_launch(*c->jit, RKB_JIT_ROLLOUT, c->sp, &A, nullptr, A.n_samples, 0, s);
  else if (c->serial_ok && c->sk) {
    if (table) e = c->sk->rollout_rk(c->sp, A, *table, s);
    else if (use_split(c, A.n_samples) && !A.active) e = c->sk->rollout_duo(c->sp, A, s);
    else e = c->sk->rollout(c->sp, A, s);
  }
  else if (c->generic_ok) e = rkb_generic_rollout(ctx->d_prog, c->gp, A, table, s);
  else return RKB_ERR_UNSUPPORTED;
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  c->launches += 1;
  return RKB_OK;
}

}  // namespace

extern "C" {

int rkb_version(void) { return RKB_VERSION; }

const char* rkb_build_id(void) { return RKB_BUILD_ID; }

const char* rkb_strerror(int code) {
  switch (code) {
    case RKB_OK: return "ok";
    case RKB_ERR_INVALID: return "invalid argument or malformed chain descriptor";
    case RKB_ERR_UNSUPPORTED: return "chain or element outside the compiled path";
    case RKB_ERR_DIMENSION: return "state or input vector dimension mismatch";
    case RKB_ERR_CUDA: return "CUDA failure or no usable sm_100 device (there is no CPU fallback)";
    case RKB_ERR_NOMEM: return "out of memory";
    case RKB_ERR_INTEGRATION: return "impossible integration (zero step or negative step count)";
    default: return "unknown error";
  }
}

const char* rkb_last_cuda_error(void) { return g_cuda_err; }

int rkb_chain_create_ex(const rkb_chain_desc* desc, unsigned create_flags, rkb_chain** out) {
  if (!out) return RKB_ERR_INVALID;
  *out = nullptr;
  if (create_flags & ~(RKB_CREATE_INTERPRETER | RKB_CREATE_GENERAL)) return RKB_ERR_INVALID;
  int rc = validate(desc);
  if (rc) return rc;
  rkb_chain* c = new (std::nothrow) rkb_chain();
  if (!c) return RKB_ERR_NOMEM;
  c->desc = *desc;
  c->elements.assign(desc->elements, desc->elements + desc->n_elements);
  c->desc.elements = c->elements.data();
  c->n = desc->n_coords;
  c->nu = desc->n_inputs;
  for (int e = 0; e < desc->n_elements; ++e) if (desc->elements[e].kind == RKB_FREE_3D || desc->elements[e].kind == RKB_FREE_2D) c->n_free += 1;
  c->nx = 2 * c->n + (desc->dim == 3 ? 13 : 7) * c->n_free;  // kte_nl_system::get_state_dimensions, kte_nl_system.hpp:145-147
  c->na = c->n + (desc->dim == 3 ? 6 : 3) * c->n_free;
  c->create_flags = create_flags;
  c->serial_ok = lower_serial(c->desc, c->sp, c->serial_fl, c->serial_shape);
  if (c->serial_ok) {
    c->sk = find_serial(c->n, c->serial_fl, c->serial_shape, (create_flags & RKB_CREATE_GENERAL) != 0);
    if (!c->sk) c->serial_ok = false;
  }
  c->generic_ok = lower_generic(c->desc, c->gp);
  if ((create_flags & RKB_CREATE_INTERPRETER) && c->generic_ok) { c->serial_ok = false; c->sk = nullptr; }
  if (!c->serial_ok && !c->generic_ok) { delete c; return RKB_ERR_UNSUPPORTED; }
  *out = c;
  return RKB_OK;
}

int rkb_chain_create(const rkb_chain_desc* desc, rkb_chain** out) { return rkb_chain_create_ex(desc, 0u, out); }

int rkb_chain_set_option(rkb_chain* c, int option, long long value) {
  if (!c) return RKB_ERR_INVALID;
  std::lock_guard<std::mutex> lock(c->mu);
  switch (option) {
    case RKB_OPT_SPLIT_MAX_SAMPLES: if (value < -1) return RKB_ERR_INVALID; c->split_max = value; return RKB_OK;
    case RKB_OPT_FUSED_STEER: c->fused_steer = value != 0; return RKB_OK;
    case RKB_OPT_FUSED_SEQUENCE: c->fused_sequence = value != 0; return RKB_OK;
    case RKB_OPT_HOST_PIPELINE: c->host_pipeline = value != 0; return RKB_OK;
    case RKB_OPT_AUTO_SPECIALIZE: c->auto_specialize = value != 0; return RKB_OK;
    default: return RKB_ERR_INVALID;
  }
}
long long rkb_chain_get_option(const rkb_chain* c, int option) {
  if (!c) return RKB_ERR_INVALID;
  switch (option) {
    case RKB_OPT_SPLIT_MAX_SAMPLES: return c->split_max >= 0 ? c->split_max : (c->n >= kSplitMinCoords ? kSplitMaxDefault : 0);
    case RKB_OPT_FUSED_STEER: return c->fused_steer ? 1 : 0;
    case RKB_OPT_FUSED_SEQUENCE: return c->fused_sequence ? 1 : 0;
    case RKB_OPT_HOST_PIPELINE: return c->host_pipeline ? 1 : 0;
    case RKB_OPT_AUTO_SPECIALIZE: return c->auto_specialize ? 1 : 0;
    default: return RKB_ERR_INVALID;
  }
}

void rkb_chain_destroy(rkb_chain* c) {
  if (!c) return;
  for (DeviceCtx* x : c->ctx) {
    DeviceGuard g(x->device);
    if (x->d_prog) cudaFree(x->d_prog);
    DevBuf* bufs[] = {&x->in_x, &x->in_u, &x->out_a, &x->out_b, &x->st, &x->scratch_x, &x->scratch_u, &x->scratch_o,
                      &x->scratch_s, &x->in_goal, &x->out_idx, &x->out_cost, &x->in_bias, &x->in_gain, &x->io_uprev,
                      &x->out_ndone, &x->act};
    for (DevBuf* b : bufs) b->release();
    if (x->ev0) cudaEventDestroy(x->ev0);
    if (x->ev1) cudaEventDestroy(x->ev1);
    if (x->s_in) {
      cudaStreamDestroy(x->s_in); cudaStreamDestroy(x->s_k[0]); cudaStreamDestroy(x->s_k[1]); cudaStreamDestroy(x->s_out);
      for (int i = 0; i < 16; ++i) { if (x->ev_in[i]) cudaEventDestroy(x->ev_in[i]); if (x->ev_k[i]) cudaEventDestroy(x->ev_k[i]); }
      if (x->ev_join) cudaEventDestroy(x->ev_join);
    }
    delete x;
  }
  delete c;
}

int rkb_chain_state_dim(const rkb_chain* c) { return c ? c->nx : RKB_ERR_INVALID; }
int rkb_chain_input_dim(const rkb_chain* c) { return c ? c->nu : RKB_ERR_INVALID; }
int rkb_chain_dof(const rkb_chain* c) { return c ? c->n : RKB_ERR_INVALID; }
/* 1 when the chain runs on the register-resident serial kernels, 0 on the interpreter */
int rkb_chain_is_serial(const rkb_chain* c) { return c ? (c->serial_ok ? 1 : 0) : RKB_ERR_INVALID; }
/* structural promises of the kernels picked for this chain (0 = general code) and those the chain would allow */
unsigned long long rkb_chain_kernel_shape(const rkb_chain* c) { return !c ? 0ull : c->jit ? c->jit->shape : c->sk ? c->sk->shape : 0ull; }

/* Compile the serial kernels for exactly this chain's structure (NVRTC, a few seconds, cached per shape in the
 * process) and route the handle's launches to them.  RKB_ERR_UNSUPPORTED: the chain is not serial or
 * libnvrtc is not installed; RKB_ERR_CUDA: compilation or loading failed (text in rkb_last_cuda_error). */
int rkb_chain_specialize(rkb_chain* c, int device) {
  if (!c) return RKB_ERR_INVALID;
  if (!c->serial_ok || !c->sk) return RKB_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  const JitKernels* J = nullptr;
  rc = rkb_jit_get(c->n, c->serial_fl, c->serial_shape, &J);
  if (rc) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "%.250s", rkb_jit_log()); return rc; }
  for (DeviceCtx* x : c->ctx) {  // every device this handle has been used on so far; later ones are prepared in get_ctx
    DeviceGuard g(x->device);
    const cudaError_t e = rkb_jit_prepare(*J);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  }
  c->jit = J;
  return RKB_OK;
}
int rkb_chain_is_specialized(const rkb_chain* c) { return c ? (c->jit ? 1 : 0) : RKB_ERR_INVALID; }
unsigned long long rkb_chain_shape(const rkb_chain* c) { return c ? c->serial_shape : 0ull; }

int rkb_eval(rkb_chain* c, int device, size_t N, const double* x, const double* u, double* xdot, int32_t* status,
             unsigned flags, void* stream) {
  return run_eval_like(c, OP_EVAL, device, N, x, u, xdot, nullptr, status, flags, stream);
}

int rkb_gen_forces(rkb_chain* c, int device, size_t N, const double* x, const double* u, double* f, unsigned flags, void* stream) {
  return run_eval_like(c, OP_FORCES, device, N, x, u, f, nullptr, nullptr, flags, stream);
}

int rkb_mass_matrix(rkb_chain* c, int device, size_t N, const double* x, double* M, double* Mdot, unsigned flags, void* stream) {
  return run_eval_like(c, OP_MASS, device, N, x, nullptr, M, Mdot, nullptr, flags, stream);
}

int rkb_chain_frame_count(const rkb_chain* c) { return c ? c->desc.n_frames : RKB_ERR_INVALID; }

/* Samples in one full wave of the RK4 rollout kernel on `device`: SMs x resident CTAs per SM x 128.  A caller that cuts
 * a batch into pieces (to overlap their transfer with the integration of the next piece) should cut at multiples of it:
 * every launch ends with a partial wave, and pieces of, say, 1.7 waves each would cost two. */
long long rkb_chain_wave_samples(rkb_chain* c, int device) {
  if (!c) return RKB_ERR_INVALID;
  if (!c->serial_ok || !c->sk) return 0;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) return RKB_ERR_CUDA;
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); return RKB_ERR_CUDA; }
  return (long long)sms * c->sk->rollout_ctas_per_sm() * c->sk->block;
}

int rkb_frames(rkb_chain* c, int device, size_t N, const double* x, const double* u, double* frames, unsigned flags, void* stream) {
  return run_eval_like(c, OP_FRAMES, device, N, x, u, frames, nullptr, nullptr, flags, stream);
}

/* ---- proximity (kte_proximity.cuh) ---------------------------------------------------------------- */
}  // extern "C"

// CTAs of 128 threads per SM the generated proximity kernels are compiled for (measured: profiles/r2_proximity.md)
constexpr int kProxSpecMinBlocks = 6;

struct rkb_proxy {
  ProxProgram prog;
  std::vector<std::pair<int, int> > finders;
  int n_frames;
  // run-time specialisation (rkb_prox_jit.cu): the chain's program is kept to write the kernel source from
  GenericProgram gp;
  int min_blocks = kProxSpecMinBlocks;
  bool auto_specialize = true;
  mutable std::mutex mu;
  mutable const SourceKernels* spec = nullptr;
  mutable bool auto_done = false;
  mutable long long seen = 0;  // states queried so far: the background compilation starts once 4096 have gone by
  unsigned long long serial = 0;  // identity for the chain's table of checked-steering kernels (a pointer can be reused)
};
std::atomic<unsigned long long> g_proxy_serial{1};

namespace {
const char* const kProxSpecNames[1] = {"rkb_prox_spec_d"};

// The proximity query of one pair at every state of A: the kernels compiled for this chain and pair when they are
// there (rkb_proxy_specialize, or in the background from the first call of >= 4096 states on), else the interpreter.
cudaError_t launch_proximity(const rkb_chain* c, const DeviceCtx* ctx, const rkb_proxy* p, const EvalArgs& A, cudaStream_t s) {
  const SourceKernels* K = nullptr;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    if (!p->spec && !p->auto_done) p->seen += A.n_samples;  // a planner's many small queries count like one large one
    if (!p->spec && p->auto_specialize && c->auto_specialize && !p->auto_done && p->seen >= 4096) {
      const std::string src = rkb_prox_source(p->gp, p->prog, p->min_blocks);
      const SourceKernels* J = nullptr;
      if (src.empty() || rkb_jit_source_poll("prox", src, kProxSpecNames, 1, false, &J) != RKB_OK) p->auto_done = true;  // no NVRTC here: stay as we are
      else if (J) { p->spec = J; p->auto_done = true; }
    }
    K = p->spec;
  }
  if (!K || A.out2.p) return rkb_generic_proximity(ctx->d_prog, c->gp, A, p->prog, s);  // (the two points: interpreter, see kte_prox_spec.cuh)
  if (A.n_samples <= 0) return cudaSuccess;
  void* argv[1] = {const_cast<EvalArgs*>(&A)};
  const unsigned grid = (unsigned)((A.n_samples + 127) / 128);
  return cudaLaunchKernel(K->kernel[0], dim3(grid), dim3(128), argv, 0, s);
}

bool prox_pair_has_finder(int ka, int kb) {  // proxy_query_model.cpp:212-384
  const int lo = ka < kb ? ka : kb, hi = ka < kb ? kb : ka;
  if (lo == RKB_SHAPE_PLANE || lo == RKB_SHAPE_SPHERE) return true;
  if (lo == RKB_SHAPE_CCYLINDER) return hi == RKB_SHAPE_CCYLINDER || hi == RKB_SHAPE_BOX;
  return false;
}
bool lower_shape_2d(const rkb_shape& in, int n_frames, ProxShape* out) {
  if (in.kind < RKB_SHAPE_CIRCLE || in.kind > RKB_SHAPE_RECTANGLE) return false;
  if (in.anchor < -1 || in.anchor >= n_frames) return false;
  const int nd = in.kind == RKB_SHAPE_CIRCLE ? 1 : 2;
  for (int k = 0; k < nd; ++k)
    if (!(in.dims[k] > 0.0) || !std::isfinite(in.dims[k])) return false;
  if (!std::isfinite(in.position[0]) || !std::isfinite(in.position[1]) || !std::isfinite(in.quat[0]) || !std::isfinite(in.quat[1])) return false;
  if (std::fabs(in.quat[0] * in.quat[0] + in.quat[1] * in.quat[1] - 1.0) > 1e-6) return false;
  std::memset(out, 0, sizeof *out);
  out->kind = in.kind;
  out->anchor = in.anchor;
  out->pos[0] = in.position[0]; out->pos[1] = in.position[1];
  out->quat[0] = in.quat[0]; out->quat[1] = in.quat[1];  // (cos, sin) as rot_mat_2D holds them
  for (int k = 0; k < nd; ++k) out->dims[k] = in.dims[k];
  // getBoundingRadius: circle.cpp:31, capped_rectangle.cpp:31, rectangle.cpp:31
  out->brad = in.kind == RKB_SHAPE_CIRCLE ? in.dims[0] : std::sqrt(in.dims[0] * in.dims[0] + in.dims[1] * in.dims[1]) * 0.5;
  return true;
}
bool lower_shape(const rkb_shape& in, int n_frames, ProxShape* out) {
  if (in.kind < RKB_SHAPE_PLANE || in.kind > RKB_SHAPE_BOX) return false;
  if (in.anchor < -1 || in.anchor >= n_frames) return false;
  const int nd = in.kind == RKB_SHAPE_SPHERE ? 1 : in.kind == RKB_SHAPE_BOX ? 3 : 2;
  for (int k = 0; k < nd; ++k)
    if (!(in.dims[k] > 0.0) || !std::isfinite(in.dims[k])) return false;
  double qn = 0.0;
  for (int k = 0; k < 4; ++k) { if (!std::isfinite(in.quat[k])) return false; qn += in.quat[k] * in.quat[k]; }
  if (std::fabs(qn - 1.0) > 1e-6) return false;
  for (int k = 0; k < 3; ++k) if (!std::isfinite(in.position[k])) return false;
  out->kind = in.kind;
  out->anchor = in.anchor;
  for (int k = 0; k < 3; ++k) { out->pos[k] = in.position[k]; out->dims[k] = k < nd ? in.dims[k] : 0.0; }
  // quaternion(const Vector&) normalises what it is given (rotations_3D.hpp:917-920)
  for (int k = 0; k < 4; ++k) out->quat[k] = in.quat[k] / std::sqrt(qn);
  {  // quaternion * vect forms these nine sums first (rotations_3D.hpp:1139-1150); see rot_table in kte_proximity.cuh
    const double w = out->quat[0], x = out->quat[1], y = out->quat[2], z = out->quat[3];
    const double t0 = w * x, t1 = w * y, t2 = w * z, t3 = -x * x, t4 = x * y, t5 = x * z, t6 = -y * y, t7 = y * z, t8 = -z * z;
    double* m = out->rot;
    m[0] = t6 + t8; m[1] = t4 - t2; m[2] = t1 + t5;
    m[3] = t2 + t4; m[4] = t3 + t8; m[5] = t7 - t0;
    m[6] = t5 - t1; m[7] = t0 + t7; m[8] = t3 + t6;
  }
  const double* d = out->dims;
  // getBoundingRadius: plane.cpp:32, sphere.cpp:31, capped_cylinder.cpp:30, cylinder.cpp:33, box.cpp:31
  switch (in.kind) {
    case RKB_SHAPE_PLANE: out->brad = std::sqrt(d[0] * d[0] + d[1] * d[1]) * 0.5; break;
    case RKB_SHAPE_SPHERE: out->brad = d[0]; break;
    case RKB_SHAPE_CCYLINDER: out->brad = d[0] * 0.5 + d[1]; break;
    case RKB_SHAPE_CYLINDER: out->brad = std::sqrt(d[1] * d[1] + 0.25 * d[0] * d[0]); break;
    default: out->brad = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * 0.5; break;
  }
  return true;
}
}  // namespace

extern "C" {

int rkb_proxy_create(const rkb_chain* c, const rkb_shape* m1, int n1, const rkb_shape* m2, int n2, rkb_proxy** out) {
  if (!c || !out || n1 < 0 || n2 < 0 || (n1 > 0 && !m1) || (n2 > 0 && !m2)) return RKB_ERR_INVALID;
  if (n1 > RKB_PROX_MAX_SHAPES || n2 > RKB_PROX_MAX_SHAPES) return RKB_ERR_UNSUPPORTED;
  if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  const bool planar = c->desc.dim == 2;
  for (int k = 0; k < n1 + n2; ++k) {  // planar shapes ride on planar chains, spatial shapes on spatial ones
    const int kind = (k < n1 ? m1[k] : m2[k - n1]).kind;
    if (kind >= RKB_SHAPE_PLANE && kind <= RKB_SHAPE_RECTANGLE && (kind >= RKB_SHAPE_CIRCLE) != planar) return RKB_ERR_UNSUPPORTED;
  }
  rkb_proxy* p = new (std::nothrow) rkb_proxy();
  if (!p) return RKB_ERR_NOMEM;
  std::memset(&p->prog, 0, sizeof p->prog);
  p->prog.n1 = n1;
  p->prog.n2 = n2;
  p->n_frames = c->desc.n_frames;
  p->gp = c->gp;
  p->serial = g_proxy_serial.fetch_add(1);
  for (int k = 0; k < n1 + n2; ++k) {
    const rkb_shape& in = k < n1 ? m1[k] : m2[k - n1];
    if (!(planar ? lower_shape_2d(in, c->desc.n_frames, &p->prog.s[k]) : lower_shape(in, c->desc.n_frames, &p->prog.s[k]))) { delete p; return RKB_ERR_INVALID; }
  }
  for (int a = 0; a < n1; ++a)
    for (int b = 0; b < n2; ++b)
      if (planar || prox_pair_has_finder(m1[a].kind, m2[b].kind)) p->finders.push_back(std::make_pair(a, b));
  {  // which frames outlive the element after their writer (motion_pose in kte_generic.cu walks the same loop)
    bool keep[RKB_GEN_MAX_FRAMES] = {};
    for (int k = 0; k < n1 + n2; ++k)
      if (p->prog.s[k].anchor >= 0) keep[p->prog.s[k].anchor] = true;
    int last = c->gp.base_frame;
    for (int e = 0; e < c->gp.n_elements; ++e) {
      const GenericElement& E = c->gp.el[e];
      if (E.kind != RKB_REVOLUTE_3D && E.kind != RKB_PRISMATIC_3D && E.kind != RKB_RIGID_LINK_3D && E.kind != RKB_FREE_3D &&
          E.kind != RKB_REVOLUTE_2D && E.kind != RKB_PRISMATIC_2D && E.kind != RKB_RIGID_LINK_2D && E.kind != RKB_FREE_2D) continue;
      if (E.fa != last) keep[E.fa] = true;
      last = E.fb;
    }
    p->prog.n_slots = 0;
    for (int f = 0; f < RKB_GEN_MAX_FRAMES; ++f)
      p->prog.slot_of[f] = (f < c->desc.n_frames && keep[f]) ? (int8_t)p->prog.n_slots++ : (int8_t)-1;
  }
  *out = p;
  return RKB_OK;
}

void rkb_proxy_destroy(rkb_proxy* p) { delete p; }

int rkb_proxy_set_option(rkb_proxy* p, int option, long long value) {
  if (!p) return RKB_ERR_INVALID;
  std::lock_guard<std::mutex> lock(p->mu);
  switch (option) {
    case RKB_PROXY_OPT_AUTO_SPECIALIZE: p->auto_specialize = value != 0; return RKB_OK;
    case RKB_PROXY_OPT_MIN_BLOCKS:
      if (value < 1 || value > 8) return RKB_ERR_INVALID;
      p->min_blocks = (int)value;
      return RKB_OK;
    default: return RKB_ERR_INVALID;
  }
}

int rkb_proxy_specialize(rkb_proxy* p, int device) {
  if (!p) return RKB_ERR_INVALID;
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  std::lock_guard<std::mutex> lock(p->mu);
  const std::string src = rkb_prox_source(p->gp, p->prog, p->min_blocks);
  if (src.empty()) return RKB_ERR_UNSUPPORTED;
  const SourceKernels* J = nullptr;
  const int rc = rkb_jit_source_get("prox", src, kProxSpecNames, 1, false, &J);
  if (rc) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "run-time compilation failed: %.200s", rkb_jit_log()); return rc; }
  p->spec = J;
  return RKB_OK;
}

int rkb_proxy_is_specialized(const rkb_proxy* p) {
  if (!p) return 0;
  std::lock_guard<std::mutex> lock(p->mu);
  return p->spec ? 1 : 0;
}

int rkb_proxy_source(const rkb_proxy* p, char* out, size_t size) {
  if (!p) return RKB_ERR_INVALID;
  const std::string src = rkb_prox_source(p->gp, p->prog, p->min_blocks);
  if (src.empty()) return RKB_ERR_UNSUPPORTED;
  if (out && size > src.size()) std::memcpy(out, src.c_str(), src.size() + 1);
  else if (out) return RKB_ERR_INVALID;
  return (int)src.size() + 1;
}

int rkb_proxy_finder_count(const rkb_proxy* p) { return p ? (int)p->finders.size() : RKB_ERR_INVALID; }

int rkb_proxy_finder(const rkb_proxy* p, int k, int* i1, int* i2) {
  if (!p || k < 0 || k >= (int)p->finders.size()) return RKB_ERR_INVALID;
  if (i1) *i1 = p->finders[k].first;
  if (i2) *i2 = p->finders[k].second;
  return RKB_OK;
}

int rkb_chain_program(const rkb_chain* c, void* out, size_t size) {
  if (!c || !out || size < sizeof(GenericProgram)) return RKB_ERR_INVALID;
  if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  std::memcpy(out, &c->gp, sizeof(GenericProgram));
  return (int)sizeof(GenericProgram);
}

int rkb_proxy_program(const rkb_proxy* p, void* out, size_t size) {
  if (!p || !out || size < sizeof(ProxProgram)) return RKB_ERR_INVALID;
  std::memcpy(out, &p->prog, sizeof(ProxProgram));
  return (int)sizeof(ProxProgram);
}

int rkb_min_distance(rkb_chain* c, const rkb_proxy* p, int device, size_t N, const double* x, double* distance, int32_t* finder,
                     double* points, unsigned flags, void* stream) {
  if (!c || !p) return RKB_ERR_INVALID;
  if (N == 0) return RKB_OK;
  if (!x || !distance) return RKB_ERR_INVALID;
  if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  if (p->n_frames != c->desc.n_frames) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const void* dx = nullptr;
  void *dd = nullptr, *dp = nullptr, *df = nullptr;
  if ((rc = stage_in(ctx->in_x, x, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if ((rc = stage_out(ctx->out_a, distance, N * sizeof(double), L.device, &dd))) return rc;
  if ((rc = stage_out(ctx->out_b, points, N * 6 * sizeof(double), L.device, &dp))) return rc;
  if ((rc = stage_out(ctx->st, finder, N * sizeof(int32_t), L.device, &df))) return rc;
  EvalArgs A;
  A.x = cview((const double*)dx, (long long)N, nx, L.soa, L.blocked);
  A.u = cview((const double*)dx, (long long)N, 1, L.soa);
  A.out = view((double*)dd, (long long)N, 1, L.soa);
  A.out2 = view((double*)dp, (long long)N, 6, L.soa);
  A.status = (int32_t*)df;
  A.n_samples = (long long)N;
  CU(cudaEventRecord(ctx->ev0, s));
  const cudaError_t e = launch_proximity(c, ctx, p, A, s);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  c->launches += 1;
  if ((rc = unstage_out(dd, distance, N * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dp, points, N * 6 * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(df, finder, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

int rkb_collision_points(rkb_chain* c, const rkb_proxy* p, int device, size_t N, const double* x, int max_records, int32_t* count,
                         int32_t* finder, double* records, unsigned flags, void* stream) {
  if (!c || !p) return RKB_ERR_INVALID;
  if (max_records < 1 || max_records > 2 * RKB_PROX_MAX_SHAPES * RKB_PROX_MAX_SHAPES) return RKB_ERR_INVALID;
  if (N == 0) return RKB_OK;
  if (!x || !count || !records) return RKB_ERR_INVALID;
  if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  if (p->n_frames != c->desc.n_frames) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  if (L.soa) return RKB_ERR_UNSUPPORTED;  // records are per state: AoS only
  const int nx = c->nx;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const void* dx = nullptr;
  void *dr = nullptr, *df = nullptr, *dc = nullptr;
  const size_t M = (size_t)max_records;
  if ((rc = stage_in(ctx->in_x, x, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if ((rc = stage_out(ctx->out_a, records, N * M * 7 * sizeof(double), L.device, &dr))) return rc;
  if ((rc = stage_out(ctx->out_b, finder, N * M * sizeof(int32_t), L.device, &df))) return rc;
  if ((rc = stage_out(ctx->st, count, N * sizeof(int32_t), L.device, &dc))) return rc;
  EvalArgs A;
  A.x = cview((const double*)dx, (long long)N, nx, false, L.blocked);
  A.u = cview((const double*)dx, (long long)N, 1, false);
  A.out = view((double*)dr, (long long)N, (int)(M * 7), false);
  A.out2 = view((double*)nullptr, (long long)N, 1, false);
  A.status = (int32_t*)dc;
  A.n_samples = (long long)N;
  CU(cudaEventRecord(ctx->ev0, s));
  const cudaError_t e = rkb_generic_collisions(ctx->d_prog, c->gp, A, p->prog, max_records, (int32_t*)df, s);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  c->launches += 1;
  if ((rc = unstage_out(dr, records, N * M * 7 * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(df, finder, N * M * sizeof(int32_t), L.device, s))) return rc;
  if ((rc = unstage_out(dc, count, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

int rkb_is_free(rkb_chain* c, int device, size_t N, const double* x, const rkb_proxy* const* pairs, int n_pairs, int32_t* is_free,
                unsigned flags, void* stream) {
  if (!c || n_pairs < 1 || !pairs) return RKB_ERR_INVALID;
  if (N == 0) return RKB_OK;
  if (!x || !is_free) return RKB_ERR_INVALID;
  if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  for (int p = 0; p < n_pairs; ++p)
    if (!pairs[p] || pairs[p]->n_frames != c->desc.n_frames) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const void* dx = nullptr;
  void* dfree = nullptr;
  if ((rc = stage_in(ctx->in_x, x, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if ((rc = stage_out(ctx->st, is_free, N * sizeof(int32_t), L.device, &dfree))) return rc;
  if ((rc = ctx->scratch_o.ensure(N * sizeof(double) * (size_t)n_pairs))) return rc;
  CU(cudaEventRecord(ctx->ev0, s));
  for (int p = 0; p < n_pairs; ++p) {
    EvalArgs E;
    E.x = cview((const double*)dx, (long long)N, nx, L.soa, L.blocked);
    E.u = cview((const double*)dx, (long long)N, 1, L.soa);
    E.out = view((double*)ctx->scratch_o.p + (size_t)p * N, (long long)N, 1, false);
    E.out2 = view((double*)nullptr, (long long)N, 6, false);
    E.status = nullptr;
    E.n_samples = (long long)N;
    const cudaError_t e = launch_proximity(c, ctx, pairs[p], E, s);
    if (e != cudaSuccess) return cuda_fail(e, "proximity kernel");
    c->launches += 1;
  }
  const cudaError_t e = rkb_free_combine((const double*)ctx->scratch_o.p, n_pairs, (long long)N, (int32_t*)dfree, s);
  if (e != cudaSuccess) return cuda_fail(e, "free combine");
  c->launches += 1;
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  if ((rc = unstage_out(dfree, is_free, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

int rkb_twist_shaping_rows(const rkb_chain* c) { return c ? tmt_rows(c->desc) : RKB_ERR_INVALID; }

/* Mcm of get_TMT_TdMT (mass_matrix_calculator.cpp:265-285): block diagonal, the masses and the tensors */
int rkb_twist_shaping_mcm(const rkb_chain* c, double* Mcm) {
  if (!c || !Mcm) return RKB_ERR_INVALID;
  const int rows = tmt_rows(c->desc);
  std::memset(Mcm, 0, sizeof(double) * (size_t)rows * (size_t)rows);
  int row = 0;
  for (int pass = 0; pass < 3; ++pass)
    for (int e = 0; e < c->desc.n_elements; ++e) {
      const rkb_element& E = c->elements[e];
      if (pass == 0 && E.kind == RKB_INERTIA_GEN) { Mcm[row * rows + row] = E.p[0]; row += 1; }
      else if (pass == 1 && E.kind == RKB_INERTIA_2D) {
        Mcm[row * rows + row] = Mcm[(row + 1) * rows + row + 1] = E.p[0];
        Mcm[(row + 2) * rows + row + 2] = E.p[1];
        row += 3;
      } else if (pass == 2 && E.kind == RKB_INERTIA_3D) {
        for (int k = 0; k < 3; ++k) Mcm[(row + k) * rows + row + k] = E.p[0];
        const double* I = &E.p[1];
        const int r = row + 3;
        const double t[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Mcm[(r + a) * rows + r + b] = t[3 * a + b];
        row += 6;
      }
    }
  return RKB_OK;
}

/* Jacobian of one frame with respect to the coordinates and its time derivative (see kte_generic.cu: frame_jac) */
int rkb_frame_jacobian(rkb_chain* c, int device, size_t N, const double* x, int frame, uint64_t upstream, double* J, double* Jdot,
                       unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  {  // bits 0 .. n-1: coordinates; bits 32 .. 32 + n_free - 1: free joints (as rkb_element::upstream)
    const uint64_t allowed = (c->n >= 32 ? 0xffffffffull : ((1ull << c->n) - 1ull)) | (((1ull << c->n_free) - 1ull) << 32);
    if (frame < 0 || frame >= c->desc.n_frames || (upstream & ~allowed)) return RKB_ERR_INVALID;
  }
  if (N == 0) return RKB_OK;
  if (!x || !J) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx, rows = c->desc.dim == 3 ? 6 : 3, dim = rows * c->na;
  if (dim == 0) return RKB_OK;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const void* dx = nullptr;
  void *dJ = nullptr, *dJd = nullptr;
  if ((rc = stage_in(ctx->in_x, x, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if ((rc = stage_out(ctx->out_a, J, N * dim * sizeof(double), L.device, &dJ))) return rc;
  if ((rc = stage_out(ctx->out_b, Jdot, N * dim * sizeof(double), L.device, &dJd))) return rc;
  EvalArgs A;
  A.x = cview((const double*)dx, (long long)N, nx, L.soa, L.blocked);
  A.u = cview((const double*)dx, (long long)N, 1, L.soa);
  A.out = view((double*)dJ, (long long)N, dim, L.soa);
  A.out2 = view((double*)dJd, (long long)N, dim, L.soa);
  A.status = nullptr;
  A.n_samples = (long long)N;
  CU(cudaEventRecord(ctx->ev0, s));
  const cudaError_t e = rkb_generic_frame_jac(ctx->d_prog, c->gp, A, frame, (unsigned)(upstream & 0xffffull) | ((unsigned)((upstream >> 32) & 0xffull) << RKB_GEN_FREE_BIT), s);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  c->launches += 1;
  if ((rc = unstage_out(dJ, J, N * dim * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dJd, Jdot, N * dim * sizeof(double), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

int rkb_twist_shaping(rkb_chain* c, int device, size_t N, const double* x, double* Tcm, double* Tcm_dot, unsigned flags, void* stream) {
  return run_eval_like(c, OP_TMT, device, N, x, nullptr, Tcm, Tcm_dot, nullptr, flags, stream);
}

}  // extern "C"

namespace {

// What to integrate: scheme (table == false: the dedicated RK4 kernels), step, steps per control
// interval, number of intervals.
struct RolloutPlan {
  bool use_table = false;
  RkTable table;
  double dt = 0.0;
  int n_steps = 0, n_intervals = 1;
};

// The schemes of fixed_step_integrators.hpp rewritten as "start of step + weighted increments"
// (the reference updates the state incrementally; the weights below are those sums in closed form).
int make_plan(const rkb_rollout_opts* o, RolloutPlan& pl) {
  if (!o || o->reserved != 0) return RKB_ERR_INVALID;
  if (o->dt == 0.0 || !std::isfinite(o->dt) || o->steps_per_interval < 0 || o->n_intervals < 1) return RKB_ERR_INTEGRATION;
  pl.dt = o->dt; pl.n_steps = o->steps_per_interval; pl.n_intervals = o->n_intervals;
  std::memset(&pl.table, 0, sizeof pl.table);
  double (*c)[RKB_RK_MAX_STAGES] = pl.table.c;
  switch (o->scheme) {
    case RKB_SCHEME_RK4: pl.use_table = false; return RKB_OK;
    case RKB_SCHEME_EULER:  // :78  x += f dt
      pl.table.stages = 1; c[0][0] = 1.0; break;
    case RKB_SCHEME_MIDPOINT:  // :193-199  w = x + f dt/2 ; x += f(w) dt
      pl.table.stages = 2; c[0][0] = 0.5; c[1][1] = 1.0; break;
    case RKB_SCHEME_RK5:  // :368-395
      pl.table.stages = 6;
      c[0][0] = 0.25;
      c[1][0] = 0.25 - 5.0 / 32.0; c[1][1] = 9.0 / 32.0;
      c[2][0] = c[1][0] + 276165.0 / 351520.0; c[2][1] = c[1][1] - 1250865.0 / 351520.0; c[2][2] = 1167360.0 / 351520.0;
      c[3][0] = 439.0 / 216.0; c[3][1] = -8.0; c[3][2] = 3680.0 / 513.0; c[3][3] = -845.0 / 4104.0;
      c[4][0] = -8.0 / 27.0; c[4][1] = 2.0; c[4][2] = -3544.0 / 2565.0; c[4][3] = 1859.0 / 4104.0; c[4][4] = -11.0 / 40.0;
      c[5][0] = 16.0 / 135.0; c[5][2] = 6656.0 / 12825.0; c[5][3] = 28561.0 / 56430.0; c[5][4] = -9.0 / 50.0; c[5][5] = 2.0 / 55.0;
      break;
    default: return RKB_ERR_INVALID;
  }
  pl.use_table = true;
  return RKB_OK;
}

// One launch per control interval over device-resident AoS/SoA views.  Interval 0 reads x0 and every
// later one the previous end state, in place in xout; the end state of interval j also goes to slot j
// of the trajectory when one is wanted; status bits accumulate.
//   u:    input k of interval j of sample i at u.p[i * u.si + j * u_sj + k * u.sk]
//   traj: likewise with traj_sj
int launch_intervals(rkb_chain* c, DeviceCtx* ctx, const RolloutPlan& pl, long long n, ConstBatchView x0, ConstBatchView u, long long u_sj,
                     BatchView xout, BatchView traj, long long traj_sj, int32_t* status, cudaStream_t s) {
  if (pl.n_intervals > 1 && !pl.use_table && c->serial_ok && c->sk && c->fused_sequence) {
    RolloutSeqArgs A;  // RK4 on the serial kernels: the whole sequence in one launch
    A.x0 = x0; A.u = u; A.xout = xout; A.traj = traj; A.status = status;
    A.n_samples = n; A.u_sj = u_sj; A.traj_sj = traj_sj; A.dt = pl.dt; A.n_steps = pl.n_steps; A.n_intervals = pl.n_intervals;
    A.half_step_nodes = 0; A.pad = 0;
    cudaError_t e = c->jit ? rkb_jit_launch(*c->jit, RKB_JIT_ROLLOUT_SEQ, c->sp, &A, nullptr, A.n_samples, 0, s)
                    : use_split(c, A.n_samples) ? c->sk->rollout_seq_duo(c->sp, A, s) : c->sk->rollout_seq(c->sp, A, s);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    c->launches += 1;
    return RKB_OK;
  }
  for (int j = 0; j < pl.n_intervals; ++j) {
    RolloutArgs A;
    A.x0 = j == 0 ? x0 : ConstBatchView{xout.p, xout.si, xout.sk, xout.blocked};
    A.u = ConstBatchView{u.p + j * u_sj, u.si, u.sk};
    A.xout = xout;
    A.traj = traj.p ? BatchView{traj.p + j * traj_sj, traj.si, traj.sk, traj.blocked} : BatchView{nullptr, 0, 0, 0};
    A.status = status;
    A.n_samples = n;
    A.x0_div = 1;
    A.dt = pl.dt;
    A.n_steps = pl.n_steps;
    A.status_or = j > 0;
    A.active = nullptr;
    A.u_node_stride = 0;
    const int rc = launch_rollout(c, ctx, A, pl.use_table ? &pl.table : nullptr, s);
    if (rc) return rc;
  }
  return RKB_OK;
}

constexpr int kPipeChunks = 8;
constexpr size_t kPipeMinSamples = 1u << 16;

int ensure_pipe(DeviceCtx* ctx) {
  if (ctx->s_in) return 0;
  CU(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&ctx->s_k[0], cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&ctx->s_k[1], cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < kPipeChunks; ++i) {
    CU(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  return 0;
}

// Host AoS buffers, large batch: the batch is cut into chunks whose host->device copy, rollout
// kernel and device->host copy overlap (copy-in stream, two alternating compute streams so that
// one chunk's tail wave overlaps the next chunk's head, copy-out stream).  Only the first copy-in
// and the last copy-out stay exposed.  Pinned host memory is what makes the copies asynchronous.
int rollout_host_issue(rkb_chain* c, DeviceCtx* ctx, size_t N, const double* x0, const double* u, const RolloutPlan& pl,
                       double* x_out, double* x_traj, int32_t* status, cudaStream_t s, bool join_caller, bool blocked = false) {
  const int nx = c->nx;
  const size_t nu = (size_t)c->nu * pl.n_intervals;  // doubles of input per sample
  const size_t nt = (size_t)nx * pl.n_intervals;     // doubles of trajectory per sample
  int rc;
  if ((rc = ensure_pipe(ctx))) return rc;
  if ((rc = ctx->in_x.ensure(N * nx * sizeof(double)))) return rc;
  if (nu > 0 && (rc = ctx->in_u.ensure(N * nu * sizeof(double)))) return rc;
  if ((rc = ctx->out_a.ensure(N * nx * sizeof(double)))) return rc;
  if (x_traj && (rc = ctx->out_b.ensure(N * nt * sizeof(double)))) return rc;
  if ((rc = ctx->st.ensure(N * sizeof(int32_t)))) return rc;
  double* dx = (double*)ctx->in_x.p;
  double* du = (double*)ctx->in_u.p;
  double* dout = (double*)ctx->out_a.p;
  double* dtraj = x_traj ? (double*)ctx->out_b.p : nullptr;
  int32_t* dst = (int32_t*)ctx->st.p;
  if (join_caller) {  // order after whatever the caller queued on its stream
    CU(cudaEventRecord(ctx->ev_join, s));
    CU(cudaStreamWaitEvent(ctx->s_in, ctx->ev_join, 0));
    CU(cudaStreamWaitEvent(ctx->s_k[0], ctx->ev_join, 0));
  }
  CU(cudaEventRecord(ctx->ev0, ctx->s_k[0]));
  const size_t per = ((N + kPipeChunks - 1) / kPipeChunks + 127) / 128 * 128;
  int last_k = 0;
  for (int i = 0; i < kPipeChunks; ++i) {
    const size_t lo = per * i;
    if (lo >= N) break;
    const size_t m = (lo + per <= N) ? per : N - lo;
    CU(cudaMemcpyAsync(dx + lo * nx, x0 + lo * nx, m * nx * sizeof(double), cudaMemcpyHostToDevice, ctx->s_in));
    if (nu > 0) CU(cudaMemcpyAsync(du + lo * nu, u + lo * nu, m * nu * sizeof(double), cudaMemcpyHostToDevice, ctx->s_in));
    CU(cudaEventRecord(ctx->ev_in[i], ctx->s_in));
    cudaStream_t sk = ctx->s_k[i & 1];
    CU(cudaStreamWaitEvent(sk, ctx->ev_in[i], 0));
    if ((rc = launch_intervals(c, ctx, pl, (long long)m, cview(dx + lo * nx, (long long)m, nx, false, blocked),
                               ConstBatchView{nu > 0 ? du + lo * nu : dx, (long long)(nu > 0 ? nu : 1), 1, 0}, c->nu,
                               view(dout + lo * nx, (long long)m, nx, false, blocked),
                               BatchView{dtraj ? dtraj + lo * nt : nullptr, (long long)nt, 1, blocked ? 1 : 0}, nx, dst + lo, sk)))
      return rc;
    CU(cudaEventRecord(ctx->ev_k[i], sk));
    CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[i], 0));
    CU(cudaMemcpyAsync(x_out + lo * nx, dout + lo * nx, m * nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_out));
    if (x_traj) CU(cudaMemcpyAsync(x_traj + lo * nt, dtraj + lo * nt, m * nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_out));
    if (status) CU(cudaMemcpyAsync(status + lo, dst + lo, m * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_out));
    last_k = i;
  }
  // ev1 closes the timed span on the stream that ran the last kernel, after the other one joined it
  cudaStream_t sl = ctx->s_k[last_k & 1];
  if (last_k > 0) CU(cudaStreamWaitEvent(sl, ctx->ev_k[last_k - 1], 0));
  CU(cudaEventRecord(ctx->ev1, sl));
  ctx->timed = true;
  c->last = ctx;
  return RKB_OK;
}

int rollout_host_wait(DeviceCtx* ctx) {
  CU(cudaStreamSynchronize(ctx->s_out));
  CU(cudaStreamSynchronize(ctx->s_k[0]));
  CU(cudaStreamSynchronize(ctx->s_k[1]));
  return RKB_OK;
}

int rollout_host_pipelined(rkb_chain* c, DeviceCtx* ctx, size_t N, const double* x0, const double* u, const RolloutPlan& pl,
                           double* x_out, double* x_traj, int32_t* status, cudaStream_t s, bool blocked) {
  int rc = rollout_host_issue(c, ctx, N, x0, u, pl, x_out, x_traj, status, s, true, blocked);
  if (rc) return rc;
  if ((rc = rollout_host_wait(ctx))) return rc;
  // let the caller's stream observe completion as well
  CU(cudaEventRecord(ctx->ev_join, ctx->s_out));
  CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));
  return RKB_OK;
}

}  // namespace

extern "C" {

static int do_rollout(rkb_chain* c, int device, size_t N, const double* x0, const double* u, const RolloutPlan& pl,
                      double* x_out, double* x_traj, int32_t* status, unsigned flags, void* stream) {
  if (N == 0) return RKB_OK;
  if (!x0 || !x_out || (c->nu > 0 && !u)) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx, nu = c->nu, J = pl.n_intervals;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  maybe_auto_specialize(c, (long long)N);
  cudaStream_t s = (cudaStream_t)stream;
  if (!L.device && !L.soa && N >= kPipeMinSamples && pl.n_steps > 0 && c->host_pipeline)
    return rollout_host_pipelined(c, ctx, N, x0, u, pl, x_out, x_traj, status, s, L.blocked);
  const void *dx = nullptr, *du = nullptr;
  void *dout = nullptr, *dtraj = nullptr, *dst = nullptr;
  if ((rc = stage_in(ctx->in_x, x0, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if (nu > 0) { if ((rc = stage_in(ctx->in_u, u, N * nu * J * sizeof(double), L.device, s, &du))) return rc; }
  if ((rc = stage_out(ctx->out_a, x_out, N * nx * sizeof(double), L.device, &dout))) return rc;
  if ((rc = stage_out(ctx->out_b, x_traj, N * nx * J * sizeof(double), L.device, &dtraj))) return rc;
  if ((rc = stage_out(ctx->st, status, N * sizeof(int32_t), L.device, &dst))) return rc;
  // AoS [N][J][nu]: sample stride J nu, interval stride nu; SoA [J][nu][N]: sample stride 1, interval stride nu N
  const ConstBatchView uv = L.soa ? ConstBatchView{(const double*)(du ? du : dx), 1, (long long)N}
                                  : ConstBatchView{(const double*)(du ? du : dx), (long long)(nu > 0 ? nu * J : 1), 1};
  const BatchView tv = L.soa ? BatchView{(double*)dtraj, 1, (long long)N, L.blocked ? 1 : 0}
                             : BatchView{(double*)dtraj, (long long)nx * J, 1, L.blocked ? 1 : 0};
  CU(cudaEventRecord(ctx->ev0, s));
  if ((rc = launch_intervals(c, ctx, pl, (long long)N, cview((const double*)dx, (long long)N, nx, L.soa, L.blocked), uv,
                             L.soa ? (long long)nu * (long long)N : (long long)nu, view((double*)dout, (long long)N, nx, L.soa, L.blocked), tv,
                             L.soa ? (long long)nx * (long long)N : (long long)nx, (int32_t*)dst, s)))
    return rc;
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  if ((rc = unstage_out(dout, x_out, N * nx * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dtraj, x_traj, N * nx * J * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dst, status, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

int rkb_rollout_rk4(rkb_chain* c, int device, size_t N, const double* x0, const double* u, double dt, int n_steps,
                    double* x_out, int32_t* status, unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (dt == 0.0 || n_steps < 0 || !std::isfinite(dt)) return RKB_ERR_INTEGRATION;  // fixed_step_integrators.hpp:258-266
  RolloutPlan pl;
  pl.dt = dt; pl.n_steps = n_steps; pl.n_intervals = 1;
  return do_rollout(c, device, N, x0, u, pl, x_out, nullptr, status, flags, stream);
}

int rkb_rollout(rkb_chain* c, int device, size_t N, const double* x0, const double* u, const rkb_rollout_opts* opts,
                double* x_out, double* x_traj, int32_t* status, unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  RolloutPlan pl;
  const int rc = make_plan(opts, pl);
  if (rc) return rc;
  return do_rollout(c, device, N, x0, u, pl, x_out, x_traj, status, flags, stream);
}

/* ctrl::detail::runge_kutta4_integrate_impl (ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp:50-97): RK4 with an
 * input TRAJECTORY — the input is read at t for the first evaluation of a step, at t + dt/2 for the second and third
 * and at t + dt for the fourth.  The batched form takes the trajectory sampled at those instants: 2 n_steps + 1 nodes. */
int rkb_rollout_rk4_inputs(rkb_chain* c, int device, size_t N, const double* x0, const double* u_nodes, double dt, int n_steps,
                           double* x_out, int32_t* status, unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (dt == 0.0 || n_steps < 0 || !std::isfinite(dt)) return RKB_ERR_INTEGRATION;
  if (c->nu == 0) return rkb_rollout_rk4(c, device, N, x0, nullptr, dt, n_steps, x_out, status, flags, stream);  // nothing to sample
  if (N == 0) return RKB_OK;
  if (!x0 || !x_out || !u_nodes) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx, nu = c->nu;
  const long long J = 2LL * n_steps + 1;
  std::lock_guard<std::mutex> lock(c->mu);
  if (!(c->serial_ok && c->sk) && !c->generic_ok) return RKB_ERR_UNSUPPORTED;
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const void *dx = nullptr, *du = nullptr;
  void *dout = nullptr, *dst = nullptr;
  if ((rc = stage_in(ctx->in_x, x0, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if ((rc = stage_in(ctx->in_u, u_nodes, N * (size_t)nu * (size_t)J * sizeof(double), L.device, s, &du))) return rc;
  if ((rc = stage_out(ctx->out_a, x_out, N * nx * sizeof(double), L.device, &dout))) return rc;
  if ((rc = stage_out(ctx->st, status, N * sizeof(int32_t), L.device, &dst))) return rc;
  // AoS [N][J][nu]: sample stride J nu, node stride nu; SoA [J][nu][N]: sample stride 1, node stride nu N, input stride N
  const ConstBatchView uv = L.soa ? ConstBatchView{(const double*)du, 1, (long long)N, 0} : ConstBatchView{(const double*)du, (long long)nu * J, 1, 0};
  const long long u_sj = L.soa ? (long long)nu * (long long)N : (long long)nu;
  CU(cudaEventRecord(ctx->ev0, s));
  cudaError_t e;
  if (c->serial_ok && c->sk) {  // (a run-time specialised set has no node kernel: the shipped one serves)
    RolloutSeqArgs A;
    A.x0 = cview((const double*)dx, (long long)N, nx, L.soa, L.blocked); A.u = uv;
    A.xout = view((double*)dout, (long long)N, nx, L.soa, L.blocked); A.traj = BatchView{nullptr, 0, 0, 0};
    A.status = (int32_t*)dst; A.n_samples = (long long)N; A.u_sj = u_sj; A.traj_sj = 0; A.dt = dt; A.n_steps = n_steps; A.n_intervals = 1;
    A.half_step_nodes = 1; A.pad = 0;
    e = c->sk->rollout_seq(c->sp, A, s);
  } else {
    RolloutArgs A;
    A.x0 = cview((const double*)dx, (long long)N, nx, L.soa, L.blocked); A.u = uv;
    A.xout = view((double*)dout, (long long)N, nx, L.soa, L.blocked); A.traj = BatchView{nullptr, 0, 0, 0};
    A.status = (int32_t*)dst; A.n_samples = (long long)N; A.x0_div = 1; A.dt = dt; A.n_steps = n_steps; A.status_or = 0; A.active = nullptr;
    A.u_node_stride = u_sj;
    e = rkb_generic_rollout(ctx->d_prog, c->gp, A, nullptr, s);
  }
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  c->launches += 1;
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  if ((rc = unstage_out(dout, x_out, N * nx * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dst, status, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

/* Rollout of this rank's block of a sample-sharded batch with the all-gather folded into the kernel: the end states go
 * straight to their rows in n_dest copies of the gathered batch — the caller's own and the peers', mapped into this
 * process (CUDA IPC / symmetric memory) — as coalesced stores over NVLink.  Device buffers, AoS, serial chains. */
int rkb_rollout_rk4_scatter(rkb_chain* c, int device, size_t N, const double* x0, const double* u, double dt, int n_steps,
                            int n_dest, double* const* x_out_dest, int32_t* const* status_dest, size_t row_offset,
                            unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (dt == 0.0 || n_steps < 0 || !std::isfinite(dt)) return RKB_ERR_INTEGRATION;
  if (n_dest < 1 || n_dest > RKB_MAX_DEST || !x_out_dest) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  if (!L.device || L.soa) return RKB_ERR_UNSUPPORTED;
  if (!c->serial_ok || !c->sk) return RKB_ERR_UNSUPPORTED;
  if (N == 0) return RKB_OK;
  if (!x0 || (c->nu > 0 && !u)) return RKB_ERR_INVALID;
  for (int d = 0; d < n_dest; ++d) if (!x_out_dest[d]) return RKB_ERR_INVALID;
  const int nx = c->nx, nu = c->nu;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  RolloutScatterArgs A;
  std::memset(&A, 0, sizeof A);
  A.x0 = cview(x0, (long long)N, nx, false, L.blocked);
  A.u = cview(nu > 0 ? u : x0, (long long)N, nu > 0 ? nu : 1, false);
  for (int d = 0; d < n_dest; ++d) { A.xout[d] = x_out_dest[d]; A.status[d] = status_dest ? status_dest[d] : nullptr; }
  A.n_samples = (long long)N; A.row_offset = (long long)row_offset; A.dt = dt; A.n_steps = n_steps; A.n_dest = n_dest;
  A.blocked = L.blocked ? 1 : 0;
  CU(cudaEventRecord(ctx->ev0, s));
  const cudaError_t e = c->sk->rollout_scatter(c->sp, A, s);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  c->launches += 1;
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  return RKB_OK;
}

/* Host-buffer rollout sharded over several GPUs of one box from one process: contiguous block
 * partition by sample index (sizes differ by at most one 128-sample tile), every device runs its
 * own copy-in / compute / copy-out pipeline concurrently, no inter-GPU communication — the
 * "gather" is the device->host copies landing in the caller's output buffer. */
int rkb_rollout_rk4_multi(rkb_chain* c, int n_devices, const int* devices, size_t N, const double* x0, const double* u,
                          double dt, int n_steps, double* x_out, int32_t* status) {
  if (!c || n_devices < 1 || !devices) return RKB_ERR_INVALID;
  if (dt == 0.0 || n_steps < 0 || !std::isfinite(dt)) return RKB_ERR_INTEGRATION;
  if (N == 0) return RKB_OK;
  if (!x0 || !x_out || (c->nu > 0 && !u)) return RKB_ERR_INVALID;
  const int nx = c->nx, nu = c->nu;
  std::lock_guard<std::mutex> lock(c->mu);
  int prev = -1;
  cudaGetDevice(&prev);
  std::vector<DeviceCtx*> used(n_devices, nullptr);
  const size_t tiles = (N + 127) / 128;
  int rc = RKB_OK;
  for (int g = 0; g < n_devices && rc == RKB_OK; ++g) {
    const size_t lo = std::min(N, (tiles * g / n_devices) * 128), hi = std::min(N, (tiles * (g + 1) / n_devices) * 128);
    if (hi <= lo) continue;
    if (cudaSetDevice(devices[g]) != cudaSuccess) { cudaGetLastError(); rc = RKB_ERR_CUDA; break; }
    DeviceCtx* ctx = nullptr;
    if ((rc = get_ctx(c, devices[g], &ctx))) break;
    if (n_steps == 0) {
      std::memcpy(x_out + lo * nx, x0 + lo * nx, (hi - lo) * nx * sizeof(double));
      if (status) std::memset(status + lo, 0, (hi - lo) * sizeof(int32_t));
      continue;
    }
    RolloutPlan pl;
    pl.dt = dt; pl.n_steps = n_steps; pl.n_intervals = 1;
    rc = rollout_host_issue(c, ctx, hi - lo, x0 + lo * nx, nu > 0 ? u + lo * nu : nullptr, pl, x_out + lo * nx, nullptr,
                            status ? status + lo : nullptr, nullptr, false);
    used[g] = ctx;
  }
  for (int g = 0; g < n_devices; ++g) {
    if (!used[g]) continue;
    cudaSetDevice(devices[g]);
    const int w = rollout_host_wait(used[g]);
    if (rc == RKB_OK) rc = w;
  }
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

int rkb_steer_batch(rkb_chain* c, int device, size_t P, size_t R, const double* x0, const double* goal, const double* u,
                    double dt, int n_steps, int32_t* best_idx, double* best_x, double* best_cost, int32_t* status,
                    unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (dt == 0.0 || n_steps < 0 || !std::isfinite(dt)) return RKB_ERR_INTEGRATION;
  if (P == 0) return RKB_OK;
  if (R == 0) return RKB_ERR_INVALID;
  if (!x0 || !goal || !best_idx || !best_x || (c->nu > 0 && !u)) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx, nu = c->nu;
  const size_t T = P * R;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  maybe_auto_specialize(c, (long long)P * R);
  cudaStream_t s = (cudaStream_t)stream;
  const void *dx0 = nullptr, *dgoal = nullptr, *du = nullptr;
  void *didx = nullptr, *dbx = nullptr, *dbc = nullptr, *dst = nullptr;
  if ((rc = stage_in(ctx->in_x, x0, P * nx * sizeof(double), L.device, s, &dx0))) return rc;
  if ((rc = stage_in(ctx->in_goal, goal, P * nx * sizeof(double), L.device, s, &dgoal))) return rc;
  if (nu > 0) { if ((rc = stage_in(ctx->in_u, u, T * nu * sizeof(double), L.device, s, &du))) return rc; }
  if ((rc = stage_out(ctx->out_idx, best_idx, P * sizeof(int32_t), L.device, &didx))) return rc;
  if ((rc = stage_out(ctx->out_a, best_x, P * nx * sizeof(double), L.device, &dbx))) return rc;
  if ((rc = ctx->out_cost.ensure(P * sizeof(double)))) return rc;
  dbc = (L.device && best_cost) ? (void*)best_cost : ctx->out_cost.p;
  if ((rc = stage_out(ctx->st, status, T * sizeof(int32_t), L.device, &dst))) return rc;
  // every rollout of a pair starts from the pair's state (x0_div = R); end states go to a
  // [P*R][nx] scratch that the per-pair arg-min reads.
  if ((rc = ctx->scratch_o.ensure(T * nx * sizeof(double)))) return rc;
  CU(cudaEventRecord(ctx->ev0, s));
  cudaError_t e;
  RolloutArgs A;
  A.x0 = cview((const double*)dx0, (long long)P, nx, false, L.blocked);
  A.u = (L.soa && nu > 0) ? cview((const double*)du, (long long)T, nu, true) : cview((const double*)(du ? du : dx0), (long long)T, nu > 0 ? nu : 1, false);
  A.xout = view((double*)ctx->scratch_o.p, (long long)T, nx, false, L.blocked);  // rows in the caller's component order
  A.status = (int32_t*)dst;
  A.n_samples = (long long)T;
  A.x0_div = (long long)R;
  A.traj = BatchView{nullptr, 0, 0, 0};
  A.dt = dt;
  A.n_steps = n_steps;
  A.status_or = 0;
  A.active = nullptr;
  A.u_node_stride = 0;
  if ((rc = launch_rollout(c, ctx, A, nullptr, s))) return rc;
  e = rkb_steer_reduce(nx, (long long)P, (long long)R, (const double*)ctx->scratch_o.p, (const double*)dgoal, (int32_t*)didx,
                       (double*)dbx, (double*)dbc, s);
  if (e != cudaSuccess) return cuda_fail(e, "steer reduce");
  c->launches += 1;
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  if ((rc = unstage_out(didx, best_idx, P * sizeof(int32_t), L.device, s))) return rc;
  if ((rc = unstage_out(dbx, best_x, P * nx * sizeof(double), L.device, s))) return rc;
  if (!L.device && best_cost) CU(cudaMemcpyAsync(best_cost, dbc, P * sizeof(double), cudaMemcpyDeviceToHost, s));
  if ((rc = unstage_out(dst, status, T * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

}  // extern "C"

namespace {
// The steering kernel with the collision test of `pairs` compiled in (c->mu held).  sync: compile now if need be; else
// from the first call of >= 4096 tuples on in the background, NULL until it is there (the caller runs interval by interval).
const SourceKernels* checked_steer_kernel(rkb_chain* c, const rkb_proxy* const* pairs, int n_pairs, long long n_samples, bool sync, int* rc_out) {
  if (rc_out) *rc_out = RKB_OK;
  if (!c->serial_ok || !c->sk || c->n_free || n_pairs < 1) { if (rc_out) *rc_out = RKB_ERR_UNSUPPORTED; return nullptr; }
  std::vector<unsigned long long> key;
  bool may_auto = c->auto_specialize;
  for (int p = 0; p < n_pairs; ++p) {
    key.push_back(pairs[p]->serial);
    std::lock_guard<std::mutex> lock(pairs[p]->mu);
    may_auto = may_auto && pairs[p]->auto_specialize;
  }
  rkb_chain::CheckedSteer* entry = nullptr;
  for (auto& e : c->checked)
    if (e.key == key) entry = &e;
  if (entry && (entry->K || (entry->done && !sync))) return entry->K;
  if (!entry) {
    if (c->checked.size() >= 64) c->checked.erase(c->checked.begin());  // (pairs come and go: the table stays small; the kernels themselves are cached by rkb_jit.cu)
    c->checked.push_back(rkb_chain::CheckedSteer{key, nullptr, false, 0});
    entry = &c->checked.back();
  }
  entry->seen += n_samples;  // many small calls count like one large one
  if (!sync && (!may_auto || entry->seen < 4096)) return nullptr;
  int coords[RKB_SERIAL_MAX_DOF];
  for (int s = 0; s < c->n; ++s) coords[s] = c->sp.st[s].coord;
  std::vector<const ProxProgram*> progs;
  for (int p = 0; p < n_pairs; ++p) progs.push_back(&pairs[p]->prog);
  std::string expr;
  const std::string src = rkb_steer_checked_source(c->n, c->serial_fl, c->serial_shape, coords, c->gp, progs.data(), n_pairs, &expr);
  if (src.empty()) { entry->done = true; if (rc_out) *rc_out = RKB_ERR_UNSUPPORTED; return nullptr; }
  const char* names[1] = {expr.c_str()};
  const SourceKernels* K = nullptr;
  const int rc = sync ? rkb_jit_source_get("steerchk", src, names, 1, true, &K) : rkb_jit_source_poll("steerchk", src, names, 1, true, &K);
  if (rc != RKB_OK) { entry->done = true; if (rc_out) *rc_out = rc; return nullptr; }  // no NVRTC here, or it failed: stay as we are
  if (K) { entry->K = K; entry->done = true; }
  return K;
}

int steer_feedback_impl(rkb_chain* c, int device, size_t N, const double* x0, const double* x_goal, const double* u_bias,
                        const double* gain, double* u_prev, const rkb_steer_opts* o, const rkb_proxy* const* pairs, int n_pairs,
                        double* x_out, int32_t* n_done, int32_t* collided, double* x_traj, int32_t* status, unsigned flags,
                        void* stream) {
  if (!c || !o || o->reserved != 0) return RKB_ERR_INVALID;
  if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !collided))) return RKB_ERR_INVALID;
  for (int p = 0; p < n_pairs; ++p) {
    if (!pairs[p] || pairs[p]->n_frames != c->desc.n_frames) return RKB_ERR_INVALID;
    if (!c->generic_ok) return RKB_ERR_UNSUPPORTED;
  }
  if (o->dt == 0.0 || !std::isfinite(o->dt) || o->substeps < 1 || o->max_intervals < 0) return RKB_ERR_INTEGRATION;
  if (!(o->time_step > 0.0) || !std::isfinite(o->time_step) || !std::isfinite(o->goal_proximity)) return RKB_ERR_INVALID;
  if ((o->u_lower == nullptr) != (o->u_upper == nullptr) || (o->du_lower == nullptr) != (o->du_upper == nullptr)) return RKB_ERR_INVALID;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  if (L.soa) return RKB_ERR_UNSUPPORTED;
  if (N == 0) return RKB_OK;
  const int nx = c->nx, nu = c->nu;
  if (!x0 || !x_goal || !x_out || !n_done || (nu > 0 && (!u_bias || !gain || !u_prev))) return RKB_ERR_INVALID;
  for (int r = 0; r < nu; ++r) {
    if (o->u_lower && !(o->u_lower[r] < o->u_upper[r])) return RKB_ERR_INVALID;
    if (o->du_lower && !(o->du_lower[r] < o->du_upper[r])) return RKB_ERR_INVALID;
  }
  const int J = o->max_intervals;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  maybe_auto_specialize(c, (long long)N);
  cudaStream_t s = (cudaStream_t)stream;
  const void *dx0 = nullptr, *dgoal = nullptr, *dbias = nullptr, *dgain = nullptr, *dup_in = nullptr;
  void *dxo = nullptr, *dnd = nullptr, *dtraj = nullptr, *dst = nullptr;
  const size_t bx = N * nx * sizeof(double), bu = N * (size_t)nu * sizeof(double);
  if ((rc = stage_in(ctx->in_x, x0, bx, L.device, s, &dx0))) return rc;
  if ((rc = stage_in(ctx->in_goal, x_goal, bx, L.device, s, &dgoal))) return rc;
  if (nu > 0) {
    if ((rc = stage_in(ctx->in_bias, u_bias, bu, L.device, s, &dbias))) return rc;
    if ((rc = stage_in(ctx->in_gain, gain, bu * nx, L.device, s, &dgain))) return rc;
    if ((rc = stage_in(ctx->io_uprev, u_prev, bu, L.device, s, &dup_in))) return rc;
  }
  if ((rc = stage_out(ctx->out_a, x_out, bx, L.device, &dxo))) return rc;
  if ((rc = stage_out(ctx->out_ndone, n_done, N * sizeof(int32_t), L.device, &dnd))) return rc;
  if ((rc = stage_out(ctx->out_b, x_traj, bx * (size_t)J, L.device, &dtraj))) return rc;  // J == 0: x_traj holds nothing, nothing is staged
  if ((rc = ctx->st.ensure(N * sizeof(int32_t)))) return rc;  // status is accumulated on the device even if not wanted
  dst = (L.device && status) ? (void*)status : ctx->st.p;
  if ((rc = ctx->act.ensure(N * sizeof(int32_t)))) return rc;
  CU(cudaMemsetAsync(dst, 0, N * sizeof(int32_t), s));
  const SourceKernels* CK = (n_pairs > 0 && c->fused_steer) ? checked_steer_kernel(c, pairs, n_pairs, (long long)N, false, nullptr) : nullptr;
  const bool fused = (n_pairs == 0 || CK) && c->serial_ok && c->sk && c->fused_steer;
  void *dcol = nullptr, *dxn = nullptr, *dun = nullptr, *ddist = nullptr;
  if (CK) {
    if ((rc = stage_out(ctx->out_idx, collided, N * sizeof(int32_t), L.device, &dcol))) return rc;
  } else if (n_pairs > 0) {
    // the interval is integrated into x_next, tested, and only then accepted (MEAQR_topology.hpp:550-559)
    if ((rc = stage_out(ctx->out_idx, collided, N * sizeof(int32_t), L.device, &dcol))) return rc;
    if ((rc = ctx->scratch_x.ensure(bx))) return rc;
    if ((rc = ctx->scratch_u.ensure(bu > 0 ? bu : sizeof(double)))) return rc;
    if ((rc = ctx->scratch_o.ensure(N * sizeof(double) * (size_t)n_pairs))) return rc;
    dxn = ctx->scratch_x.p; dun = ctx->scratch_u.p; ddist = ctx->scratch_o.p;
    CU(cudaMemsetAsync(dcol, 0, N * sizeof(int32_t), s));
    CU(cudaMemsetAsync(dxn, 0, bx, s));
  }
  if (fused) {
    // serial chains: the whole loop in one launch
    SteerArgs F;
    F.x0 = (const double*)dx0; F.goal = (const double*)dgoal; F.u_bias = (const double*)dbias; F.gain = (const double*)dgain;
    F.u_prev = (double*)dup_in; F.xout = (double*)dxo; F.traj = (double*)dtraj; F.n_done = (int32_t*)dnd;
    F.status = (int32_t*)dst;
    F.n_samples = (long long)N; F.nu = nu; F.max_intervals = J; F.substeps = o->substeps; F.saturate_first = o->saturate_first ? 1 : 0;
    F.have_u_box = o->u_lower ? 1 : 0; F.have_du_box = o->du_lower ? 1 : 0; F.blocked = L.blocked ? 1 : 0; F.pad = 0;
    F.time_step = o->time_step; F.dt = o->dt; F.proximity = o->goal_proximity;
    F.collided = (int32_t*)dcol;
    for (int r = 0; r < RKB_MAX_COORDS; ++r) {
      F.u_lo[r] = (o->u_lower && r < nu) ? o->u_lower[r] : 0.0; F.u_hi[r] = (o->u_upper && r < nu) ? o->u_upper[r] : 0.0;
      F.du_lo[r] = (o->du_lower && r < nu) ? o->du_lower[r] : 0.0; F.du_hi[r] = (o->du_upper && r < nu) ? o->du_upper[r] : 0.0;
    }
    CU(cudaEventRecord(ctx->ev0, s));
    cudaError_t e = CK ? rkb_jit_launch_source_rollout(*CK, 0, c->sp, &F, F.n_samples, c->n, s)
                    : c->jit ? rkb_jit_launch(*c->jit, RKB_JIT_STEER, c->sp, &F, nullptr, F.n_samples, 0, s)
                    : use_split(c, F.n_samples) ? c->sk->steer_duo(c->sp, F, s) : c->sk->steer(c->sp, F, s);
    if (e != cudaSuccess) return cuda_fail(e, "steer kernel");
    c->launches += 1;
  } else {
  SteerLawArgs W;
  W.x0 = (const double*)dx0; W.x = (double*)dxo; W.goal = (const double*)dgoal;
  W.u_bias = (const double*)dbias; W.gain = (const double*)dgain;
  W.u_prev = (double*)dup_in;  // device memory: updated in place (the caller's buffer, or the staging copy)
  W.u_next = n_pairs > 0 ? (double*)dun : nullptr;
  W.n_done = (int32_t*)dnd; W.active = (int32_t*)ctx->act.p;
  W.n_samples = (long long)N; W.nx = nx; W.nu = nu; W.saturate_first = o->saturate_first ? 1 : 0;
  W.have_u_box = o->u_lower ? 1 : 0; W.have_du_box = o->du_lower ? 1 : 0;
  W.time_step = o->time_step; W.proximity = o->goal_proximity;
  for (int r = 0; r < RKB_MAX_COORDS; ++r) {
    W.u_lo[r] = (o->u_lower && r < nu) ? o->u_lower[r] : 0.0; W.u_hi[r] = (o->u_upper && r < nu) ? o->u_upper[r] : 0.0;
    W.du_lo[r] = (o->du_lower && r < nu) ? o->du_lower[r] : 0.0; W.du_hi[r] = (o->du_upper && r < nu) ? o->du_upper[r] : 0.0;
  }
  CU(cudaEventRecord(ctx->ev0, s));
  if (J == 0) {  // the reference loop would not run at all
    if (dxo != dx0) CU(cudaMemcpyAsync(dxo, dx0, bx, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemsetAsync(dnd, 0, N * sizeof(int32_t), s));
  }
  for (int k = 0; k < J; ++k) {
    W.interval = k;
    cudaError_t e = rkb_steer_law(W, s);
    if (e != cudaSuccess) return cuda_fail(e, "steer law");
    c->launches += 1;
    RolloutArgs A;
    A.x0 = cview((const double*)dxo, (long long)N, nx, false, L.blocked);
    A.u = cview((const double*)(nu > 0 ? (n_pairs > 0 ? dun : (void*)dup_in) : dxo), (long long)N, nu > 0 ? nu : 1, false);
    A.xout = view((double*)(n_pairs > 0 ? dxn : dxo), (long long)N, nx, false, L.blocked);
    A.traj = (dtraj && n_pairs == 0) ? BatchView{(double*)dtraj + (size_t)k * nx, (long long)nx * J, 1, L.blocked ? 1 : 0}
                                     : BatchView{nullptr, 0, 0, 0};
    A.status = (int32_t*)dst;
    A.n_samples = (long long)N;
    A.x0_div = 1;
    A.dt = o->dt;
    A.n_steps = o->substeps;
    A.status_or = 1;
    A.active = (const int32_t*)ctx->act.p;
    A.u_node_stride = 0;
    if ((rc = launch_rollout(c, ctx, A, nullptr, s))) return rc;
    if (n_pairs > 0) {
      for (int p = 0; p < n_pairs; ++p) {
        EvalArgs E;
        E.x = cview((const double*)dxn, (long long)N, nx, false, L.blocked);
        E.u = cview((const double*)dxn, (long long)N, 1, false);
        E.out = view((double*)ddist + (size_t)p * N, (long long)N, 1, false);
        E.out2 = view((double*)nullptr, (long long)N, 6, false);
        E.status = nullptr;
        E.n_samples = (long long)N;
        const cudaError_t pe = launch_proximity(c, ctx, pairs[p], E, s);
        if (pe != cudaSuccess) return cuda_fail(pe, "proximity kernel");
        c->launches += 1;
      }
      SteerCommitArgs K;
      K.x = (double*)dxo; K.x_next = (const double*)dxn; K.u_prev = (double*)dup_in; K.u_next = (const double*)dun;
      K.traj = (double*)dtraj; K.dist = (const double*)ddist; K.n_done = (int32_t*)dnd; K.active = (int32_t*)ctx->act.p;
      K.collided = (int32_t*)dcol; K.n_samples = (long long)N; K.nx = nx; K.nu = nu; K.interval = k; K.max_intervals = J;
      K.n_pairs = n_pairs; K.pad = 0;
      const cudaError_t ke = rkb_steer_commit(K, s);
      if (ke != cudaSuccess) return cuda_fail(ke, "steer commit");
      c->launches += 1;
    }
  }
  }
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  if ((rc = unstage_out(dxo, x_out, bx, L.device, s))) return rc;
  if ((rc = unstage_out(dnd, n_done, N * sizeof(int32_t), L.device, s))) return rc;
  if ((rc = unstage_out(dtraj, x_traj, bx * (size_t)J, L.device, s))) return rc;
  if (n_pairs > 0 && (rc = unstage_out(dcol, collided, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device && nu > 0) CU(cudaMemcpyAsync(u_prev, dup_in, bu, cudaMemcpyDeviceToHost, s));
  if (!L.device && status) CU(cudaMemcpyAsync(status, dst, N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}
}  // namespace

extern "C" {

int rkb_steer_feedback(rkb_chain* c, int device, size_t N, const double* x0, const double* x_goal, const double* u_bias,
                       const double* gain, double* u_prev, const rkb_steer_opts* o, double* x_out, int32_t* n_done,
                       double* x_traj, int32_t* status, unsigned flags, void* stream) {
  return steer_feedback_impl(c, device, N, x0, x_goal, u_bias, gain, u_prev, o, nullptr, 0, x_out, n_done, nullptr, x_traj, status,
                             flags, stream);
}

int rkb_steer_feedback_checked(rkb_chain* c, int device, size_t N, const double* x0, const double* x_goal, const double* u_bias,
                               const double* gain, double* u_prev, const rkb_steer_opts* o, const rkb_proxy* const* pairs,
                               int n_pairs, double* x_out, int32_t* n_done, int32_t* collided, double* x_traj, int32_t* status,
                               unsigned flags, void* stream) {
  if (n_pairs <= 0) return RKB_ERR_INVALID;
  return steer_feedback_impl(c, device, N, x0, x_goal, u_bias, gain, u_prev, o, pairs, n_pairs, x_out, n_done, collided, x_traj,
                             status, flags, stream);
}

int rkb_steer_checked_specialize(rkb_chain* c, int device, const rkb_proxy* const* pairs, int n_pairs) {
  if (!c || !pairs || n_pairs < 1) return RKB_ERR_INVALID;
  for (int p = 0; p < n_pairs; ++p)
    if (!pairs[p] || pairs[p]->n_frames != c->desc.n_frames) return RKB_ERR_INVALID;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  int rc = RKB_OK;
  const SourceKernels* K = checked_steer_kernel(c, pairs, n_pairs, 0, true, &rc);
  if (!K && rc == RKB_OK) rc = RKB_ERR_UNSUPPORTED;
  if (rc && rc != RKB_ERR_UNSUPPORTED) std::snprintf(g_cuda_err, sizeof g_cuda_err, "run-time compilation failed: %.200s", rkb_jit_log());
  return rc;
}

int rkb_steer_checked_source(rkb_chain* c, const rkb_proxy* const* pairs, int n_pairs, char* out, size_t size) {
  if (!c || !pairs || n_pairs < 1) return RKB_ERR_INVALID;
  for (int p = 0; p < n_pairs; ++p)
    if (!pairs[p] || pairs[p]->n_frames != c->desc.n_frames) return RKB_ERR_INVALID;
  if (!c->serial_ok || !c->sk || c->n_free) return RKB_ERR_UNSUPPORTED;
  int coords[RKB_SERIAL_MAX_DOF];
  for (int s = 0; s < c->n; ++s) coords[s] = c->sp.st[s].coord;
  std::vector<const ProxProgram*> progs;
  for (int p = 0; p < n_pairs; ++p) progs.push_back(&pairs[p]->prog);
  std::string expr;
  std::string src = rkb_steer_checked_source(c->n, c->serial_fl, c->serial_shape, coords, c->gp, progs.data(), n_pairs, &expr);
  if (src.empty()) return RKB_ERR_UNSUPPORTED;
  src += "// kernel: " + expr + "\n";
  if (out && size > src.size()) std::memcpy(out, src.c_str(), src.size() + 1);
  else if (out) return RKB_ERR_INVALID;
  return (int)src.size() + 1;
}

int rkb_steer_checked_is_specialized(rkb_chain* c, const rkb_proxy* const* pairs, int n_pairs) {
  if (!c || !pairs || n_pairs < 1) return 0;
  std::lock_guard<std::mutex> lock(c->mu);
  std::vector<unsigned long long> key;
  for (int p = 0; p < n_pairs; ++p) { if (!pairs[p]) return 0; key.push_back(pairs[p]->serial); }
  for (const auto& e : c->checked)
    if (e.key == key && e.K) return 1;
  return 0;
}

/* A = d xdot / d x (2n x 2n) and B = d xdot / d u (2n x n_inputs) about (x[i], u[i]) by central differences of
 * get_state_derivative — the linearisation the LQR steering of examples/misc/IHAQR_topology.hpp:240-258 asks its system
 * for (get_linear_blocks).  One perturbation kernel, the ordinary evaluation kernels on the 2 (2n + n_inputs) perturbed
 * copies of every sample, one differencing kernel; the batch goes in slices so that the scratch stays below ~150 MB. */
int rkb_linearize(rkb_chain* c, int device, size_t N, const double* x, const double* u, double eps, double* A, double* B,
                  int32_t* status, unsigned flags, void* stream) {
  if (!c) return RKB_ERR_INVALID;
  if (N == 0) return RKB_OK;
  const Layout L = parse_flags(flags);
  if (L.blocked && c->n_free) return RKB_ERR_UNSUPPORTED;
  if (L.soa || L.blocked) return RKB_ERR_UNSUPPORTED;
  const int nx = c->nx, nu = c->nu, D = nx + nu;
  if (!x || (nu > 0 && !u) || (!A && !B) || (B && nu == 0)) return RKB_ERR_INVALID;
  if (!(eps > 0.0) || !std::isfinite(eps)) eps = 1.0e-6;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard guard(device);
  if (!guard.ok) { std::snprintf(g_cuda_err, sizeof g_cuda_err, "cudaSetDevice(%d) failed", device); return RKB_ERR_CUDA; }
  DeviceCtx* ctx = nullptr;
  int rc = get_ctx(c, device, &ctx);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const void *dx = nullptr, *du = nullptr;
  void *dA = nullptr, *dB = nullptr, *dst = nullptr;
  if ((rc = stage_in(ctx->in_x, x, N * nx * sizeof(double), L.device, s, &dx))) return rc;
  if (nu > 0 && (rc = stage_in(ctx->in_u, u, N * nu * sizeof(double), L.device, s, &du))) return rc;
  if ((rc = stage_out(ctx->out_a, A, N * nx * nx * sizeof(double), L.device, &dA))) return rc;
  if ((rc = stage_out(ctx->out_b, B, N * nx * (size_t)nu * sizeof(double), L.device, &dB))) return rc;
  if ((rc = stage_out(ctx->st, status, N * sizeof(int32_t), L.device, &dst))) return rc;
  const size_t slice = 1u << 14, rows = (N < slice ? N : slice) * (size_t)D * 2;
  if ((rc = ctx->scratch_x.ensure(rows * nx * sizeof(double)))) return rc;
  if ((rc = ctx->scratch_u.ensure(rows * (nu > 0 ? nu : 1) * sizeof(double)))) return rc;
  if ((rc = ctx->scratch_o.ensure(rows * nx * sizeof(double)))) return rc;
  if ((rc = ctx->scratch_s.ensure(rows * sizeof(int32_t)))) return rc;
  if (dst) CU(cudaMemsetAsync(dst, 0, N * sizeof(int32_t), s));
  CU(cudaEventRecord(ctx->ev0, s));
  for (size_t lo = 0; lo < N; lo += slice) {
    const size_t m = N - lo < slice ? N - lo : slice, r = m * (size_t)D * 2;
    cudaError_t e = rkb_lin_perturb((long long)m, nx, nu, eps, (const double*)dx + lo * nx, nu > 0 ? (const double*)du + lo * nu : nullptr,
                                    (double*)ctx->scratch_x.p, (double*)ctx->scratch_u.p, s);
    if (e != cudaSuccess) return cuda_fail(e, "linearize: perturb");
    EvalArgs E;
    E.x = cview((const double*)ctx->scratch_x.p, (long long)r, nx, false);
    E.u = cview((const double*)(nu > 0 ? ctx->scratch_u.p : ctx->scratch_x.p), (long long)r, nu > 0 ? nu : 1, false);
    E.out = view((double*)ctx->scratch_o.p, (long long)r, nx, false);
    E.out2 = view((double*)nullptr, (long long)r, nx, false);
    E.status = (int32_t*)ctx->scratch_s.p;
    E.n_samples = (long long)r;
    if (c->serial_ok && c->sk) e = c->jit ? rkb_jit_launch(*c->jit, RKB_JIT_EVAL, c->sp, &E, nullptr, E.n_samples, 0, s) : c->sk->eval(c->sp, E, s);
    else if (c->generic_ok) e = rkb_generic_eval(ctx->d_prog, c->gp, E, s);
    else return RKB_ERR_UNSUPPORTED;
    if (e != cudaSuccess) return cuda_fail(e, "linearize: evaluation");
    e = rkb_lin_combine((long long)m, nx, nu, (const double*)ctx->scratch_x.p, (const double*)ctx->scratch_u.p, (const double*)ctx->scratch_o.p,
                        (const int32_t*)ctx->scratch_s.p, dA ? (double*)dA + lo * nx * nx : nullptr, dB ? (double*)dB + lo * nx * (size_t)nu : nullptr,
                        dst ? (int32_t*)dst + lo : nullptr, s);
    if (e != cudaSuccess) return cuda_fail(e, "linearize: differences");
    c->launches += 3;
  }
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->timed = true;
  c->last = ctx;
  if ((rc = unstage_out(dA, A, N * nx * nx * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dB, B, N * nx * (size_t)nu * sizeof(double), L.device, s))) return rc;
  if ((rc = unstage_out(dst, status, N * sizeof(int32_t), L.device, s))) return rc;
  if (!L.device) CU(cudaStreamSynchronize(s));
  return RKB_OK;
}

double rkb_last_kernel_ms(rkb_chain* c) {
  if (!c) return -1.0;
  std::lock_guard<std::mutex> lock(c->mu);
  if (!c->last || !c->last->timed) return -1.0;
  DeviceGuard guard(c->last->device);
  float ms = -1.0f;
  if (cudaEventSynchronize(c->last->ev1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
  if (cudaEventElapsedTime(&ms, c->last->ev0, c->last->ev1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
  return (double)ms;
}

uint64_t rkb_launch_count(const rkb_chain* c) { return c ? c->launches : 0; }

/* A ReaK caller holds its states in pageable std::vectors: cudaMemcpyAsync from those is a synchronous staged copy.
 * Pinning the buffer once (cudaHostRegister) makes every later RKB_MEM_HOST call on it a plain DMA. */
int rkb_host_pin(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return RKB_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return RKB_OK; }
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
  return RKB_OK;
}
int rkb_host_unpin(void* ptr) {
  if (!ptr) return RKB_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return RKB_OK; }
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
  return RKB_OK;
}

}  // extern "C"
