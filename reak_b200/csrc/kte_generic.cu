// kte_generic.cu — interpreter kernels: any chain the descriptor can express (2D and 3D frames,
// any element order, two-anchor springs/dampers, partial upstream-joint sets).  One thread per
// sample; the frames live in per-thread local memory, the element list is read from a
// GenericProgram in global memory (uniform loads).  This is the compatibility path — canonical
// serial chains run on the register-resident kernels of kte_serial.cuh instead.
//
// Reference semantics followed (paths relative to ReaK's source tree):
//   doMotion / clearForce / doForce      ctrl/mbd_kte/kte_map_chain.hpp:71-89 and the element .cpp files
//   mass matrix                           ctrl/mbd_kte/mass_matrix_calculator.cpp:80-98,100-287 with the
//                                         Jacobian columns of core/kinetostatics/motion_jacobians.hpp:139-147,
//                                         238-251 written in closed form from the world-frame kinematics
//   solve                                 core/lin_alg/mat_cholesky.hpp:63-84,160-179
//   RK4                                   core/integrators/fixed_step_integrators.hpp:256-293
#include <cuda_runtime.h>
#include <math.h>
#include <cstdlib>
#include "rkb_internal.h"

namespace {

#define GD __device__ __forceinline__
#ifdef RKB_HOST_TEST
#define GNI static                    // (tests/host_build: an ordinary function)
#else
#define GNI __device__ __noinline__   // a real call on the device: see rate()
#endif
#define MAXC RKB_MAX_COORDS

#include "kte_math.cuh"

struct Fr3 { V3 p; Q4 q; V3 v, w, a, al, F, T; };
struct Fr2 { V2 p; R2 R; V2 v; double w; V2 a; double al; V2 F; double T; };

template <int DIM> struct FrameOf;
template <> struct FrameOf<3> { typedef Fr3 type; };
template <> struct FrameOf<2> { typedef Fr2 type; };

template <int DIM, int MAXF>
struct Work {
  typename FrameOf<DIM>::type fr[MAXF];
  // coordinate frame of the chain's free_joint_3D (kte_nl_system::dofs_3D[0]): Position, Quat (normalised), Velocity,
  // AngVelocity from the state, zero accelerations, Force / Torque collected by doForce.  Unused by other chains.
  typename FrameOf<DIM>::type fc;
  double fraw[2];  // free_joint_2D: (cos, sin) of the state as given (its derivative uses the raw values)
  double q[MAXC], qd[MAXC], f[RKB_GEN_MAX_ACC], u[MAXC];
};
#define MAXA RKB_GEN_MAX_ACC
#define MAXX RKB_GEN_MAX_STATE

GD void set_base(const GenericProgram* G, Fr3& B) {
  const double* b = G->base;
  B.p = ldv(b); B.q.w = b[3]; B.q.x = b[4]; B.q.y = b[5]; B.q.z = b[6];
  B.v = ldv(b + 7); B.w = ldv(b + 10); B.a = ldv(b + 13); B.al = ldv(b + 16);
}
GD void set_base(const GenericProgram* G, Fr2& B) {
  const double* b = G->base;
  B.p = v2(b[0], b[1]); B.R.c = b[3]; B.R.s = b[4];
  B.v = v2(b[7], b[8]); B.w = b[10]; B.a = v2(b[13], b[14]); B.al = b[16];
}

// ---- doMotion ---------------------------------------------------------------------------------
template <int MAXF>
GD void motion(const GenericProgram* G, Work<3, MAXF>& W) {
  set_base(G, W.fr[G->base_frame]);
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind == RKB_RIGID_LINK_GEN) {  // rigid_link.cpp:34-40
      W.q[E.aux] = W.q[E.coord] + E.p[0];
      W.qd[E.aux] = W.qd[E.coord];
    } else if (E.kind == RKB_REVOLUTE_3D) {  // revolute_joint.cpp:121-152 (q_ddot = 0)
      const Fr3 B = W.fr[E.fa];
      Fr3& N = W.fr[E.fb];
      const V3 ax = ldv(E.p), an = unit_axis(ax);
      double sh, ch;
      sincos(0.5 * W.q[E.coord], &sh, &ch);
      Q4 tq; tq.w = ch; tq.x = an.x * sh; tq.y = an.y * sh; tq.z = an.z * sh;
      const M3 R = qrot(tq);
      const V3 wt = tmul(R, B.w), qda = W.qd[E.coord] * ax;
      N.p = B.p; N.v = B.v; N.a = B.a;
      N.q = qmul(B.q, tq);
      N.w = wt + qda;
      N.al = tmul(R, B.al) + cross(wt, qda);
    } else if (E.kind == RKB_PRISMATIC_3D) {  // prismatic_joint.cpp:129-161
      const Fr3 B = W.fr[E.fa];
      Fr3& N = W.fr[E.fb];
      const V3 ax = ldv(E.p);
      const M3 R = qrot(B.q);
      const V3 tp = W.q[E.coord] * ax, tv = W.qd[E.coord] * ax;
      N.p = B.p + mul(R, tp);
      N.v = B.v + mul(R, cross(B.w, tp) + tv);
      N.a = B.a + mul(R, cross(B.w, cross(B.w, tp)) + 2.0 * cross(B.w, tv) + cross(B.al, tp));
      N.q = B.q; N.w = B.w; N.al = B.al;
    } else if (E.kind == RKB_FREE_3D) {  // free_joints.cpp:123-131: *mEnd = (*mBase) * (*mCoord), frame_3D.hpp:219-234
      const Fr3 B = W.fr[E.fa];
      Fr3& N = W.fr[E.fb];
      const Fr3& C = W.fc;
      const M3 R = qrot(B.q), R2 = qrot(C.q);
      N.p = B.p + mul(R, C.p);
      N.v = B.v + mul(R, cross(B.w, C.p) + C.v);
      N.a = B.a + mul(R, cross(B.w, cross(B.w, C.p)) + 2.0 * cross(B.w, C.v) + cross(B.al, C.p) + C.a);
      N.q = qmul(B.q, C.q);
      const V3 wt = tmul(R2, B.w);
      N.al = tmul(R2, B.al) + cross(wt, C.w) + C.al;
      N.w = wt + C.w;
    } else if (E.kind == RKB_RIGID_LINK_3D) {  // rigid_link.cpp:156 -> frame_3D.hpp:236-251
      const Fr3 B = W.fr[E.fa];
      Fr3& N = W.fr[E.fb];
      const V3 po = ldv(E.p);
      Q4 qo; qo.w = E.p[3]; qo.x = E.p[4]; qo.y = E.p[5]; qo.z = E.p[6];
      const M3 R = qrot(B.q), Ro = qrot(qo);
      N.p = B.p + mul(R, po);
      N.v = B.v + mul(R, cross(B.w, po));
      N.a = B.a + mul(R, cross(B.w, cross(B.w, po)) + cross(B.al, po));
      N.q = qmul(B.q, qo);
      N.al = tmul(Ro, B.al);
      N.w = tmul(Ro, B.w);
    }
  }
}
template <int MAXF>
GD void motion(const GenericProgram* G, Work<2, MAXF>& W) {
  set_base(G, W.fr[G->base_frame]);
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind == RKB_RIGID_LINK_GEN) {  // rigid_link.cpp:34-40
      W.q[E.aux] = W.q[E.coord] + E.p[0];
      W.qd[E.aux] = W.qd[E.coord];
    } else if (E.kind == RKB_FREE_2D) {  // free_joints.cpp:33-41: *mEnd = (*mBase) * (*mCoord), frame_2D.hpp:288-300
      const Fr2 B = W.fr[E.fa];
      Fr2& N = W.fr[E.fb];
      const Fr2& C = W.fc;
      N.p = B.p + rmul(B.R, C.p);
      N.v = B.v + rmul(B.R, crs(B.w, C.p) + C.v);
      N.a = B.a + rmul(B.R, (-B.w * B.w) * C.p + crs(2.0 * B.w, C.v) + crs(B.al, C.p) + C.a);
      N.R = rr(B.R, C.R);
      N.w = B.w + C.w;
      N.al = B.al + C.al;
    } else if (E.kind == RKB_REVOLUTE_2D) {  // revolute_joint.cpp:32-58
      const Fr2 B = W.fr[E.fa];
      Fr2& N = W.fr[E.fb];
      R2 rq;
      sincos(W.q[E.coord], &rq.s, &rq.c);
      N.p = B.p; N.v = B.v; N.a = B.a;
      N.R = rr(B.R, rq);
      N.w = B.w + W.qd[E.coord];
      N.al = B.al;
    } else if (E.kind == RKB_PRISMATIC_2D) {  // prismatic_joint.cpp:33-67
      const Fr2 B = W.fr[E.fa];
      Fr2& N = W.fr[E.fb];
      const V2 ax = v2(E.p[0], E.p[1]);
      const V2 tp = W.q[E.coord] * ax, tv = W.qd[E.coord] * ax;
      N.p = B.p + rmul(B.R, tp);
      N.v = B.v + rmul(B.R, crs(B.w, tp) + tv);
      N.a = B.a + rmul(B.R, (-B.w * B.w) * tp + crs(2.0 * B.w, tv) + crs(B.al, tp));
      N.R = B.R; N.w = B.w; N.al = B.al;
    } else if (E.kind == RKB_RIGID_LINK_2D) {  // rigid_link.cpp:87-99
      const Fr2 B = W.fr[E.fa];
      Fr2& N = W.fr[E.fb];
      const V2 po = v2(E.p[0], E.p[1]);
      R2 ro; ro.c = E.p[3]; ro.s = E.p[4];
      N.p = B.p + rmul(B.R, po);
      N.v = B.v + rmul(B.R, crs(B.w, po));
      N.a = B.a + rmul(B.R, (-B.w * B.w) * po + crs(B.al, po));
      N.R = rr(B.R, ro);
      N.w = B.w; N.al = B.al;
    }
  }
}

// ---- clearForce + doForce ------------------------------------------------------------------------
// elements that act on generalized coordinates alone (the same in 2D and 3D chains); f is indexed like q: system
// coordinates, then the auxiliary gen_coords.  Returns false for any other kind.
GD bool force_gen(const GenericElement& E, const double* q, const double* qd, double* f) {
  switch (E.kind) {
    case RKB_RIGID_LINK_GEN:  // rigid_link.cpp:53-57
      f[E.coord] += f[E.aux];
      return true;
    case RKB_SPRING_GEN: {  // spring.cpp:50-84
      const double rest = E.p[0], k = E.p[1], sat = E.p[2];
      const bool first = q[E.coord] > q[E.aux];
      const double fm = ((first ? q[E.coord] - q[E.aux] : q[E.aux] - q[E.coord]) - rest) * k;
      double t = fm;
      if (sat > 0.0 && fabs(fm) > sat) t = fm > 0.0 ? sat : -sat;
      if (first) { f[E.coord] -= t; f[E.aux] += t; } else { f[E.coord] += t; f[E.aux] -= t; }
      return true;
    }
    case RKB_DAMPER_GEN: {  // damper.cpp:48-57
      const double fm = (qd[E.coord] - qd[E.aux]) * E.p[0];
      f[E.coord] -= fm; f[E.aux] += fm;
      return true;
    }
    default: return false;
  }
}

template <int MAXF>
GD void force(const GenericProgram* G, Work<3, MAXF>& W) {
  for (int i = 0; i < G->n_frames; ++i) { W.fr[i].F = v3(0, 0, 0); W.fr[i].T = v3(0, 0, 0); }
  for (int i = 0; i < G->n_coords + G->n_aux; ++i) W.f[i] = 0.0;
  W.fc.F = v3(0, 0, 0); W.fc.T = v3(0, 0, 0);
  for (int e = G->n_elements - 1; e >= 0; --e) {
    const GenericElement& E = G->el[e];
    switch (E.kind) {
      case RKB_REVOLUTE_3D: {  // revolute_joint.cpp:172-184
        const Fr3& N = W.fr[E.fb];
        Fr3& B = W.fr[E.fa];
        const V3 ax = ldv(E.p);
        const M3 R = aa_rot(W.q[E.coord], unit_axis(ax));
        const double ta = dot(N.T, ax);
        B.F = B.F + mul(R, N.F);
        W.f[E.coord] += ta;
        B.T = B.T + mul(R, N.T - ta * ax);
        break;
      }
      case RKB_PRISMATIC_3D: {  // prismatic_joint.cpp:181-193
        const Fr3& N = W.fr[E.fb];
        Fr3& B = W.fr[E.fa];
        const V3 ax = ldv(E.p);
        const double tf = dot(N.F, ax);
        W.f[E.coord] += tf;
        B.F = B.F + (N.F - tf * ax);
        B.T = B.T + (N.T + cross(W.q[E.coord] * ax, N.F));
        break;
      }
      case RKB_FREE_3D: {  // free_joints.cpp:164-172: the end frame's wrench lands on the coordinate frame
        const Fr3& N = W.fr[E.fb];
        W.fc.F = W.fc.F + N.F;
        W.fc.T = W.fc.T + N.T;
        break;
      }
      case RKB_RIGID_LINK_3D: {  // rigid_link.cpp:170-177
        const Fr3& N = W.fr[E.fb];
        Fr3& B = W.fr[E.fa];
        Q4 qo; qo.w = E.p[3]; qo.x = E.p[4]; qo.y = E.p[5]; qo.z = E.p[6];
        const M3 Ro = qrot(qo);
        const V3 tf = mul(Ro, N.F);
        B.F = B.F + tf;
        B.T = B.T + (mul(Ro, N.T) + cross(ldv(E.p), tf));
        break;
      }
      case RKB_INERTIA_3D: {  // inertia.cpp:111-121
        Fr3& Gf = W.fr[E.fa];
        const double* I = &E.p[1];
        const V3 Ial = v3(I[0] * Gf.al.x + I[1] * Gf.al.y + I[2] * Gf.al.z, I[1] * Gf.al.x + I[3] * Gf.al.y + I[4] * Gf.al.z,
                          I[2] * Gf.al.x + I[4] * Gf.al.y + I[5] * Gf.al.z);
        const V3 Iw = v3(I[0] * Gf.w.x + I[1] * Gf.w.y + I[2] * Gf.w.z, I[1] * Gf.w.x + I[3] * Gf.w.y + I[4] * Gf.w.z,
                         I[2] * Gf.w.x + I[4] * Gf.w.y + I[5] * Gf.w.z);
        Gf.F = Gf.F - E.p[0] * tmul(qrot(Gf.q), Gf.a);
        Gf.T = Gf.T - (Ial + cross(Gf.w, Iw));
        break;
      }
      case RKB_ACTUATOR_GEN: {  // driving_actuator.cpp:31-38 + applyReactionForce
        const GenericElement& J = G->el[E.fb];
        const double drive = W.u[E.aux];
        W.f[E.coord] += drive;
        Fr3& B = W.fr[J.fa];
        if (J.kind == RKB_REVOLUTE_3D) B.T = B.T - drive * ldv(J.p);        // revolute_joint.cpp:210-213
        else if (J.kind == RKB_PRISMATIC_3D) B.F = B.F - drive * ldv(J.p);  // prismatic_joint.cpp:219-222
        break;
      }
      case RKB_TORSION_SPRING_3D: {  // torsion_spring.cpp:106-129, axis_angle(quaternion) rotations_3D.hpp:1985-2010
        Fr3& A1 = W.fr[E.fa];
        Fr3& A2 = W.fr[E.fb];
        Q4 d = qmul(qconj(A1.q), A2.q);
        const double nq = sqrt(d.w * d.w + d.x * d.x + d.y * d.y + d.z * d.z);
        d.w /= nq; d.x /= nq; d.y /= nq; d.z /= nq;
        const double tmp = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
        V3 ax = v3(1.0, 0.0, 0.0);
        double angle = 0.0;
        if (tmp > 0.0000001) {
          ax = v3(d.x / tmp, d.y / tmp, d.z / tmp);
          // 2 acos(|w|) is ill-conditioned near w = 1; 2 atan2(|v|, |w|) is the same angle, well-conditioned
          angle = 2.0 * atan2(tmp, fabs(d.w));
          if (d.w < 0.0) ax = v3(-ax.x, -ax.y, -ax.z);
        }
        const double mag = E.p[0] * angle, sat = E.p[1];
        V3 t = mag * ax;
        if (sat > 0.0 && fabs(mag) > sat) t = (mag > 0.0 ? sat : -sat) * ax;
        A1.T = A1.T + t;
        A2.T = A2.T - t;
        break;
      }
      case RKB_TORSION_DAMPER_3D: {  // torsion_damper.cpp:93-104
        Fr3& A1 = W.fr[E.fa];
        Fr3& A2 = W.fr[E.fb];
        const M3 R1 = qrot(A1.q), R2_ = qrot(A2.q);
        const V3 diff = E.p[0] * (mul(R1, A1.w) - mul(R2_, A2.w));
        A1.T = A1.T - tmul(R1, diff);
        A2.T = A2.T + tmul(R2_, diff);
        break;
      }
      case RKB_SPRING_3D: {  // spring.cpp:178-207
        Fr3& A1 = W.fr[E.fa];
        Fr3& A2 = W.fr[E.fb];
        V3 diff = A1.p - A2.p;
        const double mag = sqrt(dot(diff, diff));
        if (mag > 1E-7) {
          const double fm = (mag - E.p[0]) * E.p[1], sat = E.p[2];
          double sc = fm / mag;
          if (sat > 0.0 && fabs(fm) > sat) sc = (fm > 0.0 ? sat : -sat) / mag;
          diff = sc * diff;
          A1.F = A1.F - tmul(qrot(A1.q), diff);
          A2.F = A2.F + tmul(qrot(A2.q), diff);
        }
        break;
      }
      case RKB_DAMPER_3D: {  // damper.cpp:136-149
        Fr3& A1 = W.fr[E.fa];
        Fr3& A2 = W.fr[E.fb];
        V3 diff = A1.p - A2.p;
        const double sq = dot(diff, diff);
        if (sq > 1E-7) {
          diff = (dot(A1.v - A2.v, diff) * E.p[0] / sq) * diff;
          A1.F = A1.F - tmul(qrot(A1.q), diff);
          A2.F = A2.F + tmul(qrot(A2.q), diff);
        }
        break;
      }
      default: force_gen(E, W.q, W.qd, W.f); break;  // (inertia_gen: f -= q_ddot * m with q_ddot = 0)
    }
  }
}
template <int MAXF>
GD void force(const GenericProgram* G, Work<2, MAXF>& W) {
  for (int i = 0; i < G->n_frames; ++i) { W.fr[i].F = v2(0, 0); W.fr[i].T = 0.0; }
  for (int i = 0; i < G->n_coords + G->n_aux; ++i) W.f[i] = 0.0;
  W.fc.F = v2(0, 0); W.fc.T = 0.0;
  for (int e = G->n_elements - 1; e >= 0; --e) {
    const GenericElement& E = G->el[e];
    switch (E.kind) {
      case RKB_FREE_2D: {  // free_joints.cpp:75-85
        const Fr2& N = W.fr[E.fb];
        W.fc.F = W.fc.F + N.F;
        W.fc.T += N.T;
        break;
      }
      case RKB_REVOLUTE_2D: {  // revolute_joint.cpp:78-90 (torque is not passed to the base)
        const Fr2& N = W.fr[E.fb];
        Fr2& B = W.fr[E.fa];
        R2 rq;
        sincos(W.q[E.coord], &rq.s, &rq.c);
        B.F = B.F + rmul(rq, N.F);
        W.f[E.coord] += N.T;
        break;
      }
      case RKB_PRISMATIC_2D: {  // prismatic_joint.cpp:83-95
        const Fr2& N = W.fr[E.fb];
        Fr2& B = W.fr[E.fa];
        const V2 ax = v2(E.p[0], E.p[1]);
        const double tf = dot(N.F, ax);
        W.f[E.coord] += tf;
        B.F = B.F + (N.F - tf * ax);
        B.T += N.T + cross(W.q[E.coord] * ax, N.F);
        break;
      }
      case RKB_RIGID_LINK_2D: {  // rigid_link.cpp:117-125
        const Fr2& N = W.fr[E.fb];
        Fr2& B = W.fr[E.fa];
        R2 ro; ro.c = E.p[3]; ro.s = E.p[4];
        const V2 tf = rmul(ro, N.F);
        B.F = B.F + tf;
        B.T += N.T + cross(v2(E.p[0], E.p[1]), tf);
        break;
      }
      case RKB_INERTIA_2D: {  // inertia.cpp:77-86
        Fr2& Gf = W.fr[E.fa];
        Gf.F = Gf.F - E.p[0] * rtmul(Gf.R, Gf.a);
        Gf.T -= E.p[1] * Gf.al;
        break;
      }
      case RKB_ACTUATOR_GEN: {
        const GenericElement& J = G->el[E.fb];
        const double drive = W.u[E.aux];
        W.f[E.coord] += drive;
        Fr2& B = W.fr[J.fa];
        if (J.kind == RKB_REVOLUTE_2D) B.T -= drive;                                 // revolute_joint.cpp:113-116
        else if (J.kind == RKB_PRISMATIC_2D) B.F = B.F - drive * v2(J.p[0], J.p[1]); // prismatic_joint.cpp:120-123
        break;
      }
      case RKB_TORSION_SPRING_2D: {  // torsion_spring.cpp:50-71
        Fr2& A1 = W.fr[E.fa];
        Fr2& A2 = W.fr[E.fb];
        R2 inv; inv.c = A1.R.c; inv.s = -A1.R.s;
        const R2 rel = rr(inv, A2.R);
        const double ad = atan2(rel.s, rel.c) * E.p[0], sat = E.p[1];
        double t = ad;
        if (sat > 0.0 && fabs(ad) > sat) t = ad > 0.0 ? sat : -sat;
        A1.T += t; A2.T -= t;
        break;
      }
      case RKB_TORSION_DAMPER_2D: {  // torsion_damper.cpp:49-58
        Fr2& A1 = W.fr[E.fa];
        Fr2& A2 = W.fr[E.fb];
        const double tm = (A1.w - A2.w) * E.p[0];
        A1.T -= tm; A2.T += tm;
        break;
      }
      case RKB_SPRING_2D: {  // spring.cpp:116-143
        Fr2& A1 = W.fr[E.fa];
        Fr2& A2 = W.fr[E.fb];
        V2 diff = A1.p - A2.p;
        const double mag = sqrt(dot(diff, diff));
        if (mag > 1E-7) {
          const double fm = (mag - E.p[0]) * E.p[1], sat = E.p[2];
          double sc = fm / mag;
          if (sat > 0.0 && fabs(fm) > sat) sc = (fm > 0.0 ? sat : -sat) / mag;
          diff = sc * diff;
          A1.F = A1.F - rtmul(A1.R, diff);
          A2.F = A2.F + rtmul(A2.R, diff);
        }
        break;
      }
      case RKB_DAMPER_2D: {  // damper.cpp:88-102
        Fr2& A1 = W.fr[E.fa];
        Fr2& A2 = W.fr[E.fb];
        V2 diff = A1.p - A2.p;
        const double sq = dot(diff, diff);
        if (sq > 1E-7) {
          diff = (dot(A1.v - A2.v, diff) * E.p[0] / sq) * diff;
          A1.F = A1.F - rtmul(A1.R, diff);
          A2.F = A2.F + rtmul(A2.R, diff);
        }
        break;
      }
      default: force_gen(E, W.q, W.qd, W.f); break;
    }
  }
}

// ---- mass matrix ---------------------------------------------------------------------------------
// Column of the twist-shaping matrix (and of its time derivative) of coordinate `c` seen from the
// inertia frame F: jacobian_gen_3D::get_jac_relative_to (motion_jacobians.hpp:238-251) with
// f2 = (~E) * F (frame_3D.hpp:183-188, 219-234, 376-388) written out in world-frame quantities:
//   a_g = R_E a,  dp = p_F - p_E,  dv = v_F - v_E,  adot_g = (R_E w_E) x a_g
//   revolute : Tv = R_F^T (a_g x dp), Tw = R_F^T a_g,
//              Tvd = R_F^T (adot_g x dp + a_g x dv) - w_F x Tv,  Twd = R_F^T adot_g - w_F x Tw
//   prismatic: Tv = R_F^T a_g, Tw = 0,  Tvd = R_F^T adot_g - w_F x Tv, Twd = 0
template <int MAXF>
GD void jac_col(const GenericProgram* G, const Work<3, MAXF>& W, int c, const Fr3& F, const M3& RF, double* T, double* Td, bool want_dot) {
  const GenericElement& J = G->el[G->jelem[c]];
  const Fr3& E = W.fr[J.fb];
  const M3 RE = qrot(E.q);
  const V3 ag = mul(RE, ldv(J.p));
  const V3 adg = cross(mul(RE, E.w), ag);
  V3 Tv, Tw, Tvd = v3(0, 0, 0), Twd = v3(0, 0, 0);
  if (J.kind == RKB_REVOLUTE_3D) {
    const V3 dp = F.p - E.p;
    Tv = tmul(RF, cross(ag, dp));
    Tw = tmul(RF, ag);
    if (want_dot) {
      const V3 dv = F.v - E.v;
      Tvd = tmul(RF, cross(adg, dp) + cross(ag, dv)) - cross(F.w, Tv);
      Twd = tmul(RF, adg) - cross(F.w, Tw);
    }
  } else {
    Tv = tmul(RF, ag);
    Tw = v3(0, 0, 0);
    if (want_dot) Tvd = tmul(RF, adg) - cross(F.w, Tv);
  }
  T[0] = Tv.x; T[1] = Tv.y; T[2] = Tv.z; T[3] = Tw.x; T[4] = Tw.y; T[5] = Tw.z;
  Td[0] = Tvd.x; Td[1] = Tvd.y; Td[2] = Tvd.z; Td[3] = Twd.x; Td[4] = Twd.y; Td[5] = Twd.z;
}

// The six columns a free_joint_3D contributes for inertia frame F: jacobian_3D_3D::get_jac_relative_to
// (motion_jacobians.hpp:1077-1140) on the joint's two identity blocks, with f2 = (~E) * F (E the joint's end frame)
// in world quantities: g_k = R_E e_k,  dp = p_F - p_E,  dv = v_F - v_E,  W_E = R_E w_E,  w_rel = w_F - R_F^T W_E
//   velocity input k        : Tv = R_F^T g_k,          Tw = 0,          Tvd = -w_rel x Tv,  Twd = 0
//   angular-velocity input k: Tv = R_F^T (g_k x dp),   Tw = R_F^T g_k,  Tvd = R_F^T (g_k x (dv - W_E x dp)) - w_rel x Tv,  Twd = -w_rel x Tw
// T[l] / Td[l]: column l (0-2 velocity, 3-5 angular velocity) as (v3, w3), the order write_to_matrices uses (:1142-1203).
template <int MAXF>
GD void jac_free_block(const GenericProgram* G, const Work<3, MAXF>& W, int fj, const Fr3& F, const M3& RF, double (*T)[6], double (*Td)[6], bool want_dot) {
  const Fr3& E = W.fr[G->el[G->free_elem[fj]].fb];
  const M3 RE = qrot(E.q);
  const V3 dp = F.p - E.p, WE = mul(RE, E.w);
  const V3 wrel = F.w - tmul(RF, WE);
  const V3 dvr = (F.v - E.v) - cross(WE, dp);
  for (int k = 0; k < 3; ++k) {
    const V3 g = v3(RE.m[k], RE.m[3 + k], RE.m[6 + k]);
    const V3 Tg = tmul(RF, g);
    const V3 Tv = tmul(RF, cross(g, dp));
    T[k][0] = Tg.x; T[k][1] = Tg.y; T[k][2] = Tg.z; T[k][3] = 0.0; T[k][4] = 0.0; T[k][5] = 0.0;
    T[3 + k][0] = Tv.x; T[3 + k][1] = Tv.y; T[3 + k][2] = Tv.z; T[3 + k][3] = Tg.x; T[3 + k][4] = Tg.y; T[3 + k][5] = Tg.z;
    if (want_dot) {
      const V3 a = v3(0, 0, 0) - cross(wrel, Tg);
      const V3 b = tmul(RF, cross(g, dvr)) - cross(wrel, Tv);
      Td[k][0] = a.x; Td[k][1] = a.y; Td[k][2] = a.z; Td[k][3] = 0.0; Td[k][4] = 0.0; Td[k][5] = 0.0;
      Td[3 + k][0] = b.x; Td[3 + k][1] = b.y; Td[3 + k][2] = b.z; Td[3 + k][3] = a.x; Td[3 + k][4] = a.y; Td[3 + k][5] = a.z;
    }
  }
}

// does an inertia with upstream mask `up` depend on column c of the (n_coords + 6 n_free)-column twist-shaping matrix?
GD bool col_upstream(unsigned up, int n, int c, int per_free = 6) { return c < n ? ((up >> c) & 1u) : ((up >> (RKB_GEN_FREE_BIT + (c - n) / per_free)) & 1u); }

// M and (optionally) S with Mdot = S + S^T; both n x n row-major in local memory.
template <int MAXF>
GD void mass(const GenericProgram* G, const Work<3, MAXF>& W, double* M, double* S, bool want_dot) {
  const int nc = G->n_coords, n = nc + 6 * G->n_free;
  for (int i = 0; i < n * n; ++i) { M[i] = 0.0; if (want_dot) S[i] = 0.0; }
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind == RKB_INERTIA_GEN) {
      M[E.coord * n + E.coord] += E.p[0];  // jacobian_gen_gen(1, 0): Tcm = 1, Tcm_dot = 0
    } else if (E.kind == RKB_INERTIA_3D) {
      const Fr3& F = W.fr[E.fa];
      const M3 RF = qrot(F.q);
      const double m = E.p[0];
      const double* I = &E.p[1];
      double T[MAXA][6], Td[MAXA][6], MT[MAXA][6];
      for (int c = 0; c < nc; ++c) {
        if (!((E.upstream >> c) & 1u)) continue;
        jac_col(G, W, c, F, RF, T[c], Td[c], want_dot);
      }
      for (int fj = 0; fj < G->n_free; ++fj)
        if ((E.upstream >> (RKB_GEN_FREE_BIT + fj)) & 1u) jac_free_block(G, W, fj, F, RF, &T[nc + 6 * fj], &Td[nc + 6 * fj], want_dot);
      for (int c = 0; c < n; ++c) {
        if (!col_upstream(E.upstream, nc, c)) continue;
        MT[c][0] = m * T[c][0]; MT[c][1] = m * T[c][1]; MT[c][2] = m * T[c][2];
        MT[c][3] = I[0] * T[c][3] + I[1] * T[c][4] + I[2] * T[c][5];
        MT[c][4] = I[1] * T[c][3] + I[3] * T[c][4] + I[4] * T[c][5];
        MT[c][5] = I[2] * T[c][3] + I[4] * T[c][4] + I[5] * T[c][5];
      }
      for (int a = 0; a < n; ++a) {
        if (!col_upstream(E.upstream, nc, a)) continue;
        for (int b = 0; b < n; ++b) {
          if (!col_upstream(E.upstream, nc, b)) continue;
          double s = 0.0, sd = 0.0;
          for (int k = 0; k < 6; ++k) { s += T[a][k] * MT[b][k]; if (want_dot) sd += Td[a][k] * MT[b][k]; }
          M[a * n + b] += s;
          if (want_dot) S[a * n + b] += sd;
        }
      }
    }
  }
}
// 2D: jacobian_gen_2D::get_jac_relative_to (motion_jacobians.hpp:139-147) with f2 = (~E) * F
// (frame_2D.hpp:288-300, 350-360): f2.p = R_E^T dp, f2.R = R_E^T R_F, f2.w = w_F - w_E,
// f2.v = R_E^T (dv - w_E % dp).  Column c of the twist-shaping matrix seen from the inertia frame F: a joint's column for
// c < n_coords; else column (c - n_coords) % 3 of a free_joint_2D (jacobian_2D_2D::get_jac_relative_to on the identity,
// motion_jacobians.hpp:448-470): a velocity input e_k gives Tv = R_F^T R_E e_k, Tw = 0, Tvd = -w_rel % Tv; the
// angular-velocity input gives what a revolute joint at E gives.
template <int MAXF>
GD void jac_col2(const GenericProgram* G, const Work<2, MAXF>& W, int c, const Fr2& F, double* col, double* cold) {
  const int nc = G->n_coords;
  // kind of column: 0 = a rotation about E (revolute joint, angular velocity of a free joint), 1 = a translation along
  // `ax` expressed in E (prismatic joint, velocity component of a free joint)
  int elem, translation;
  V2 ax = v2(0.0, 0.0);
  if (c < nc) {
    elem = G->jelem[c];
    translation = G->el[elem].kind == RKB_REVOLUTE_2D ? 0 : 1;
    if (translation) ax = v2(G->el[elem].p[0], G->el[elem].p[1]);
  } else {
    const int l = (c - nc) % 3;
    elem = G->free_elem[(c - nc) / 3];
    translation = l == 2 ? 0 : 1;
    if (translation) ax = l == 0 ? v2(1.0, 0.0) : v2(0.0, 1.0);
  }
  const Fr2 Ej = W.fr[G->el[elem].fb];
  const V2 dp = F.p - Ej.p, dv = F.v - Ej.v;
  const double wrel = F.w - Ej.w;
  V2 Tv, Tvd;
  double Tw;
  if (!translation) {
    Tv = rtmul(F.R, crs(1.0, dp));
    Tw = 1.0;
    Tvd = rtmul(F.R, crs(1.0, dv - crs(Ej.w, dp))) - crs(wrel, Tv);
  } else {
    Tv = rtmul(F.R, rmul(Ej.R, ax));
    Tw = 0.0;
    Tvd = v2(0, 0) - crs(wrel, Tv);
  }
  col[0] = Tv.x; col[1] = Tv.y; col[2] = Tw;
  cold[0] = Tvd.x; cold[1] = Tvd.y; cold[2] = 0.0;
}

template <int MAXF>
GD void mass(const GenericProgram* G, const Work<2, MAXF>& W, double* M, double* S, bool want_dot) {
  const int nc = G->n_coords, n = nc + 3 * G->n_free;
  for (int i = 0; i < n * n; ++i) { M[i] = 0.0; if (want_dot) S[i] = 0.0; }
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind == RKB_INERTIA_GEN) {
      M[E.coord * n + E.coord] += E.p[0];
    } else if (E.kind == RKB_INERTIA_2D) {
      const Fr2 F = W.fr[E.fa];
      double T[MAXA][3], Td[MAXA][3], MT[MAXA][3];
      for (int c = 0; c < n; ++c) {
        T[c][0] = 0.0; T[c][1] = 0.0; T[c][2] = 0.0; Td[c][0] = 0.0; Td[c][1] = 0.0; Td[c][2] = 0.0;
        if (col_upstream(E.upstream, nc, c, 3)) jac_col2(G, W, c, F, T[c], Td[c]);
        MT[c][0] = E.p[0] * T[c][0]; MT[c][1] = E.p[0] * T[c][1]; MT[c][2] = E.p[1] * T[c][2];
      }
      for (int a = 0; a < n; ++a) {
        if (!col_upstream(E.upstream, nc, a, 3)) continue;
        for (int b = 0; b < n; ++b) {
          if (!col_upstream(E.upstream, nc, b, 3)) continue;
          double s = 0.0, sd = 0.0;
          for (int k = 0; k < 3; ++k) { s += T[a][k] * MT[b][k]; if (want_dot) sd += Td[a][k] * MT[b][k]; }
          M[a * n + b] += s;
          if (want_dot) S[a * n + b] += sd;
        }
      }
    }
  }
}

// Tcm / Tcm_dot of mass_matrix_calc::get_TMT_TdMT written out (rows x n per sample; `row` of every inertia was
// fixed at lowering: gen inertias, then 2D, then 3D).  Entries of coordinates outside an inertia's upstream
// set stay zero.
template <int MAXF>
GD void tmt(const GenericProgram* G, const Work<3, MAXF>& W, const BatchView& T, const BatchView& Td, long long i) {
  const int nc = G->n_coords, n = nc + 6 * G->n_free;
  const bool want_dot = Td.p != (double*)0;
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind == RKB_INERTIA_GEN) {
      for (int c = 0; c < n; ++c) {
        const long long k = (long long)E.row * n + c;
        T.p[i * T.si + k * T.sk] = (c < nc && ((E.upstream >> c) & 1u)) ? 1.0 : 0.0;
        if (want_dot) Td.p[i * Td.si + k * Td.sk] = 0.0;
      }
    } else if (E.kind == RKB_INERTIA_3D) {
      const Fr3& F = W.fr[E.fa];
      const M3 RF = qrot(F.q);
      for (int c = 0; c < nc; ++c) {
        double col[6] = {0, 0, 0, 0, 0, 0}, cold[6] = {0, 0, 0, 0, 0, 0};
        if ((E.upstream >> c) & 1u) jac_col(G, W, c, F, RF, col, cold, want_dot);
        for (int r = 0; r < 6; ++r) {
          const long long k = (long long)(E.row + r) * n + c;
          T.p[i * T.si + k * T.sk] = col[r];
          if (want_dot) Td.p[i * Td.si + k * Td.sk] = cold[r];
        }
      }
      for (int fj = 0; fj < G->n_free; ++fj) {
        double blk[6][6], blkd[6][6];
        const bool up = (E.upstream >> (RKB_GEN_FREE_BIT + fj)) & 1u;
        if (up) jac_free_block(G, W, fj, F, RF, blk, blkd, want_dot);
        for (int l = 0; l < 6; ++l)
          for (int r = 0; r < 6; ++r) {
            const long long k = (long long)(E.row + r) * n + nc + 6 * fj + l;
            T.p[i * T.si + k * T.sk] = up ? blk[l][r] : 0.0;
            if (want_dot) Td.p[i * Td.si + k * Td.sk] = up ? blkd[l][r] : 0.0;
          }
      }
    }
  }
}
template <int MAXF>
GD void tmt(const GenericProgram* G, const Work<2, MAXF>& W, const BatchView& T, const BatchView& Td, long long i) {
  const int nc = G->n_coords, n = nc + 3 * G->n_free;
  const bool want_dot = Td.p != (double*)0;
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind == RKB_INERTIA_GEN) {
      for (int c = 0; c < n; ++c) {
        const long long k = (long long)E.row * n + c;
        T.p[i * T.si + k * T.sk] = (c < nc && ((E.upstream >> c) & 1u)) ? 1.0 : 0.0;
        if (want_dot) Td.p[i * Td.si + k * Td.sk] = 0.0;
      }
    } else if (E.kind == RKB_INERTIA_2D) {
      const Fr2& F = W.fr[E.fa];
      for (int c = 0; c < n; ++c) {
        double col[3] = {0, 0, 0}, cold[3] = {0, 0, 0};
        if (col_upstream(E.upstream, nc, c, 3)) jac_col2(G, W, c, F, col, cold);
        for (int r = 0; r < 3; ++r) {
          const long long k = (long long)(E.row + r) * n + c;
          T.p[i * T.si + k * T.sk] = col[r];
          if (want_dot) Td.p[i * Td.si + k * Td.sk] = cold[r];
        }
      }
    }
  }
}

// Jacobian of ONE frame with respect to the coordinates (and its time derivative): the rows a 3D / 2D inertia on that
// frame would have in Tcm — jacobian_gen_3D / _2D::get_jac_relative_to(frame) written by write_to_matrices — which is
// what manip_kin_mdl_jac_calculator::getJacobianMatrixAndDerivative stacks for the dependent (end-effector) frames of
// a manipulator (ctrl/mbd_kte/manipulator_model_helper.hpp:342-...).  Rows: v (3 / 2), then w (3 / 1), expressed in
// the frame's own coordinates like every twist in ReaK.
template <int MAXF>
GD void frame_jac(const GenericProgram* G, const Work<3, MAXF>& W, int frame, unsigned upstream, const BatchView& T, const BatchView& Td, long long i) {
  const int nc = G->n_coords, n = nc + 6 * G->n_free;
  const bool want_dot = Td.p != (double*)0;
  const Fr3& F = W.fr[frame];
  const M3 RF = qrot(F.q);
  for (int c = 0; c < nc; ++c) {
    double col[6] = {0, 0, 0, 0, 0, 0}, cold[6] = {0, 0, 0, 0, 0, 0};
    if ((upstream >> c) & 1u) jac_col(G, W, c, F, RF, col, cold, want_dot);
    for (int r = 0; r < 6; ++r) {
      const long long k = (long long)r * n + c;
      T.p[i * T.si + k * T.sk] = col[r];
      if (want_dot) Td.p[i * Td.si + k * Td.sk] = cold[r];
    }
  }
  for (int fj = 0; fj < G->n_free; ++fj) {  // jacobian_3D_3D columns of a free joint upstream of the frame
    double blk[6][6], blkd[6][6];
    const bool up = (upstream >> (RKB_GEN_FREE_BIT + fj)) & 1u;
    if (up) jac_free_block(G, W, fj, F, RF, blk, blkd, want_dot);
    for (int l = 0; l < 6; ++l)
      for (int r = 0; r < 6; ++r) {
        const long long k = (long long)r * n + nc + 6 * fj + l;
        T.p[i * T.si + k * T.sk] = up ? blk[l][r] : 0.0;
        if (want_dot) Td.p[i * Td.si + k * Td.sk] = up ? blkd[l][r] : 0.0;
      }
  }
}
template <int MAXF>
GD void frame_jac(const GenericProgram* G, const Work<2, MAXF>& W, int frame, unsigned upstream, const BatchView& T, const BatchView& Td, long long i) {
  const int nc = G->n_coords, n = nc + 3 * G->n_free;
  const bool want_dot = Td.p != (double*)0;
  const Fr2& F = W.fr[frame];
  for (int c = 0; c < n; ++c) {
    double col[3] = {0, 0, 0}, cold[3] = {0, 0, 0};
    if (col_upstream(upstream, nc, c, 3)) jac_col2(G, W, c, F, col, cold);
    for (int r = 0; r < 3; ++r) {
      const long long k = (long long)r * n + c;
      T.p[i * T.si + k * T.sk] = col[r];
      if (want_dot) Td.p[i * Td.si + k * Td.sk] = cold[r];
    }
  }
}

// linsolve_Cholesky (mat_cholesky.hpp:63-84, 160-179) on a row-major n x n matrix, in place
GD int cholesky_solve(int n, double* A, double* b) {
  int st = 0;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < i; ++j) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / A[j * n + j];
    }
    double d = A[i * n + i];
    for (int k = 0; k < i; ++k) d -= A[i * n + k] * A[i * n + k];
    if (!(d >= 1.0e-8)) st = RKB_STATUS_SINGULAR;
    A[i * n + i] = sqrt(d);
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i * n + k] * b[k];
    b[i] = s / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = n - 1; k > i; --k) s -= A[k * n + i] * b[k];
    b[i] = s / A[i * n + i];
  }
  return st;
}

// accelerations into W.f (the coordinates', then 6 per free joint: kte_nl_system.hpp:256-273); returns status
template <int MAXF>
GD void pack_free_forces(const GenericProgram* G, Work<3, MAXF>& W) {
  if (G->n_free) {
    double* f = &W.f[G->n_coords];
    f[0] = W.fc.F.x; f[1] = W.fc.F.y; f[2] = W.fc.F.z; f[3] = W.fc.T.x; f[4] = W.fc.T.y; f[5] = W.fc.T.z;
  }
}
template <int MAXF>
GD void pack_free_forces(const GenericProgram* G, Work<2, MAXF>& W) {
  if (G->n_free) {
    double* f = &W.f[G->n_coords];
    f[0] = W.fc.F.x; f[1] = W.fc.F.y; f[2] = W.fc.T;
  }
}

template <int DIM, int MAXF>
GD int accel(const GenericProgram* G, Work<DIM, MAXF>& W) {
  double M[MAXA * MAXA];
  motion(G, W);
  force(G, W);
  pack_free_forces(G, W);
  mass(G, W, M, (double*)0, false);
  return cholesky_solve(G->n_coords + G->free_acc * G->n_free, M, W.f);
}

// kte_nl_system::apply_states_and_inputs for the free joint's 13 states (kte_nl_system.hpp:205-219): the quaternion is
// normalised (explicit quaternion(Vector), rotations_3D.hpp:917-920), the accelerations are zero
template <int MAXF>
GD void apply_free(Work<3, MAXF>& W, const double* s) {
  Fr3& C = W.fc;
  C.p = v3(s[0], s[1], s[2]);
  const double nq = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
  C.q.w = s[3] / nq; C.q.x = s[4] / nq; C.q.y = s[5] / nq; C.q.z = s[6] / nq;
  C.v = v3(s[7], s[8], s[9]);
  C.w = v3(s[10], s[11], s[12]);
  C.a = v3(0, 0, 0); C.al = v3(0, 0, 0);
}
// ... and for a free_joint_2D's 7 states (kte_nl_system.hpp:194-204): rot_mat_2D(vect<2>) normalises (rotations_2D.hpp:119-123)
template <int MAXF>
GD void apply_free(Work<2, MAXF>& W, const double* s) {
  Fr2& C = W.fc;
  C.p = v2(s[0], s[1]);
  const double nr = sqrt(s[2] * s[2] + s[3] * s[3]);
  C.R.c = s[2] / nr; C.R.s = s[3] / nr;
  W.fraw[0] = s[2]; W.fraw[1] = s[3];
  C.v = v2(s[4], s[5]);
  C.w = s[6];
  C.a = v2(0, 0); C.al = 0.0;
}

// the free joint's 13 state derivatives after accel(): Velocity, QuatDot (quaternion::getQuaternionDot,
// rotations_3D.hpp:1206-1211, of the normalised quaternion), the six accelerations (kte_nl_system.hpp:293-308)
template <int MAXF>
GD void free_derivative(const GenericProgram* G, const Work<3, MAXF>& W, double* o) {
  const Fr3& C = W.fc;
  o[0] = C.v.x; o[1] = C.v.y; o[2] = C.v.z;
  o[3] = -0.5 * (C.q.x * C.w.x + C.q.y * C.w.y + C.q.z * C.w.z);
  o[4] = 0.5 * (C.q.w * C.w.x - C.q.z * C.w.y + C.q.y * C.w.z);
  o[5] = 0.5 * (C.q.w * C.w.y + C.q.z * C.w.x - C.q.x * C.w.z);
  o[6] = 0.5 * (C.q.w * C.w.z - C.q.y * C.w.x + C.q.x * C.w.y);
  for (int k = 0; k < 6; ++k) o[7 + k] = W.f[G->n_coords + k];
}
// free_joint_2D (kte_nl_system.hpp:282-291): Velocity, (-sin, cos) AngVelocity from the RAW state, three accelerations
template <int MAXF>
GD void free_derivative(const GenericProgram* G, const Work<2, MAXF>& W, double* o) {
  const Fr2& C = W.fc;
  o[0] = C.v.x; o[1] = C.v.y;
  o[2] = -W.fraw[1] * C.w;
  o[3] = W.fraw[0] * C.w;
  for (int k = 0; k < 3; ++k) o[4 + k] = W.f[G->n_coords + k];
}

template <int DIM, int MAXF>
GD void load(const GenericProgram* G, Work<DIM, MAXF>& W, const ConstBatchView& x, const ConstBatchView& u, long long ix, long long iu, bool with_u) {
  for (int c = 0; c < G->n_coords; ++c) {
    W.q[c] = x.p[ix * x.si + rkb_state_q(x.blocked, G->n_coords, c) * x.sk];
    W.qd[c] = x.p[ix * x.si + rkb_state_qd(x.blocked, G->n_coords, c) * x.sk];
  }
  for (int k = 0; k < G->n_aux; ++k) { W.q[G->n_coords + k] = G->aux_q[k]; W.qd[G->n_coords + k] = G->aux_qd[k]; }
  if (G->n_free) {  // the 13 states of the free joint follow the coordinates' (never blocked: rejected by the host)
    double s[13];
    for (int k = 0; k < G->free_states; ++k) s[k] = x.p[ix * x.si + (2 * G->n_coords + k) * x.sk];
    apply_free(W, s);
  }
  for (int k = 0; k < G->n_inputs; ++k) W.u[k] = with_u ? u.p[iu * u.si + k * u.sk] : 0.0;
}

#define GEN_BLOCK 128

template <int DIM, int MAXF>
__global__ void __launch_bounds__(GEN_BLOCK) generic_eval_kernel(const GenericProgram* __restrict__ G, const EvalArgs A) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Work<DIM, MAXF> W;
  load(G, W, A.x, A.u, i, i, true);
  int st = accel(G, W);
  bool finite = true;
  for (int c = 0; c < G->n_coords; ++c) {
    A.out.p[i * A.out.si + rkb_state_q(A.out.blocked, G->n_coords, c) * A.out.sk] = W.qd[c];
    A.out.p[i * A.out.si + rkb_state_qd(A.out.blocked, G->n_coords, c) * A.out.sk] = W.f[c];
    finite = finite && isfinite(W.f[c]) && isfinite(W.qd[c]);
  }
  if (G->n_free) {
    double o[13];
    free_derivative(G, W, o);
    for (int k = 0; k < G->free_states; ++k) {
      A.out.p[i * A.out.si + (2 * G->n_coords + k) * A.out.sk] = o[k];
      finite = finite && isfinite(o[k]);
    }
  }
  if (!finite) st |= RKB_STATUS_NONFINITE;
  if (A.status) A.status[i] = st;
}

template <int DIM, int MAXF>
__global__ void __launch_bounds__(GEN_BLOCK) generic_forces_kernel(const GenericProgram* __restrict__ G, const EvalArgs A) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Work<DIM, MAXF> W;
  load(G, W, A.x, A.u, i, i, true);
  motion(G, W);
  force(G, W);
  pack_free_forces(G, W);
  for (int c = 0; c < G->n_coords + G->free_acc * G->n_free; ++c) A.out.p[i * A.out.si + c * A.out.sk] = W.f[c];
}

template <int DIM, int MAXF>
__global__ void __launch_bounds__(GEN_BLOCK) generic_mass_kernel(const GenericProgram* __restrict__ G, const EvalArgs A) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Work<DIM, MAXF> W;
  load(G, W, A.x, A.u, i, i, false);
  motion(G, W);
  double M[MAXA * MAXA], S[MAXA * MAXA];
  const bool want_dot = A.out2.p != (double*)0;
  mass(G, W, M, S, want_dot);
  const int n = G->n_coords + G->free_acc * G->n_free;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      // mat<symmetric> converting ctor averages the two halves (mat_alg_symmetric.hpp:171-200)
      const double m = (a == b) ? M[a * n + a] : 0.5 * (M[a * n + b] + M[b * n + a]);
      A.out.p[i * A.out.si + (long long)(a * n + b) * A.out.sk] = m;
      if (want_dot) A.out2.p[i * A.out2.si + (long long)(a * n + b) * A.out2.sk] = S[a * n + b] + S[b * n + a];
    }
}

GD void put_frame(const Fr3& F, const BatchView& o, long long base) {
  const double v[25] = {F.p.x, F.p.y, F.p.z, F.q.w, F.q.x, F.q.y, F.q.z, F.v.x, F.v.y, F.v.z, F.w.x, F.w.y, F.w.z,
                        F.a.x, F.a.y, F.a.z, F.al.x, F.al.y, F.al.z, F.F.x, F.F.y, F.F.z, F.T.x, F.T.y, F.T.z};
  for (int k = 0; k < 25; ++k) o.p[base + k * o.sk] = v[k];
}
GD void put_frame(const Fr2& F, const BatchView& o, long long base) {
  const double v[25] = {F.p.x, F.p.y, 0, F.R.c, F.R.s, 0, 0, F.v.x, F.v.y, 0, F.w, 0, 0, F.a.x, F.a.y, 0, F.al, 0, 0, F.F.x, F.F.y, 0, F.T, 0, 0};
  for (int k = 0; k < 25; ++k) o.p[base + k * o.sk] = v[k];
}

template <int DIM, int MAXF>
__global__ void __launch_bounds__(GEN_BLOCK) generic_frames_kernel(const GenericProgram* __restrict__ G, const EvalArgs A) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Work<DIM, MAXF> W;
  load(G, W, A.x, A.u, i, i, true);
  motion(G, W);
  force(G, W);
  for (int f = 0; f < G->n_frames; ++f) put_frame(W.fr[f], A.out, i * A.out.si + (long long)(25 * f) * A.out.sk);
}

#include "kte_proximity.cuh"

// doMotion at position level only (the poses are all a proximity query reads): the Position / Quat lines of
// motion() above, operation for operation.  The frame being built lives in registers; only the frames the
// program marks (ProxProgram::slot_of) are written to the thread's local array.
template <int MAXS>
GD void motion_pose(const GenericProgram* G, const ProxProgram& P, const double* q, const Pose& freec, Pose (&slots)[MAXS]) {
  Pose cur;
  {
    const double* b = G->base;
    cur.p = ldv(b); cur.q.w = b[3]; cur.q.x = b[4]; cur.q.y = b[5]; cur.q.z = b[6];
  }
  int last = G->base_frame;
  if (P.slot_of[last] >= 0) slots[P.slot_of[last]] = cur;
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind != RKB_REVOLUTE_3D && E.kind != RKB_PRISMATIC_3D && E.kind != RKB_RIGID_LINK_3D && E.kind != RKB_FREE_3D) continue;
    Pose B = cur;
    if (E.fa != last) B = slots[P.slot_of[E.fa]];
    if (E.kind == RKB_REVOLUTE_3D) {  // revolute_joint.cpp:121-131
      const V3 an = unit_axis(ldv(E.p));
      double sh, ch;
      sincos(0.5 * q[E.coord], &sh, &ch);
      Q4 tq; tq.w = ch; tq.x = an.x * sh; tq.y = an.y * sh; tq.z = an.z * sh;
      cur.p = B.p;
      cur.q = qmul(B.q, tq);
    } else if (E.kind == RKB_PRISMATIC_3D) {  // prismatic_joint.cpp:129-140
      cur.p = B.p + mul(qrot(B.q), q[E.coord] * ldv(E.p));
      cur.q = B.q;
    } else if (E.kind == RKB_FREE_3D) {  // free_joints.cpp:127: End = Base * Coord
      cur.p = B.p + mul(qrot(B.q), freec.p);
      cur.q = qmul(B.q, freec.q);
    } else {  // rigid_link.cpp:156 -> pose_3D::addBefore
      Q4 qo; qo.w = E.p[3]; qo.x = E.p[4]; qo.y = E.p[5]; qo.z = E.p[6];
      cur.p = B.p + mul(qrot(B.q), ldv(E.p));
      cur.q = qmul(B.q, qo);
    }
    last = E.fb;
    if (P.slot_of[last] >= 0) slots[P.slot_of[last]] = cur;
  }
}

// proxy_query_pair_3D::findMinimumDistance (proxy_query_model.cpp:388-412) at the chain's pose for
// state x[i]: distance, the index of the finder that gave it and its two points.
template <int DIM, int MAXS, int MINB>
__global__ void __launch_bounds__(GEN_BLOCK, MINB) generic_proximity_kernel(const GenericProgram* __restrict__ G, const EvalArgs A,
                                                                       const __grid_constant__ ProxProgram P) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  double q[MAXC];
  for (int c = 0; c < G->n_coords; ++c) q[c] = A.x.p[i * A.x.si + rkb_state_q(A.x.blocked, G->n_coords, c) * A.x.sk];
  Pose fr[MAXS], freec;
  freec.p = v3(0, 0, 0); freec.q.w = 1.0; freec.q.x = 0.0; freec.q.y = 0.0; freec.q.z = 0.0;
  if (G->n_free) {  // pose states of the free joint, the quaternion normalised as apply_states_and_inputs does
    double s[7];
    for (int k = 0; k < 7; ++k) s[k] = A.x.p[i * A.x.si + (2 * G->n_coords + k) * A.x.sk];
    const double nq = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
    freec.p = v3(s[0], s[1], s[2]);
    freec.q.w = s[3] / nq; freec.q.x = s[4] / nq; freec.q.y = s[5] / nq; freec.q.z = s[6] / nq;
  }
  motion_pose(G, P, q, freec, fr);
  ProxRecord bestR;
  const int best = prox_min_distance(P, fr, A.out2.p != (double*)0, bestR);
  A.out.p[i * A.out.si] = bestR.d;
  if (A.status) A.status[i] = best;
  if (A.out2.p) {
    const double v[6] = {bestR.p1.x, bestR.p1.y, bestR.p1.z, bestR.p2.x, bestR.p2.y, bestR.p2.z};
    for (int k = 0; k < 6; ++k) A.out2.p[i * A.out2.si + k * A.out2.sk] = v[k];
  }
}

// proxy_query_pair_3D::gatherCollisionPoints at the chain's pose for state x[i]: count into status[i], records into out
// ([N][max_records][7]: distance, point 1, point 2) and the finder index of each into out2's int view ([N][max_records])
template <int DIM, int MAXS>
__global__ void __launch_bounds__(GEN_BLOCK) generic_collision_kernel(const GenericProgram* __restrict__ G, const EvalArgs A,
                                                                       const __grid_constant__ ProxProgram P, int max_records,
                                                                       int32_t* __restrict__ finder) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  double q[MAXC];
  for (int c = 0; c < G->n_coords; ++c) q[c] = A.x.p[i * A.x.si + rkb_state_q(A.x.blocked, G->n_coords, c) * A.x.sk];
  Pose fr[MAXS], freec;
  freec.p = v3(0, 0, 0); freec.q.w = 1.0; freec.q.x = 0.0; freec.q.y = 0.0; freec.q.z = 0.0;
  if (G->n_free) {
    double s[7];
    for (int k = 0; k < 7; ++k) s[k] = A.x.p[i * A.x.si + (2 * G->n_coords + k) * A.x.sk];
    const double nq = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
    freec.p = v3(s[0], s[1], s[2]);
    freec.q.w = s[3] / nq; freec.q.x = s[4] / nq; freec.q.y = s[5] / nq; freec.q.z = s[6] / nq;
  }
  motion_pose(G, P, q, freec, fr);
  double* rec = A.out.p + i * (long long)max_records * 7;
  int32_t* fnd = finder ? finder + i * (long long)max_records : (int32_t*)0;
  const int n = prox_gather_collisions(P, fr, max_records, [&](int r, int f, const ProxRecord& R) {
    double* o = rec + 7 * r;
    o[0] = R.d; o[1] = R.p1.x; o[2] = R.p1.y; o[3] = R.p1.z; o[4] = R.p2.x; o[5] = R.p2.y; o[6] = R.p2.z;
    if (fnd) fnd[r] = f;
  });
  for (int r = n; r < max_records; ++r) {
    double* o = rec + 7 * r;
    o[0] = INFINITY;
    for (int k = 1; k < 7; ++k) o[k] = 0.0;
    if (fnd) fnd[r] = -1;
  }
  A.status[i] = n;
}

#include "kte_proximity2d.cuh"

// doMotion of a planar chain at position level (revolute_joint.cpp:32-58, prismatic_joint.cpp:33-67, free_joints.cpp:33-41,
// rigid_link.cpp:87-99: the Position / Rotation lines of motion() above), frames kept as motion_pose keeps them
template <int MAXS>
GD void motion_pose2(const GenericProgram* G, const ProxProgram& P, const double* q, const Pose2& freec, Pose2 (&slots)[MAXS]) {
  Pose2 cur;
  {
    const double* b = G->base;
    cur.p = v2(b[0], b[1]); cur.R.c = b[3]; cur.R.s = b[4];
  }
  int last = G->base_frame;
  if (P.slot_of[last] >= 0) slots[P.slot_of[last]] = cur;
  for (int e = 0; e < G->n_elements; ++e) {
    const GenericElement& E = G->el[e];
    if (E.kind != RKB_REVOLUTE_2D && E.kind != RKB_PRISMATIC_2D && E.kind != RKB_RIGID_LINK_2D && E.kind != RKB_FREE_2D) continue;
    Pose2 B = cur;
    if (E.fa != last) B = slots[P.slot_of[E.fa]];
    if (E.kind == RKB_REVOLUTE_2D) {
      R2 rq;
      sincos(q[E.coord], &rq.s, &rq.c);
      cur.p = B.p;
      cur.R = rr(B.R, rq);
    } else if (E.kind == RKB_PRISMATIC_2D) {
      cur.p = B.p + rmul(B.R, q[E.coord] * v2(E.p[0], E.p[1]));
      cur.R = B.R;
    } else if (E.kind == RKB_FREE_2D) {
      cur.p = B.p + rmul(B.R, freec.p);
      cur.R = rr(B.R, freec.R);
    } else {
      R2 ro; ro.c = E.p[3]; ro.s = E.p[4];
      cur.p = B.p + rmul(B.R, v2(E.p[0], E.p[1]));
      cur.R = rr(B.R, ro);
    }
    last = E.fb;
    if (P.slot_of[last] >= 0) slots[P.slot_of[last]] = cur;
  }
}

template <int MAXS>
GD void prox_load2(const GenericProgram* G, const ProxProgram& P, const EvalArgs& A, long long i, Pose2 (&fr)[MAXS]) {
  double q[MAXC];
  for (int c = 0; c < G->n_coords; ++c) q[c] = A.x.p[i * A.x.si + rkb_state_q(A.x.blocked, G->n_coords, c) * A.x.sk];
  Pose2 freec;
  freec.p = v2(0, 0); freec.R.c = 1.0; freec.R.s = 0.0;
  if (G->n_free) {  // position and (cos, sin) states of the free_joint_2D, normalised as apply_free does
    double s[4];
    for (int k = 0; k < 4; ++k) s[k] = A.x.p[i * A.x.si + (2 * G->n_coords + k) * A.x.sk];
    const double nr = sqrt(s[2] * s[2] + s[3] * s[3]);
    freec.p = v2(s[0], s[1]);
    freec.R.c = s[2] / nr; freec.R.s = s[3] / nr;
  }
  motion_pose2(G, P, q, freec, fr);
}

// proxy_query_pair_2D::findMinimumDistance (proxy_query_model.cpp:163-190) at the pose of state x[i]; points as (x, y, 0)
template <int MAXS>
__global__ void __launch_bounds__(GEN_BLOCK) generic_proximity2d_kernel(const GenericProgram* __restrict__ G, const EvalArgs A,
                                                                         const __grid_constant__ ProxProgram P) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Pose2 fr[MAXS];
  prox_load2(G, P, A, i, fr);
  ProxRecord2 R;
  const int best = prox_min_distance2(P, fr, R);
  A.out.p[i * A.out.si] = R.d;
  if (A.status) A.status[i] = best;
  if (A.out2.p) {
    const double v[6] = {R.p1.x, R.p1.y, 0.0, R.p2.x, R.p2.y, 0.0};
    for (int k = 0; k < 6; ++k) A.out2.p[i * A.out2.si + k * A.out2.sk] = v[k];
  }
}

// proxy_query_pair_2D::gatherCollisionPoints (proxy_query_model.cpp:192-212); buffers as generic_collision_kernel
template <int MAXS>
__global__ void __launch_bounds__(GEN_BLOCK) generic_collision2d_kernel(const GenericProgram* __restrict__ G, const EvalArgs A,
                                                                         const __grid_constant__ ProxProgram P, int max_records,
                                                                         int32_t* __restrict__ finder) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Pose2 fr[MAXS];
  prox_load2(G, P, A, i, fr);
  double* rec = A.out.p + i * (long long)max_records * 7;
  int32_t* fnd = finder ? finder + i * (long long)max_records : (int32_t*)0;
  const int n = prox_gather_collisions2(P, fr, max_records, [&](int r, int f, const ProxRecord2& R) {
    double* o = rec + 7 * r;
    o[0] = R.d; o[1] = R.p1.x; o[2] = R.p1.y; o[3] = 0.0; o[4] = R.p2.x; o[5] = R.p2.y; o[6] = 0.0;
    if (fnd) fnd[r] = f;
  });
  for (int r = n; r < max_records; ++r) {
    double* o = rec + 7 * r;
    o[0] = INFINITY;
    for (int k = 1; k < 7; ++k) o[k] = 0.0;
    if (fnd) fnd[r] = -1;
  }
  A.status[i] = n;
}

template <int DIM, int MAXF>
__global__ void __launch_bounds__(GEN_BLOCK) generic_tmt_kernel(const GenericProgram* __restrict__ G, const EvalArgs A) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Work<DIM, MAXF> W;
  load(G, W, A.x, A.u, i, i, false);
  motion(G, W);
  tmt(G, W, A.out, A.out2, i);
}

template <int DIM, int MAXF>
__global__ void __launch_bounds__(GEN_BLOCK) generic_frame_jac_kernel(const GenericProgram* __restrict__ G, const EvalArgs A, int frame, unsigned upstream) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  Work<DIM, MAXF> W;
  load(G, W, A.x, A.u, i, i, false);
  motion(G, W);
  frame_jac(G, W, frame, upstream, A.out, A.out2, i);
}

// The integrators work on the state VECTOR as the reference's do (vect_n<double>; a free joint's quaternion is only
// normalised when the state is applied to the model): xs = (q0, qd0, q1, qd1, ..., then the 13 states of a free joint).
GD void load_flat(const GenericProgram* G, const ConstBatchView& x, long long ix, double* xs) {
  const int n = G->n_coords;
  for (int c = 0; c < n; ++c) {
    xs[2 * c] = x.p[ix * x.si + rkb_state_q(x.blocked, n, c) * x.sk];
    xs[2 * c + 1] = x.p[ix * x.si + rkb_state_qd(x.blocked, n, c) * x.sk];
  }
  for (int k = 0; k < G->free_states * G->n_free; ++k) xs[2 * n + k] = x.p[ix * x.si + (2 * n + k) * x.sk];
}
GD void store_flat(const GenericProgram* G, const double* xs, const BatchView& o, long long off) {
  const int n = G->n_coords;
  for (int c = 0; c < n; ++c) {
    o.p[off + rkb_state_q(o.blocked, n, c) * o.sk] = xs[2 * c];
    o.p[off + rkb_state_qd(o.blocked, n, c) * o.sk] = xs[2 * c + 1];
  }
  for (int k = 0; k < G->free_states * G->n_free; ++k) o.p[off + (2 * n + k) * o.sk] = xs[2 * n + k];
}
// apply_states_and_inputs + get_state_derivative at the state vector xs: xd = f(xs, u); returns the status bits
// A real call, not inlined: the RK4 loop evaluates the chain four times per step, and four inlined copies of the whole
// interpreter (motion, force, mass matrix, solve) made the kernel's code four times as large and cost it a CTA per SM
// (160 -> 126 registers).  Measured on B200, 2^18 samples x 10 steps: free-base arm 139.6 -> 86.8 ms, arm with two-anchor
// springs 52.7 -> 25.9 ms (profiles/r2_interpreter.md).  Same arithmetic: the body is compiled once instead of four times.
template <int DIM, int MAXF>
GNI int rate(const GenericProgram* G, Work<DIM, MAXF>& W, const double* xs, double* xd) {
  const int n = G->n_coords;
  for (int c = 0; c < n; ++c) { W.q[c] = xs[2 * c]; W.qd[c] = xs[2 * c + 1]; }
  if (G->n_free) apply_free(W, xs + 2 * n);
  const int st = accel(G, W);
  for (int c = 0; c < n; ++c) { xd[2 * c] = W.qd[c]; xd[2 * c + 1] = W.f[c]; }
  if (G->n_free) free_derivative(G, W, xd + 2 * n);
  return st;
}

// n_steps RK4 steps (fixed_step_integrators.hpp:277-289, the reference's own operation order) with
// the input held constant; TABLE selects the table-driven form for the other schemes.
template <int DIM, int MAXF, bool TABLE>
__global__ void __launch_bounds__(GEN_BLOCK) generic_rollout_kernel(const GenericProgram* __restrict__ G, const RolloutArgs A, const RkTable T) {
  const long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  if (A.active && !A.active[i]) return;
  Work<DIM, MAXF> W;
  const long long i0 = A.x0_div > 1 ? i / A.x0_div : i;
  const int nx = 2 * G->n_coords + G->free_states * G->n_free;
  double xs[MAXX], xd[MAXX];
  load_flat(G, A.x0, i0, xs);
  for (int k = 0; k < G->n_inputs; ++k) W.u[k] = A.u.p[i * A.u.si + k * A.u.sk];
  for (int k = 0; k < G->n_aux; ++k) { W.q[G->n_coords + k] = G->aux_q[k]; W.qd[G->n_coords + k] = G->aux_qd[k]; }
  const double dt = A.dt;
  double w[MAXX], acc[MAXX], k3[MAXX];
  double ks[TABLE ? RKB_RK_MAX_STAGES : 1][MAXX];
  int st = 0;
  {
    for (int step = 0; step < A.n_steps; ++step) {
      if (TABLE) {
        for (int k = 0; k < nx; ++k) w[k] = xs[k];
        for (int s = 0; s < T.stages; ++s) {
          st |= rate(G, W, xs, xd);
          for (int k = 0; k < nx; ++k) ks[s][k] = xd[k] * dt;
          for (int k = 0; k < nx; ++k) {
            double v = w[k];
            for (int j = 0; j <= s; ++j) v = fma(T.c[s][j], ks[j][k], v);
            xs[k] = v;
          }
        }
        continue;
      }
      // inputs sampled at every half step (runge_kutta4_integrate_impl, runge_kutta4_integrator_sys.hpp:50-97):
      // evaluation 1 reads node 2 step, evaluations 2 and 3 node 2 step + 1, evaluation 4 node 2 step + 2
      const long long un = A.u_node_stride;
      if (un) for (int k = 0; k < G->n_inputs; ++k) W.u[k] = A.u.p[i * A.u.si + (2LL * step) * un + k * A.u.sk];
      st |= rate(G, W, xs, xd);
      for (int k = 0; k < nx; ++k) {
        const double kk = xd[k] * dt;
        w[k] = xs[k];
        acc[k] = kk;
        xs[k] += kk * 0.5;
      }
      if (un) for (int k = 0; k < G->n_inputs; ++k) W.u[k] = A.u.p[i * A.u.si + (2LL * step + 1) * un + k * A.u.sk];
      st |= rate(G, W, xs, xd);
      for (int k = 0; k < nx; ++k) {
        const double kk = xd[k] * dt;
        acc[k] += kk * 2.0;
        xs[k] = w[k] + kk * 0.5;
      }
      st |= rate(G, W, xs, xd);
      for (int k = 0; k < nx; ++k) {
        const double kk = xd[k] * dt;
        k3[k] = kk;
        xs[k] = w[k] + kk;
      }
      if (un) for (int k = 0; k < G->n_inputs; ++k) W.u[k] = A.u.p[i * A.u.si + (2LL * step + 2) * un + k * A.u.sk];
      st |= rate(G, W, xs, xd);
      for (int k = 0; k < nx; ++k) xs[k] += (acc[k] + xd[k] * dt) / 6.0 - k3[k] * (2.0 / 3.0);
    }
  }
  bool finite = true;
  for (int k = 0; k < nx; ++k) finite = finite && isfinite(xs[k]);
  store_flat(G, xs, A.xout, i * A.xout.si);
  if (A.traj.p) store_flat(G, xs, A.traj, i * A.traj.si);
  if (!finite) st |= RKB_STATUS_NONFINITE;
  if (A.status) A.status[i] = A.status_or ? (A.status[i] | st) : st;
}

#ifndef RKB_HOST_TEST  // (tests/host_build/generic_host.cpp compiles everything above for the host: no launches there)
unsigned grid_of(long long n) { return (unsigned)((n + GEN_BLOCK - 1) / GEN_BLOCK); }

#define DISPATCH(kernel, host, ...)                                                              \
  do {                                                                                           \
    if ((host).dim == 3) {                                                                       \
      if ((host).n_frames <= 16) kernel<3, 16><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);     \
      else kernel<3, RKB_GEN_MAX_FRAMES><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);           \
    } else {                                                                                     \
      if ((host).n_frames <= 16) kernel<2, 16><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);     \
      else kernel<2, RKB_GEN_MAX_FRAMES><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);           \
    }                                                                                            \
  } while (0)

#define DISPATCH2(kernel, flag, host, ...)                                                             \
  do {                                                                                                 \
    if ((host).dim == 3) {                                                                             \
      if ((host).n_frames <= 16) kernel<3, 16, flag><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);     \
      else kernel<3, RKB_GEN_MAX_FRAMES, flag><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);           \
    } else {                                                                                           \
      if ((host).n_frames <= 16) kernel<2, 16, flag><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);     \
      else kernel<2, RKB_GEN_MAX_FRAMES, flag><<<grid_of(n), GEN_BLOCK, 0, s>>>(__VA_ARGS__);           \
    }                                                                                                  \
  } while (0)

// ---- steer: per-pair arg-min over the rollout end states -------------------------------------------
// One CTA per pair.  cost = || x_end - goal ||_2 over the 2n state components; the lowest
// rollout index wins ties.
__global__ void __launch_bounds__(256) steer_reduce_kernel(int nx, long long n_rollouts, const double* __restrict__ xend,
                                                            const double* __restrict__ goal, int32_t* best_idx,
                                                            double* best_x, double* best_cost) {
  __shared__ double s_cost[256];
  __shared__ long long s_idx[256];
  const long long pair = blockIdx.x;
  double best = INFINITY;
  long long bi = -1;
  for (long long r = threadIdx.x; r < n_rollouts; r += blockDim.x) {
    const double* xe = xend + (pair * n_rollouts + r) * nx;
    double s = 0.0;
    for (int k = 0; k < nx; ++k) { const double d = xe[k] - goal[pair * nx + k]; s += d * d; }
    double c = sqrt(s);
    if (!isfinite(c)) c = INFINITY;  // a diverged rollout (NaN / inf end state) never beats a finite one; all diverged: index 0, cost +inf
    if (bi < 0 || c < best) { best = c; bi = r; }
  }
  s_cost[threadIdx.x] = best;
  s_idx[threadIdx.x] = bi;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const double c2 = s_cost[threadIdx.x + off];
      const long long i2 = s_idx[threadIdx.x + off];
      const double c1 = s_cost[threadIdx.x];
      const long long i1 = s_idx[threadIdx.x];
      if (i2 >= 0 && (i1 < 0 || c2 < c1 || (c2 == c1 && i2 < i1))) { s_cost[threadIdx.x] = c2; s_idx[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  const long long win = s_idx[0] < 0 ? 0 : s_idx[0];
  if (threadIdx.x == 0) {
    best_idx[pair] = (int32_t)win;
    if (best_cost) best_cost[pair] = s_cost[0];
  }
  for (int k = threadIdx.x; k < nx; k += blockDim.x) best_x[pair * nx + k] = xend[(pair * n_rollouts + win) * nx + k];
}
#endif  // RKB_HOST_TEST

}  // namespace

#ifndef RKB_HOST_TEST

cudaError_t rkb_generic_eval(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  DISPATCH(generic_eval_kernel, host, prog, a);
  return cudaGetLastError();
}
cudaError_t rkb_generic_forces(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  DISPATCH(generic_forces_kernel, host, prog, a);
  return cudaGetLastError();
}
cudaError_t rkb_generic_mass(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  DISPATCH(generic_mass_kernel, host, prog, a);
  return cudaGetLastError();
}
cudaError_t rkb_generic_frames(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  DISPATCH(generic_frames_kernel, host, prog, a);
  return cudaGetLastError();
}
cudaError_t rkb_generic_proximity(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, const ProxProgram& pp, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  if (host.dim == 2) {  // planar models (kte_proximity2d.cuh)
    if (pp.n_slots <= 16) generic_proximity2d_kernel<16><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp);
    else generic_proximity2d_kernel<RKB_GEN_MAX_FRAMES><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp);
    return cudaGetLastError();
  }
  // 6 resident CTAs per SM (80 registers): measured best of 3 / 4 / 5 / 6
  if (pp.n_slots <= 8) generic_proximity_kernel<3, 8, 6><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp);
  else generic_proximity_kernel<3, RKB_GEN_MAX_FRAMES, 4><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp);
  return cudaGetLastError();
}
cudaError_t rkb_generic_collisions(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, const ProxProgram& pp,
                                   int max_records, int32_t* finder, cudaStream_t s) {
  (void)host;
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  if (host.dim == 2) {
    if (pp.n_slots <= 16) generic_collision2d_kernel<16><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp, max_records, finder);
    else generic_collision2d_kernel<RKB_GEN_MAX_FRAMES><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp, max_records, finder);
    return cudaGetLastError();
  }
  if (pp.n_slots <= 8) generic_collision_kernel<3, 8><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp, max_records, finder);
  else if (pp.n_slots <= 16) generic_collision_kernel<3, 16><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp, max_records, finder);
  else generic_collision_kernel<3, RKB_GEN_MAX_FRAMES><<<grid_of(n), GEN_BLOCK, 0, s>>>(prog, a, pp, max_records, finder);
  return cudaGetLastError();
}

cudaError_t rkb_generic_tmt(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  DISPATCH(generic_tmt_kernel, host, prog, a);
  return cudaGetLastError();
}
cudaError_t rkb_generic_frame_jac(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, int frame, unsigned upstream, cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  DISPATCH(generic_frame_jac_kernel, host, prog, a, frame, upstream);
  return cudaGetLastError();
}
cudaError_t rkb_generic_rollout(const GenericProgram* prog, const GenericProgram& host, const RolloutArgs& a, const RkTable* table,
                                cudaStream_t s) {
  const long long n = a.n_samples;
  if (n <= 0) return cudaSuccess;
  RkTable none;
  none.stages = 0;
  if (table) DISPATCH2(generic_rollout_kernel, true, host, prog, a, *table);
  else DISPATCH2(generic_rollout_kernel, false, host, prog, a, none);
  return cudaGetLastError();
}
cudaError_t rkb_steer_reduce(int nx, long long n_pairs, long long n_rollouts, const double* xend, const double* goal,
                             int32_t* best_idx, double* best_x, double* best_cost, cudaStream_t s) {
  if (n_pairs <= 0) return cudaSuccess;
  steer_reduce_kernel<<<(unsigned)n_pairs, 256, 0, s>>>(nx, n_rollouts, xend, goal, best_idx, best_x, best_cost);
  return cudaGetLastError();
}
#endif  // RKB_HOST_TEST
