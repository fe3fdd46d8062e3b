/* kte_oracle.c — TEST INFRASTRUCTURE, never part of the product path.
 *
 * Plain-C restatement of the one ReaK path reak_b200 accelerates: evaluate a serial
 * kte_map_chain (doMotion / clearForce / doForce), build the mass matrix the way
 * mass_matrix_calc does (twist-shaping matrix through per-pair relative frames), solve with
 * the reference's Cholesky and integrate with its fixed-step RK4.  It is deliberately literal:
 * every function names the reference file:line it follows (paths relative to
 * /root/reference/src/ReaK) and keeps the reference's operation order, including the places
 * where that differs from the textbook (see "gotcha" notes).  It is a generic interpreter of the
 * flat descriptor in include/reak_b200.h, scalar, single-threaded, built with
 * -ffp-contract=off so that x86-64 evaluates it in plain IEEE double like the reference build.
 *
 * Parity pin: tests/test_oracle.py checks this file (a) against oracle/_ref/libreak_ref.so —
 * the UNMODIFIED reference sources compiled here — on every preset chain, (b) against the
 * golden vectors under tests/golden/ generated from that library (tests/golden/make_golden.py)
 * and (c) against the known answers the reference's own tests/demos hold (pendulum of
 * ctrl/mbd_kte/test_bm.cpp, the 3x3 SPD system of core/lin_alg/unit_test_mat_num.cpp:48-92).
 */
#define _GNU_SOURCE
#include "kte_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * small vector algebra (core/lin_alg/vect_alg.hpp)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double x[3]; } v3;
typedef struct { double x[2]; } v2;
typedef struct { double q[4]; } quat;  /* (w,x,y,z), rotations_3D.hpp:551 */
typedef struct { double q[9]; } rot3;  /* column-major, rotations_3D.hpp:99-109 */
typedef struct { double q[2]; } rot2;  /* (cos, sin), rotations_2D.hpp:89 */

static v3 V3(double a, double b, double c) { v3 r = {{a, b, c}}; return r; }
static v3 add3(v3 a, v3 b) { return V3(a.x[0] + b.x[0], a.x[1] + b.x[1], a.x[2] + b.x[2]); }
static v3 sub3(v3 a, v3 b) { return V3(a.x[0] - b.x[0], a.x[1] - b.x[1], a.x[2] - b.x[2]); }
static v3 neg3(v3 a) { return V3(-a.x[0], -a.x[1], -a.x[2]); }
static v3 scl3(double s, v3 a) { return V3(a.x[0] * s, a.x[1] * s, a.x[2] * s); }
static double dot3(v3 a, v3 b) { return a.x[0] * b.x[0] + a.x[1] * b.x[1] + a.x[2] * b.x[2]; }
/* vect_alg.hpp:1215-1221 */
static v3 cross3(v3 a, v3 b) {
  return V3(a.x[1] * b.x[2] - a.x[2] * b.x[1], a.x[2] * b.x[0] - a.x[0] * b.x[2], a.x[0] * b.x[1] - a.x[1] * b.x[0]);
}
static v2 V2(double a, double b) { v2 r = {{a, b}}; return r; }
static v2 add2(v2 a, v2 b) { return V2(a.x[0] + b.x[0], a.x[1] + b.x[1]); }
static v2 sub2(v2 a, v2 b) { return V2(a.x[0] - b.x[0], a.x[1] - b.x[1]); }
static v2 neg2(v2 a) { return V2(-a.x[0], -a.x[1]); }
static v2 scl2(double s, v2 a) { return V2(a.x[0] * s, a.x[1] * s); }
static double dot2(v2 a, v2 b) { return a.x[0] * b.x[0] + a.x[1] * b.x[1]; }
static double cross22(v2 a, v2 b) { return a.x[0] * b.x[1] - a.x[1] * b.x[0]; } /* vect_alg.hpp:1142 */
static v2 cross_s2(double s, v2 v) { return V2(-v.x[1] * s, v.x[0] * s); }       /* vect_alg.hpp:1171 */

/* ------------------------------------------------------------------------------------------
 * rotations (core/kinetostatics/rotations_3D.hpp, rotations_2D.hpp)
 * ---------------------------------------------------------------------------------------- */
/* rot_mat_3D(a11,a12,a13,a21,...) stores column-major */
static rot3 R3(double a11, double a12, double a13, double a21, double a22, double a23, double a31, double a32, double a33) {
  rot3 r = {{a11, a21, a31, a12, a22, a32, a13, a23, a33}};
  return r;
}
/* rotations_3D.hpp:372-376 : R * V */
static v3 rmul(rot3 R, v3 V) {
  return V3(R.q[0] * V.x[0] + R.q[3] * V.x[1] + R.q[6] * V.x[2],
            R.q[1] * V.x[0] + R.q[4] * V.x[1] + R.q[7] * V.x[2],
            R.q[2] * V.x[0] + R.q[5] * V.x[1] + R.q[8] * V.x[2]);
}
/* rotations_3D.hpp:379-383 : V * R  (= R^T V) */
static v3 vmulr(v3 V, rot3 R) {
  return V3(R.q[0] * V.x[0] + R.q[1] * V.x[1] + R.q[2] * V.x[2],
            R.q[3] * V.x[0] + R.q[4] * V.x[1] + R.q[5] * V.x[2],
            R.q[6] * V.x[0] + R.q[7] * V.x[1] + R.q[8] * V.x[2]);
}
static quat Q4(double w, double x, double y, double z) { quat r = {{w, x, y, z}}; return r; }
/* explicit quaternion(Vector) normalises, rotations_3D.hpp:917-920 */
static quat quat_unit(double w, double x, double y, double z) {
  double n = sqrt(w * w + x * x + y * y + z * z);
  return Q4(w / n, x / n, y / n, z / n);
}
/* quaternion::getRotMat, rotations_3D.hpp:986-1000 */
static rot3 quat_rotmat(quat Q) {
  const double* q = Q.q;
  double t01 = 2.0 * q[0] * q[1], t02 = 2.0 * q[0] * q[2], t03 = 2.0 * q[0] * q[3];
  double t11 = 2.0 * q[1] * q[1], t12 = 2.0 * q[1] * q[2], t13 = 2.0 * q[1] * q[3];
  double t22 = 2.0 * q[2] * q[2], t23 = 2.0 * q[2] * q[3], t33 = 2.0 * q[3] * q[3];
  return R3(1.0 - t22 - t33, t12 - t03, t02 + t13,
            t12 + t03, 1.0 - t11 - t33, t23 - t01,
            t13 - t02, t01 + t23, 1.0 - t11 - t22);
}
/* quaternion product, rotations_3D.hpp:1093-1098 */
static quat qmul(quat A, quat B) {
  const double *a = A.q, *b = B.q;
  return Q4(b[0] * a[0] - b[1] * a[1] - b[2] * a[2] - b[3] * a[3],
            b[0] * a[1] + b[3] * a[2] - b[2] * a[3] + b[1] * a[0],
            b[0] * a[2] - b[3] * a[1] + b[1] * a[3] + b[2] * a[0],
            b[0] * a[3] + b[2] * a[1] - b[1] * a[2] + b[3] * a[0]);
}
static quat qinv(quat A) { return Q4(A.q[0], -A.q[1], -A.q[2], -A.q[3]); } /* rotations_3D.hpp:1280-1282 */
/* quaternion * vect, rotations_3D.hpp:1137-1150 */
static v3 qrot(quat Q, v3 V) {
  const double* q = Q.q;
  double t[9];
  t[0] = q[0] * q[1]; t[1] = q[0] * q[2]; t[2] = q[0] * q[3];
  t[3] = -q[1] * q[1]; t[4] = q[1] * q[2]; t[5] = q[1] * q[3];
  t[6] = -q[2] * q[2]; t[7] = q[2] * q[3]; t[8] = -q[3] * q[3];
  return V3(2.0 * ((t[6] + t[8]) * V.x[0] + (t[4] - t[2]) * V.x[1] + (t[1] + t[5]) * V.x[2]) + V.x[0],
            2.0 * ((t[2] + t[4]) * V.x[0] + (t[3] + t[8]) * V.x[1] + (t[7] - t[0]) * V.x[2]) + V.x[1],
            2.0 * ((t[5] - t[1]) * V.x[0] + (t[0] + t[7]) * V.x[1] + (t[3] + t[6]) * V.x[2]) + V.x[2]);
}
/* axis_angle(angle, axis): the axis is normalised, rotations_3D.hpp:1962-1974 */
static v3 aa_axis(v3 a) {
  double n = sqrt(a.x[0] * a.x[0] + a.x[1] * a.x[1] + a.x[2] * a.x[2]);
  if (n > 0.0000001) return V3(a.x[0] / n, a.x[1] / n, a.x[2] / n);
  return V3(1.0, 0.0, 0.0);
}
/* axis_angle::getQuaternion, rotations_3D.hpp:2107-2115 */
static quat aa_quat(double angle, v3 axis_raw) {
  v3 a = aa_axis(axis_raw);
  double t = sin(0.5 * angle);
  return Q4(cos(0.5 * angle), a.x[0] * t, a.x[1] * t, a.x[2] * t);
}
/* axis_angle::getRotMat, rotations_3D.hpp:2159-2178 */
static rot3 aa_rotmat(double angle, v3 axis_raw) {
  v3 a = aa_axis(axis_raw);
  double ca = cos(angle), omc = 1.0 - ca;
  double t11 = ca + omc * a.x[0] * a.x[0], t22 = ca + omc * a.x[1] * a.x[1], t33 = ca + omc * a.x[2] * a.x[2];
  double t12 = omc * a.x[0] * a.x[1], t13 = omc * a.x[0] * a.x[2], t23 = omc * a.x[1] * a.x[2];
  double sa = sin(angle);
  double t01 = sa * a.x[0], t02 = sa * a.x[1], t03 = sa * a.x[2];
  return R3(t11, t12 - t03, t13 + t02, t12 + t03, t22, t23 - t01, t13 - t02, t23 + t01, t33);
}
/* axis_angle(quaternion), rotations_3D.hpp:1985-2010 */
static void aa_from_quat(quat Q, double* angle, v3* axis) {
  double n = sqrt(Q.q[0] * Q.q[0] + Q.q[1] * Q.q[1] + Q.q[2] * Q.q[2] + Q.q[3] * Q.q[3]);
  double v0 = Q.q[0] / n, v1 = Q.q[1] / n, v2_ = Q.q[2] / n, v3_ = Q.q[3] / n;
  double tmp = sqrt(v1 * v1 + v2_ * v2_ + v3_ * v3_);
  if (tmp > 0.0000001) {
    *axis = V3(v1 / tmp, v2_ / tmp, v3_ / tmp);
    if (v0 < 0.0) { *angle = 2.0 * acos(-v0); *axis = neg3(*axis); }
    else          { *angle = 2.0 * acos(v0); }
  } else {
    *axis = V3(1.0, 0.0, 0.0);
    *angle = 0.0;
  }
}
static rot2 rot2_angle(double a) { rot2 r = {{cos(a), sin(a)}}; return r; }               /* rotations_2D.hpp:108-112 */
static rot2 rot2_mul(rot2 A, rot2 B) {                                                      /* rotations_2D.hpp:264-267 */
  rot2 r = {{A.q[0] * B.q[0] - A.q[1] * B.q[1], A.q[1] * B.q[0] + A.q[0] * B.q[1]}};
  return r;
}
static rot2 rot2_inv(rot2 A) { rot2 r = {{A.q[0], -A.q[1]}}; return r; }
static v2 r2mul(rot2 R, v2 V) { return V2(V.x[0] * R.q[0] - V.x[1] * R.q[1], V.x[0] * R.q[1] + V.x[1] * R.q[0]); } /* :292 */
static v2 v2mulr(v2 V, rot2 R) { return V2(V.x[0] * R.q[0] + V.x[1] * R.q[1], V.x[1] * R.q[0] - V.x[0] * R.q[1]); } /* :300 */

/* ------------------------------------------------------------------------------------------
 * frames (core/kinetostatics/frame_3D.hpp, frame_2D.hpp).  Position/Velocity/Acceleration in
 * parent coordinates, AngVelocity/AngAcceleration/Force/Torque in local coordinates.
 * ---------------------------------------------------------------------------------------- */
typedef struct { v3 p; quat Q; v3 v, w, a, al, F, T; } frame3;
typedef struct { v2 p; rot2 R; v2 v; double w; v2 a; double al; v2 F; double T; } frame2;
typedef struct { double q, qd, qdd, f; } gcoord; /* gen_coord.hpp:44-178 */

/* frame_3D::operator~, frame_3D.hpp:376-388 (forces not needed by the callers here) */
static frame3 f3_inverse(const frame3* f) {
  rot3 R = quat_rotmat(f->Q);
  frame3 r;
  memset(&r, 0, sizeof r);
  r.Q = qinv(f->Q);
  r.w = rmul(R, neg3(f->w));
  r.al = rmul(R, neg3(f->al));
  r.p = vmulr(neg3(f->p), R);
  r.v = vmulr(neg3(add3(cross3(r.w, f->p), f->v)), R);
  /* "value_type(2.0) * result.AngVelocity % Velocity" parses as (2 w) % v */
  r.a = vmulr(neg3(add3(add3(add3(cross3(r.w, cross3(r.w, f->p)), cross3(scl3(2.0, r.w), f->v)), cross3(r.al, f->p)), f->a)), R);
  return r;
}
/* frame_3D::addBefore(const frame_3D&), frame_3D.hpp:219-234 */
static frame3 f3_compose(frame3 s, const frame3* f) {
  rot3 R = quat_rotmat(s.Q);
  s.p = add3(s.p, rmul(R, f->p));
  s.v = add3(s.v, rmul(R, add3(cross3(s.w, f->p), f->v)));
  s.a = add3(s.a, rmul(R, add3(add3(add3(cross3(s.w, cross3(s.w, f->p)), scl3(2.0, cross3(s.w, f->v))), cross3(s.al, f->p)), f->a)));
  rot3 R2 = quat_rotmat(f->Q);
  s.Q = qmul(s.Q, f->Q);
  s.al = add3(add3(vmulr(s.al, R2), cross3(vmulr(s.w, R2), f->w)), f->al);
  s.w = add3(vmulr(s.w, R2), f->w);
  return s;
}
/* frame_3D::addBefore(const pose_3D&), frame_3D.hpp:236-251 */
static frame3 f3_compose_pose(frame3 s, v3 po, quat qo) {
  rot3 R = quat_rotmat(s.Q);
  s.p = add3(s.p, rmul(R, po));
  s.v = add3(s.v, rmul(R, cross3(s.w, po)));
  s.a = add3(s.a, rmul(R, add3(cross3(s.w, cross3(s.w, po)), cross3(s.al, po))));
  rot3 R2 = quat_rotmat(qo);
  s.Q = qmul(s.Q, qo);
  s.al = vmulr(s.al, R2);
  s.w = vmulr(s.w, R2);
  return s;
}
/* frame_2D::operator~, frame_2D.hpp:350-360 */
static frame2 f2_inverse(const frame2* f) {
  frame2 r;
  memset(&r, 0, sizeof r);
  r.p = v2mulr(neg2(f->p), f->R);
  r.R = rot2_inv(f->R);
  r.v = v2mulr(sub2(cross_s2(f->w, f->p), f->v), f->R);
  r.w = -f->w;
  r.a = v2mulr(sub2(add2(add2(scl2(f->w * f->w, f->p), cross_s2(2.0 * f->w, f->v)), cross_s2(f->al, f->p)), f->a), f->R);
  r.al = -f->al;
  return r;
}
/* operator*(frame_2D, frame_2D), frame_2D.hpp:288-300 */
static frame2 f2_compose(const frame2* A, const frame2* B) {
  frame2 r;
  memset(&r, 0, sizeof r);
  r.p = add2(A->p, r2mul(A->R, B->p));
  r.v = add2(A->v, r2mul(A->R, add2(cross_s2(A->w, B->p), B->v)));
  r.a = add2(A->a, r2mul(A->R, add2(add2(add2(scl2(-A->w * A->w, B->p), cross_s2(2.0 * A->w, B->v)), cross_s2(A->al, B->p)), B->a)));
  r.R = rot2_mul(A->R, B->R);
  r.w = A->w + B->w;
  r.al = A->al + B->al;
  return r;
}

/* ------------------------------------------------------------------------------------------
 * the model
 * ---------------------------------------------------------------------------------------- */
/* jacobian_gen_3D / jacobian_gen_2D as filled by the joints (motion_jacobians.hpp:108-330) */
typedef struct { int parent; v3 qd_vel, qd_avel, qd_acc, qd_aacc; } jac3;
typedef struct { int parent; v2 qd_vel; double qd_avel; v2 qd_acc; double qd_aacc; } jac2;

/* jacobian_2D_2D as filled by free_joint_2D::doMotion (motion_jacobians.hpp:411-446) */
typedef struct { int parent; v2 vel_vel[2]; double vel_avel[2]; v2 avel_vel; double avel_avel; v2 vel_acc[2]; double vel_aacc[2]; v2 avel_acc; double avel_aacc; } jac22;

/* jacobian_3D_3D as filled by free_joint_3D::doMotion (motion_jacobians.hpp:1035-1076): eight triples of vectors */
typedef struct { int parent; v3 vel_vel[3], vel_avel[3], avel_vel[3], avel_avel[3], vel_acc[3], vel_aacc[3], avel_acc[3], avel_aacc[3]; } jac33;

#define KTO_MAX_FREE 1
#define KTO_MAX_ACC (RKB_MAX_COORDS + 6 * KTO_MAX_FREE)
#define KTO_MAX_STATE (2 * RKB_MAX_COORDS + 13 * KTO_MAX_FREE)

typedef struct {
  rkb_chain_desc d;
  rkb_element* el;
  frame3* f3;
  frame2* f2;
  gcoord* c;
  jac3* j3;
  jac2* j2;
  double* u; /* inputs[k]->mDriveForce */
  int n, nu, m_rows;
  int n_aux; /* auxiliary gen_coords (RKB_COORD_GEN): c[n .. n + n_aux - 1], never part of the state */
  /* free_joint_3D coordinate frames (kte_nl_system::dofs_3D): 13 states and 6 accelerations each */
  int nfree, nx, na;
  int free_states, free_acc; /* per free joint: 13 / 6 (free_joint_3D), 7 / 3 (free_joint_2D) */
  frame3 fc[KTO_MAX_FREE];
  jac33 jf[KTO_MAX_FREE];
  /* free_joint_2D: coordinate frames (kte_nl_system::dofs_2D), the raw (cos, sin) of the state, jacobian_2D_2D */
  frame2 fc2[KTO_MAX_FREE];
  double fc2_raw[KTO_MAX_FREE][2];
  jac22 jf2[KTO_MAX_FREE];
} model;

static void model_free(model* m) {
  if (!m) return;
  free(m->el); free(m->f3); free(m->f2); free(m->c); free(m->j3); free(m->j2); free(m->u);
  free(m);
}

static int count_rows(const model* m) {
  int rows = 0, e;
  for (e = 0; e < m->d.n_elements; ++e) {
    int k = m->el[e].kind;
    if (k == RKB_INERTIA_GEN) rows += 1;
    else if (k == RKB_INERTIA_2D) rows += 3;
    else if (k == RKB_INERTIA_3D) rows += 6;
  }
  return rows;
}

void* kto_create(const rkb_chain_desc* desc) {
  model* m;
  int e;
  if (!desc || !desc->elements || (desc->dim != 2 && desc->dim != 3)) return NULL;
  if (desc->n_coords < 0 || desc->n_coords > RKB_MAX_COORDS || desc->n_frames < 1) return NULL;
  m = (model*)calloc(1, sizeof(model));
  m->d = *desc;
  m->n = desc->n_coords;
  m->nu = desc->n_inputs;
  m->el = (rkb_element*)malloc(sizeof(rkb_element) * (size_t)(desc->n_elements > 0 ? desc->n_elements : 1));
  memcpy(m->el, desc->elements, sizeof(rkb_element) * (size_t)desc->n_elements);
  m->d.elements = m->el;
  m->f3 = (frame3*)calloc((size_t)desc->n_frames, sizeof(frame3));
  m->f2 = (frame2*)calloc((size_t)desc->n_frames, sizeof(frame2));
  for (e = 0; e < desc->n_elements; ++e) if (m->el[e].kind == RKB_COORD_GEN) m->n_aux += 1;
  if (m->n + m->n_aux > RKB_MAX_COORDS) { model_free(m); return NULL; }
  m->c = (gcoord*)calloc((size_t)(m->n + m->n_aux) + 1, sizeof(gcoord));
  m->j3 = (jac3*)calloc((size_t)m->n + 1, sizeof(jac3));
  m->j2 = (jac2*)calloc((size_t)m->n + 1, sizeof(jac2));
  m->u = (double*)calloc((size_t)m->nu + 1, sizeof(double));
  for (e = 0; e < desc->n_frames; ++e) { m->f3[e].Q = Q4(1, 0, 0, 0); m->f2[e].R.q[0] = 1.0; }
  for (e = 0; e < desc->n_elements; ++e) {
    const rkb_element* E = &m->el[e];
    int bad = 0;
    switch (E->kind) {
      case RKB_REVOLUTE_3D: case RKB_PRISMATIC_3D: case RKB_REVOLUTE_2D: case RKB_PRISMATIC_2D:
        bad = E->coord < 0 || E->coord >= m->n; /* fallthrough to frame checks */
        bad |= E->frame_a < 0 || E->frame_a >= desc->n_frames || E->frame_b < 0 || E->frame_b >= desc->n_frames;
        break;
      case RKB_COORD_GEN: /* an auxiliary gen_coord holding constant values unless a rigid_link_gen writes it */
        bad = E->coord < m->n || E->coord >= m->n + m->n_aux;
        if (!bad) { m->c[E->coord].q = E->p[0]; m->c[E->coord].qd = E->p[1]; m->c[E->coord].qdd = E->p[2]; }
        break;
      case RKB_RIGID_LINK_GEN:
        bad = E->coord < 0 || E->coord >= m->n + m->n_aux || E->aux < m->n || E->aux >= m->n + m->n_aux;
        break;
      case RKB_SPRING_GEN: case RKB_DAMPER_GEN:
        bad = E->coord < 0 || E->coord >= m->n + m->n_aux || E->aux < 0 || E->aux >= m->n + m->n_aux;
        break;
      case RKB_FREE_2D:
        bad = desc->dim != 2 || E->coord != m->nfree || m->nfree >= KTO_MAX_FREE;
        bad |= E->frame_a < 0 || E->frame_a >= desc->n_frames || E->frame_b < 0 || E->frame_b >= desc->n_frames;
        if (!bad) m->nfree += 1;
        break;
      case RKB_FREE_3D:
        bad = desc->dim != 3 || E->coord != m->nfree || m->nfree >= KTO_MAX_FREE;
        bad |= E->frame_a < 0 || E->frame_a >= desc->n_frames || E->frame_b < 0 || E->frame_b >= desc->n_frames;
        if (!bad) m->nfree += 1;
        break;
      case RKB_RIGID_LINK_3D: case RKB_RIGID_LINK_2D:
      case RKB_TORSION_SPRING_3D: case RKB_TORSION_DAMPER_3D: case RKB_SPRING_3D: case RKB_DAMPER_3D:
      case RKB_TORSION_SPRING_2D: case RKB_TORSION_DAMPER_2D: case RKB_SPRING_2D: case RKB_DAMPER_2D:
        bad = E->frame_a < 0 || E->frame_a >= desc->n_frames || E->frame_b < 0 || E->frame_b >= desc->n_frames;
        break;
      case RKB_INERTIA_3D: case RKB_INERTIA_2D:
        bad = E->frame_a < 0 || E->frame_a >= desc->n_frames;
        break;
      case RKB_INERTIA_GEN:
        bad = E->coord < 0 || E->coord >= m->n;
        break;
      case RKB_ACTUATOR_GEN:
        bad = E->coord < 0 || E->coord >= m->n || E->aux < 0 || E->aux >= m->nu ||
              E->frame_b < 0 || E->frame_b >= desc->n_elements;
        break;
      default: bad = 1;
    }
    if (bad) { model_free(m); return NULL; }
  }
  /* base frame (robot_base of CRS_A465_models.cpp:298-301 / base_frame of test_bm.cpp:50-52) */
  {
    const rkb_base_frame* b = &desc->base;
    if (desc->dim == 3) {
      frame3* B = &m->f3[desc->base_frame];
      B->p = V3(b->position[0], b->position[1], b->position[2]);
      B->Q = quat_unit(b->quat[0], b->quat[1], b->quat[2], b->quat[3]);
      B->v = V3(b->velocity[0], b->velocity[1], b->velocity[2]);
      B->w = V3(b->ang_velocity[0], b->ang_velocity[1], b->ang_velocity[2]);
      B->a = V3(b->acceleration[0], b->acceleration[1], b->acceleration[2]);
      B->al = V3(b->ang_acceleration[0], b->ang_acceleration[1], b->ang_acceleration[2]);
    } else {
      frame2* B = &m->f2[desc->base_frame];
      B->p = V2(b->position[0], b->position[1]);
      B->R = rot2_angle(b->quat[0]);
      B->v = V2(b->velocity[0], b->velocity[1]);
      B->w = b->ang_velocity[0];
      B->a = V2(b->acceleration[0], b->acceleration[1]);
      B->al = b->ang_acceleration[0];
    }
  }
  m->m_rows = count_rows(m);
  m->free_states = desc->dim == 3 ? 13 : 7;
  m->free_acc = desc->dim == 3 ? 6 : 3;
  m->nx = 2 * m->n + m->free_states * m->nfree; /* kte_nl_system.hpp:145-147 */
  m->na = m->n + m->free_acc * m->nfree;
  for (e = 0; e < KTO_MAX_FREE; ++e) { m->fc[e].Q = Q4(1, 0, 0, 0); m->fc2[e].R.q[0] = 1.0; }
  return m;
}

void kto_destroy(void* h) { model_free((model*)h); }

/* kte_nl_system::apply_states_and_inputs, ctrl/ctrl_sys/kte_nl_system.hpp:180-225 */
static void apply_states_and_inputs(model* m, const double* p, const double* u) {
  int j;
  for (j = 0; j < m->n; ++j) {
    m->c[j].q = p[2 * j];
    m->c[j].qd = p[2 * j + 1];
    m->c[j].qdd = 0.0;
  }
  for (j = 0; j < m->nfree && m->d.dim == 2; ++j) { /* :194-204: rotation_type(vect<2>) normalises (rotations_2D.hpp:119-123) */
    const double* s = p + 2 * m->n + 7 * j;
    frame2* F = &m->fc2[j];
    double nrm = sqrt(s[2] * s[2] + s[3] * s[3]);
    F->p = V2(s[0], s[1]);
    F->R.q[0] = s[2] / nrm; F->R.q[1] = s[3] / nrm;
    m->fc2_raw[j][0] = s[2]; m->fc2_raw[j][1] = s[3];
    F->v = V2(s[4], s[5]);
    F->w = s[6];
    F->a = V2(0, 0);
    F->al = 0.0;
  }
  for (j = 0; j < m->nfree && m->d.dim == 3; ++j) { /* :205-219: the quaternion is normalised by quaternion(vect<4>) */
    const double* s = p + 2 * m->n + 13 * j;
    frame3* F = &m->fc[j];
    F->p = V3(s[0], s[1], s[2]);
    F->Q = quat_unit(s[3], s[4], s[5], s[6]);
    F->v = V3(s[7], s[8], s[9]);
    F->w = V3(s[10], s[11], s[12]);
    F->a = V3(0, 0, 0);
    F->al = V3(0, 0, 0);
  }
  for (j = 0; j < m->nu; ++j) m->u[j] = u ? u[j] : 0.0;
}

/* kte_map_chain::doMotion, ctrl/mbd_kte/kte_map_chain.hpp:71-76 */
static void do_motion(model* m) {
  int e;
  for (e = 0; e < m->d.n_elements; ++e) {
    const rkb_element* E = &m->el[e];
    switch (E->kind) {
      case RKB_REVOLUTE_3D: { /* revolute_joint.cpp:121-152 */
        const frame3* B = &m->f3[E->frame_a];
        frame3* N = &m->f3[E->frame_b];
        const gcoord* c = &m->c[E->coord];
        v3 axis = V3(E->p[0], E->p[1], E->p[2]);
        quat tq = aa_quat(c->q, axis);
        rot3 R2 = quat_rotmat(tq);
        v3 wB = B->w, alB = B->al;
        N->p = B->p; N->v = B->v; N->a = B->a;
        N->Q = qmul(B->Q, tq);
        N->w = add3(vmulr(wB, R2), scl3(c->qd, axis));
        N->al = add3(add3(vmulr(alB, R2), cross3(vmulr(wB, R2), scl3(c->qd, axis))), scl3(c->qdd, axis));
        m->j3[E->coord].parent = E->frame_b;
        m->j3[E->coord].qd_vel = V3(0, 0, 0);
        m->j3[E->coord].qd_avel = axis;
        m->j3[E->coord].qd_acc = V3(0, 0, 0);
        m->j3[E->coord].qd_aacc = V3(0, 0, 0);
        break;
      }
      case RKB_PRISMATIC_3D: { /* prismatic_joint.cpp:129-161 */
        const frame3* B = &m->f3[E->frame_a];
        frame3* N = &m->f3[E->frame_b];
        const gcoord* c = &m->c[E->coord];
        v3 axis = V3(E->p[0], E->p[1], E->p[2]);
        rot3 R = quat_rotmat(B->Q);
        v3 tp = scl3(c->q, axis), tv = scl3(c->qd, axis);
        v3 np = add3(B->p, rmul(R, tp));
        v3 nv = add3(B->v, rmul(R, add3(cross3(B->w, tp), tv)));
        v3 na = add3(B->a, rmul(R, add3(add3(add3(cross3(B->w, cross3(B->w, tp)), scl3(2.0, cross3(B->w, tv))),
                                             cross3(B->al, tp)), scl3(c->qdd, axis))));
        N->p = np; N->v = nv; N->a = na;
        N->Q = B->Q; N->w = B->w; N->al = B->al;
        m->j3[E->coord].parent = E->frame_b;
        m->j3[E->coord].qd_vel = axis;
        m->j3[E->coord].qd_avel = V3(0, 0, 0);
        m->j3[E->coord].qd_acc = V3(0, 0, 0);
        m->j3[E->coord].qd_aacc = V3(0, 0, 0);
        break;
      }
      case RKB_RIGID_LINK_GEN: { /* rigid_link.cpp:34-40 */
        const gcoord* B = &m->c[E->coord];
        gcoord* N = &m->c[E->aux];
        N->q = B->q + E->p[0];
        N->qd = B->qd;
        N->qdd = B->qdd;
        break;
      }
      case RKB_FREE_3D: { /* free_joints.cpp:123-146: *mEnd = (*mBase) * (*mCoord); the Jacobian is two identity blocks */
        frame3 r = f3_compose(m->f3[E->frame_a], &m->fc[E->coord]);
        frame3* N = &m->f3[E->frame_b];
        jac33* J = &m->jf[E->coord];
        int k;
        N->p = r.p; N->Q = r.Q; N->v = r.v; N->w = r.w; N->a = r.a; N->al = r.al; /* frame_3D.hpp:296-308 */
        memset(J, 0, sizeof *J);
        J->parent = E->frame_b;
        for (k = 0; k < 3; ++k) { J->vel_vel[k].x[k] = 1.0; J->avel_avel[k].x[k] = 1.0; }
        break;
      }
      case RKB_RIGID_LINK_3D: { /* rigid_link.cpp:152-156: *mEnd = *mBase * mPoseOffset */
        frame3 r = f3_compose_pose(m->f3[E->frame_a], V3(E->p[0], E->p[1], E->p[2]),
                                   quat_unit(E->p[3], E->p[4], E->p[5], E->p[6]));
        frame3* N = &m->f3[E->frame_b];
        N->p = r.p; N->Q = r.Q; N->v = r.v; N->w = r.w; N->a = r.a; N->al = r.al; /* frame_3D.hpp:296-308 */
        break;
      }
      case RKB_FREE_2D: { /* free_joints.cpp:33-56: *mEnd = (*mBase) * (*mCoord); the Jacobian is the identity */
        frame2 r = f2_compose(&m->f2[E->frame_a], &m->fc2[E->coord]);
        frame2* N = &m->f2[E->frame_b];
        jac22* J = &m->jf2[E->coord];
        N->p = r.p; N->R = r.R; N->v = r.v; N->w = r.w; N->a = r.a; N->al = r.al;
        memset(J, 0, sizeof *J);
        J->parent = E->frame_b;
        J->vel_vel[0] = V2(1, 0); J->vel_vel[1] = V2(0, 1); J->avel_avel = 1.0;
        break;
      }
      case RKB_REVOLUTE_2D: { /* revolute_joint.cpp:32-58 */
        const frame2* B = &m->f2[E->frame_a];
        frame2* N = &m->f2[E->frame_b];
        const gcoord* c = &m->c[E->coord];
        N->p = B->p; N->v = B->v; N->a = B->a;
        N->R = rot2_mul(B->R, rot2_angle(c->q));
        N->w = B->w + c->qd;
        N->al = B->al + c->qdd;
        m->j2[E->coord].parent = E->frame_b;
        m->j2[E->coord].qd_vel = V2(0, 0);
        m->j2[E->coord].qd_avel = 1.0;
        m->j2[E->coord].qd_acc = V2(0, 0);
        m->j2[E->coord].qd_aacc = 0.0;
        break;
      }
      case RKB_PRISMATIC_2D: { /* prismatic_joint.cpp:33-67 */
        const frame2* B = &m->f2[E->frame_a];
        frame2* N = &m->f2[E->frame_b];
        const gcoord* c = &m->c[E->coord];
        v2 axis = V2(E->p[0], E->p[1]);
        v2 tp = scl2(c->q, axis), tv = scl2(c->qd, axis);
        v2 np = add2(B->p, r2mul(B->R, tp));
        v2 nv = add2(B->v, r2mul(B->R, add2(cross_s2(B->w, tp), tv)));
        v2 na = add2(B->a, r2mul(B->R, add2(add2(add2(scl2(-B->w * B->w, tp), cross_s2(2.0 * B->w, tv)),
                                                 cross_s2(B->al, tp)), scl2(c->qdd, axis))));
        N->p = np; N->v = nv; N->a = na;
        N->R = B->R; N->w = B->w; N->al = B->al;
        m->j2[E->coord].parent = E->frame_b;
        m->j2[E->coord].qd_vel = axis;
        m->j2[E->coord].qd_avel = 0.0;
        m->j2[E->coord].qd_acc = V2(0, 0);
        m->j2[E->coord].qd_aacc = 0.0;
        break;
      }
      case RKB_RIGID_LINK_2D: { /* rigid_link.cpp:87-99 */
        const frame2* B = &m->f2[E->frame_a];
        frame2* N = &m->f2[E->frame_b];
        v2 po = V2(E->p[0], E->p[1]);
        v2 np = add2(B->p, r2mul(B->R, po));
        v2 nv = add2(B->v, r2mul(B->R, cross_s2(B->w, po)));
        v2 na = add2(B->a, r2mul(B->R, add2(scl2(-B->w * B->w, po), cross_s2(B->al, po))));
        N->R = rot2_mul(B->R, rot2_angle(E->p[2]));
        N->p = np; N->v = nv; N->a = na;
        N->w = B->w; N->al = B->al;
        break;
      }
      default: break; /* inertias, springs, dampers, actuators: doMotion does nothing */
    }
  }
}

/* kte_map_chain::clearForce, kte_map_chain.hpp:85-89.  Every frame and coordinate of the
 * descriptor is touched by at least one element's clearForce, so "zero everything" is the same. */
static void clear_force(model* m) {
  int i;
  for (i = 0; i < m->d.n_frames; ++i) {
    m->f3[i].F = V3(0, 0, 0); m->f3[i].T = V3(0, 0, 0);
    m->f2[i].F = V2(0, 0); m->f2[i].T = 0.0;
  }
  for (i = 0; i < m->n + m->n_aux; ++i) m->c[i].f = 0.0;
  for (i = 0; i < m->nfree; ++i) { m->fc[i].F = V3(0, 0, 0); m->fc[i].T = V3(0, 0, 0); } /* free_joints.cpp:184-197 */
  for (i = 0; i < m->nfree; ++i) { m->fc2[i].F = V2(0, 0); m->fc2[i].T = 0.0; }                  /* free_joints.cpp:93-106 */
}

/* kte_map_chain::doForce, kte_map_chain.hpp:78-83 (REVERSE order) */
static void do_force(model* m) {
  int e;
  for (e = m->d.n_elements - 1; e >= 0; --e) {
    const rkb_element* E = &m->el[e];
    switch (E->kind) {
      case RKB_REVOLUTE_3D: { /* revolute_joint.cpp:172-184 */
        frame3* B = &m->f3[E->frame_a];
        const frame3* N = &m->f3[E->frame_b];
        gcoord* c = &m->c[E->coord];
        v3 axis = V3(E->p[0], E->p[1], E->p[2]);
        rot3 R = aa_rotmat(c->q, axis);
        B->F = add3(B->F, rmul(R, N->F));
        c->f += dot3(N->T, axis);
        B->T = add3(B->T, rmul(R, sub3(N->T, scl3(dot3(N->T, axis), axis))));
        break;
      }
      case RKB_PRISMATIC_3D: { /* prismatic_joint.cpp:181-193 */
        frame3* B = &m->f3[E->frame_a];
        const frame3* N = &m->f3[E->frame_b];
        gcoord* c = &m->c[E->coord];
        v3 axis = V3(E->p[0], E->p[1], E->p[2]);
        double tf = dot3(N->F, axis);
        c->f += tf;
        B->F = add3(B->F, sub3(N->F, scl3(tf, axis)));
        B->T = add3(B->T, add3(N->T, cross3(scl3(c->q, axis), N->F)));
        break;
      }
      case RKB_RIGID_LINK_GEN: /* rigid_link.cpp:53-57 */
        m->c[E->coord].f += m->c[E->aux].f;
        break;
      case RKB_SPRING_GEN: { /* spring.cpp:50-84 */
        gcoord* A1 = &m->c[E->coord];
        gcoord* A2 = &m->c[E->aux];
        double rest = E->p[0], k = E->p[1], sat = E->p[2];
        if (A1->q > A2->q) {
          double fm = (A1->q - A2->q - rest) * k;
          if (sat > 0 && fabs(fm) > sat) {
            if (fm > 0) { A1->f -= sat; A2->f += sat; } else { A1->f += sat; A2->f -= sat; }
          } else { A1->f -= fm; A2->f += fm; }
        } else {
          double fm = (A2->q - A1->q - rest) * k;
          if (sat > 0 && fabs(fm) > sat) {
            if (fm > 0) { A1->f += sat; A2->f -= sat; } else { A1->f -= sat; A2->f += sat; }
          } else { A1->f += fm; A2->f -= fm; }
        }
        break;
      }
      case RKB_DAMPER_GEN: { /* damper.cpp:48-57 */
        gcoord* A1 = &m->c[E->coord];
        gcoord* A2 = &m->c[E->aux];
        double fm = (A1->qd - A2->qd) * E->p[0];
        A1->f -= fm;
        A2->f += fm;
        break;
      }
      case RKB_FREE_3D: { /* free_joints.cpp:164-172: the wrench at the end frame lands on the coordinate frame */
        frame3* C = &m->fc[E->coord];
        const frame3* N = &m->f3[E->frame_b];
        C->F = add3(C->F, N->F);
        C->T = add3(C->T, N->T);
        break;
      }
      case RKB_RIGID_LINK_3D: { /* rigid_link.cpp:170-177 */
        frame3* B = &m->f3[E->frame_a];
        const frame3* N = &m->f3[E->frame_b];
        rot3 R = quat_rotmat(quat_unit(E->p[3], E->p[4], E->p[5], E->p[6]));
        v3 tf = rmul(R, N->F);
        B->F = add3(B->F, tf);
        B->T = add3(B->T, add3(rmul(R, N->T), cross3(V3(E->p[0], E->p[1], E->p[2]), tf)));
        break;
      }
      case RKB_INERTIA_3D: { /* inertia.cpp:111-121; getGlobalFrame() is the frame itself (Parent expired) */
        frame3* G = &m->f3[E->frame_a];
        const double* I = &E->p[1]; /* Ixx Ixy Ixz Iyy Iyz Izz */
        v3 Ial = V3(I[0] * G->al.x[0] + I[1] * G->al.x[1] + I[2] * G->al.x[2],
                    I[1] * G->al.x[0] + I[3] * G->al.x[1] + I[4] * G->al.x[2],
                    I[2] * G->al.x[0] + I[4] * G->al.x[1] + I[5] * G->al.x[2]);
        v3 Iw = V3(I[0] * G->w.x[0] + I[1] * G->w.x[1] + I[2] * G->w.x[2],
                   I[1] * G->w.x[0] + I[3] * G->w.x[1] + I[4] * G->w.x[2],
                   I[2] * G->w.x[0] + I[4] * G->w.x[1] + I[5] * G->w.x[2]);
        G->F = sub3(G->F, scl3(E->p[0], qrot(qinv(G->Q), G->a)));
        G->T = sub3(G->T, add3(Ial, cross3(G->w, Iw)));
        break;
      }
      case RKB_INERTIA_GEN: /* inertia.cpp:47-53 */
        m->c[E->coord].f -= m->c[E->coord].qdd * E->p[0];
        break;
      case RKB_ACTUATOR_GEN: { /* driving_actuator.cpp:31-38 + applyReactionForce of the joint */
        const rkb_element* J = &m->el[E->frame_b];
        double drive = m->u[E->aux];
        m->c[E->coord].f += drive;
        if (J->kind == RKB_REVOLUTE_3D) {        /* revolute_joint.cpp:210-213 */
          frame3* B = &m->f3[J->frame_a];
          B->T = sub3(B->T, scl3(drive, V3(J->p[0], J->p[1], J->p[2])));
        } else if (J->kind == RKB_PRISMATIC_3D) { /* prismatic_joint.cpp:219-222 */
          frame3* B = &m->f3[J->frame_a];
          B->F = sub3(B->F, scl3(drive, V3(J->p[0], J->p[1], J->p[2])));
        } else if (J->kind == RKB_REVOLUTE_2D) {  /* revolute_joint.cpp:113-116 */
          m->f2[J->frame_a].T -= drive;
        } else if (J->kind == RKB_PRISMATIC_2D) { /* prismatic_joint.cpp:120-123 */
          frame2* B = &m->f2[J->frame_a];
          B->F = sub2(B->F, scl2(drive, V2(J->p[0], J->p[1])));
        }
        break;
      }
      case RKB_TORSION_SPRING_3D: { /* torsion_spring.cpp:106-129 */
        frame3* A1 = &m->f3[E->frame_a];
        frame3* A2 = &m->f3[E->frame_b];
        double angle, k = E->p[0], sat = E->p[1], mag;
        v3 ax;
        aa_from_quat(qmul(qinv(A1->Q), A2->Q), &angle, &ax);
        mag = k * angle;
        if (sat > 0 && fabs(mag) > sat) {
          if (mag > 0) { A1->T = add3(A1->T, scl3(sat, ax)); A2->T = sub3(A2->T, scl3(sat, ax)); }
          else         { A1->T = sub3(A1->T, scl3(sat, ax)); A2->T = add3(A2->T, scl3(sat, ax)); }
        } else {
          A1->T = add3(A1->T, scl3(k * angle, ax));
          A2->T = sub3(A2->T, scl3(k * angle, ax));
        }
        break;
      }
      case RKB_TORSION_DAMPER_3D: { /* torsion_damper.cpp:93-104 */
        frame3* A1 = &m->f3[E->frame_a];
        frame3* A2 = &m->f3[E->frame_b];
        rot3 R1 = quat_rotmat(A1->Q), R2 = quat_rotmat(A2->Q);
        v3 diff = scl3(E->p[0], sub3(rmul(R1, A1->w), rmul(R2, A2->w)));
        A1->T = sub3(A1->T, vmulr(diff, R1));
        A2->T = add3(A2->T, vmulr(diff, R2));
        break;
      }
      case RKB_SPRING_3D: { /* spring.cpp:178-207 */
        frame3* A1 = &m->f3[E->frame_a];
        frame3* A2 = &m->f3[E->frame_b];
        double rest = E->p[0], k = E->p[1], sat = E->p[2];
        v3 diff = sub3(A1->p, A2->p);
        double mag = sqrt(dot3(diff, diff));
        if (mag > 1E-7) {
          double fm = (mag - rest) * k;
          if (sat > 0 && fabs(fm) > sat) {
            diff = scl3(sat / mag, diff);
            if (fm > 0) { A1->F = sub3(A1->F, qrot(qinv(A1->Q), diff)); A2->F = add3(A2->F, qrot(qinv(A2->Q), diff)); }
            else        { A1->F = add3(A1->F, qrot(qinv(A1->Q), diff)); A2->F = sub3(A2->F, qrot(qinv(A2->Q), diff)); }
          } else {
            diff = scl3(fm / mag, diff);
            A1->F = sub3(A1->F, qrot(qinv(A1->Q), diff));
            A2->F = add3(A2->F, qrot(qinv(A2->Q), diff));
          }
        }
        break;
      }
      case RKB_DAMPER_3D: { /* damper.cpp:136-149 */
        frame3* A1 = &m->f3[E->frame_a];
        frame3* A2 = &m->f3[E->frame_b];
        v3 diff = sub3(A1->p, A2->p);
        double sq = dot3(diff, diff);
        if (sq > 1E-7) {
          diff = scl3(dot3(sub3(A1->v, A2->v), diff) * E->p[0] / sq, diff);
          A1->F = sub3(A1->F, qrot(qinv(A1->Q), diff));
          A2->F = add3(A2->F, qrot(qinv(A2->Q), diff));
        }
        break;
      }
      case RKB_FREE_2D: { /* free_joints.cpp:75-85 */
        frame2* C = &m->fc2[E->coord];
        const frame2* N = &m->f2[E->frame_b];
        C->F = add2(C->F, N->F);
        C->T += N->T;
        break;
      }
      case RKB_REVOLUTE_2D: { /* revolute_joint.cpp:78-90; the torque is NOT passed to the base */
        frame2* B = &m->f2[E->frame_a];
        const frame2* N = &m->f2[E->frame_b];
        gcoord* c = &m->c[E->coord];
        B->F = add2(B->F, r2mul(rot2_angle(c->q), N->F));
        c->f += N->T;
        break;
      }
      case RKB_PRISMATIC_2D: { /* prismatic_joint.cpp:83-95 */
        frame2* B = &m->f2[E->frame_a];
        const frame2* N = &m->f2[E->frame_b];
        gcoord* c = &m->c[E->coord];
        v2 axis = V2(E->p[0], E->p[1]);
        double tf = dot2(N->F, axis);
        c->f += tf;
        B->F = add2(B->F, sub2(N->F, scl2(tf, axis)));
        B->T += N->T + cross22(scl2(c->q, axis), N->F);
        break;
      }
      case RKB_RIGID_LINK_2D: { /* rigid_link.cpp:117-125 */
        frame2* B = &m->f2[E->frame_a];
        const frame2* N = &m->f2[E->frame_b];
        v2 tf = r2mul(rot2_angle(E->p[2]), N->F);
        B->F = add2(B->F, tf);
        B->T += N->T + cross22(V2(E->p[0], E->p[1]), tf);
        break;
      }
      case RKB_INERTIA_2D: { /* inertia.cpp:77-86 */
        frame2* G = &m->f2[E->frame_a];
        G->F = sub2(G->F, scl2(E->p[0], v2mulr(G->a, G->R)));
        G->T -= E->p[1] * G->al;
        break;
      }
      case RKB_TORSION_SPRING_2D: { /* torsion_spring.cpp:50-71 */
        frame2* A1 = &m->f2[E->frame_a];
        frame2* A2 = &m->f2[E->frame_b];
        rot2 rel = rot2_mul(rot2_inv(A1->R), A2->R);
        double ad = atan2(rel.q[1], rel.q[0]) * E->p[0], sat = E->p[1];
        if (sat > 0 && fabs(ad) > sat) {
          if (ad > 0) { A1->T += sat; A2->T -= sat; }
          else        { A1->T -= sat; A2->T += sat; }
        } else { A1->T += ad; A2->T -= ad; }
        break;
      }
      case RKB_TORSION_DAMPER_2D: { /* torsion_damper.cpp:49-58 */
        frame2* A1 = &m->f2[E->frame_a];
        frame2* A2 = &m->f2[E->frame_b];
        double tm = (A1->w - A2->w) * E->p[0];
        A1->T -= tm; A2->T += tm;
        break;
      }
      case RKB_SPRING_2D: { /* spring.cpp:116-143 */
        frame2* A1 = &m->f2[E->frame_a];
        frame2* A2 = &m->f2[E->frame_b];
        double rest = E->p[0], k = E->p[1], sat = E->p[2];
        v2 diff = sub2(A1->p, A2->p);
        double mag = sqrt(dot2(diff, diff));
        if (mag > 1E-7) {
          double fm = (mag - rest) * k;
          if (sat > 0 && fabs(fm) > sat) {
            diff = scl2(sat / mag, diff);
            if (fm > 0) { A1->F = sub2(A1->F, v2mulr(diff, A1->R)); A2->F = add2(A2->F, v2mulr(diff, A2->R)); }
            else        { A1->F = add2(A1->F, v2mulr(diff, A1->R)); A2->F = sub2(A2->F, v2mulr(diff, A2->R)); }
          } else {
            diff = scl2(fm / mag, diff);
            A1->F = sub2(A1->F, v2mulr(diff, A1->R));
            A2->F = add2(A2->F, v2mulr(diff, A2->R));
          }
        }
        break;
      }
      case RKB_DAMPER_2D: { /* damper.cpp:88-102 */
        frame2* A1 = &m->f2[E->frame_a];
        frame2* A2 = &m->f2[E->frame_b];
        v2 diff = sub2(A1->p, A2->p);
        double sq = dot2(diff, diff);
        if (sq > 1E-7) {
          diff = scl2(dot2(sub2(A1->v, A2->v), diff) * E->p[0] / sq, diff);
          A1->F = sub2(A1->F, v2mulr(diff, A1->R));
          A2->F = add2(A2->F, v2mulr(diff, A2->R));
        }
        break;
      }
      default: break;
    }
  }
}

/* jacobian_gen_3D::get_jac_relative_to + write_to_matrices, motion_jacobians.hpp:238-279, with
 * frame_3D::getFrameRelativeTo taking its first branch (frame_3D.hpp:183-188): every chain
 * frame shares the base's expired Parent, so f2 = (~E) * F. */
static void jac3_rel(const model* m, const jac3* J, int frame, double* col, double* coldot) {
  frame3 inv = f3_inverse(&m->f3[J->parent]);
  frame3 f2 = f3_compose(inv, &m->f3[frame]);
  rot3 R = quat_rotmat(f2.Q);
  v3 w_tmp = vmulr(J->qd_avel, R);
  v3 v_tmp = vmulr(add3(cross3(J->qd_avel, f2.p), J->qd_vel), R);
  v3 acc = sub3(vmulr(add3(add3(cross3(J->qd_avel, f2.v), cross3(J->qd_aacc, f2.p)), J->qd_acc), R), cross3(f2.w, v_tmp));
  v3 aacc = sub3(vmulr(J->qd_aacc, R), cross3(f2.w, w_tmp));
  int k;
  for (k = 0; k < 3; ++k) { col[k] = v_tmp.x[k]; col[3 + k] = w_tmp.x[k]; coldot[k] = acc.x[k]; coldot[3 + k] = aacc.x[k]; }
}
/* jacobian_gen_2D::get_jac_relative_to, motion_jacobians.hpp:139-147 */
static void jac2_rel(const model* m, const jac2* J, int frame, double* col, double* coldot) {
  frame2 inv = f2_inverse(&m->f2[J->parent]);
  frame2 f2 = f2_compose(&inv, &m->f2[frame]);
  v2 v_tmp = v2mulr(add2(cross_s2(J->qd_avel, f2.p), J->qd_vel), f2.R);
  v2 acc = sub2(v2mulr(add2(add2(cross_s2(J->qd_avel, f2.v), cross_s2(J->qd_aacc, f2.p)), J->qd_acc), f2.R), cross_s2(f2.w, v_tmp));
  col[0] = v_tmp.x[0]; col[1] = v_tmp.x[1]; col[2] = J->qd_avel;
  coldot[0] = acc.x[0]; coldot[1] = acc.x[1]; coldot[2] = J->qd_aacc;
}

/* jacobian_3D_3D::get_jac_relative_to + write_to_matrices, motion_jacobians.hpp:1077-1203: a 6 x 6 block, column k
 * (k < 3) from vel_*[k], column 3 + k from avel_*[k]; rows v (3) then w (3).  blk / blkdot: row-major 6 x 6. */
static void jac33_rel(const model* m, const jac33* J, int frame, double* blk, double* blkdot) {
  frame3 inv = f3_inverse(&m->f3[J->parent]);
  frame3 f2 = f3_compose(inv, &m->f3[frame]);
  rot3 R = quat_rotmat(f2.Q);
  int k, r;
  for (k = 0; k < 3; ++k) {
    v3 n_vel_avel = vmulr(J->vel_avel[k], R);
    v3 n_vel_aacc = sub3(vmulr(J->vel_aacc[k], R), cross3(f2.w, n_vel_avel));
    v3 n_avel_avel = vmulr(J->avel_avel[k], R);
    v3 n_avel_aacc = sub3(vmulr(J->avel_aacc[k], R), cross3(f2.w, n_avel_avel));
    v3 n_vel_vel = vmulr(add3(cross3(J->vel_avel[k], f2.p), J->vel_vel[k]), R);
    v3 n_vel_acc = sub3(vmulr(add3(add3(cross3(J->vel_avel[k], f2.v), cross3(J->vel_aacc[k], f2.p)), J->vel_acc[k]), R),
                        cross3(f2.w, n_vel_vel));
    v3 n_avel_vel = vmulr(add3(cross3(J->avel_avel[k], f2.p), J->avel_vel[k]), R);
    v3 n_avel_acc = sub3(vmulr(add3(add3(cross3(J->avel_avel[k], f2.v), cross3(J->avel_aacc[k], f2.p)), J->avel_acc[k]), R),
                         cross3(f2.w, n_avel_vel));
    for (r = 0; r < 3; ++r) {
      blk[r * 6 + k] = n_vel_vel.x[r];        blk[(3 + r) * 6 + k] = n_vel_avel.x[r];
      blk[r * 6 + 3 + k] = n_avel_vel.x[r];   blk[(3 + r) * 6 + 3 + k] = n_avel_avel.x[r];
      blkdot[r * 6 + k] = n_vel_acc.x[r];     blkdot[(3 + r) * 6 + k] = n_vel_aacc.x[r];
      blkdot[r * 6 + 3 + k] = n_avel_acc.x[r]; blkdot[(3 + r) * 6 + 3 + k] = n_avel_aacc.x[r];
    }
  }
}

/* jacobian_2D_2D::get_jac_relative_to + write_to_matrices, motion_jacobians.hpp:448-502: a 3 x 3 block, column k (k < 2)
 * from vel_*[k], column 2 from avel_*; rows v (2) then w.  blk / blkdot: row-major 3 x 3. */
static void jac22_rel(const model* m, const jac22* J, int frame, double* blk, double* blkdot) {
  frame2 inv = f2_inverse(&m->f2[J->parent]);
  frame2 f2 = f2_compose(&inv, &m->f2[frame]);
  v2 n_avel_vel = v2mulr(add2(cross_s2(J->avel_avel, f2.p), J->avel_vel), f2.R);
  v2 n_avel_acc = sub2(v2mulr(add2(add2(cross_s2(J->avel_avel, f2.v), cross_s2(J->avel_aacc, f2.p)), J->avel_acc), f2.R),
                       cross_s2(f2.w, n_avel_vel));
  int k;
  for (k = 0; k < 2; ++k) {
    v2 n_vel_vel = v2mulr(add2(cross_s2(J->vel_avel[k], f2.p), J->vel_vel[k]), f2.R);
    v2 n_vel_acc = sub2(v2mulr(add2(add2(cross_s2(J->vel_avel[k], f2.v), cross_s2(J->vel_aacc[k], f2.p)), J->vel_acc[k]), f2.R),
                        cross_s2(f2.w, n_vel_vel));
    blk[0 * 3 + k] = n_vel_vel.x[0]; blk[1 * 3 + k] = n_vel_vel.x[1]; blk[2 * 3 + k] = J->vel_avel[k];
    blkdot[0 * 3 + k] = n_vel_acc.x[0]; blkdot[1 * 3 + k] = n_vel_acc.x[1]; blkdot[2 * 3 + k] = J->vel_aacc[k];
  }
  blk[0 * 3 + 2] = n_avel_vel.x[0]; blk[1 * 3 + 2] = n_avel_vel.x[1]; blk[2 * 3 + 2] = J->avel_avel;
  blkdot[0 * 3 + 2] = n_avel_acc.x[0]; blkdot[1 * 3 + 2] = n_avel_acc.x[1]; blkdot[2 * 3 + 2] = J->avel_aacc;
}

/* mass_matrix_calc::get_TMT_TdMT, ctrl/mbd_kte/mass_matrix_calculator.cpp:100-287.
 * Rows: gen inertias, then 2D inertias (vx, vy, w), then 3D inertias (v3, w3), each group in
 * registration (= chain) order; columns: the coordinates.  T, Td are rows x n, Mc rows x rows. */
static void get_tmt(const model* m, double* T, double* Mc, double* Td) {
  const int nc = m->n, n = m->na, rows = m->m_rows; /* columns: the coordinates, then 6 per free-joint frame (:232-276) */
  int e, i, row = 0, pass;
  memset(T, 0, sizeof(double) * (size_t)rows * (size_t)n);
  memset(Td, 0, sizeof(double) * (size_t)rows * (size_t)n);
  memset(Mc, 0, sizeof(double) * (size_t)rows * (size_t)rows);
  for (pass = 0; pass < 3; ++pass) {
    for (e = 0; e < m->d.n_elements; ++e) {
      const rkb_element* E = &m->el[e];
      if (pass == 0 && E->kind == RKB_INERTIA_GEN) {
        for (i = 0; i < nc; ++i)
          if ((E->upstream >> i) & 1u) { T[row * n + i] = 1.0; Td[row * n + i] = 0.0; } /* jacobian_gen_gen(1,0), motion_jacobians.hpp:49-107 */
        Mc[row * rows + row] = E->p[0];
        row += 1;
      } else if (pass == 1 && E->kind == RKB_INERTIA_2D) {
        for (i = 0; i < nc; ++i)
          if ((E->upstream >> i) & 1u) {
            double c[3], cd[3];
            int k;
            jac2_rel(m, &m->j2[i], E->frame_a, c, cd);
            for (k = 0; k < 3; ++k) { T[(row + k) * n + i] = c[k]; Td[(row + k) * n + i] = cd[k]; }
          }
        for (i = 0; i < m->nfree; ++i)
          if ((E->upstream >> (32 + i)) & 1u) { /* mUpStream2DJoints, :184-196 */
            double b[9], bd[9];
            int k, l;
            jac22_rel(m, &m->jf2[i], E->frame_a, b, bd);
            for (k = 0; k < 3; ++k)
              for (l = 0; l < 3; ++l) { T[(row + k) * n + nc + 3 * i + l] = b[k * 3 + l]; Td[(row + k) * n + nc + 3 * i + l] = bd[k * 3 + l]; }
          }
        Mc[row * rows + row] = E->p[0];
        Mc[(row + 1) * rows + row + 1] = E->p[0];
        Mc[(row + 2) * rows + row + 2] = E->p[1];
        row += 3;
      } else if (pass == 2 && E->kind == RKB_INERTIA_3D) {
        const double* I = &E->p[1];
        for (i = 0; i < nc; ++i)
          if ((E->upstream >> i) & 1u) {
            double c[6], cd[6];
            int k;
            jac3_rel(m, &m->j3[i], E->frame_a, c, cd);
            for (k = 0; k < 6; ++k) { T[(row + k) * n + i] = c[k]; Td[(row + k) * n + i] = cd[k]; }
          }
        for (i = 0; i < m->nfree; ++i)
          if ((E->upstream >> (32 + i)) & 1u) { /* mUpStream3DJoints, :264-274 */
            double b[36], bd[36];
            int k, l;
            jac33_rel(m, &m->jf[i], E->frame_a, b, bd);
            for (k = 0; k < 6; ++k)
              for (l = 0; l < 6; ++l) { T[(row + k) * n + nc + 6 * i + l] = b[k * 6 + l]; Td[(row + k) * n + nc + 6 * i + l] = bd[k * 6 + l]; }
          }
        Mc[row * rows + row] = E->p[0];
        Mc[(row + 1) * rows + row + 1] = E->p[0];
        Mc[(row + 2) * rows + row + 2] = E->p[0];
        {
          int r = row + 3;
          Mc[r * rows + r] = I[0]; Mc[r * rows + r + 1] = I[1]; Mc[r * rows + r + 2] = I[2];
          Mc[(r + 1) * rows + r] = I[1]; Mc[(r + 1) * rows + r + 1] = I[3]; Mc[(r + 1) * rows + r + 2] = I[4];
          Mc[(r + 2) * rows + r] = I[2]; Mc[(r + 2) * rows + r + 1] = I[4]; Mc[(r + 2) * rows + r + 2] = I[5];
        }
        row += 6;
      }
    }
  }
}

/* mass_matrix_calc::getMassMatrix / getMassMatrixAndDerivative, mass_matrix_calculator.cpp:80-98.
 * M (n x n row-major) = sym(T^T (Mc T)) with the 1/2 (Mij + Mji) average of the
 * mat<symmetric> converting constructor (core/lin_alg/mat_alg_symmetric.hpp:171-200);
 * Mdot = Td^T (Mc T) + its transpose. */
static void mass_matrix(const model* m, double* M, double* Mdot) {
  const int n = m->na, rows = m->m_rows;
  double* T = (double*)malloc(sizeof(double) * (size_t)(rows * n + 1));
  double* Td = (double*)malloc(sizeof(double) * (size_t)(rows * n + 1));
  double* Mc = (double*)malloc(sizeof(double) * (size_t)(rows * rows + 1));
  double* MT = (double*)malloc(sizeof(double) * (size_t)(rows * n + 1));
  double* G = (double*)malloc(sizeof(double) * (size_t)(n * n + 1));
  int i, j, k;
  get_tmt(m, T, Mc, Td);
  for (i = 0; i < rows; ++i)
    for (j = 0; j < n; ++j) {
      double s = 0.0;
      for (k = 0; k < rows; ++k) s += Mc[i * rows + k] * T[k * n + j];
      MT[i * n + j] = s;
    }
  for (i = 0; i < n; ++i)
    for (j = 0; j < n; ++j) {
      double s = 0.0;
      for (k = 0; k < rows; ++k) s += T[k * n + i] * MT[k * n + j];
      G[i * n + j] = s;
    }
  for (i = 0; i < n; ++i) {
    for (j = 0; j < i; ++j) M[i * n + j] = M[j * n + i] = 0.5 * (G[j * n + i] + G[i * n + j]);
    M[i * n + i] = G[i * n + i];
  }
  if (Mdot) {
    for (i = 0; i < n; ++i)
      for (j = 0; j < n; ++j) {
        double s = 0.0;
        for (k = 0; k < rows; ++k) s += Td[k * n + i] * MT[k * n + j];
        G[i * n + j] = s;
      }
    for (i = 0; i < n; ++i)
      for (j = 0; j < n; ++j) Mdot[i * n + j] = G[i * n + j] + G[j * n + i];
  }
  free(T); free(Td); free(Mc); free(MT); free(G);
}

/* decompose_Cholesky_impl + backsub_Cholesky_impl, core/lin_alg/mat_cholesky.hpp:63-84, 160-179.
 * Returns 1 where the reference throws singularity_error (pivot < tol, tested BEFORE the sqrt). */
static int cholesky_solve(int n, const double* A, double* b, int nrhs, double tol) {
  double L[KTO_MAX_ACC * KTO_MAX_ACC];
  int i, j, k, c;
  memset(L, 0, sizeof L);
  for (i = 0; i < n; ++i) {
    for (j = 0; j < i; ++j) {
      L[i * n + j] = A[i * n + j];
      for (k = 0; k < j; ++k) L[i * n + j] -= L[i * n + k] * L[j * n + k];
      L[i * n + j] /= L[j * n + j];
    }
    L[i * n + i] = A[i * n + i];
    for (k = 0; k < i; ++k) L[i * n + i] -= L[i * n + k] * L[i * n + k];
    if (L[i * n + i] < tol) return 1;
    L[i * n + i] = sqrt(L[i * n + i]);
  }
  for (c = 0; c < nrhs; ++c) {
    for (i = 0; i < n; ++i) {
      for (k = 0; k < i; ++k) b[i * nrhs + c] -= L[i * n + k] * b[k * nrhs + c];
      b[i * nrhs + c] /= L[i * n + i];
    }
    for (i = n - 1; i >= 0; --i) {
      for (k = n - 1; k > i; --k) b[i * nrhs + c] -= L[k * n + i] * b[k * nrhs + c];
      b[i * nrhs + c] /= L[i * n + i];
    }
  }
  return 0;
}
int kto_cholesky_solve(int n, const double* A, double* b, int nrhs, double tol) {
  if (n < 1 || n > RKB_MAX_COORDS) return -1;
  return cholesky_solve(n, A, b, nrhs, tol);
}
/* decompose_LDL_impl + backsub_LDL_impl, mat_cholesky.hpp:134-157, 207-227 */
int kto_ldl_solve(int n, const double* Ain, double* b, int nrhs, double tol) {
  double A[RKB_MAX_COORDS * RKB_MAX_COORDS], v[RKB_MAX_COORDS];
  int i, j, k, c;
  if (n < 1 || n > RKB_MAX_COORDS) return -1;
  memcpy(A, Ain, sizeof(double) * (size_t)(n * n));
  for (i = 0; i < n; ++i) {
    for (j = 0; j < i; ++j) v[j] = A[i * n + j] * A[j * n + j];
    v[i] = A[i * n + i];
    for (j = 0; j < i; ++j) v[i] -= A[i * n + j] * v[j];
    A[i * n + i] = v[i];
    if (fabs(v[i]) < tol) return 1;
    for (j = i + 1; j < n; ++j) {
      for (k = 0; k < i; ++k) A[j * n + i] -= A[j * n + k] * v[k];
      A[j * n + i] /= v[i];
    }
  }
  for (c = 0; c < nrhs; ++c) {
    for (i = 0; i < n; ++i)
      for (k = 0; k < i; ++k) b[i * nrhs + c] -= A[i * n + k] * b[k * nrhs + c];
    for (i = 0; i < n; ++i) b[i * nrhs + c] /= A[i * n + i];
    for (i = n; i > 0;) {
      --i;
      for (k = n - 1; k > i; --k) b[i * nrhs + c] -= A[k * n + i] * b[k * nrhs + c];
    }
  }
  return 0;
}

/* kte_nl_system::get_state_derivative, ctrl/ctrl_sys/kte_nl_system.hpp:238-346 (gen coords and 3D free frames) */
static int state_derivative(model* m, const double* x, const double* u, double* xd) {
  double M[KTO_MAX_ACC * KTO_MAX_ACC], f[KTO_MAX_ACC];
  int i, k, st;
  apply_states_and_inputs(m, x, u);
  do_motion(m);
  clear_force(m);
  do_force(m);
  for (i = 0; i < m->n; ++i) f[i] = m->c[i].f;
  for (i = 0; i < m->nfree && m->d.dim == 3; ++i)                                    /* :262-270 */
    for (k = 0; k < 3; ++k) { f[m->n + 6 * i + k] = m->fc[i].F.x[k]; f[m->n + 6 * i + 3 + k] = m->fc[i].T.x[k]; }
  for (i = 0; i < m->nfree && m->d.dim == 2; ++i) {                                  /* :255-260 */
    f[m->n + 3 * i] = m->fc2[i].F.x[0]; f[m->n + 3 * i + 1] = m->fc2[i].F.x[1]; f[m->n + 3 * i + 2] = m->fc2[i].T;
  }
  mass_matrix(m, M, NULL);
  st = cholesky_solve(m->na, M, f, 1, 1E-8);
  if (st) return RKB_STATUS_SINGULAR;
  for (i = 0; i < m->n; ++i) { xd[2 * i] = m->c[i].qd; xd[2 * i + 1] = f[i]; }
  for (i = 0; i < m->nfree && m->d.dim == 2; ++i) {                                  /* :282-291: (cos, sin)' from the RAW state */
    double* o = xd + 2 * m->n + 7 * i;
    const frame2* F = &m->fc2[i];
    o[0] = F->v.x[0]; o[1] = F->v.x[1];
    o[2] = -m->fc2_raw[i][1] * F->w;
    o[3] = m->fc2_raw[i][0] * F->w;
    for (k = 0; k < 3; ++k) o[4 + k] = f[m->n + 3 * i + k];
  }
  for (i = 0; i < m->nfree && m->d.dim == 3; ++i) {                                  /* :293-308 */
    double* o = xd + 2 * m->n + 13 * i;
    const frame3* F = &m->fc[i];
    const double* q = F->Q.q;
    const double* W = F->w.x;
    for (k = 0; k < 3; ++k) o[k] = F->v.x[k];
    /* quaternion::getQuaternionDot, rotations_3D.hpp:1206-1211 */
    o[3] = -0.5 * (q[1] * W[0] + q[2] * W[1] + q[3] * W[2]);
    o[4] = 0.5 * (q[0] * W[0] - q[3] * W[1] + q[2] * W[2]);
    o[5] = 0.5 * (q[0] * W[1] + q[3] * W[0] - q[1] * W[2]);
    o[6] = 0.5 * (q[0] * W[2] - q[2] * W[0] + q[1] * W[1]);
    for (k = 0; k < 6; ++k) o[7 + k] = f[m->n + 6 * i + k];
  }
  return 0;
}

int kto_eval(void* h, size_t N, const double* x, const double* u, double* xdot, int32_t* status) {
  model* m = (model*)h;
  const int nx = m->nx;
  size_t i;
  int k;
  for (i = 0; i < N; ++i) {
    int st = state_derivative(m, x + i * nx, u ? u + i * m->nu : NULL, xdot + i * nx);
    if (st) for (k = 0; k < nx; ++k) xdot[i * nx + k] = NAN;
    if (status) status[i] = st;
  }
  return 0;
}

int kto_gen_forces(void* h, size_t N, const double* x, const double* u, double* f) {
  model* m = (model*)h;
  const int nx = m->nx;
  size_t i;
  int k;
  for (i = 0; i < N; ++i) {
    apply_states_and_inputs(m, x + i * nx, u ? u + i * m->nu : NULL);
    do_motion(m); clear_force(m); do_force(m);
    for (k = 0; k < m->n; ++k) f[i * m->na + k] = m->c[k].f;
    for (k = 0; k < 6 * m->nfree && m->d.dim == 3; ++k)  /* Force, Torque of the free joints' coordinate frames (kte_nl_system.hpp:262-270) */
      f[i * m->na + m->n + k] = (k % 6 < 3) ? m->fc[k / 6].F.x[k % 6] : m->fc[k / 6].T.x[k % 6 - 3];
    for (k = 0; k < 3 * m->nfree && m->d.dim == 2; ++k)
      f[i * m->na + m->n + k] = (k % 3 < 2) ? m->fc2[k / 3].F.x[k % 3] : m->fc2[k / 3].T;
  }
  return 0;
}

/* doMotion / clearForce / doForce with caller-chosen q_ddot (the demo of ctrl/mbd_kte/test_bm.cpp:103-121
 * evaluates the chain at q_ddot = 0 and q_ddot = 1 to read the mass matrix off the force difference) */
int kto_gen_forces_qdd(void* h, const double* x, const double* u, const double* qdd, double* f) {
  model* m = (model*)h;
  int k;
  apply_states_and_inputs(m, x, u);
  for (k = 0; k < m->n; ++k) m->c[k].qdd = qdd[k];
  do_motion(m); clear_force(m); do_force(m);
  for (k = 0; k < m->n; ++k) f[k] = m->c[k].f;
  return 0;
}

int kto_mass(void* h, size_t N, const double* x, double* M, double* Mdot) {
  model* m = (model*)h;
  const int nx = m->nx, nn = m->na * m->na;
  size_t i;
  for (i = 0; i < N; ++i) {
    apply_states_and_inputs(m, x + i * nx, NULL);
    do_motion(m);
    mass_matrix(m, M + i * nn, Mdot ? Mdot + i * nn : NULL);
  }
  return 0;
}

int kto_tmt(void* h, const double* x, double* Tcm, double* Mcm, double* Tcm_dot) {
  model* m = (model*)h;
  if (!Tcm || !Mcm || !Tcm_dot) return m->m_rows;
  apply_states_and_inputs(m, x, NULL);
  do_motion(m);
  get_tmt(m, Tcm, Mcm, Tcm_dot);
  return m->m_rows;
}

int kto_frames(void* h, const double* x, const double* u, double* out) {
  model* m = (model*)h;
  int i, k;
  apply_states_and_inputs(m, x, u);
  do_motion(m); clear_force(m); do_force(m);
  for (i = 0; i < m->d.n_frames; ++i) {
    double* o = out + 25 * i;
    for (k = 0; k < 25; ++k) o[k] = 0.0;
    if (m->d.dim == 3) {
      const frame3* F = &m->f3[i];
      for (k = 0; k < 3; ++k) { o[k] = F->p.x[k]; o[7 + k] = F->v.x[k]; o[10 + k] = F->w.x[k]; o[13 + k] = F->a.x[k];
                                o[16 + k] = F->al.x[k]; o[19 + k] = F->F.x[k]; o[22 + k] = F->T.x[k]; }
      for (k = 0; k < 4; ++k) o[3 + k] = F->Q.q[k];
    } else {
      const frame2* F = &m->f2[i];
      for (k = 0; k < 2; ++k) { o[k] = F->p.x[k]; o[3 + k] = F->R.q[k]; o[7 + k] = F->v.x[k]; o[13 + k] = F->a.x[k]; o[19 + k] = F->F.x[k]; }
      o[10] = F->w; o[16] = F->al; o[22] = F->T;
    }
  }
  return 0;
}

/* runge_kutta4_integrator<T>::integrate, core/integrators/fixed_step_integrators.hpp:256-293,
 * driven for an explicit number of steps (the reference loop is time-driven; ref_lib.cpp feeds it
 * an end time of (n_steps - 0.5) dt so both run exactly n_steps steps).  The rate function is
 * num_int_dtnl_sys::rate_function_impl (ctrl/ctrl_sys/num_int_dtnl_system.hpp:85-99): the input
 * is held constant.  The last rate evaluation of the run (:291) does not change the state and
 * is skipped; a singular mass matrix there would still raise in the reference, so it is kept
 * for the status word only. */
static void rk4_range(model* m, size_t i0, size_t i1, const double* x0, const double* u, double dt, int n_steps,
                      double* xout, int32_t* status) {
  const int nx = m->nx;
  double x[KTO_MAX_STATE], w[KTO_MAX_STATE], f[KTO_MAX_STATE];
  double k1[KTO_MAX_STATE], k2[KTO_MAX_STATE], k3[KTO_MAX_STATE];
  size_t i;
  int s, k;
  for (i = i0; i < i1; ++i) {
    const double* ui = u ? u + i * m->nu : NULL;
    int st = 0;
    for (k = 0; k < nx; ++k) x[k] = x0[i * nx + k];
    if (n_steps > 0) {
      st |= state_derivative(m, x, ui, f);                                    /* :273 */
      for (s = 0; s < n_steps && !st; ++s) {
        for (k = 0; k < nx; ++k) { w[k] = x[k]; k1[k] = f[k] * dt; x[k] += k1[k] * 0.5; }  /* :277-279 */
        st |= state_derivative(m, x, ui, f);                                  /* :281 */
        if (st) break;
        for (k = 0; k < nx; ++k) { k2[k] = f[k] * dt; x[k] = w[k] + k2[k] * 0.5; }          /* :282-283 */
        st |= state_derivative(m, x, ui, f);                                  /* :284 */
        if (st) break;
        for (k = 0; k < nx; ++k) { k3[k] = f[k] * dt; x[k] = w[k] + k3[k]; }                /* :285-286 */
        st |= state_derivative(m, x, ui, f);                                  /* :288 */
        if (st) break;
        for (k = 0; k < nx; ++k)                                                            /* :289 */
          x[k] += (k1[k] + k2[k] * 2.0 + f[k] * dt) / 6.0 - k3[k] * (2.0 / 3.0);
        st |= state_derivative(m, x, ui, f);                                  /* :291 */
      }
    }
    for (k = 0; k < nx; ++k) {
      xout[i * nx + k] = x[k];
      if (!isfinite(x[k])) st |= RKB_STATUS_NONFINITE;
    }
    if (status) status[i] = st;
  }
}

/* ctrl::detail::runge_kutta4_integrate_impl (ctrl/sys_integrators/runge_kutta4_integrator_sys.hpp:50-97): RK4 whose
 * input comes from a trajectory, read at t (:66, carried over from the previous step's :91), t + dt/2 (:81, used for the
 * second and third evaluations) and t + dt (:91).  u_nodes: the trajectory sampled at every half step,
 * [N][2 n_steps + 1][nu].  The final combination is written as the reference writes it (:93), which differs from
 * fixed_step_integrators.hpp:289 in the order of the additions. */
void kto_rk4_inputs(void* h, size_t N, const double* x0, const double* u_nodes, double dt, int n_steps, double* xout, int32_t* status) {
  model* m = (model*)h;
  const int nx = m->nx;
  const size_t J = 2 * (size_t)n_steps + 1;
  double x[KTO_MAX_STATE], w[KTO_MAX_STATE], f[KTO_MAX_STATE];
  double k1[KTO_MAX_STATE], k2[KTO_MAX_STATE], k3[KTO_MAX_STATE];
  size_t i;
  int s, k;
  for (i = 0; i < N; ++i) {
    const double* ui = u_nodes ? u_nodes + i * J * (size_t)m->nu : NULL;
    int st = 0;
    for (k = 0; k < nx; ++k) x[k] = x0[i * nx + k];
    if (n_steps > 0) {
      st |= state_derivative(m, x, ui, f);                                                             /* :68 */
      for (s = 0; s < n_steps && !st; ++s) {
        const double* u_mid = ui ? ui + (2 * (size_t)s + 1) * (size_t)m->nu : NULL;
        const double* u_end = ui ? ui + (2 * (size_t)s + 2) * (size_t)m->nu : NULL;
        for (k = 0; k < nx; ++k) { w[k] = x[k]; k1[k] = dt * f[k]; x[k] = x[k] + 0.5 * k1[k]; }        /* :75-77 */
        st |= state_derivative(m, x, u_mid, f);                                                        /* :79-81 */
        if (st) break;
        for (k = 0; k < nx; ++k) { k2[k] = dt * f[k]; x[k] = w[k] + 0.5 * k2[k]; }                      /* :82-83 */
        st |= state_derivative(m, x, u_mid, f);                                                        /* :85 */
        if (st) break;
        for (k = 0; k < nx; ++k) { k3[k] = dt * f[k]; x[k] = w[k] + k3[k]; }                            /* :86-87 */
        st |= state_derivative(m, x, u_end, f);                                                        /* :89-92 */
        if (st) break;
        for (k = 0; k < nx; ++k)                                                                       /* :93 */
          x[k] = x[k] + ((((1.0 / 6.0) * k1[k] + (2.0 / 6.0) * k2[k]) + (dt / 6.0) * f[k]) - (2.0 / 3.0) * k3[k]);
        st |= state_derivative(m, x, u_end, f);                                                        /* :95 */
      }
    }
    for (k = 0; k < nx; ++k) {
      xout[i * nx + k] = x[k];
      if (!isfinite(x[k])) st |= RKB_STATUS_NONFINITE;
    }
    if (status) status[i] = st;
  }
}

/* euler_integrator<T>::integrate (fixed_step_integrators.hpp:64-84), midpoint_integrator<T>::integrate
 * (:177-202) and runge_kutta5_integrator<T>::integrate (:351-399), same conventions as rk4_range:
 * explicit step count, input held constant, the trailing rate evaluation kept for the status only. */
static void scheme_range(model* m, int scheme, size_t i0, size_t i1, const double* x0, const double* u, double dt, int n_steps,
                         double* xout, int32_t* status) {
  const int nx = m->nx;
  double x[KTO_MAX_STATE], w[KTO_MAX_STATE], f[KTO_MAX_STATE];
  double k1[KTO_MAX_STATE], k2[KTO_MAX_STATE], k3[KTO_MAX_STATE], k4[KTO_MAX_STATE], k5[KTO_MAX_STATE];
  size_t i;
  int s, k;
  if (scheme == RKB_SCHEME_RK4) { rk4_range(m, i0, i1, x0, u, dt, n_steps, xout, status); return; }
  for (i = i0; i < i1; ++i) {
    const double* ui = u ? u + i * m->nu : NULL;
    int st = 0;
    for (k = 0; k < nx; ++k) x[k] = x0[i * nx + k];
    if (n_steps > 0) {
      st |= state_derivative(m, x, ui, f);
      for (s = 0; s < n_steps && !st; ++s) {
        if (scheme == RKB_SCHEME_EULER) {
          for (k = 0; k < nx; ++k) x[k] += f[k] * dt;                                   /* :78 */
          st |= state_derivative(m, x, ui, f);                                          /* :82 */
        } else if (scheme == RKB_SCHEME_MIDPOINT) {
          for (k = 0; k < nx; ++k) w[k] = x[k] + f[k] * (dt * 0.5);                     /* :193 */
          st |= state_derivative(m, w, ui, f);                                          /* :195 */
          if (st) break;
          for (k = 0; k < nx; ++k) x[k] += f[k] * dt;                                   /* :197 */
          st |= state_derivative(m, x, ui, f);                                          /* :199 */
        } else {
          for (k = 0; k < nx; ++k) { w[k] = x[k]; k1[k] = f[k] * dt; x[k] += k1[k] * 0.25; }              /* :367-368 */
          st |= state_derivative(m, x, ui, f);
          if (st) break;
          for (k = 0; k < nx; ++k) { k2[k] = f[k] * dt; x[k] += (k2[k] * 9.0 - k1[k] * 5.0) / 32.0; }    /* :372 */
          st |= state_derivative(m, x, ui, f);
          if (st) break;
          for (k = 0; k < nx; ++k) {                                                                     /* :376 */
            k3[k] = f[k] * dt;
            x[k] += (k1[k] * 276165.0 - k2[k] * 1250865.0 + k3[k] * 1167360.0) / 351520.0;
          }
          st |= state_derivative(m, x, ui, f);
          if (st) break;
          for (k = 0; k < nx; ++k) {                                                                     /* :380 */
            k4[k] = f[k] * dt;
            x[k] = w[k] + k1[k] * (439.0 / 216.0) - k2[k] * 8.0 + k3[k] * (3680.0 / 513.0) - k4[k] * (845.0 / 4104.0);
          }
          st |= state_derivative(m, x, ui, f);
          if (st) break;
          for (k = 0; k < nx; ++k) {                                                                     /* :384 */
            k5[k] = f[k] * dt;
            x[k] = w[k] - k1[k] * (8.0 / 27.0) + k2[k] * 2.0 - k3[k] * (3544.0 / 2565.0) + k4[k] * (1859.0 / 4104.0) - k5[k] * (11.0 / 40.0);
          }
          st |= state_derivative(m, x, ui, f);
          if (st) break;
          for (k = 0; k < nx; ++k)                                                                       /* :388 */
            x[k] = w[k] + k1[k] * (16.0 / 135.0) + k3[k] * (6656.0 / 12825.0) + k4[k] * (28561.0 / 56430.0) - k5[k] * (9.0 / 50.0)
                   + f[k] * (2.0 * dt / 55.0);
          st |= state_derivative(m, x, ui, f);                                                           /* :391 */
        }
      }
    }
    for (k = 0; k < nx; ++k) {
      xout[i * nx + k] = x[k];
      if (!isfinite(x[k])) st |= RKB_STATUS_NONFINITE;
    }
    if (status) status[i] = st;
  }
}

#include "steer_law.h"

/* test hook: steer_bounded_input for `count` (u_prev, u_bias, u_correction) triples — what tests/test_oracle.py holds against
 * IHAQR_topology::get_bounded_input itself (oracle/ref_steer_law.cpp) */
int kto_bounded_input(int nu, double T, const double* lo, const double* hi, const double* dlo, const double* dhi, int count,
                      const double* u_prev, const double* u_bias, const double* u_corr, double* u_out) {
  int i;
  if (nu < 1 || nu > STEER_MAX_INPUTS) return -1;
  for (i = 0; i < count; ++i)
    steer_bounded_input(nu, T, lo, hi, dlo, dhi, u_prev + (size_t)i * nu, u_bias + (size_t)i * nu, u_corr + (size_t)i * nu, u_out + (size_t)i * nu);
  return 0;
}

/* The steering loop of MEAQR_topology.hpp:503-561 / IHAQR_topology.hpp:349-378 for each sample (see
 * steer_law.h for the feedback law); one control interval = `substeps` RK4 steps of `dt`. */
int kto_steer_feedback(void* h, size_t N, const double* x0, const double* goal, const double* u_bias, const double* gain,
                       double* u_prev, double T, double dt, int substeps, int max_intervals, double proximity, int saturate_first,
                       const double* lo, const double* hi, const double* dlo, const double* dhi,
                       double* x_out, int32_t* n_done, double* traj, int32_t* status) {
  model* m = (model*)h;
  const int nx = m->nx, nu = m->nu;
  size_t i;
  if (nu > STEER_MAX_INPUTS) return -1;
  for (i = 0; i < N; ++i) {
    double x[KTO_MAX_STATE], xn[KTO_MAX_STATE], u[STEER_MAX_INPUTS], up[STEER_MAX_INPUTS];
    int k = 0, j, st = 0;
    for (j = 0; j < nx; ++j) x[j] = x0[i * nx + j];
    for (j = 0; j < nu; ++j) up[j] = u_prev[i * nu + j];
    while (k < max_intervals) {
      int32_t s1 = 0;
      if (!steer_next_input(nx, nu, T, proximity, (!saturate_first && k == 0), lo, hi, dlo, dhi, x, goal + i * nx,
                            u_bias + i * nu, gain + i * (size_t)nu * nx, up, u))
        break;
      rk4_range(m, 0, 1, x, u, dt, substeps, xn, &s1);
      st |= s1;
      for (j = 0; j < nx; ++j) x[j] = xn[j];
      for (j = 0; j < nu; ++j) up[j] = u[j];
      if (traj) for (j = 0; j < nx; ++j) traj[(i * (size_t)max_intervals + k) * nx + j] = x[j];
      ++k;
    }
    for (j = 0; j < nx; ++j) x_out[i * nx + j] = x[j];
    for (j = 0; j < nu; ++j) u_prev[i * nu + j] = up[j];
    if (n_done) n_done[i] = k;
    if (status) status[i] = st;
  }
  return 0;
}

static model* model_clone(const model* m) { return (model*)kto_create(&m->d); }

double kto_rk4(void* h, size_t N, const double* x0, const double* u, double dt, int n_steps,
               double* xout, int32_t* status, int n_workers) {
  return kto_integrate(h, N, x0, u, RKB_SCHEME_RK4, dt, n_steps, xout, status, n_workers);
}

double kto_integrate(void* h, size_t N, const double* x0, const double* u, int scheme, double dt, int n_steps,
                     double* xout, int32_t* status, int n_workers) {
  model* m = (model*)h;
  const int nx = m->nx;
  struct timespec t0, t1;
  if (n_workers < 1) n_workers = 1;
  if ((size_t)n_workers > N && N > 0) n_workers = (int)N;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (n_workers == 1 || N == 0) {
    scheme_range(m, scheme, 0, N, x0, u, dt, n_steps, xout, status);
  } else {
    /* one forked worker per block of samples; results come back through a shared mapping */
    size_t bx = N * (size_t)nx * sizeof(double), bs = N * sizeof(int32_t);
    void* shm = mmap(NULL, bx + bs, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    double* sx;
    int32_t* ss;
    pid_t pids[1024];
    int t, np = 0, ok = 1;
    if (shm == MAP_FAILED) return -1.0;
    if (n_workers > 1024) n_workers = 1024;
    sx = (double*)shm;
    ss = (int32_t*)((char*)shm + bx);
    for (t = 0; t < n_workers; ++t) {
      size_t i0 = N * (size_t)t / (size_t)n_workers, i1 = N * (size_t)(t + 1) / (size_t)n_workers;
      pid_t pid = fork();
      if (pid == 0) {
        model* mine = model_clone(m);
        scheme_range(mine, scheme, i0, i1, x0, u, dt, n_steps, sx, ss);
        _exit(0);
      }
      if (pid < 0) { ok = 0; break; }
      pids[np++] = pid;
    }
    for (t = 0; t < np; ++t) {
      int wst = 0;
      if (waitpid(pids[t], &wst, 0) < 0 || !WIFEXITED(wst) || WEXITSTATUS(wst) != 0) ok = 0;
    }
    if (ok) {
      memcpy(xout, sx, bx);
      if (status) memcpy(status, ss, bs);
    }
    munmap(shm, bx + bs);
    if (!ok) return -1.0;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------
 * nearest neighbours: ReaK::pp::min_dist_linear_search (ctrl/path_planning/topological_search.hpp:91-112 and :238-270)
 * over points of a vect_n topology, distance = norm_2(difference) (core/lin_alg/vect_alg.hpp:2314-2333: the sum of
 * squares accumulated in index order from 0.0, then sqrt).  A candidate must compare less than the running k-th
 * distance (initially `radius`) to enter; the output is sorted by ascending distance.  Restated as "the k smallest by
 * (distance, index)": identical to the reference whenever no two candidate distances are bit-equal (the reference's
 * heap leaves the order of equal keys to std::push_heap / pop_heap; its single-neighbour form :102-110 keeps the first
 * of equal minima, as this does).
 * ---------------------------------------------------------------------------------------- */
int kto_nearest(size_t V, const double* vertices, size_t Q, const double* queries, int dim, int k, double radius,
                int32_t* index, double* distance, int32_t* count) {
  size_t i, v;
  int c, r;
  double* dl = (double*)malloc(sizeof(double) * (size_t)(k + 1));
  int32_t* il = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k + 1));
  for (i = 0; i < Q; ++i) {
    int cnt = 0;
    double bound = radius;
    for (v = 0; v < V; ++v) {
      double sum = 0.0, d;
      int pos;
      for (c = 0; c < dim; ++c) {
        double t = vertices[v * (size_t)dim + c] - queries[i * (size_t)dim + c];
        sum += t * t;
      }
      d = sqrt(sum);
      if (!(d < bound)) continue;
      pos = cnt < k ? cnt : k - 1;
      while (pos > 0 && dl[pos - 1] > d) { dl[pos] = dl[pos - 1]; il[pos] = il[pos - 1]; --pos; }
      dl[pos] = d; il[pos] = (int32_t)v;
      if (cnt < k) ++cnt;
      if (cnt == k) bound = dl[k - 1];
    }
    for (r = 0; r < k; ++r) {
      index[i * (size_t)k + r] = r < cnt ? il[r] : -1;
      if (distance) distance[i * (size_t)k + r] = r < cnt ? dl[r] : INFINITY;
    }
    if (count) count[i] = cnt;
  }
  free(dl); free(il);
  return 0;
}
