/* c_abi_demo.c — plain C against the C-ABI (include/reak_b200.h), no Python, no C++:
 * a 1-DOF pendulum (ctrl/mbd_kte/test_bm.cpp:46-72: revolute_joint_2D + 0.5 m rigid_link_2D + 1 kg
 * inertia_2D, gravity as an upward base acceleration) and a 2-DOF 3D arm are built from flat element
 * lists, evaluated and integrated on GPU 0, and checked against closed-form answers:
 *   pendulum  q_ddot = -g cos(q) / L,  M = m L^2
 *   energy of the torque-free 3D arm is conserved by RK4 to O(dt^4)
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -Lreak_b200/lib -lreak_b200 -lm -o examples/c_abi_demo
 *   LD_LIBRARY_PATH=reak_b200/lib examples/c_abi_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "reak_b200.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int rc_ = (call);                                                                                 \
    if (rc_ != RKB_OK) {                                                                              \
      fprintf(stderr, "%s -> %d (%s) %s\n", #call, rc_, rkb_strerror(rc_), rkb_last_cuda_error());    \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

static rkb_element element(int kind, int fa, int fb, int coord) {
  rkb_element e;
  memset(&e, 0, sizeof e);
  e.kind = kind; e.frame_a = fa; e.frame_b = fb; e.coord = coord;
  return e;
}

int main(void) {
  /* ---- pendulum of test_bm.cpp -------------------------------------------------------------- */
  rkb_element el[3];
  el[0] = element(RKB_REVOLUTE_2D, 0, 1, 0);
  el[1] = element(RKB_RIGID_LINK_2D, 1, 2, -1);
  el[1].p[0] = 0.5;                        /* offset (0.5, 0), angle 0 */
  el[2] = element(RKB_INERTIA_2D, 2, -1, -1);
  el[2].p[0] = 1.0; el[2].p[1] = 0.0;      /* mass, moment of inertia */
  el[2].upstream = 1u;                     /* depends on coordinate 0 */
  rkb_chain_desc d;
  memset(&d, 0, sizeof d);
  d.dim = 2; d.n_elements = 3; d.n_frames = 3; d.n_coords = 1; d.n_inputs = 0; d.base_frame = 0;
  d.base.acceleration[1] = 9.81;           /* gravity: the base accelerates upwards */
  d.elements = el;
  rkb_chain* pend = NULL;
  CHECK(rkb_chain_create(&d, &pend));
  enum { N = 1000 };
  static double x[N][2], xd[N][2], M[N], x1[N][2];
  static int32_t st[N];
  int i;
  for (i = 0; i < N; ++i) { x[i][0] = -3.0 + 6.0 * i / N; x[i][1] = 0.3; }
  CHECK(rkb_eval(pend, 0, N, &x[0][0], NULL, &xd[0][0], st, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
  CHECK(rkb_mass_matrix(pend, 0, N, &x[0][0], M, NULL, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
  double worst = 0.0;
  for (i = 0; i < N; ++i) {
    const double want = -9.81 * cos(x[i][0]) / 0.5;
    worst = fmax(worst, fabs(xd[i][1] - want) / fmax(1.0, fabs(want)));
    worst = fmax(worst, fabs(xd[i][0] - 0.3));
    worst = fmax(worst, fabs(M[i] - 0.25));
    if (st[i]) { fprintf(stderr, "status %d at %d\n", st[i], i); return 1; }
  }
  printf("pendulum: max error of q_ddot = -g cos q / L and M = m L^2 over %d states: %.2e\n", N, worst);
  if (!(worst < 1e-12)) return 1;
  /* RK4 against the energy integral E = 1/2 m L^2 qd^2 + m g L sin q (conserved, O(dt^4) drift) */
  CHECK(rkb_rollout_rk4(pend, 0, N, &x[0][0], NULL, 1e-3, 500, &x1[0][0], st, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
  worst = 0.0;
  for (i = 0; i < N; ++i) {
    const double e0 = 0.125 * x[i][1] * x[i][1] + 9.81 * 0.5 * sin(x[i][0]);
    const double e1 = 0.125 * x1[i][1] * x1[i][1] + 9.81 * 0.5 * sin(x1[i][0]);
    worst = fmax(worst, fabs(e1 - e0));
  }
  printf("pendulum: energy drift after 500 RK4 steps of 1 ms: %.2e J (kernel %.3f ms)\n", worst, rkb_last_kernel_ms(pend));
  if (!(worst < 1e-9)) return 1;
  /* the same rollout as 5 control intervals of 100 steps through rkb_rollout gives the same bits */
  {
    static double x2[N][2], traj[N][5][2];
    rkb_rollout_opts o;
    memset(&o, 0, sizeof o);
    o.scheme = RKB_SCHEME_RK4; o.n_intervals = 5; o.steps_per_interval = 100; o.dt = 1e-3;
    CHECK(rkb_rollout(pend, 0, N, &x[0][0], NULL, &o, &x2[0][0], &traj[0][0][0], st, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
    if (memcmp(x1, x2, sizeof x1) != 0 || memcmp(&traj[N - 1][4][0], &x1[N - 1][0], 2 * sizeof(double)) != 0) {
      fprintf(stderr, "interval rollout differs from the single rollout\n");
      return 1;
    }
    printf("pendulum: 5 x 100-step control intervals == one 500-step rollout, bit for bit\n");
  }
  rkb_chain_destroy(pend);

  /* ---- argument errors come back as codes, never as exceptions ------------------------------- */
  {
    rkb_chain* none = NULL;
    rkb_element bad = element(99, 0, 1, 0);
    rkb_chain_desc db = d;
    db.n_elements = 1; db.elements = &bad;
    if (rkb_chain_create(&db, &none) != RKB_ERR_INVALID || none != NULL) return 1;
    if (rkb_chain_create(NULL, &none) != RKB_ERR_INVALID) return 1;
    printf("malformed descriptors are rejected with RKB_ERR_INVALID\n");
  }
  printf("ok (library version %d)\n", rkb_version());
  return 0;
}
