/* proximity_demo.c — plain C against the C-ABI (include/reak_b200.h): the planners' collision test on propagated states,
 * checked against closed-form distances.
 *
 *   (1) planar: the pendulum of ctrl/mbd_kte/test_bm.cpp with a circle (geom::circle) at its tip against a wall
 *       (geom::rectangle) fixed in the world — proxy_query_pair_2D::findMinimumDistance, prox_circle_rectangle.cpp;
 *   (2) spatial: a one-joint arm with a sphere at its tip above a floor (geom::plane) — proxy_query_pair_3D, prox_plane_sphere.cpp
 *       — first on the interpreter kernel, then on the kernel generated for this chain and pair (rkb_proxy_specialize),
 *       and the is_free test of manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/proximity_demo.c -Lreak_b200/lib -lreak_b200 -lm -o examples/proximity_demo
 *   LD_LIBRARY_PATH=reak_b200/lib examples/proximity_demo
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
static rkb_shape shape(int kind, int anchor, double d0, double d1, double d2) {
  rkb_shape s;
  memset(&s, 0, sizeof s);
  s.kind = kind; s.anchor = anchor;
  s.quat[0] = 1.0;  /* identity: (w, x, y, z) = (1, 0, 0, 0) in 3D, (cos, sin) = (1, 0) in 2D */
  s.dims[0] = d0; s.dims[1] = d1; s.dims[2] = d2;
  return s;
}

enum { N = 4096 };
static double x[N][2], dist[N], pts[N][6];
static int32_t finder[N], is_free[N];

int main(void) {
  int i;
  double worst;
  /* ---- (1) planar pendulum, circle of radius 0.05 at the tip, wall x in [0.7, 0.9], y in [-2, 2] -------------------- */
  {
    rkb_element el[3];
    rkb_chain_desc d;
    rkb_chain* pend = NULL;
    rkb_proxy* pair = NULL;
    rkb_shape tip = shape(RKB_SHAPE_CIRCLE, 2, 0.05, 0.0, 0.0);
    rkb_shape wall = shape(RKB_SHAPE_RECTANGLE, -1, 0.2, 4.0, 0.0);
    el[0] = element(RKB_REVOLUTE_2D, 0, 1, 0);
    el[1] = element(RKB_RIGID_LINK_2D, 1, 2, -1);
    el[1].p[0] = 0.5;                      /* offset (0.5, 0), angle 0 */
    el[2] = element(RKB_INERTIA_2D, 2, -1, -1);
    el[2].p[0] = 1.0;
    el[2].upstream = 1u;
    memset(&d, 0, sizeof d);
    d.dim = 2; d.n_elements = 3; d.n_frames = 3; d.n_coords = 1; d.base_frame = 0;   /* (base at the origin, angle 0) */
    d.elements = el;
    wall.position[0] = 0.8;
    CHECK(rkb_chain_create(&d, &pend));
    CHECK(rkb_proxy_create(pend, &tip, 1, &wall, 1, &pair));
    for (i = 0; i < N; ++i) { x[i][0] = -1.2 + 2.4 * i / N; x[i][1] = 0.0; }
    CHECK(rkb_min_distance(pend, pair, 0, N, &x[0][0], dist, finder, &pts[0][0], RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
    worst = 0.0;
    for (i = 0; i < N; ++i) {
      /* the tip stays left of the wall and within its height: nearest point on the face x = 0.7 */
      const double want = 0.7 - 0.5 * cos(x[i][0]) - 0.05;
      worst = fmax(worst, fabs(dist[i] - want));
      worst = fmax(worst, fabs(pts[i][3] - 0.7) + fabs(pts[i][4] - 0.5 * sin(x[i][0])));
      if (finder[i] != 0) { fprintf(stderr, "finder %d at %d\n", finder[i], i); return 1; }
    }
    printf("planar: circle on the pendulum's tip against a wall, max error over %d states: %.2e\n", N, worst);
    if (!(worst < 1e-12)) return 1;
    rkb_proxy_destroy(pair);
    rkb_chain_destroy(pend);
  }
  /* ---- (2) one-joint arm about e_y, 0.5 m link along x, sphere of radius 0.1 at the tip, floor at z = -0.3 ---------- */
  {
    rkb_element el[3];
    rkb_chain_desc d;
    rkb_chain* arm = NULL;
    rkb_proxy* pair = NULL;
    const rkb_proxy* pairs[1];
    rkb_shape tip = shape(RKB_SHAPE_SPHERE, 2, 0.1, 0.0, 0.0);
    rkb_shape floor_ = shape(RKB_SHAPE_PLANE, -1, 10.0, 10.0, 0.0);
    int n_blocked = 0, rc;
    el[0] = element(RKB_REVOLUTE_3D, 0, 1, 0);
    el[0].p[1] = 1.0;                      /* axis e_y */
    el[1] = element(RKB_RIGID_LINK_3D, 1, 2, -1);
    el[1].p[0] = 0.5;                      /* offset (0.5, 0, 0) */
    el[1].p[3] = 1.0;                      /* no rotation */
    el[2] = element(RKB_INERTIA_3D, 2, -1, -1);
    el[2].p[0] = 1.0;                      /* mass; tensor = identity */
    el[2].p[1] = 1.0; el[2].p[4] = 1.0; el[2].p[6] = 1.0;
    el[2].upstream = 1u;
    memset(&d, 0, sizeof d);
    d.dim = 3; d.n_elements = 3; d.n_frames = 3; d.n_coords = 1; d.base_frame = 0;
    d.base.quat[0] = 1.0;
    d.elements = el;
    floor_.position[2] = -0.3;
    CHECK(rkb_chain_create(&d, &arm));
    CHECK(rkb_proxy_create(arm, &tip, 1, &floor_, 1, &pair));
    pairs[0] = pair;
    for (i = 0; i < N; ++i) { x[i][0] = -3.0 + 6.0 * i / N; x[i][1] = 0.0; }
    CHECK(rkb_min_distance(arm, pair, 0, N, &x[0][0], dist, finder, NULL, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
    worst = 0.0;
    for (i = 0; i < N; ++i) worst = fmax(worst, fabs(dist[i] - (-0.5 * sin(x[i][0]) + 0.3 - 0.1)));
    printf("spatial: sphere on the arm's tip above a floor, interpreter kernel, max error: %.2e\n", worst);
    if (!(worst < 1e-12)) return 1;
    /* the same query as CUDA generated for this chain and pair (needs libnvrtc.so.12; optional) */
    rc = rkb_proxy_specialize(pair, 0);
    if (rc == RKB_ERR_UNSUPPORTED) {
      printf("spatial: libnvrtc.so.12 not installed, the interpreter kernel keeps serving the pair\n");
    } else {
      CHECK(rc);
      CHECK(rkb_min_distance(arm, pair, 0, N, &x[0][0], dist, finder, NULL, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
      worst = 0.0;
      for (i = 0; i < N; ++i) worst = fmax(worst, fabs(dist[i] - (-0.5 * sin(x[i][0]) + 0.3 - 0.1)));
      printf("spatial: generated kernel (rkb_proxy_specialize), max error: %.2e, specialised: %d\n", worst, rkb_proxy_is_specialized(pair));
      if (!(worst < 1e-12) || !rkb_proxy_is_specialized(pair)) return 1;
    }
    CHECK(rkb_is_free(arm, 0, N, &x[0][0], pairs, 1, is_free, RKB_MEM_HOST | RKB_LAYOUT_AOS, NULL));
    for (i = 0; i < N; ++i) {
      const int want = !(-0.5 * sin(x[i][0]) + 0.2 < 0.0);
      if (fabs(-0.5 * sin(x[i][0]) + 0.2) > 1e-9 && is_free[i] != want) { fprintf(stderr, "is_free[%d] = %d\n", i, is_free[i]); return 1; }
      n_blocked += !is_free[i];
    }
    printf("spatial: is_free agrees with sin q <= 0.4 on every state (%d of %d in collision)\n", n_blocked, N);
    if (n_blocked == 0 || n_blocked == N) return 1;
    rkb_proxy_destroy(pair);
    rkb_chain_destroy(arm);
  }
  printf("ok\n");
  return 0;
}
