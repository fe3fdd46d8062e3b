/* steer_law.h — TEST INFRASTRUCTURE.  Restatement of the feedback law inside the reference's steering
 * loops, shared by the two checkers (kte_oracle.c and ref_lib.cpp); nothing in reak_b200/ includes it.
 *
 *   IHAQR_topology::get_bounded_input            examples/misc/IHAQR_topology.hpp:304-327
 *   hyperbox_topology::bring_point_in_bounds     ctrl/topologies/hyperbox_topology.hpp:114-128
 *   hyperbox_topology::is_in_bounds              ctrl/topologies/hyperbox_topology.hpp:178-189
 *   the loop around them                         examples/misc/MEAQR_topology.hpp:503-561 (steer_with_constant_control),
 *                                                examples/misc/IHAQR_topology.hpp:349-378 (move_position_toward_impl)
 *
 * PARITY: PINNED against the reference's own classes (oracle/ref_steer_law.cpp instantiates the unmodified
 * examples/misc/IHAQR_topology.hpp and MEAQR_topology.hpp over the live kte_nl_system and reaches their protected members
 * through derived classes; tests/test_oracle.py):
 *   steer_bounded_input  == IHAQR_topology::get_bounded_input, bit for bit over every regime (request inside / outside the
 *                           box, bias outside, bisection, rate limit acting or not, values on the faces of the box);
 *   the loop, every interval saturated      == IHAQR_topology::move_position_toward_impl (end state, 1e-16);
 *   the loop, first interval unsaturated    == MEAQR_topology::steer_with_constant_control with H = I, eta = 0
 *                                              (end state, last input, number of intervals; with_collision_check = true
 *                                              as well, its virtual is_free_impl answered by the reference's proxy pairs:
 *                                              collision flag and the state the loop stops on).
 * The reference integrates an interval with a time-driven loop (T / 100 resp. T / 10 steps, 100 or 101 resp. 10 or 11 of
 * them depending on rounding); the batch call takes the step count explicitly, and the tests use a T for which the
 * reference's count is the same in every interval.
 */
#ifndef RKB_ORACLE_STEER_LAW_H
#define RKB_ORACLE_STEER_LAW_H

#include <math.h>

#define STEER_MAX_INPUTS 16

/* lo/hi may be NULL: unbounded.  (lower_corner < upper_corner is required by the batch API.) */
static inline void steer_clamp(int nu, const double* lo, const double* hi, double* a) {
  int i;
  if (!lo || !hi) return;
  for (i = 0; i < nu; ++i) {
    if (a[i] < lo[i]) a[i] = lo[i];
    else if (a[i] > hi[i]) a[i] = hi[i];
  }
}
static inline int steer_in_bounds(int nu, const double* lo, const double* hi, const double* a) {
  int i;
  if (!lo || !hi) return 1;
  for (i = 0; i < nu; ++i)
    if ((a[i] < lo[i]) || (a[i] > hi[i])) return 0;
  return 1;
}

/* u_out = get_bounded_input(u_prev, u_bias, u_correction), IHAQR_topology.hpp:304-327 */
static inline void steer_bounded_input(int nu, double T, const double* lo, const double* hi, const double* dlo, const double* dhi,
                                       const double* u_prev, const double* u_bias_in, const double* u_corr_in, double* u_out) {
  double u_bias[STEER_MAX_INPUTS], u_corr[STEER_MAX_INPUTS], u_cur[STEER_MAX_INPUTS], du[STEER_MAX_INPUTS];
  int i, j;
  for (i = 0; i < nu; ++i) { u_bias[i] = u_bias_in[i]; u_corr[i] = u_corr_in[i]; }
  steer_clamp(nu, lo, hi, u_bias);                                     /* :306 */
  for (i = 0; i < nu; ++i) u_cur[i] = u_bias[i] + u_corr[i];            /* :308 */
  if (steer_in_bounds(nu, lo, hi, u_cur)) {                             /* :309-313 */
    for (i = 0; i < nu; ++i) du[i] = (u_cur[i] - u_prev[i]) * (1.0 / T);
    steer_clamp(nu, dlo, dhi, du);
    for (i = 0; i < nu; ++i) u_out[i] = u_prev[i] + T * du[i];
    return;
  }
  for (j = 0; j < 10; ++j) {                                            /* :315-322 */
    for (i = 0; i < nu; ++i) { u_corr[i] *= 0.5; u_cur[i] -= u_corr[i]; }
    if (steer_in_bounds(nu, lo, hi, u_cur)) {
      for (i = 0; i < nu; ++i) { u_bias[i] = u_cur[i]; u_cur[i] += u_corr[i]; }
    }
  }
  for (i = 0; i < nu; ++i) du[i] = (u_bias[i] - u_prev[i]) * (1.0 / T);  /* :324-326 */
  steer_clamp(nu, dlo, dhi, du);
  for (i = 0; i < nu; ++i) u_out[i] = u_prev[i] + T * du[i];
}

/* One pass of the loop head: returns 0 when the loop would stop (goal reached), else 1 with the
 * input of the next interval in u_out.  gain is nu x nx row-major; correction = -gain (x - goal). */
static inline int steer_next_input(int nx, int nu, double T, double proximity, int first_unsaturated,
                                   const double* lo, const double* hi, const double* dlo, const double* dhi,
                                   const double* x, const double* goal, const double* u_bias, const double* gain,
                                   const double* u_prev, double* u_out) {
  double corr[STEER_MAX_INPUTS], d2 = 0.0;
  int i, k;
  for (k = 0; k < nx; ++k) d2 += (x[k] - goal[k]) * (x[k] - goal[k]);
  if (!(sqrt(d2) > proximity)) return 0;                                /* MEAQR_topology.hpp:513-514 */
  for (i = 0; i < nu; ++i) {
    double s = 0.0;
    for (k = 0; k < nx; ++k) s += gain[i * nx + k] * (x[k] - goal[k]);
    corr[i] = -s;
  }
  if (first_unsaturated) {                                              /* MEAQR_topology.hpp:521-522 */
    for (i = 0; i < nu; ++i) u_out[i] = u_bias[i] + corr[i];
  } else {
    steer_bounded_input(nu, T, lo, hi, dlo, dhi, u_prev, u_bias, corr, u_out);
  }
  return 1;
}

#endif
