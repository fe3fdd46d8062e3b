/* kte_oracle.h — TEST INFRASTRUCTURE.  Plain-C restatement of ReaK's KTE-chain forward-dynamics
 * path (see kte_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load libkte_oracle.so; nothing in reak_b200/ links or calls it. */
#ifndef KTE_ORACLE_H
#define KTE_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/reak_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void*  kto_create(const rkb_chain_desc* desc);
void   kto_destroy(void* h);
int    kto_eval(void* h, size_t n, const double* x, const double* u, double* xdot, int32_t* status);
int    kto_gen_forces(void* h, size_t n, const double* x, const double* u, double* f);
int    kto_gen_forces_qdd(void* h, const double* x, const double* u, const double* qdd, double* f);
int    kto_mass(void* h, size_t n, const double* x, double* M, double* Mdot);
int    kto_frames(void* h, const double* x, const double* u, double* out);
/* returns wall seconds (< 0 on failure); n_workers > 1 forks worker processes */
void kto_rk4_inputs(void* h, size_t n, const double* x0, const double* u_nodes, double dt, int n_steps, double* xout, int32_t* status);
double kto_rk4(void* h, size_t n, const double* x0, const double* u, double dt, int n_steps,
               double* xout, int32_t* status, int n_workers);
/* the same for any scheme of enum rkb_scheme (euler, midpoint, runge_kutta4, runge_kutta5) */
double kto_integrate(void* h, size_t n, const double* x0, const double* u, int scheme, double dt, int n_steps,
                     double* xout, int32_t* status, int n_workers);
/* steer_bounded_input of oracle/steer_law.h for `count` triples (test hook: pinned against IHAQR_topology::get_bounded_input) */
int    kto_bounded_input(int nu, double T, const double* lo, const double* hi, const double* dlo, const double* dhi, int count,
                         const double* u_prev, const double* u_bias, const double* u_corr, double* u_out);
/* closed-loop steering (oracle/steer_law.h); lo/hi/dlo/dhi may be NULL */
int    kto_steer_feedback(void* h, size_t n, const double* x0, const double* goal, const double* u_bias, const double* gain,
                          double* u_prev, double T, double dt, int substeps, int max_intervals, double proximity, int saturate_first,
                          const double* lo, const double* hi, const double* dlo, const double* dhi,
                          double* x_out, int32_t* n_done, double* traj, int32_t* status);
/* raw twist-shaping matrices of mass_matrix_calc::get_TMT_TdMT for one state: Tcm, Tcm_dot are
 * m x n row-major, Mcm m x m; returns m (rows) or < 0.  Pass NULL to query m only. */
int    kto_tmt(void* h, const double* x, double* Tcm, double* Mcm, double* Tcm_dot);
/* core/lin_alg/mat_cholesky.hpp pieces on a dense row-major n x n matrix (known-answer tests) */
int    kto_cholesky_solve(int n, const double* A, double* b, int nrhs, double tol);
int    kto_ldl_solve(int n, const double* A, double* b, int nrhs, double tol);

/* k nearest vertices per query within `radius` (min_dist_linear_search, topological_search.hpp:91-112, 238-270) */
int    kto_nearest(size_t n_vertices, const double* vertices, size_t n_queries, const double* queries, int dim, int k, double radius,
                   int32_t* index, double* distance, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif
