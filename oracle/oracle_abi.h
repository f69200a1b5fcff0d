/*
 * oracle_abi.h -- plain-C problem description shared by the two CPU checkers:
 *   - oracle/tinympc_oracle.c   ("port": a C restatement of the reference algorithm)
 *   - oracle/ref_driver.cpp     ("reference": a batch driver linked against the UNMODIFIED
 *                                reference sources compiled from /root/reference into oracle/_ref/)
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Conventions (follow the reference, tinympc/TinyMPC/src/tinympc/types.hpp:15-17, 86-187):
 *   matrices are column-major double; a trajectory "nx x N" is therefore N consecutive nx-vectors.
 *   Batched arrays are float32 on input (the same float32 values the GPU receives, widened to
 *   double) and double on output.
 */
#ifndef TINYMPC_ORACLE_ABI_H
#define TINYMPC_ORACLE_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int nx, nu, N;
    /* tiny_setup arguments (tiny_api.cpp:21-23); Q, R are the user's diagonals (before +rho) */
    const double *A;      /* nx*nx col-major */
    const double *B;      /* nx*nu col-major */
    const double *f;      /* nx */
    const double *Qdiag;  /* nx */
    const double *Rdiag;  /* nu */
    double rho;
    /* TinySettings (types.hpp:63-80) */
    double abs_pri_tol, abs_dua_tol;
    int max_iter, check_termination;
    int en_state_bound, en_input_bound;
    int en_state_soc, en_input_soc;
    int en_state_linear, en_input_linear;
    int adaptive_rho;
    double adaptive_rho_min, adaptive_rho_max;
    int adaptive_rho_enable_clipping;
    /* shared bounds, nx*N / nu*(N-1) col-major; may be NULL when per-problem bounds are given
       or the corresponding en_* flag is 0 */
    const double *x_min, *x_max, *u_min, *u_max;
    /* cones exactly as they must land in TinyWorkspace (work->Acx.. = "state", work->Acu.. = "input");
       any argument-order swap is the caller's business */
    int n_state_cones; const int *Acx; const int *qcx; const double *cx;
    int n_input_cones; const int *Acu; const int *qcu; const double *cu;
    /* linear inequality rows  Alin_x (n_state_lin x nx, col-major), blin_x ... */
    int n_state_lin; const double *Alin_x; const double *blin_x;
    int n_input_lin; const double *Alin_u; const double *blin_u;
    /* adaptive-rho sensitivities: 0 = none (zeros), 1 = the hard-coded quadrotor tables of
       tiny_initialize_sensitivity_matrices (tiny_api.cpp:411-472), 2 = explicit arrays below */
    int sens_mode;
    const double *dKinf, *dPinf, *dC1, *dC2;   /* nu*nx, nx*nx, nu*nu, nx*nx col-major */
} oracle_problem;

typedef struct {
    int batch;
    const float *x0;       /* batch*nx                      (required) */
    const float *Xref;     /* batch*nx*N     or NULL = zeros */
    const float *Uref;     /* batch*nu*(N-1) or NULL = zeros */
    /* optional per-problem bounds (override the shared ones when non-NULL) */
    const float *x_min, *x_max;   /* batch*nx*N     */
    const float *u_min, *u_max;   /* batch*nu*(N-1) */
} oracle_batch_in;

typedef struct {
    double *x;          /* batch*nx*N      solution->x  (= vnew, admm.cpp:370,386) */
    double *u;          /* batch*nu*(N-1)  solution->u  (= znew) */
    int    *iter;       /* batch           solution->iter */
    int    *status;     /* batch           work->status: 1 solved, 11 unsolved (admm.cpp:279,365) */
    double *residuals;  /* batch*4 or NULL: pri_state, dua_state, pri_input, dua_input (admm.cpp:257-260) */
    double *rho;        /* batch or NULL: cache->rho at exit (adaptive rho) */
} oracle_batch_out;

/* cache computed by tiny_precompute_and_set_cache (tiny_api.cpp:244-318); all arrays caller-allocated */
typedef struct {
    double *Kinf;     /* nu*nx */
    double *Pinf;     /* nx*nx */
    double *Quu_inv;  /* nu*nu */
    double *AmBKt;    /* nx*nx */
    double *APf;      /* nx */
    double *BPf;      /* nu */
    double *dKinf, *dPinf, *dC1, *dC2;  /* sensitivities as stored in the cache (may be NULL) */
} oracle_cache_out;

#ifdef __cplusplus
}
#endif
#endif
