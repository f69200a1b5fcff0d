/*
 * tinympc_oracle.h -- C restatement ("port") of the reference ADMM hot path, double precision.
 * TEST INFRASTRUCTURE ONLY; see oracle/README.md.  Exposes the same entry points as
 * oracle/ref_driver.cpp with the prefix port_ instead of ref_.
 */
#ifndef TINYMPC_ORACLE_H
#define TINYMPC_ORACLE_H
#include "oracle_abi.h"
#ifdef __cplusplus
extern "C" {
#endif

int port_solve_batch(const oracle_problem* d, const oracle_batch_in* in, const oracle_batch_out* out, int threads);
int port_get_cache(const oracle_problem* d, oracle_cache_out* c);

void* port_session_create(const oracle_problem* d);
void port_session_destroy(void* h);
int port_session_set_x0(void* h, const double* x0);
int port_session_set_x_ref(void* h, const double* xr);
int port_session_set_u_ref(void* h, const double* ur);
int port_session_solve(void* h, double* x, double* u, int* iter, int* status, double* work_u0);

#ifdef __cplusplus
}
#endif
#endif
