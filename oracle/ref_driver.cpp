/*
 * ref_driver.cpp -- batch driver around the UNMODIFIED reference solver.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  This file contains no solver arithmetic:
 * it calls the reference's own tiny_setup / tiny_set_* / tiny_solve, which oracle/Makefile
 * compiles straight from /root/reference/tinympc/TinyMPC/src/tinympc/{admm,tiny_api,
 * rho_benchmark,codegen}.cpp into oracle/_ref/ (never copied into this repository).
 *
 * What it does per problem (SURVEY.md section 8c "cold-start reset"):
 *   zero the workspace exactly as tiny_setup leaves it (tiny_api.cpp:68-105), restore the pristine
 *   cache when adaptive rho is on, set x0 / Xref / Uref / bounds, call tiny_solve, copy out
 *   solution->x,u,iter and work->status.  std::cout is silenced because the reference prints a
 *   line on every converged solve (admm.cpp:373).
 */
#include <atomic>
#include <cstring>
#include <iostream>
#include <thread>
#include <vector>

#include "tinympc/tiny_api.hpp"
#include "tinympc/codegen.hpp"

#include "oracle_abi.h"

namespace {

tinyMatrix map_mat(const double* p, int r, int c) {
    if (!p) return tinyMatrix::Zero(r, c);
    return Eigen::Map<const tinyMatrix>(p, r, c);
}

tinyMatrix widen(const float* p, int r, int c) {
    tinyMatrix m(r, c);
    for (int j = 0; j < c; ++j)
        for (int i = 0; i < r; ++i) m(i, j) = static_cast<double>(p[(size_t)j * r + i]);
    return m;
}

struct Pristine {
    tinytype rho;
    tinyMatrix Kinf, Pinf, C1, C2;
};

/* Build one reference solver from the plain-C description, through the reference's public API
   only (tiny_api.hpp:10-50), then poke settings the way the reference's own examples do
   (quadrotor_hovering.cpp:54, rocket_landing_mpc.cpp:97-98). */
TinySolver* make_solver(const oracle_problem* d) {
    const int nx = d->nx, nu = d->nu, N = d->N;
    TinySolver* s = nullptr;
    tinyMatrix A = map_mat(d->A, nx, nx), B = map_mat(d->B, nx, nu), f = map_mat(d->f, nx, 1);
    tinyMatrix Q = tinyMatrix::Zero(nx, nx), R = tinyMatrix::Zero(nu, nu);
    for (int i = 0; i < nx; ++i) Q(i, i) = d->Qdiag[i];
    for (int i = 0; i < nu; ++i) R(i, i) = d->Rdiag[i];
    if (tiny_setup(&s, A, B, f, Q, R, d->rho, nx, nu, N, 0) != 0) return nullptr;

    if (d->x_min && d->x_max && d->u_min && d->u_max)
        tiny_set_bound_constraints(s, map_mat(d->x_min, nx, N), map_mat(d->x_max, nx, N),
                                   map_mat(d->u_min, nu, N - 1), map_mat(d->u_max, nu, N - 1));

    if (d->n_state_cones > 0 || d->n_input_cones > 0) {
        VectorXi Acx(d->n_state_cones), qcx(d->n_state_cones), Acu(d->n_input_cones), qcu(d->n_input_cones);
        tinyVector cx(d->n_state_cones), cu(d->n_input_cones);
        for (int k = 0; k < d->n_state_cones; ++k) { Acx(k) = d->Acx[k]; qcx(k) = d->qcx[k]; cx(k) = d->cx[k]; }
        for (int k = 0; k < d->n_input_cones; ++k) { Acu(k) = d->Acu[k]; qcu(k) = d->qcu[k]; cu(k) = d->cu[k]; }
        /* The DEFINITION's positional order is state-first (tiny_api.cpp:166-168); the description
           already says where each spec must land in the workspace. */
        tiny_set_cone_constraints(s, Acx, qcx, cx, Acu, qcu, cu);
    }
    if (d->n_state_lin > 0 || d->n_input_lin > 0) {
        tinyMatrix Ax = map_mat(d->Alin_x, d->n_state_lin, nx), Au = map_mat(d->Alin_u, d->n_input_lin, nu);
        tinyVector bx = map_mat(d->blin_x, d->n_state_lin, 1), bu = map_mat(d->blin_u, d->n_input_lin, 1);
        tiny_set_linear_constraints(s, Ax, bx, Au, bu);
    }

    TinySettings* st = s->settings;
    st->abs_pri_tol = d->abs_pri_tol;
    st->abs_dua_tol = d->abs_dua_tol;
    st->max_iter = d->max_iter;
    st->check_termination = d->check_termination;
    st->en_state_bound = d->en_state_bound;
    st->en_input_bound = d->en_input_bound;
    st->en_state_soc = d->en_state_soc;
    st->en_input_soc = d->en_input_soc;
    st->en_state_linear = d->en_state_linear;
    st->en_input_linear = d->en_input_linear;
    st->adaptive_rho = d->adaptive_rho;
    st->adaptive_rho_min = d->adaptive_rho_min;
    st->adaptive_rho_max = d->adaptive_rho_max;
    st->adaptive_rho_enable_clipping = d->adaptive_rho_enable_clipping;

    /* sensitivities (only these two routes ever fill cache->d*) */
    s->cache->dKinf_drho = tinyMatrix::Zero(nu, nx);
    s->cache->dPinf_drho = tinyMatrix::Zero(nx, nx);
    s->cache->dC1_drho = tinyMatrix::Zero(nu, nu);
    s->cache->dC2_drho = tinyMatrix::Zero(nx, nx);
    if (d->sens_mode == 1) {
        tiny_initialize_sensitivity_matrices(s);
    } else if (d->sens_mode == 2) {
        s->cache->dKinf_drho = map_mat(d->dKinf, nu, nx);
        s->cache->dPinf_drho = map_mat(d->dPinf, nx, nx);
        s->cache->dC1_drho = map_mat(d->dC1, nu, nu);
        s->cache->dC2_drho = map_mat(d->dC2, nx, nx);
    }
    return s;
}

void free_solver(TinySolver* s) {
    if (!s) return;
    delete s->solution; delete s->cache; delete s->settings; delete s->work; delete s;
}

/* state after tiny_setup, tiny_api.cpp:68-105 */
void cold_reset(TinySolver* s) {
    TinyWorkspace* w = s->work;
    w->x.setZero(); w->u.setZero(); w->q.setZero(); w->r.setZero(); w->p.setZero(); w->d.setZero();
    w->v.setZero(); w->vnew.setZero(); w->z.setZero(); w->znew.setZero(); w->g.setZero(); w->y.setZero();
    w->vc.setZero(); w->vcnew.setZero(); w->zc.setZero(); w->zcnew.setZero(); w->gc.setZero(); w->yc.setZero();
    w->vl.setZero(); w->vlnew.setZero(); w->zl.setZero(); w->zlnew.setZero(); w->gl.setZero(); w->yl.setZero();
    w->primal_residual_state = w->primal_residual_input = 0;
    w->dual_residual_state = w->dual_residual_input = 0;
    w->status = 0; w->iter = 0;
}

struct CoutSilencer {
    std::streambuf* old;
    CoutSilencer() : old(std::cout.rdbuf(nullptr)) {}
    ~CoutSilencer() { std::cout.rdbuf(old); std::cout.clear(); }
};

void solve_range(const oracle_problem* d, const oracle_batch_in* in, const oracle_batch_out* out,
                 std::atomic<int>* next, int chunk, std::atomic<int>* err) {
    const int nx = d->nx, nu = d->nu, N = d->N;
    TinySolver* s = make_solver(d);
    if (!s) { err->store(1); return; }
    Pristine pr{s->cache->rho, s->cache->Kinf, s->cache->Pinf, s->cache->C1, s->cache->C2};
    const size_t sx = (size_t)nx * N, su = (size_t)nu * (N - 1);
    for (;;) {
        int b0 = next->fetch_add(chunk);
        if (b0 >= in->batch) break;
        int b1 = b0 + chunk < in->batch ? b0 + chunk : in->batch;
        for (int b = b0; b < b1; ++b) {
            cold_reset(s);
            if (d->adaptive_rho) {
                s->cache->rho = pr.rho; s->cache->Kinf = pr.Kinf; s->cache->Pinf = pr.Pinf;
                s->cache->C1 = pr.C1; s->cache->C2 = pr.C2;
            }
            tiny_set_x0(s, widen(in->x0 + (size_t)b * nx, nx, 1));
            if (in->Xref) tiny_set_x_ref(s, widen(in->Xref + b * sx, nx, N)); else s->work->Xref.setZero();
            if (in->Uref) tiny_set_u_ref(s, widen(in->Uref + b * su, nu, N - 1)); else s->work->Uref.setZero();
            if (in->x_min) s->work->x_min = widen(in->x_min + b * sx, nx, N);
            if (in->x_max) s->work->x_max = widen(in->x_max + b * sx, nx, N);
            if (in->u_min) s->work->u_min = widen(in->u_min + b * su, nu, N - 1);
            if (in->u_max) s->work->u_max = widen(in->u_max + b * su, nu, N - 1);

            tiny_solve(s);

            if (out->x) std::memcpy(out->x + b * sx, s->solution->x.data(), sx * sizeof(double));
            if (out->u) std::memcpy(out->u + b * su, s->solution->u.data(), su * sizeof(double));
            if (out->iter) out->iter[b] = s->solution->iter;
            if (out->status) out->status[b] = s->work->status;
            if (out->residuals) {
                out->residuals[4 * (size_t)b + 0] = s->work->primal_residual_state;
                out->residuals[4 * (size_t)b + 1] = s->work->dual_residual_state;
                out->residuals[4 * (size_t)b + 2] = s->work->primal_residual_input;
                out->residuals[4 * (size_t)b + 3] = s->work->dual_residual_input;
            }
            if (out->rho) out->rho[b] = s->cache->rho;
        }
    }
    free_solver(s);
}

}  // namespace

extern "C" {

/* Cold-start solve of a whole batch on `threads` host threads (<=0: hardware_concurrency).
   Returns 0 on success. */
int ref_solve_batch(const oracle_problem* d, const oracle_batch_in* in, const oracle_batch_out* out, int threads) {
    if (!d || !in || !out || in->batch < 0) return 1;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > in->batch) threads = in->batch > 0 ? in->batch : 1;
    CoutSilencer quiet;
    std::atomic<int> next(0), err(0);
    int chunk = in->batch / (threads * 8);
    if (chunk < 1) chunk = 1;
    if (chunk > 256) chunk = 256;
    if (threads == 1) {
        solve_range(d, in, out, &next, chunk, &err);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(solve_range, d, in, out, &next, chunk, &err);
        for (auto& t : pool) t.join();
    }
    return err.load();
}

/* The cache tiny_setup computes for this description (tiny_api.cpp:126, 244-318). */
int ref_get_cache(const oracle_problem* d, oracle_cache_out* c) {
    CoutSilencer quiet;
    TinySolver* s = make_solver(d);
    if (!s) return 1;
    const int nx = d->nx, nu = d->nu;
    auto cp = [](double* dst, const tinyMatrix& m) { if (dst) std::memcpy(dst, m.data(), sizeof(double) * m.size()); };
    cp(c->Kinf, s->cache->Kinf); cp(c->Pinf, s->cache->Pinf); cp(c->Quu_inv, s->cache->Quu_inv);
    cp(c->AmBKt, s->cache->AmBKt); cp(c->APf, s->cache->APf); cp(c->BPf, s->cache->BPf);
    cp(c->dKinf, s->cache->dKinf_drho); cp(c->dPinf, s->cache->dPinf_drho);
    cp(c->dC1, s->cache->dC1_drho); cp(c->dC2, s->cache->dC2_drho);
    (void)nx; (void)nu;
    free_solver(s);
    return 0;
}

/* ---- warm-started session: the closed-loop pattern of quadrotor_hovering.cpp:73-93 ---- */
void* ref_session_create(const oracle_problem* d) {
    CoutSilencer quiet;
    return make_solver(d);
}
void ref_session_destroy(void* h) { free_solver(static_cast<TinySolver*>(h)); }
int ref_session_set_x0(void* h, const double* x0) {
    TinySolver* s = static_cast<TinySolver*>(h);
    return tiny_set_x0(s, map_mat(x0, s->work->nx, 1));
}
int ref_session_set_x_ref(void* h, const double* xr) {
    TinySolver* s = static_cast<TinySolver*>(h);
    return tiny_set_x_ref(s, map_mat(xr, s->work->nx, s->work->N));
}
int ref_session_set_u_ref(void* h, const double* ur) {
    TinySolver* s = static_cast<TinySolver*>(h);
    return tiny_set_u_ref(s, map_mat(ur, s->work->nu, s->work->N - 1));
}
/* returns tiny_solve's return value (0 converged / 1 max_iter) */
int ref_session_solve(void* h, double* x, double* u, int* iter, int* status, double* work_u0) {
    TinySolver* s = static_cast<TinySolver*>(h);
    CoutSilencer quiet;
    int rc = tiny_solve(s);
    const TinyWorkspace* w = s->work;
    if (x) std::memcpy(x, s->solution->x.data(), sizeof(double) * w->nx * w->N);
    if (u) std::memcpy(u, s->solution->u.data(), sizeof(double) * w->nu * (w->N - 1));
    if (iter) *iter = s->solution->iter;
    if (status) *status = w->status;
    if (work_u0) std::memcpy(work_u0, w->u.data(), sizeof(double) * w->nu);  /* work->u.col(0), the control the examples apply */
    return rc;
}

/* The reference's own code generator (codegen.cpp:68-80) on the solver this description sets up: writes
   <dir>/src/tiny_data.cpp, <dir>/tinympc/tiny_data.hpp, <dir>/src/tiny_main.cpp.  Used once, in this container, to produce the
   golden files tests/test_codegen.py compares the product's generator with (oracle/gen_golden.py). */
int ref_codegen(const oracle_problem* d, const char* dir) {
    CoutSilencer quiet;
    TinySolver* s = make_solver(d);
    if (!s) return 1;
    const int rc = tiny_codegen(s, dir, 0);
    free_solver(s);
    return rc;
}

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  /* extern "C" */
