/*
 * ref_b200_binding.cpp -- the reference-side binding of INTEGRATION.md (option B), in its tested form.
 *
 * TEST INFRASTRUCTURE (drop-in proof).  oracle/Makefile links this file with the UNMODIFIED reference
 * tiny_api.cpp (setup, cache precompute, setters: all Eigen, all reference code) and WITHOUT the reference's
 * admm.cpp / rho_benchmark.cpp: the one symbol tiny_solve() needs from them, `int solve(TinySolver*)`
 * (tinympc/TinyMPC/src/tinympc/admm.hpp:9, called at tiny_api.cpp:322), is defined here and forwards the live
 * solver to the B200 C ABI (include/tinympc_b200.h).  The result, oracle/_ref/libtinympc_refhost_b200.so, is the
 * reference library with exactly its hot path swapped for the GPU; tests/test_gpu_dropin.py runs the
 * reference's golden cases through it.  There is no solver arithmetic in this file.
 */
#include <cstdio>

#include "tinympc/admm.hpp"
#include "tinympc/tiny_api.hpp"

#include "../include/tinympc_b200.h"

namespace {

tinympc_cuda_solver* g_cuda = nullptr;

void fill_family(const TinySolver* s, tinympc_cuda_family* f) {   // every Eigen matrix here is column-major double
    const TinyWorkspace* w = s->work;
    const TinyCache* c = s->cache;
    const TinySettings* st = s->settings;
    *f = tinympc_cuda_family{};
    f->nx = w->nx; f->nu = w->nu; f->N = w->N;
    f->Adyn = w->Adyn.data(); f->Bdyn = w->Bdyn.data(); f->fdyn = w->fdyn.data(); f->Q = w->Q.data(); f->R = w->R.data();
    f->rho = c->rho; f->Kinf = c->Kinf.data(); f->Pinf = c->Pinf.data(); f->Quu_inv = c->Quu_inv.data();
    f->AmBKt = c->AmBKt.data(); f->APf = c->APf.data(); f->BPf = c->BPf.data();
    f->dKinf_drho = c->dKinf_drho.size() ? c->dKinf_drho.data() : nullptr;
    f->dPinf_drho = c->dPinf_drho.size() ? c->dPinf_drho.data() : nullptr;
    f->abs_pri_tol = st->abs_pri_tol; f->abs_dua_tol = st->abs_dua_tol;
    f->max_iter = st->max_iter; f->check_termination = st->check_termination;
    f->en_state_bound = st->en_state_bound; f->en_input_bound = st->en_input_bound;
    f->en_state_soc = st->en_state_soc; f->en_input_soc = st->en_input_soc;
    f->en_state_linear = st->en_state_linear; f->en_input_linear = st->en_input_linear;
    f->adaptive_rho = st->adaptive_rho; f->adaptive_rho_min = st->adaptive_rho_min;
    f->adaptive_rho_max = st->adaptive_rho_max; f->adaptive_rho_enable_clipping = st->adaptive_rho_enable_clipping;
    f->x_min = w->x_min.size() ? w->x_min.data() : nullptr; f->x_max = w->x_max.size() ? w->x_max.data() : nullptr;
    f->u_min = w->u_min.size() ? w->u_min.data() : nullptr; f->u_max = w->u_max.size() ? w->u_max.data() : nullptr;
    f->numStateCones = w->numStateCones; f->numInputCones = w->numInputCones;
    f->Acx = w->Acx.data(); f->qcx = w->qcx.data(); f->cx = w->cx.data();
    f->Acu = w->Acu.data(); f->qcu = w->qcu.data(); f->cu = w->cu.data();
    f->numStateLinear = w->numStateLinear; f->numInputLinear = w->numInputLinear;
    f->Alin_x = w->Alin_x.data(); f->blin_x = w->blin_x.data(); f->Alin_u = w->Alin_u.data(); f->blin_u = w->blin_u.data();
}

template <class M> double* ptr_or_null(M& m) { return m.size() ? m.data() : nullptr; }

}  // namespace

extern "C" int solve(TinySolver* s) {   // replaces admm.cpp:274-389
    if (!g_cuda) {
        const int rc = tinympc_cuda_create(&g_cuda, nullptr, 0);
        if (rc) { std::fprintf(stderr, "ref_b200_binding: tinympc_cuda_create failed (%d): no CPU fallback\n", rc); return -1; }
    }
    tinympc_cuda_family f;
    fill_family(s, &f);
    if (tinympc_cuda_set_family(g_cuda, &f)) { std::fprintf(stderr, "ref_b200_binding: %s\n", tinympc_cuda_last_error(g_cuda)); return -1; }
    TinyWorkspace* w = s->work;
    tinympc_cuda_workspace ws{};
    ws.x = w->x.data(); ws.u = w->u.data(); ws.q = w->q.data(); ws.r = w->r.data(); ws.p = w->p.data(); ws.d = w->d.data();
    ws.v = w->v.data(); ws.vnew = w->vnew.data(); ws.z = w->z.data(); ws.znew = w->znew.data(); ws.g = w->g.data(); ws.y = w->y.data();
    ws.vcnew = ptr_or_null(w->vcnew); ws.zcnew = ptr_or_null(w->zcnew); ws.gc = ptr_or_null(w->gc); ws.yc = ptr_or_null(w->yc);
    ws.vlnew = ptr_or_null(w->vlnew); ws.zlnew = ptr_or_null(w->zlnew); ws.gl = ptr_or_null(w->gl); ws.yl = ptr_or_null(w->yl);
    ws.Xref = w->Xref.data(); ws.Uref = w->Uref.data();
    ws.rho = &s->cache->rho; ws.Kinf = s->cache->Kinf.data(); ws.Pinf = s->cache->Pinf.data();
    ws.sol_x = s->solution->x.data(); ws.sol_u = s->solution->u.data();
    ws.iter = &s->solution->iter; ws.status = &w->status; ws.solved = &s->solution->solved;
    double res[4] = {0, 0, 0, 0};
    ws.residuals = res;
    if (tinympc_cuda_solve_workspace(g_cuda, &ws)) { std::fprintf(stderr, "ref_b200_binding: %s\n", tinympc_cuda_last_error(g_cuda)); return -1; }
    w->iter = s->solution->iter;
    w->primal_residual_state = res[0]; w->dual_residual_state = res[1];
    w->primal_residual_input = res[2]; w->dual_residual_input = res[3];
    return s->solution->solved ? 0 : 1;   // admm.cpp:376 / :388
}
