// tmpc_wpp.h -- host-visible interface of the warp-per-problem kernel (tmpc_wpp.cu)
#pragma once
#include <cuda_runtime.h>

#include "tmpc_common.h"

namespace tmpc {

// Element offsets of one problem's explicit workspace (the TinyWorkspace members, types.hpp:86-187)
struct WppLayout {
    int x, u, q, r, p, d, v, vnew, z, znew, g, y;
    int vcnew, zcnew, gc, yc, vlnew, zlnew, gl, yl;
    int tmp, scalars;      // scalars: [0] rho, [1] iter, [2] status, [3] pri_state, [4] dua_state, [5] pri_input, [6] dua_input, [7] solved
    int zero_end;          // [0, zero_end) is zeroed on a cold start
    int Xref, Uref, xmin, xmax, umin, umax;
    int Kinf, Pinf;        // per-problem cache copies (adaptive rho mutates them), row-major
    int size;
    static WppLayout make(int nx, int nu, int N);
};

// explicit_workspace = 1: `scratch` holds ONE fully initialised workspace that is iterated in place
// (tiny_solve semantics); 0: cold start per problem from p.x0/Xref/Uref, `warps` scratch workspaces;
// 2: `scratch` holds p.batch persistent workspaces (a session of warm-started solvers), each iterated in place.
template <typename T>
cudaError_t wpp_launch(const SolveParams& p, const PackLayout& L, const void* pack, const WppLayout& W, void* scratch, int warps,
                       int explicit_workspace, cudaStream_t st);

// session helpers: cold workspaces for `batch` solvers; x0 <- A x0 + B u0 + f on every workspace
template <typename T>
cudaError_t wpp_session_init(const SolveParams& p, const PackLayout& L, const void* pack, const WppLayout& W, void* wsp, int batch, cudaStream_t st);
template <typename T>
cudaError_t wpp_session_step(const PackLayout& L, const void* pack, const WppLayout& W, void* wsp, int batch, int use_solution, cudaStream_t st);

}  // namespace tmpc
