// tmpc_gpp.cuh -- "lane group per problem" ADMM kernel for sm_100a, fp64: the low-latency form of the hot path.
//
// Path (reference: tinympc/TinyMPC/src/tinympc/admm.cpp): solve :274-389 with backward_pass_grad :13-20, forward_pass :25-32,
// update_slack (box) :81-98, update_dual :181-187, update_linear_cost :214-247, termination_condition :253-271 -- direct form,
// reference order (backward, forward, slack, dual, linear cost, check), double precision.
//
// Why a second mapping.  The thread-per-problem kernels (tmpc_tpp2/3.cuh) run a problem's ~9 000 instructions per ADMM iteration
// on ONE thread: unbeatable for throughput when a million problems are waiting, but an fp64 iteration then takes ~37 us of
// dependent issue, and the exact-count mode (fp32 pass + fp64 re-solve of the few problems whose termination decision is too
// close to call) ends with a tail of up to max_iter x 37 us = 3.7 ms during which the GPU is nearly idle.  Here a problem is
// spread over a group of GS = 8 / 16 lanes of a warp -- the north-star mapping (problem per warp, shuffle reductions):
//   * lane l < NX owns ROW l of the state trajectory, lane NX + a owns row a of the input trajectory; each keeps its row of
//     the slack, the dual and the reference term for ALL time steps in registers (the time loops are fully unrolled, so the
//     arrays are statically indexed);
//   * every lane holds its rows of the matrices in registers (2 x (NX + NU) doubles).  Substituting u = -Kinf x - d into the
//     rollout, both sweeps become ONE fused product per time step in which state and input lanes run the same instruction
//     stream on different coefficients:
//        forward :  state lane r: x_{i+1,r} = (A - B Kinf)_r . x_i - B_r . d_i + f_r      input lane a: u_{i,a} = -Kinf_a . x_i - d_{i,a}
//        backward:  state lane r: p_{i,r} = q_{i,r} + AmBKt_r . p_{i+1} - Kinf'_r . r_i + APf_r
//                   input lane a: d_{i,a} = (Quu_inv B')_a . p_{i+1} + Quu_inv_a . r_i + (Quu_inv BPf)_a
//     (NX + NU = 16 DFMA per lane and step, four independent chains; the products Quu_inv B' and A - B Kinf are formed once on
//     the host in double -- 1e-16 relative to the reference's order of operations);
//   * the vector a step produces (x_{i+1}, or p_i and r_{i-1}) is handed to the group through a double-buffered shared-memory
//     slot: one 8-byte store per lane, __syncwarp, NX/2 + NU/2 broadcast 16-byte loads;
//   * slack, dual, linear-cost and residual terms are element-wise on the owning lane; the four infinity norms are shuffle
//     max-reductions over the group (admm.cpp:257-260).
// One ADMM iteration of the quadrotor shape is 18 such steps of ~45 instructions: ~2 us, 20x shorter than the thread-per-problem
// fp64 iteration, and -- because the FP64 pipe sees four chains per lane and two problems per warp -- a higher fp64 throughput
// as well.  It serves fp64 batches (precision = 64), the second pass of the exact-count mode (index list or device queue,
// same SolveParams protocol as tmpc_tpp2.cuh) and small batches of the compiled shapes.
#pragma once
#include "tmpc_tpp2.cuh"

namespace tmpc {

template <int NX_, int NU_, int NH_, int GS_, int BLOCK_ = 128>
struct GppCfg {
    using T = double;
    static constexpr int NX = NX_, NU = NU_, NH = NH_, GS = GS_, BLOCK = BLOCK_;
    static_assert(GS_ == 8 || GS_ == 16 || GS_ == 32, "group size: 8, 16 or 32 lanes");
    static_assert(NX_ + NU_ <= GS_, "one lane per state row and input row");
    static_assert(NX_ % 2 == 0 || true, "");
    static constexpr int GPW = 32 / GS_;                 // problems per warp
    static constexpr int GPB = BLOCK_ / GS_;             // problems per CTA
    static constexpr int NV = NX_ + NU_;
    static constexpr int NXP = (NX_ + 1) & ~1, NUP = (NU_ + 1) & ~1;   // exchange slots padded to 16 bytes
    // shared memory per group (doubles): forward slot x (2 buffers), backward slot [p | r] (2 buffers), table of d (N-1 steps)
    static constexpr int oXB = 0, oPB = 2 * NXP, oDT = oPB + 2 * (NXP + NUP), GWORDS = oDT + (NH_ - 1) * NUP;
    static constexpr int SX = NX_ * NH_, SU = NU_ * (NH_ - 1);
};

// Per-lane tables, built on the host in double from the family's master pack and passed in the kernel-parameter space.
template <int NX, int NU, int NH, int GS>
struct alignas(16) GppTab {
    double CF1[GS][NX], CF2[GS][NU], cf0[GS];     // forward : acc = cf0 + CF1 . x_i + CF2 . d_i
    double CB1[GS][NX], CB2[GS][NU], cb0[GS];     // backward: acc = cb0 + CB1 . p_{i+1} + CB2 . r_i (+ q_i on state lanes)
    double wref[GS];                               // Qd_r / Rd_a: weight of the lane's reference term (work->Q, work->R)
    double lo[NH][GS], hi[NH][GS];                 // box of the element the lane handles in forward slot s (state: column s + 1; input: step s); row NH-1: column 0
    double Pinf[NX][NX];                           // row-major: terminal term -(xref_N' Pinf)'
};

template <int NX, int NU, int NH, int GS>
inline void fill_gpp_tab(GppTab<NX, NU, NH, GS>& t, const double* pk, const PackLayout& L, const SolveParams& prm) {
    std::memset(&t, 0, sizeof(t));
    const double* A = pk + L.A; const double* B = pk + L.B; const double* K = pk + L.Kinf; const double* AK = pk + L.AmBKt;
    const double* Qi = pk + L.Quu_inv;
    const double inf = 1.0 / 0.0;
    for (int l = 0; l < GS; ++l) for (int s = 0; s < NH; ++s) { t.lo[s][l] = -inf; t.hi[s][l] = inf; }
    for (int r = 0; r < NX; ++r) {   // state lanes
        for (int c = 0; c < NX; ++c) {
            double acc = A[r * NX + c];
            for (int a = 0; a < NU; ++a) acc -= B[r * NU + a] * K[a * NX + c];     // (A - B Kinf)[r][c]
            t.CF1[r][c] = acc;
            t.CB1[r][c] = AK[r * NX + c];
        }
        for (int a = 0; a < NU; ++a) { t.CF2[r][a] = -B[r * NU + a]; t.CB2[r][a] = -K[a * NX + r]; }
        t.cf0[r] = pk[L.f + r];
        t.cb0[r] = pk[L.APf + r];
        t.wref[r] = pk[L.Qd + r];
        if (prm.en_state_bound)
            for (int s = 0; s < NH; ++s) {
                const int col = s == NH - 1 ? 0 : s + 1;
                t.lo[s][r] = pk[L.xmin + col * NX + r]; t.hi[s][r] = pk[L.xmax + col * NX + r];
            }
    }
    for (int a = 0; a < NU; ++a) {   // input lanes
        const int l = NX + a;
        for (int c = 0; c < NX; ++c) {
            t.CF1[l][c] = -K[a * NX + c];
            double acc = 0.0;
            for (int b = 0; b < NU; ++b) acc += Qi[a * NU + b] * B[c * NU + b];   // (Quu_inv B')[a][c]
            t.CB1[l][c] = acc;
        }
        for (int b = 0; b < NU; ++b) { t.CF2[l][b] = (a == b) ? -1.0 : 0.0; t.CB2[l][b] = Qi[a * NU + b]; }
        double acc = 0.0;
        for (int b = 0; b < NU; ++b) acc += Qi[a * NU + b] * pk[L.BPf + b];
        t.cb0[l] = acc;
        t.wref[l] = pk[L.Rd + a];
        if (prm.en_input_bound)
            for (int s = 0; s < NH - 1; ++s) { t.lo[s][l] = pk[L.umin + s * NU + a]; t.hi[s][l] = pk[L.umax + s * NU + a]; }
    }
    for (int r = 0; r < NX; ++r) for (int c = 0; c < NX; ++c) t.Pinf[r][c] = pk[L.Pinf + r * NX + c];
}

template <class C>
__global__ void __launch_bounds__(C::BLOCK)
gpp_kernel(const __grid_constant__ SolveParams prm, const __grid_constant__ GppTab<C::NX, C::NU, C::NH, C::GS> tab) {
    using T = double;
    constexpr int NX = C::NX, NU = C::NU, NH = C::NH, GS = C::GS, NXP = C::NXP, NUP = C::NUP, SXL = C::SX, SUL = C::SU;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char gpp_smem[];
    T* const gbase = reinterpret_cast<T*>(gpp_smem) + (size_t)(threadIdx.x / GS) * C::GWORDS;
    T* const lo_t = reinterpret_cast<T*>(gpp_smem) + (size_t)C::GPB * C::GWORDS;   // [NH][GS] per CTA (identical for every group)
    T* const hi_t = lo_t + NH * GS;

    const int lane = threadIdx.x & 31;
    const int l = lane % GS;                       // role within the group
    const int gl0 = lane - l;                      // first lane of this group
    const bool is_x = l < NX, is_u = l >= NX && l < NX + NU;
    const int row = is_x ? l : (is_u ? l - NX : 0);
    const int lc = (is_x || is_u) ? l : 0;         // idle lanes compute lane 0's (discarded) values

    for (int e = threadIdx.x; e < NH * GS; e += C::BLOCK) { lo_t[e] = tab.lo[e / GS][e % GS]; hi_t[e] = tab.hi[e / GS][e % GS]; }
    __syncthreads();

    // ---- this lane's rows of the matrices
    T cf1[NX], cf2[NU], cb1[NX], cb2[NU];
#pragma unroll
    for (int c = 0; c < NX; ++c) { cf1[c] = tab.CF1[lc][c]; cb1[c] = tab.CB1[lc][c]; }
#pragma unroll
    for (int a = 0; a < NU; ++a) { cf2[a] = tab.CF2[lc][a]; cb2[a] = tab.CB2[lc][a]; }
    const T cf0 = tab.cf0[lc], cb0 = tab.cb0[lc], wref = tab.wref[lc];
    const T rho = prm.rho, tol_pri = prm.abs_pri_tol, tol_dua = prm.abs_dua_tol;
    const int max_iter = prm.max_iter, check_every = prm.check_termination;
    const int n_items = prm.batch_ptr ? min(*prm.batch_ptr, prm.batch) : prm.batch;
    const bool consumer = prm.q_tail != nullptr && prm.q_consume != 0;

    // ---- per-problem state of this lane: slot s = state column s + 1 / input step s; slot N-1 (state lanes) = column 0
    T G[NH], V[NH];          // dual (g / y) and slack (v / z) of the lane's row
    float RF[NH];            // reference term before weighting: Xref / Uref of the row (slot N-2 of a state lane is replaced by PT below)
    T PT = 0;                // state lanes: -(xref_N' Pinf)'_r
    T x0r = 0;               // state lanes: x0_r
#pragma unroll
    for (int s = 0; s < NH; ++s) { G[s] = 0; V[s] = 0; RF[s] = 0.f; }

    int prob = 0, k = 0, next_check = check_every, claim = 0, seen = 0, unpub = -1;
    bool active = false, exhausted = false, pending = false;
    T m0 = 0;                // 0 on a problem's first iteration (q = r = p = 0 on the cold workspace, tiny_api.cpp:68-105), then 1
    T res_px = 0, res_dx = 0, res_pu = 0, res_du = 0;

    T* const xb = gbase + C::oXB;    // [2][NXP]
    T* const pb = gbase + C::oPB;    // [2][NXP + NUP]
    T* const dt = gbase + C::oDT;    // [NH-1][NUP]

    for (;;) {
        // ------------------------------------------------------------------ refill idle groups (decisions are group-uniform)
        {
            if (l == 0) publish_done(prm, unpub);
            const bool want = !active && !exhausted && !pending;
            const unsigned mw = __ballot_sync(FULL, want && l == 0);
            if (mw) {
                const int leader = __ffs(mw) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(prm.work_counter, __popc(mw));
                base = __shfl_sync(FULL, base, leader);
                if (want) {
                    claim = base + __popc(mw & ((1u << gl0) - 1u));
                    if (consumer) {
                        pending = true;
                    } else if (claim >= n_items) {
                        exhausted = true;
                    } else {
                        if (prm.index_list) claim = __ldg(prm.index_list + claim);
                        pending = true;
                    }
                }
            }
            // group leader resolves the claim (queue ticket / arrival watermark), the group follows
            int ok = 0, qp = claim;
            if (l == 0 && pending) {
                if (consumer) {
                    bool none = false;
                    int e = 0;
                    if (queue_take(prm, claim, e, none)) { ok = 1; qp = e; }
                    else if (none) ok = 2;
                } else {
                    ok = problem_ready(prm, claim, seen) ? 1 : 0;
                }
            }
            ok = __shfl_sync(FULL, ok, gl0);
            qp = __shfl_sync(FULL, qp, gl0);
            if (pending && ok == 2) { pending = false; exhausted = true; }
            const bool mine = pending && ok == 1;
            if (mine) {
                prob = qp;
                pending = false; active = true;
                k = 0; next_check = check_every; m0 = 0;
                res_px = res_dx = res_pu = res_du = 0;
#pragma unroll
                for (int s = 0; s < NH; ++s) { G[s] = 0; V[s] = 0; RF[s] = 0.f; }
                PT = 0;
                if (is_x) {
                    x0r = static_cast<T>(__ldg(prm.x0 + (size_t)prob * NX + row));
                    if (prm.Xref) {
                        const float* src = prm.Xref + (size_t)prob * SXL + row;
#pragma unroll
                        for (int s = 0; s < NH - 1; ++s) RF[s] = __ldg(src + (s + 1) * NX);
                    }
                } else if (is_u) {
                    if (prm.Uref) {
                        const float* src = prm.Uref + (size_t)prob * SUL + row;
#pragma unroll
                        for (int s = 0; s < NH - 1; ++s) RF[s] = __ldg(src + s * NU);
                    }
                }
            }
            if (__any_sync(FULL, mine)) {
                // terminal term: PT_r = -(xref_N' Pinf)_r = -sum_c xref_N[c] Pinf[c][r]   (admm.cpp:238-240)
                if (mine && is_x) xb[row] = static_cast<T>(RF[NH - 2]);
                __syncwarp();
                if (mine && is_x) {
                    T acc = 0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc = fma(xb[c], tab.Pinf[c][row], acc);
                    PT = -acc;
                }
                __syncwarp();
            }
            if (!__any_sync(FULL, active)) {
                if (l == 0) publish_done(prm, unpub);
                if (!__any_sync(FULL, pending)) break;
                __nanosleep(256);
                continue;
            }
        }

        // weighted reference term of slot s: -(Xref .* Q) / -(Uref .* R); the terminal slot of a state lane holds PT instead
        auto refterm = [&](int s) -> T { return (s == NH - 2 && is_x) ? PT : -(static_cast<T>(RF[s]) * wref); };
        // w_s = q / r of slot s as update_linear_cost left it (admm.cpp:218-246): ref - rho (v - g); 0 on the cold workspace
        auto lincost = [&](int s) -> T { return m0 * fma(-rho, V[s] - G[s], refterm(s)); };

        // ------------------------------------------------------------------ backward_pass_grad (admm.cpp:13-20)
        {
            // p_{N-1} (state lanes) and r_{N-2} (input lanes) open the sweep
            const T w = lincost(NH - 2);
            T* buf = pb + ((NH - 1) & 1) * (NXP + NUP);
            if (is_x) buf[row] = w; else if (is_u) buf[NXP + row] = w;
            __syncwarp();
#pragma unroll
            for (int i = NH - 2; i >= 0; --i) {
                const T* src = pb + ((i + 1) & 1) * (NXP + NUP);
                T pv[NXP], rv[NUP];
#pragma unroll
                for (int c = 0; c < NXP / 2; ++c) { const double2 t2 = *reinterpret_cast<const double2*>(src + 2 * c); pv[2 * c] = t2.x; pv[2 * c + 1] = t2.y; }
#pragma unroll
                for (int a = 0; a < NUP / 2; ++a) { const double2 t2 = *reinterpret_cast<const double2*>(src + NXP + 2 * a); rv[2 * a] = t2.x; rv[2 * a + 1] = t2.y; }
                // four chains: three over p, one over r
                T a0 = cb0, a1 = 0, a2 = 0, a3 = 0;
                constexpr int T3 = (NX + 2) / 3;
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    if (c < T3) a0 = fma(cb1[c], pv[c], a0);
                    else if (c < 2 * T3) a1 = fma(cb1[c], pv[c], a1);
                    else a2 = fma(cb1[c], pv[c], a2);
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) a3 = fma(cb2[a], rv[a], a3);
                T acc = (a0 + a1) + (a2 + a3);
                if (i > 0) {
                    const T w1 = lincost(i - 1);          // state lane: q_i (column i = slot i-1); input lane: r_{i-1}
                    T* dst = pb + (i & 1) * (NXP + NUP);
                    if (is_x) dst[row] = acc + w1; else if (is_u) dst[NXP + row] = w1;
                }
                if (is_u) dt[i * NUP + row] = acc;         // d_i
                __syncwarp();
            }
        }

        // ------------------------------------------------------------------ forward_pass + update_slack + update_dual + residuals
        T rp = 0, rd = 0;
        auto element = [&](int s, T val) {   // the lane's element of slot s: vnew = clamp(x + g), g += x - vnew (admm.cpp:85-98, 184-187)
            const T lo = lo_t[s * GS + lc], hi = hi_t[s * GS + lc];
            const T t = val + G[s];
            const T vn = fmin(hi, fmax(lo, t));
            G[s] = t - vn;
            rp = fmax(rp, fabs(val - vn));
            rd = fmax(rd, fabs(V[s] - vn));
            V[s] = vn;
        };
        {
            if (is_x) { xb[row] = x0r; element(NH - 1, x0r); }   // column 0: x_0 = x0 (never rewritten, admm.cpp:25-32)
            __syncwarp();
#pragma unroll
            for (int s = 0; s < NH - 1; ++s) {
                const T* src = xb + (s & 1) * NXP;
                T xv[NXP], dv[NUP];
#pragma unroll
                for (int c = 0; c < NXP / 2; ++c) { const double2 t2 = *reinterpret_cast<const double2*>(src + 2 * c); xv[2 * c] = t2.x; xv[2 * c + 1] = t2.y; }
#pragma unroll
                for (int a = 0; a < NUP / 2; ++a) { const double2 t2 = *reinterpret_cast<const double2*>(dt + s * NUP + 2 * a); dv[2 * a] = t2.x; dv[2 * a + 1] = t2.y; }
                T a0 = cf0, a1 = 0, a2 = 0, a3 = 0;
                constexpr int T3 = (NX + 2) / 3;
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    if (c < T3) a0 = fma(cf1[c], xv[c], a0);
                    else if (c < 2 * T3) a1 = fma(cf1[c], xv[c], a1);
                    else a2 = fma(cf1[c], xv[c], a2);
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) a3 = fma(cf2[a], dv[a], a3);
                const T val = (a0 + a1) + (a2 + a3);      // x_{s+1,r} or u_{s,a}
                if (is_x && s < NH - 2) xb[((s + 1) & 1) * NXP + row] = val;
                element(s, val);
                if (s < NH - 2) __syncwarp();
            }
        }
        m0 = 1;
        k += 1;   // work->iter += 1 (admm.cpp:328)

        // ------------------------------------------------------------------ termination_condition (admm.cpp:253-271)
        bool finish = false;
        int st = 11;
        const bool chk = (k == next_check);              // group-uniform; the groups of a warp may differ
        if (__any_sync(FULL, chk)) {                     // the shuffles are executed by the whole warp
            T px = is_x ? rp : T(0), dx = is_x ? rd : T(0), pu = is_u ? rp : T(0), du = is_u ? rd : T(0);
#pragma unroll
            for (int o = GS / 2; o > 0; o >>= 1) {
                px = fmax(px, __shfl_xor_sync(FULL, px, o));
                dx = fmax(dx, __shfl_xor_sync(FULL, dx, o));
                pu = fmax(pu, __shfl_xor_sync(FULL, pu, o));
                du = fmax(du, __shfl_xor_sync(FULL, du, o));
            }
            if (chk) {
                next_check += check_every;
                res_px = px; res_dx = dx * rho; res_pu = pu; res_du = du * rho;
                if (res_px < tol_pri && res_pu < tol_pri && res_dx < tol_dua && res_du < tol_dua) { finish = true; st = 1; }
            }
        }
        if (k >= max_iter) finish = true;
        if (active && finish) {
            // solution = (vnew, znew) (admm.cpp:364-376, 384-388)
            if (is_x) {
                float* dst = prm.x + (size_t)prob * SXL + row;
                dst[0] = static_cast<float>(V[NH - 1]);
#pragma unroll
                for (int s = 0; s < NH - 1; ++s) dst[(s + 1) * NX] = static_cast<float>(V[s]);
            } else if (is_u) {
                float* dst = prm.u + (size_t)prob * SUL + row;
#pragma unroll
                for (int s = 0; s < NH - 1; ++s) dst[s * NU] = static_cast<float>(V[s]);
            }
            if (l == 0) {
                prm.iter[prob] = k;
                prm.status[prob] = st;
                if (prm.residuals) *reinterpret_cast<float4*>(prm.residuals + 4 * (size_t)prob) =
                    make_float4(static_cast<float>(res_px), static_cast<float>(res_dx), static_cast<float>(res_pu), static_cast<float>(res_du));
                if (prm.rho_out) prm.rho_out[prob] = static_cast<float>(rho);
                if (prm.done_counters) unpub = prob;
            }
            active = false;
        }
    }
}

template <class C>
inline size_t gpp_smem_bytes(int) {
    return ((size_t)C::GPB * C::GWORDS + 2 * (size_t)C::NH * C::GS) * sizeof(double);
}

}  // namespace tmpc
