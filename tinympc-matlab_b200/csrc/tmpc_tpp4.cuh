// tmpc_tpp4.cuh -- batched ADMM throughput kernel for sm_100a, MIXED PRECISION: the rocket family (box + one second-order cone
// per side + at most one linear inequality per side), incremental form, fp32 Riccati increments, fp64 state.
//
// Same path and mapping as tmpc_tpp3.cuh (one thread = one problem, increments through the Riccati sweeps, lane refill,
// streamed-pipeline and exact-count hooks); reference: admm.cpp:274-389.  What changes is WHERE the precision sits.
//
// Why.  The rocket's thrusts are ~100 and its positions ~20, its dual tolerance is 1e-4 (rocket_landing_constraints.m): a
// float32 iterate carries 4e-6 of representation error, 4-8 % of that tolerance, and the fp32 kernel of round 1 flipped 3.3 %
// of the iteration counts and drifted 2e-4 away from the reference's thrusts.  profiles/tools/precision_lab.py locates the
// noise: it is the STORED state -- the iterates x, u that accumulate the increments, and the duals of the cone family -- not the
// arithmetic of the increments.  With x, u (and the cone duals) held in double and everything formed from them in double, while
// the Riccati sweeps keep running on float32 increments, the model reproduces 6000 / 6000 reference iteration counts.
//
// So: the increments dx, du, dq, dr, dp, dd and the mat-vecs on them are float32 (packed FFMA2, FMA pipe); the iterates x, u,
// the cone duals and the half-space multipliers are float64, and the per-element slack / dual / residual work runs on the FP64
// pipe (DADD / DSETP, 1:2 rate on B200, idle in the fp32 kernels), concurrently with the mat-vecs of the neighbouring columns.
//
// State per problem and family.  Instead of the pre-projection slack t = x + g_prev of tmpc_tpp3.cuh the kernel stores the
// OLDER DUAL s = g(k-1) next to x(k) (t = x + s exactly when s = 0, which is the common case; nothing of x's low-order bits is
// lost in a sum):
//     t_old = x_old + s,  v_old = P(t_old),  g = t_old - v_old,  t_new = x_new + g,  v_new = P(t_new),  s <- g
//   box        s float32 (an active box projects to a constant, so the rounding of its dual never reaches the test),
//              primal residual a = x_new - v_new, dual residual e = v_new - v_old, increment of (slack - dual) = e - a
//   cone       s float64 on the cone's coordinates only (the family's other coordinates have v = x, dual 0),
//              increment = 2 v_new - v_old - x_new        (t_new - t_old = x_new - v_old identically)
//   half-space one row a'v <= b: the dual is lambda a, only the multiplier is stored (float64);
//              lambda_o = max(0, (a'x_old - b)/|a|^2 + lambda_oo), lambda_n = max(0, (a'x_new - b)/|a|^2 + lambda_o),
//              increment = dx + (3 lambda_o - 2 lambda_n - lambda_oo) a
// The forward pass does not touch x: it leaves the increments dx_i (tensor memory) and du_i (in the slot of -dd_i it has just
// consumed) for the sweep, which forms x_new = x_old + dx in double.  x_0 is the problem's x0 and is never stored.
// Tensor memory per thread: columns 1 .. N-1 of [dx | x | cone dual | multiplier] + the input-side multipliers (252 words for
// the rocket, 8 warps per SM); shared memory: u, input cone dual (double), box duals, -dd / du (222 words).
// The reference terms -(Xref .* Q), -(Uref .* R), -(xref_N' Pinf)' enter the first sweep only: they are parked in the box-dual
// columns at refill (the cold dual is 0 by definition), as in tmpc_tpp3.cuh.
#pragma once
#include "tmpc_tpp3.cuh"

namespace tmpc {

template <int NX_, int NU_, int NH_, int BLOCK_, bool REFS_, bool FB_, bool AFF_, int SCS_, int SCD_, int UCS_, int UCD_, int NSL_, int NIL_>
struct Tpp4Cfg {
    using T = float;
    static constexpr int NX = NX_, NU = NU_, NH = NH_, FEAT = FEAT_CONSTR, BLOCK = BLOCK_, MINB = 1;
    static constexpr bool CONSTR = true;
    static constexpr int SCS = SCS_, SCD = SCD_, UCS = UCS_, UCD = UCD_, NSL = NSL_, NIL = NIL_;
    static_assert(SCS_ >= 0 && SCS_ + SCD_ <= NX_ && UCS_ >= 0 && UCS_ + UCD_ <= NU_, "cone block outside the vector");
    static_assert(NSL_ <= 1 && NIL_ <= 1, "one linear row per side (the multiplier form of the half-space dual)");
    static constexpr int REFMODE = REFS_ ? REFS_STATE : REFS_NONE;
    static constexpr bool REFS = REFS_, PPB = false, FB = FB_, AFF = AFF_, HYB = false;
    static constexpr int SX = NX * NH, SU = NU * (NH - 1);
    using CPack = ConstPack3<NX_, NU_, NH_, NSL_, NIL_>;
    // tensor memory, 32-bit words per thread: state columns 1 .. N-1 of [dx (NX) | x (2 NX) | cone dual (2 SCD) | multiplier (2 NSL)],
    // then the input-side multipliers of steps 0 .. N-2
    static constexpr int oDX = 0, oX = NX, oSC = 3 * NX, oSL = 3 * NX + 2 * SCD;
    static constexpr int CW = 3 * NX + 2 * SCD + 2 * NSL;
    static constexpr int oSLU = (NH - 1) * CW;
    static constexpr int TM_COLS_PER_THREAD = oSLU + 2 * NIL * (NH - 1);
    static_assert(((BLOCK_ / 32 + 3) / 4) * TM_COLS_PER_THREAD <= 512, "the state does not fit the 512 tensor-memory columns");
    // shared memory, float columns per thread: doubles first (u, input cone dual), then the box duals (state incl. column 0, input)
    // and -dd / du
    static constexpr int DCOLS = (NU + UCD) * (NH - 1);
    static constexpr int oU = 0, oSCU = NU * (NH - 1);                    // in double columns
    static constexpr int oSB = 2 * DCOLS, oSBU = oSB + SX, oND = oSBU + SU;   // in float columns
    static constexpr int COLS = oND + SU;
};

// double <-> two tensor-memory words
__device__ __forceinline__ double dbl_of(uint32_t lo, uint32_t hi) { return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo)); }
__device__ __forceinline__ void dbl_to(double v, uint32_t& lo, uint32_t& hi) { lo = static_cast<uint32_t>(__double2loint(v)); hi = static_cast<uint32_t>(__double2hiint(v)); }

// Second-order-cone projection (admm.cpp:39-60) of D coordinates in double, the reference's types: u0 double, the norm rounded
// to float (`float a = u1.norm()`), a / mu a float division, the scale factor double.
template <int D>
__device__ __forceinline__ void project_soc_d(double (&v)[D > 0 ? D : 1], float mu) {
    if constexpr (D > 0) {
        double ss = 0.0;
#pragma unroll
        for (int e = 0; e < D - 1; ++e) ss = fma(v[e], v[e], ss);
        const double u0 = v[D - 1] * static_cast<double>(mu);
        // float(sqrt(ss)): float estimate + one Newton step in double (the estimate is good to 1e-7, the step squares that)
        const float sf = static_cast<float>(ss);
        float y0f;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(y0f) : "f"(sf));
        float r0f;
        asm("rcp.approx.f32 %0, %1;" : "=f"(r0f) : "f"(y0f));
        const double y0 = static_cast<double>(y0f);
        const double y1 = (ss > 1e-30) ? fma(fma(-y0, y0, ss), static_cast<double>(0.5f * r0f), y0) : 0.0;
        const float a = static_cast<float>(y1);
        const double ad = static_cast<double>(a);
        const bool zero = ad <= -u0, inside = ad <= u0;
        // outside: 0.5 (1 + u0 / a) [u1 ; a / mu]
        // 1 / a: the float reciprocal of the estimate, one Newton step in double (1e-14 relative; a true division costs ~25 instructions)
        const double r0 = static_cast<double>(r0f);
        const double ra = r0 * fma(-ad, r0, 2.0);
        const double fct = 0.5 * fma(u0, ra, 1.0);
        const double lastv = fct * static_cast<double>(a / mu);
#pragma unroll
        for (int e = 0; e < D - 1; ++e) v[e] = zero ? 0.0 : (inside ? v[e] : fct * v[e]);
        v[D - 1] = zero ? 0.0 : (inside ? v[D - 1] : lastv);
    }
}

template <class C>
__global__ void __launch_bounds__(C::BLOCK, 1)
tpp4_kernel(const __grid_constant__ SolveParams prm, const __grid_constant__ typename C::CPack cp) {
    using T = float;
    using P = float2;
    using SP = StaticPack<C::NX, C::NU, C::NH>;
    using VX = Vec<T, C::NX>;
    using VU = Vec<T, C::NU>;
    constexpr int NX = C::NX, NU = C::NU, NH = C::NH, BLOCK = C::BLOCK, SXL = C::SX, SUL = C::SU, CW = C::CW;
    constexpr int NXP = pad2(NX), NUP = pad2(NU);
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t pack_bar;
    __shared__ uint32_t tmem_base_s;
    T* pack = reinterpret_cast<T*>(smem_raw);
    const uint32_t pack_bytes = static_cast<uint32_t>(prm.pack_elems) * sizeof(T);

    if (threadIdx.x == 0) {
        mbar_init(&pack_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&pack_bar, pack_bytes);
        tma_bulk_g2s(pack, prm.pack, pack_bytes, &pack_bar);
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    mbar_wait(&pack_bar, 0);

    T* cta_cols = pack + ((prm.pack_elems + 31) & ~31);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t w_id = static_cast<uint32_t>(tid) >> 5;
    const uint32_t tm_base = tmem_base_s + ((32u * (w_id & 3u)) << 16) + (w_id >> 2) * C::TM_COLS_PER_THREAD;
    auto col_base = [&](int i) { return tm_base + static_cast<uint32_t>((i - 1) * CW); };   // state column i >= 1
    // shared-memory columns of this thread
    double* dcol = reinterpret_cast<double*>(cta_cols) + tid;                 // double column c at dcol[c * BLOCK]
    T* fcol = cta_cols + tid;                                                 // float column c at fcol[c * BLOCK]
    auto U_at = [&](int i, int a) -> double& { return dcol[(C::oU + i * NU + a) * BLOCK]; };
    auto SCU_at = [&](int i, int c) -> double& { return dcol[(C::oSCU + i * (C::UCD > 0 ? C::UCD : 1) + c) * BLOCK]; };
    auto SB_at = [&](int i, int e) -> T& { return fcol[(C::oSB + i * NX + e) * BLOCK]; };
    auto SBU_at = [&](int i, int a) -> T& { return fcol[(C::oSBU + i * NU + a) * BLOCK]; };
    Traj<T, NU, NH - 1, C::oND, BLOCK> ND(cta_cols, tid);                     // -dd of the last sweep / du of this forward pass

    const float mu_x = prm.cx[0], mu_u = prm.cu[0];
    const bool lin_x = prm.en_state_linear != 0, lin_u = prm.en_input_linear != 0;   // an enabled family without rows still adds x to the cost
    const T* cP = pack + SP::Pinf;
    const T rho0 = static_cast<T>(prm.rho);
    const double tol_pri = prm.abs_pri_tol, tol_dua = prm.abs_dua_tol, rho_d = prm.rho;
    const int max_iter = prm.max_iter, check_every = prm.check_termination;
    // half-space rows in double (loop invariant)
    double ax[C::NSL > 0 ? NX : 1], au[C::NIL > 0 ? NU : 1];
    double bx = 0, ix = 0, bu = 0, iu = 0;
    if constexpr (C::NSL > 0) {
#pragma unroll
        for (int e = 0; e < NX; ++e) ax[e] = static_cast<double>(cp.Alx[e]);
        bx = static_cast<double>(cp.blx[0]); ix = static_cast<double>(cp.inx[0]);
    }
    if constexpr (C::NIL > 0) {
#pragma unroll
        for (int e = 0; e < NU; ++e) au[e] = static_cast<double>(cp.Alu[e]);
        bu = static_cast<double>(cp.blu[0]); iu = static_cast<double>(cp.inu[0]);
    }

    const int n_items = prm.batch_ptr ? min(*prm.batch_ptr, prm.batch) : prm.batch;
    const bool producer = prm.q_tail != nullptr;
    int prob = 0;
    bool active = false, exhausted = false, pending = false;
    int last_k = 32, claim = 0, seen = 0, unpub = -1, k = 0;
    int next_check = check_every;
    double res_px = 0, res_dx = 0, res_pu = 0, res_du = 0;
    VX x0v;
    x0v.fill(T(0));

    auto xlo = [&](int i, int e) { return static_cast<double>(cp.xmin[i * NXP + e]); };
    auto xhi = [&](int i, int e) { return static_cast<double>(cp.xmax[i * NXP + e]); };
    auto ulo = [&](int i, int a) { return static_cast<double>(cp.umin[i * NUP + a]); };
    auto uhi = [&](int i, int a) { return static_cast<double>(cp.umax[i * NUP + a]); };
    auto clampd = [](double t, double lo, double hi) { return fmin(hi, fmax(lo, t)); };
    // Box projection as its DUAL part: g = t - clamp(t) = (t - hi) above the box, (t - lo) below it, 0 inside.  The two
    // differences are DADDs, the decisions are sign tests on their upper words (integer pipe): a double-precision compare
    // (DSETP) runs at a quarter of the DADD rate on this part (profiles/microbench/fp64_bench.cu: 14.6 vs 62.7 per clock and SM).
    auto box_dual = [](double t, double lo, double hi) {
        const double dU = t - hi, dL = t - lo;
        const bool ab = __double2hiint(dU) >= 0, be = __double2hiint(dL) < 0;
        return ab ? dU : (be ? dL : 0.0);
    };
    // running maximum of |v| on (upper word, lower word): the upper words of non-negative doubles order like integers; ties in
    // the upper word (values within 2^-20 of each other) keep the first comer -- 1e-6 relative on a residual, nothing next to the
    // band of the exact-count mode
    auto amax_w = [](int& mh, int& ml, double v) {
        const int h = __double2hiint(v) & 0x7fffffff;
        const bool t = h > mh;
        mh = t ? h : mh;
        ml = t ? __double2loint(v) : ml;
    };

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes (tmpc_tpp3.cuh)
        {
            const bool want = !active && !exhausted && !pending;
            unsigned mw = __ballot_sync(FULL, want);
            if (mw && __any_sync(FULL, active)) {
                int m = prm.refill_min;
                if (m <= 0) {
                    const int sum_k = __reduce_add_sync(FULL, last_k);
                    m = __float2int_rn(sqrtf(__fdividef(471.f * 32.f, (float)max(sum_k, 32))));
                    m = min(max(m, 1), 12);
                }
                if (__popc(mw) < m) mw = 0;
            }
            if (mw) {
                const int leader = __ffs(mw) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(prm.work_counter, __popc(mw));
                base = __shfl_sync(FULL, base, leader);
                if (want) {
                    claim = base + __popc(mw & ((1u << lane) - 1u));
                    if (claim >= n_items) { exhausted = true; prob = 0; }
                    else { if (prm.index_list) claim = __ldg(prm.index_list + claim); pending = true; }
                }
            }
            const bool mine = pending && problem_ready(prm, claim, seen);
            const unsigned m = __ballot_sync(FULL, mine);
            if (m) {
                const bool have_xref = C::REFS && prm.Xref != nullptr, have_uref = C::REFS && prm.Uref != nullptr;
                if (mine) {
                    prob = claim;
                    pending = false;
                    active = true;
                    k = 0;
                    next_check = check_every;
                    res_px = res_dx = res_pu = res_du = 0;
                    if (have_xref) {
                        const char* s = reinterpret_cast<const char*>(prm.Xref + (size_t)prob * SXL);
#pragma unroll
                        for (int b = 0; b <= (SXL * 4 + 127) / 128; ++b) prefetch_l2(s + (b * 128 < SXL * 4 ? b * 128 : SXL * 4 - 4));
                    }
                    if (have_uref) {
                        const char* s = reinterpret_cast<const char*>(prm.Uref + (size_t)prob * SUL);
#pragma unroll
                        for (int b = 0; b <= (SUL * 4 + 127) / 128; ++b) prefetch_l2(s + (b * 128 < SUL * 4 ? b * 128 : SUL * 4 - 4));
                    }
                    load_span<NX, vec_width(NX, NX)>(prm.x0 + (size_t)prob * NX, [&](int i, float v) { x0v.set(i, v); });
                }
                // tensor-memory state of the refilled lanes := 0 (tcgen05.st has no lane mask: read - select - write, warp-wide)
#pragma unroll 1
                for (int i = 1; i < NH; ++i) {
                    uint32_t r[CW];
                    TmemSpan<CW>::ld(col_base(i), r);
                    TmemSpan<CW>::wait(r);
#pragma unroll
                    for (int e = 0; e < CW; ++e) r[e] = mine ? 0u : r[e];
                    TmemSpan<CW>::st(col_base(i), r);
                }
                if constexpr (C::NIL > 0) {
                    constexpr int W = 2 * C::NIL * (NH - 1);
                    uint32_t r[W];
                    TmemSpan<W>::ld(tm_base + C::oSLU, r);
                    TmemSpan<W>::wait(r);
#pragma unroll
                    for (int e = 0; e < W; ++e) r[e] = mine ? 0u : r[e];
                    TmemSpan<W>::st(tm_base + C::oSLU, r);
                }
                tmem_wait_st();
                if (mine) {
                    // box-dual columns := the reference terms of the first sweep (0 without references); column 0 of the state has no
                    // cost term (q_0 is never used); column N-1 carries the terminal term -(xref_N' Pinf)'
                    if (have_xref) {
                        constexpr int GX = steps_per_block(NH, NX, 64);
                        const float* src = prm.Xref + (size_t)prob * SXL;
                        T xr_last[NX];
#pragma unroll
                        for (int r = 0; r < NX; ++r) xr_last[r] = 0;
#pragma unroll 1
                        for (int b = 0; b < NH / GX; ++b) {
                            float buf[GX * NX];
                            load_span<GX * NX, vec_width(SXL, GX * NX)>(src + b * GX * NX, [&](int e, float v) { buf[e] = v; });
#pragma unroll
                            for (int r = 0; r < NX; ++r) xr_last[r] = buf[(GX - 1) * NX + r];
#pragma unroll
                            for (int g = 0; g < GX; ++g)
#pragma unroll
                                for (int e = 0; e < NX; ++e) SB_at(b * GX + g, e) = -(buf[g * NX + e] * cp.Qd[e]);
                        }
                        VX acc;
                        acc.fill(T(0));
#pragma unroll
                        for (int r = 0; r < NX; ++r) {
                            const T nxr = -xr_last[r];
#pragma unroll
                            for (int j = 0; j < NX / 2; ++j) acc.p[j] = fmas(mk2(cP[r * NX + 2 * j], cP[r * NX + 2 * j + 1]), nxr, acc.p[j]);
                            if constexpr (NX & 1) acc.t = fmas(cP[r * NX + NX - 1], nxr, acc.t);
                        }
#pragma unroll
                        for (int e = 0; e < NX; ++e) SB_at(NH - 1, e) = acc.get(e);
                    } else {
#pragma unroll 1
                        for (int i = 0; i < NH; ++i)
#pragma unroll
                            for (int e = 0; e < NX; ++e) SB_at(i, e) = T(0);
                    }
                    if (have_uref) {
                        constexpr int GU = steps_per_block(NH - 1, NU, 64);
                        const float* src = prm.Uref + (size_t)prob * SUL;
#pragma unroll 1
                        for (int b = 0; b < (NH - 1) / GU; ++b) {
                            float buf[GU * NU];
                            load_span<GU * NU, vec_width(SUL, GU * NU)>(src + b * GU * NU, [&](int e, float v) { buf[e] = v; });
#pragma unroll
                            for (int g = 0; g < GU; ++g)
#pragma unroll
                                for (int a = 0; a < NU; ++a) SBU_at(b * GU + g, a) = -(buf[g * NU + a] * cp.Rd[a]);
                        }
                    }
#pragma unroll 1
                    for (int i = 0; i < NH - 1; ++i) {
#pragma unroll
                        for (int a = 0; a < NU; ++a) {
                            U_at(i, a) = 0.0;
                            if (!have_uref) SBU_at(i, a) = T(0);
                        }
#pragma unroll
                        for (int c = 0; c < C::UCD; ++c) SCU_at(i, c) = 0.0;
#pragma unroll
                        for (int j = 0; j < NU / 2; ++j) ND.setp(i, j, mk2(-pack[SP::d0 + i * NU + 2 * j], -pack[SP::d0 + i * NU + 2 * j + 1]));
                        if constexpr (NU & 1) ND.sett(i, -pack[SP::d0 + i * NU + NU - 1]);
                    }
                }
            }
            if (!__any_sync(FULL, active)) {
                publish_done(prm, unpub);
                if (!__any_sync(FULL, pending)) break;
                __nanosleep(256);
                continue;
            }
        }

        // ------------------------------------------------- forward rollout of the increment (float32): leaves du_i, dx_{i+1}
        {
            const bool ff = (k == 0);                                  // first iteration: the full affine map from x0
            const bool anyff = __any_sync(FULL, active && ff);
            VX dx;
#pragma unroll
            for (int j = 0; j < NX / 2; ++j) dx.p[j] = ff ? x0v.p[j] : mk2(T(0), T(0));
            dx.t = ff ? x0v.t : T(0);
#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
                // du_i = -Kinf dx_i - dd_i (admm.cpp:29); the slot of -dd_i now holds du_i for the sweep
                VU du;
                ND.load(i, du);
                if (i > 0 || anyff) mv_acc<NU, NX>(cp.NK, 0, dx, du);
                ND.store(i, du);
                // dx_{i+1} = A dx_i + B du_i (+ f on the first iteration, admm.cpp:30)
                VX dxn;
                if constexpr (C::AFF) {
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) dxn.p[j] = ff ? mk2(cp.f[2 * j], cp.f[2 * j + 1]) : mk2(T(0), T(0));
                    if constexpr (NX & 1) dxn.t = ff ? cp.f[NX - 1] : T(0);
                } else {
                    dxn.fill(T(0));
                }
                if (i > 0 || anyff) mv_acc<NX, NX>(cp.A, 0, dx, dxn);
                mv_acc<NX, NU>(cp.B, 0, du, dxn);
                {
                    uint32_t r[NX];
#pragma unroll
                    for (int e = 0; e < NX; ++e) r[e] = __float_as_uint(dxn.get(e));
                    TmemSpan<NX>::st(col_base(i + 1) + C::oDX, r);
                }
                dx = dxn;
            }
            tmem_wait_st();
        }
        publish_done(prm, unpub);
        k += 1;   // work->iter += 1 (admm.cpp:328)

        // ------------------------------------------------- reverse sweep: x += dx, slack + dual + residuals (double) fused with the
        // Riccati step on the increment of the linear cost (float32)
        const bool first = (k == 1);
        const T nrho = -rho0;
        int rpx_h = 0, rpx_l = 0, rdx_h = 0, rdx_l = 0, rpu_h = 0, rpu_l = 0, rdu_h = 0, rdu_l = 0;

        // state column i >= 1 (cone / half-space families included): returns dq_i (first sweep: + the parked reference term)
        auto state_col = [&](int i, VX& dq) {
            uint32_t w[CW];
            TmemSpan<CW>::ld(col_base(i), w);
            T sb[NX];
#pragma unroll
            for (int e = 0; e < NX; ++e) sb[e] = SB_at(i, e);
            TmemSpan<CW>::wait(w);
            double xo[NX], xn[NX], dxd[NX], dw[NX];
#pragma unroll
            for (int e = 0; e < NX; ++e) {
                xo[e] = dbl_of(w[C::oX + 2 * e], w[C::oX + 2 * e + 1]);
                dxd[e] = static_cast<double>(__uint_as_float(w[C::oDX + e]));
                xn[e] = xo[e] + dxd[e];
                dbl_to(xn[e], w[C::oX + 2 * e], w[C::oX + 2 * e + 1]);
            }
            T ref[NX];
#pragma unroll
            for (int e = 0; e < NX; ++e) {   // box (admm.cpp:85,91-92,184)
                ref[e] = first ? sb[e] : T(0);
                const double s = static_cast<double>(first ? T(0) : sb[e]);
                const double lo = xlo(i, e), hi = xhi(i, e);
                const double to = xo[e] + s;
                double go = box_dual(to, lo, hi);
                if constexpr (!C::FB) go = first ? 0.0 : go;          // cold start: v = 0 (= t_old here) whatever the bounds are
                const double vo = to - go;
                const float gf = static_cast<float>(go);
                const double tn = xn[e] + static_cast<double>(gf);
                const double vn = tn - box_dual(tn, lo, hi);
                const double a = xn[e] - vn, ee = vn - vo;
                amax_w(rpx_h, rpx_l, a);
                amax_w(rdx_h, rdx_l, ee);
                dw[e] = ee - a;
                SB_at(i, e) = gf;
            }
            if constexpr (C::SCD > 0) {   // cone family (admm.cpp:103-117,191): v = x, dual 0 outside the cone's coordinates
                double to[C::SCD], vo[C::SCD], tn[C::SCD];
#pragma unroll
                for (int c = 0; c < C::SCD; ++c) { to[c] = xo[C::SCS + c] + dbl_of(w[C::oSC + 2 * c], w[C::oSC + 2 * c + 1]); vo[c] = to[c]; }
                project_soc_d<C::SCD>(vo, mu_x);
#pragma unroll
                for (int c = 0; c < C::SCD; ++c) {
                    const double g = to[c] - vo[c];
                    dbl_to(g, w[C::oSC + 2 * c], w[C::oSC + 2 * c + 1]);
                    tn[c] = xn[C::SCS + c] + g;
                }
                project_soc_d<C::SCD>(tn, mu_x);   // tn := v_new
#pragma unroll
                for (int e = 0; e < NX; ++e) {
                    if (e >= C::SCS && e < C::SCS + C::SCD) dw[e] += (tn[e - C::SCS] - vo[e - C::SCS]) + (tn[e - C::SCS] - xn[e]);
                    else dw[e] += dxd[e];
                }
            }
            if constexpr (C::NSL > 0) {   // half-space family, one row (admm.cpp:70-73,138-159,203-207)
                double dxo = -bx, dxn_ = -bx;
#pragma unroll
                for (int e = 0; e < NX; ++e) { dxo = fma(ax[e], xo[e], dxo); dxn_ = fma(ax[e], xn[e], dxn_); }
                const double loo = dbl_of(w[C::oSL], w[C::oSL + 1]);
                double lo_ = fmax(0.0, fma(dxo, ix, loo));
                if (bx < 0.0) lo_ = first ? 0.0 : lo_;               // cold start: the slack is 0, not the projection of 0
                const double ln = fmax(0.0, fma(dxn_, ix, lo_));
                const double c = 3.0 * lo_ - 2.0 * ln - loo;
#pragma unroll
                for (int e = 0; e < NX; ++e) dw[e] += fma(c, ax[e], dxd[e]);
                dbl_to(lo_, w[C::oSL], w[C::oSL + 1]);
            } else if (lin_x) {
#pragma unroll
                for (int e = 0; e < NX; ++e) dw[e] += dxd[e];
            }
            TmemSpan<CW - NX>::st(col_base(i) + C::oX, w + C::oX);
#pragma unroll
            for (int e = 0; e < NX; ++e) dq.set(e, fmaf(static_cast<float>(dw[e]), nrho, ref[e]));
        };
        // input column i: returns dr_i; the slot of du_i is free afterwards
        auto input_col = [&](int i, VU& dr) {
            VU du;
            ND.load(i, du);
            double uo[NU], un[NU], dud[NU], dw[NU];
            T ref[NU];
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                uo[a] = U_at(i, a);
                dud[a] = static_cast<double>(du.get(a));
                un[a] = uo[a] + dud[a];
                U_at(i, a) = un[a];
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                const T sbv = SBU_at(i, a);
                ref[a] = first ? sbv : T(0);
                const double s = static_cast<double>(first ? T(0) : sbv);
                const double lo = ulo(i, a), hi = uhi(i, a);
                const double to = uo[a] + s;
                double go = box_dual(to, lo, hi);
                if constexpr (!C::FB) go = first ? 0.0 : go;
                const double vo = to - go;
                const float gf = static_cast<float>(go);
                const double tn = un[a] + static_cast<double>(gf);
                const double vn = tn - box_dual(tn, lo, hi);
                const double pa = un[a] - vn, ee = vn - vo;
                amax_w(rpu_h, rpu_l, pa);
                amax_w(rdu_h, rdu_l, ee);
                dw[a] = ee - pa;
                SBU_at(i, a) = gf;
            }
            if constexpr (C::UCD > 0) {
                double to[C::UCD], vo[C::UCD], tn[C::UCD];
#pragma unroll
                for (int c = 0; c < C::UCD; ++c) { to[c] = uo[C::UCS + c] + SCU_at(i, c); vo[c] = to[c]; }
                project_soc_d<C::UCD>(vo, mu_u);
#pragma unroll
                for (int c = 0; c < C::UCD; ++c) {
                    const double g = to[c] - vo[c];
                    SCU_at(i, c) = g;
                    tn[c] = un[C::UCS + c] + g;
                }
                project_soc_d<C::UCD>(tn, mu_u);
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    if (a >= C::UCS && a < C::UCS + C::UCD) dw[a] += (tn[a - C::UCS] - vo[a - C::UCS]) + (tn[a - C::UCS] - un[a]);
                    else dw[a] += dud[a];
                }
            }
            if constexpr (C::NIL > 0) {
                uint32_t lw[2];
                TmemSpan<2>::ld(tm_base + C::oSLU + 2 * i, lw);
                double duo = -bu, dun = -bu;
#pragma unroll
                for (int a = 0; a < NU; ++a) { duo = fma(au[a], uo[a], duo); dun = fma(au[a], un[a], dun); }
                TmemSpan<2>::wait(lw);
                const double loo = dbl_of(lw[0], lw[1]);
                double lo_ = fmax(0.0, fma(duo, iu, loo));
                if (bu < 0.0) lo_ = first ? 0.0 : lo_;
                const double ln = fmax(0.0, fma(dun, iu, lo_));
                const double c = 3.0 * lo_ - 2.0 * ln - loo;
#pragma unroll
                for (int a = 0; a < NU; ++a) dw[a] += fma(c, au[a], dud[a]);
                dbl_to(lo_, lw[0], lw[1]);
                TmemSpan<2>::st(tm_base + C::oSLU + 2 * i, lw);
            } else if (lin_u) {
#pragma unroll
                for (int a = 0; a < NU; ++a) dw[a] += dud[a];
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) dr.set(a, fmaf(static_cast<float>(dw[a]), nrho, ref[a]));
        };

        // ONE rolled loop over the columns N-1 .. 0 with warp-uniform branches for the two special columns, so that state_col and
        // input_col are instantiated once each: with a peeled terminal column and a peeled column 0 the hot loop was 45 KB, past the
        // 32 KB instruction cache (stall_no_instruction 1.05 per issue, profiles/r02/ncu_full_rocket_tpp4.json).
        //   i = N-1 : state column only; dp_N = -(xref_N' Pinf)' [first sweep] - rho dw_N            (admm.cpp:238-246)
        //   i = 0   : input column and dd_0; of x_0 = x0 only the box slack (p_0, q_0 and its cone / half-space slacks are never used)
        VX dp;
        dp.fill(T(0));
#pragma unroll 1
        for (int i = NH - 1; i >= 0; --i) {
            const bool inner = i < NH - 1;
            VX akp;
            VU btp, dr;
            akp.fill(T(0));
            btp.fill(T(0));
            dr.fill(T(0));
            if (inner) {
                // the two products with dp_{i+1} depend on nothing of this column: they run under its loads and its FP64 work
                if (i >= 1) mv_acc<NX, NX>(cp.AK, 0, dp, akp);
                mv_acc<NU, NX>(cp.BT, 0, dp, btp);
                input_col(i, dr);
                // dd_i = Quu_inv (B' dp + dr)   (admm.cpp:17; BPf cancels in the increment)
                VU t;
#pragma unroll
                for (int j = 0; j < NU / 2; ++j) t.p[j] = addv(btp.p[j], dr.p[j]);
                if constexpr (NU & 1) t.t = btp.t + dr.t;
                VU d;
                d.fill(T(0));
                mv_acc<NU, NU>(cp.Quu, 0, t, d);
                VU nd;
#pragma unroll
                for (int j = 0; j < NU / 2; ++j) nd.p[j] = negv(d.p[j]);
                if constexpr (NU & 1) nd.t = -d.t;
                ND.store(i, nd);
            }
            if (i >= 1) {
                // dp_i = dq_i + AmBKt dp - Kinf' dr   (admm.cpp:18)
                VX dq;
                state_col(i, dq);
                if (inner) {
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) dq.p[j] = addv(dq.p[j], akp.p[j]);
                    if constexpr (NX & 1) dq.t += akp.t;
                    mv_acc<NX, NU>(cp.NKT, 0, dr, dq);
                }
                dp = dq;
            } else {
#pragma unroll
                for (int e = 0; e < NX; ++e) {
                    const double x0d = static_cast<double>(x0v.get(e));
                    const T sbv = SB_at(0, e);
                    const double s = static_cast<double>(first ? T(0) : sbv);
                    const double lo = xlo(0, e), hi = xhi(0, e);
                    const double to = x0d + s;
                    const double go = first ? 0.0 : box_dual(to, lo, hi);
                    const double vo = first ? 0.0 : (to - go);           // v(0) = 0, g(0) = 0 on the cold workspace, while x_0 = x0 from the start
                    const float gf = static_cast<float>(go);
                    const double tn = x0d + static_cast<double>(gf);
                    const double vn = tn - box_dual(tn, lo, hi);
                    amax_w(rpx_h, rpx_l, x0d - vn);
                    amax_w(rdx_h, rdx_l, vn - vo);
                    SB_at(0, e) = gf;
                }
            }
        }
        tmem_wait_st();

        // ------------------------------------------------- termination (admm.cpp:253-271, 364-388)
        bool finish = false;
        int st = 11;
        if (k == next_check) {
            next_check += check_every;
            res_px = __hiloint2double(rpx_h, rpx_l); res_dx = __hiloint2double(rdx_h, rdx_l) * rho_d;
            res_pu = __hiloint2double(rpu_h, rpu_l); res_du = __hiloint2double(rdu_h, rdu_l) * rho_d;
            if (res_px < tol_pri && res_pu < tol_pri && res_dx < tol_dua && res_du < tol_dua) { finish = true; st = 1; }
            if (prm.amb_band > 0.f) {
                const double up = 1.0 + prm.amb_band, dn = 1.0 - prm.amb_band;
                const bool below_up = res_px < tol_pri * up && res_pu < tol_pri * up && res_dx < tol_dua * up && res_du < tol_dua * up;
                const bool below_dn = res_px < tol_pri * dn && res_pu < tol_pri * dn && res_dx < tol_dua * dn && res_du < tol_dua * dn;
                if (below_up && !below_dn) { finish = true; st = kAmbiguousBit | 11; }
            }
        }
        if (k >= max_iter) finish = true;
        bool fin = active && finish;
        if (producer) {
            const bool amb = fin && (st & kAmbiguousBit);
            queue_push(prm, amb, prob, lane);
            if (amb) { fin = false; active = false; last_k = k; }
        }
        const size_t pbx = (size_t)prob * SXL, pbu = (size_t)prob * SUL;
        if (__any_sync(FULL, fin)) {
            // solution = (vnew, znew) = projection of x + (older dual) as stored (the x read is warp-collective)
#pragma unroll 1
            for (int i = 0; i < NH; ++i) {
                uint32_t w[2 * NX];
                if (i > 0) { TmemSpan<2 * NX>::ld(col_base(i) + C::oX, w); TmemSpan<2 * NX>::wait(w); }
                if (fin) {
                    float v[NX];
#pragma unroll
                    for (int e = 0; e < NX; ++e) {
                        const double xv = (i > 0) ? dbl_of(w[2 * e], w[2 * e + 1]) : static_cast<double>(x0v.get(e));
                        v[e] = static_cast<float>(clampd(xv + static_cast<double>(SB_at(i, e)), xlo(i, e), xhi(i, e)));
                    }
                    store_span<NX, vec_width(SXL, NX)>(prm.x + pbx + i * NX, [&](int r) { return v[r]; });
                }
            }
        }
        if (fin) {
#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
                float z[NU];
#pragma unroll
                for (int a = 0; a < NU; ++a) z[a] = static_cast<float>(clampd(U_at(i, a) + static_cast<double>(SBU_at(i, a)), ulo(i, a), uhi(i, a)));
                store_span<NU, vec_width(SUL, NU)>(prm.u + pbu + i * NU, [&](int a) { return z[a]; });
            }
            prm.iter[prob] = k;
            prm.status[prob] = st;
            if (prm.residuals)
                *reinterpret_cast<float4*>(prm.residuals + 4 * (size_t)prob) =
                    make_float4(static_cast<float>(res_px), static_cast<float>(res_dx), static_cast<float>(res_pu), static_cast<float>(res_du));
            if (prm.rho_out) prm.rho_out[prob] = rho0;
            if (prm.done_counters) unpub = prob;
            active = false;
            last_k = k;
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (producer && threadIdx.x == 0) queue_producer_exit(prm);
    if (threadIdx.x < 32) tmem_dealloc(tmem_base_s, 512);
}

template <class C>
inline size_t tpp4_smem_bytes(int pack_elems) {
    return ((size_t)((pack_elems + 31) & ~31) + (size_t)C::COLS * C::BLOCK) * sizeof(float);
}

}  // namespace tmpc
