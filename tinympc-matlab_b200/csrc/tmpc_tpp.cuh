// tmpc_tpp.cuh -- the batched ADMM throughput kernel for sm_100a ("thread per problem").
//
// Path implemented (reference: tinympc/TinyMPC/src/tinympc/admm.cpp):
//   solve                 :274-389   loop driver, termination, status codes
//   backward_pass_grad    :13-20     d_i = Quu_inv (B' p_{i+1} + r_i + BPf);  p_i = q_i + AmBKt p_{i+1} - Kinf' r_i + APf
//   forward_pass          :25-32     u_i = -Kinf x_i - d_i;  x_{i+1} = A x_i + B u_i + f
//   update_slack          :81-175    box clamp, second-order-cone projection (:39-60), half-space projection (:70-73)
//   update_dual           :181-208
//   update_linear_cost    :214-247   (never materialised: q, r, p_N are recomputed inside the backward sweep)
//   termination_condition :253-271
//   adaptive rho          rho_benchmark.cpp:44-250 in closed block form
//
// Mapping.  One CUDA thread owns one MPC problem for all of its ADMM iterations.  The mat-vec
// operands (A, B, Kinf, AmBKt, Quu_inv, ...) are identical for the whole batch: every CTA stages
// the family "pack" into shared memory once with a TMA bulk copy (cp.async.bulk + mbarrier) and all
// lanes read the same coefficient at the same time (a shared-memory broadcast), so a mat-vec is a
// pure FFMA stream with no shuffles.  Per-problem trajectories (duals g,y, slacks v,z, reference
// terms, d) live in registers or in conflict-free shared-memory columns ([element][thread]) for the
// whole solve; HBM is touched only to read x0/Xref/Uref once and to write the solution once.
// Problems need 1..max_iter iterations, so lanes are refilled: a lane that finishes claims the
// next unsolved problem from a global counter (warp-aggregated atomic) while its neighbours keep
// iterating -- no lane waits for the slowest problem of a warp.
//
// Loop rotation.  The reference runs backward -> forward -> slack -> dual -> linear cost -> check.
// On a cold workspace (q = r = p = 0) the first backward pass is problem independent, so its
// result d0 is precomputed on the host; each iteration here is forward+slack+dual -> check ->
// backward (for the next iteration), which is the same sequence of values and skips the backward
// pass of the final iteration.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tmpc_common.h"

namespace tmpc {

enum : int { FEAT_BOX = 0, FEAT_CONSTR = 1, FEAT_ADAPT = 2 };

// ----------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + TMA bulk copy (global -> shared), sm_90+/sm_100a
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ----------------------------------------------------------------------------------------------
// per-thread arrays: registers (static indexing only) or a shared-memory column [elem][thread]
// ----------------------------------------------------------------------------------------------
template <typename T, int LEN, bool SM, int OFF, int BLOCK>
struct Col;
template <typename T, int LEN, int OFF, int BLOCK>
struct Col<T, LEN, true, OFF, BLOCK> {
    T* base;  // thread's column origin (already + threadIdx.x)
    __device__ __forceinline__ explicit Col(T* colbase) : base(colbase + OFF * BLOCK) {}
    __device__ __forceinline__ T get(int i) const { return base[i * BLOCK]; }
    __device__ __forceinline__ void set(int i, T v) { base[i * BLOCK] = v; }
};
template <typename T, int LEN, int OFF, int BLOCK>
struct Col<T, LEN, false, OFF, BLOCK> {
    T v[LEN > 0 ? LEN : 1];
    __device__ __forceinline__ explicit Col(T*) {}
    __device__ __forceinline__ T get(int i) const { return v[i]; }
    __device__ __forceinline__ void set(int i, T val) { v[i] = val; }
};

// placement bits: 1 = shared memory column, 0 = registers
enum : unsigned {
    P_G = 1u << 0, P_V = 1u << 1, P_XRQ = 1u << 2, P_Y = 1u << 3, P_Z = 1u << 4, P_URR = 1u << 5, P_D = 1u << 6,
    P_X0 = 1u << 7, P_PT = 1u << 8, P_GC = 1u << 9, P_GL = 1u << 10, P_SX = 1u << 11, P_YC = 1u << 12, P_YL = 1u << 13,
    P_SU = 1u << 14, P_ALL = 0x7fffu
};

template <typename T_, int NX_, int NU_, int NH_, int FEAT_, int BLOCK_, bool UNROLL_, unsigned PLACE_, bool PPB_>
struct TppCfg {
    using T = T_;
    static constexpr int NX = NX_, NU = NU_, NH = NH_, FEAT = FEAT_, BLOCK = BLOCK_;
    static constexpr bool UNROLL = UNROLL_;
    static constexpr bool PPB = PPB_;                       // per-problem bounds read from global memory
    static constexpr unsigned PLACE = UNROLL_ ? PLACE_ : P_ALL;  // a rolled time loop needs dynamic indexing
    static constexpr bool CONSTR = FEAT_ == FEAT_CONSTR;
    static constexpr bool ADAPT = FEAT_ == FEAT_ADAPT;
    static constexpr int SX = NX * NH, SU = NU * (NH - 1);
    static constexpr int TU = UNROLL_ ? NH : 1;             // time-loop unroll factor
    // element offsets of the shared-memory columns
    static constexpr int sz(unsigned bit, int len, bool enabled = true) { return (enabled && (PLACE & bit)) ? len : 0; }
    static constexpr int oG = 0;
    static constexpr int oV = oG + sz(P_G, SX);
    static constexpr int oXRQ = oV + sz(P_V, SX);
    static constexpr int oY = oXRQ + sz(P_XRQ, SX);
    static constexpr int oZ = oY + sz(P_Y, SU);
    static constexpr int oURR = oZ + sz(P_Z, SU);
    static constexpr int oD = oURR + sz(P_URR, SU);
    static constexpr int oX0 = oD + sz(P_D, SU);
    static constexpr int oPT = oX0 + sz(P_X0, NX);
    static constexpr int oGC = oPT + sz(P_PT, ADAPT ? 2 * NX : NX);
    static constexpr int oGL = oGC + sz(P_GC, SX, CONSTR);
    static constexpr int oSX = oGL + sz(P_GL, SX, CONSTR);
    static constexpr int oYC = oSX + sz(P_SX, SX, CONSTR);
    static constexpr int oYL = oYC + sz(P_YC, SU, CONSTR);
    static constexpr int oSU = oYL + sz(P_YL, SU, CONSTR);
    static constexpr int oSCR = oSU + sz(P_SU, SU, CONSTR);
    static constexpr int COLS = oSCR + (CONSTR ? (NX > NU ? NX : NU) : 0);   // + cone scratch column
};

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return sqrtf(a); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double min(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
};

// vectorised, read-only loads of one problem's contiguous float chunk (LEN floats at src)
template <int LEN, typename F>
__device__ __forceinline__ void load_chunk(const float* __restrict__ src, F&& sink) {
    if constexpr (LEN % 4 == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int k = 0; k < LEN / 4; ++k) {
            float4 t = __ldg(s4 + k);
            sink(4 * k + 0, t.x); sink(4 * k + 1, t.y); sink(4 * k + 2, t.z); sink(4 * k + 3, t.w);
        }
    } else if constexpr (LEN % 2 == 0) {
        const float2* s2 = reinterpret_cast<const float2*>(src);
#pragma unroll
        for (int k = 0; k < LEN / 2; ++k) {
            float2 t = __ldg(s2 + k);
            sink(2 * k + 0, t.x); sink(2 * k + 1, t.y);
        }
    } else {
#pragma unroll
        for (int k = 0; k < LEN; ++k) sink(k, __ldg(src + k));
    }
}
template <int LEN, typename F>
__device__ __forceinline__ void store_chunk(float* __restrict__ dst, F&& src) {
    if constexpr (LEN % 4 == 0) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < LEN / 4; ++k) d4[k] = make_float4(src(4 * k), src(4 * k + 1), src(4 * k + 2), src(4 * k + 3));
    } else if constexpr (LEN % 2 == 0) {
        float2* d2 = reinterpret_cast<float2*>(dst);
#pragma unroll
        for (int k = 0; k < LEN / 2; ++k) d2[k] = make_float2(src(2 * k), src(2 * k + 1));
    } else {
#pragma unroll
        for (int k = 0; k < LEN; ++k) dst[k] = src(k);
    }
}

// second-order-cone projection of scr[start .. start+dim) in place (admm.cpp:39-60).
// mu and the norm are float in the reference (:39,:42); a/mu is a float division (:54).
template <typename T, int BLOCK>
__device__ __forceinline__ void project_soc_col(T* scr, int start, int dim, float mu) {
    using N = Num<T>;
    T* s = scr + start * BLOCK;
    const T last = s[(dim - 1) * BLOCK];
    const T u0 = last * static_cast<T>(mu);
    T ss = 0;
    for (int j = 0; j < dim - 1; ++j) { T e = s[j * BLOCK]; ss = N::fma(e, e, ss); }
    const float a = static_cast<float>(N::sqrt(ss));
    const T aT = static_cast<T>(a);
    if (aT <= -u0) {
        for (int j = 0; j < dim; ++j) s[j * BLOCK] = T(0);
    } else if (aT <= u0) {
        // inside the cone
    } else {
        const T fct = T(0.5) * (T(1) + u0 / aT);
        for (int j = 0; j < dim - 1; ++j) s[j * BLOCK] = fct * s[j * BLOCK];
        s[(dim - 1) * BLOCK] = fct * static_cast<T>(a / mu);
    }
}

template <class C>
__global__ void __launch_bounds__(C::BLOCK, 1) tpp_kernel(const SolveParams prm) {
    using T = typename C::T;
    using N = Num<T>;
    using SP = StaticPack<C::NX, C::NU, C::NH>;
    constexpr int NX = C::NX, NU = C::NU, NH = C::NH, BLOCK = C::BLOCK, SXL = C::SX, SUL = C::SU;
    constexpr unsigned PL = C::PLACE;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t pack_bar;
    T* pack = reinterpret_cast<T*>(smem_raw);
    const uint32_t pack_bytes = static_cast<uint32_t>(prm.pack_elems) * sizeof(T);

    // ---- stage the family pack into shared memory: one TMA bulk copy per CTA ----
    if (threadIdx.x == 0) {
        mbar_init(&pack_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&pack_bar, pack_bytes);
        tma_bulk_g2s(pack, prm.pack, pack_bytes, &pack_bar);
    }
    __syncthreads();
    mbar_wait(&pack_bar, 0);

    T* colbase = pack + ((prm.pack_elems + 31) & ~31) + threadIdx.x;

    Col<T, SXL, (PL & P_G) != 0, C::oG, BLOCK> G(colbase);
    Col<T, SXL, (PL & P_V) != 0, C::oV, BLOCK> V(colbase);
    Col<T, SXL, (PL & P_XRQ) != 0, C::oXRQ, BLOCK> XRQ(colbase);
    Col<T, SUL, (PL & P_Y) != 0, C::oY, BLOCK> Y(colbase);
    Col<T, SUL, (PL & P_Z) != 0, C::oZ, BLOCK> Z(colbase);
    Col<T, SUL, (PL & P_URR) != 0, C::oURR, BLOCK> URR(colbase);
    Col<T, SUL, (PL & P_D) != 0, C::oD, BLOCK> D(colbase);
    Col<T, NX, (PL & P_X0) != 0, C::oX0, BLOCK> X0(colbase);
    Col<T, C::ADAPT ? 2 * NX : NX, (PL & P_PT) != 0, C::oPT, BLOCK> PT(colbase);
    Col<T, C::CONSTR ? SXL : 0, (PL & P_GC) != 0, C::oGC, BLOCK> GC(colbase);
    Col<T, C::CONSTR ? SXL : 0, (PL & P_GL) != 0, C::oGL, BLOCK> GL(colbase);
    Col<T, C::CONSTR ? SXL : 0, (PL & P_SX) != 0, C::oSX, BLOCK> SXT(colbase);
    Col<T, C::CONSTR ? SUL : 0, (PL & P_YC) != 0, C::oYC, BLOCK> YC(colbase);
    Col<T, C::CONSTR ? SUL : 0, (PL & P_YL) != 0, C::oYL, BLOCK> YL(colbase);
    Col<T, C::CONSTR ? SUL : 0, (PL & P_SU) != 0, C::oSU, BLOCK> SUT(colbase);
    T* scr = colbase + C::oSCR * BLOCK;   // cone scratch column (CONSTR only)

    const T* cA = pack + SP::A;
    const T* cB = pack + SP::B;
    const T* cK = pack + SP::Kinf;
    const T* cAK = pack + SP::AmBKt;
    const T* cQuu = pack + SP::Quu_inv;
    const T* cP = pack + SP::Pinf;
    const T* cf = pack + SP::f;
    const T* cAPf = pack + SP::APf;
    const T* cBPf = pack + SP::BPf;
    const T* cdK = pack + SP::dKinf;
    const T* cdP = pack + SP::dPinf;

    const T rho0 = static_cast<T>(prm.rho);
    const T tol_pri = static_cast<T>(prm.abs_pri_tol), tol_dua = static_cast<T>(prm.abs_dua_tol);
    const int max_iter = prm.max_iter, check_every = prm.check_termination;
    const bool soc_x = C::CONSTR && prm.en_state_soc && prm.n_state_cones > 0;
    const bool soc_u = C::CONSTR && prm.en_input_soc && prm.n_input_cones > 0;
    const bool lin_x = C::CONSTR && prm.en_state_linear;
    const bool lin_u = C::CONSTR && prm.en_input_linear;
    const int nsl = prm.nsl, nil = prm.nil;
    const T* cAlx = pack + SP::lin;
    const T* cblx = cAlx + nsl * NX;
    const T* cnrx = cblx + nsl;
    const T* cAlu = cnrx + nsl;
    const T* cblu = cAlu + nil * NU;
    const T* cnru = cblu + nil;

    const int lane = threadIdx.x & 31;
    int prob = -1;          // problem owned by this lane
    bool active = false;    // lane holds an unfinished problem
    bool exhausted = false; // the work counter ran past the batch
    int k = 0;              // ADMM iterations done on the current problem
    T res_px = 0, res_dx = 0, res_pu = 0, res_du = 0;   // last evaluated residuals (admm.cpp:257-260)
    // adaptive rho state (cache->rho, and the Taylor offset of Kinf/Pinf): rho_lc/dl_lc are the
    // values update_linear_cost saw (it runs BEFORE the adaptation inside an iteration)
    T rho = rho0, rho_lc = rho0, dlt = 0, dlt_lc = 0;

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        {
            const bool want = !active && !exhausted;
            const unsigned m = __ballot_sync(FULL, want);
            if (m) {
                const int leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(prm.work_counter, __popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (want) {
                    prob = base + __popc(m & ((1u << lane) - 1u));
                    if (prob >= prm.batch) {
                        exhausted = true;
                        prob = 0;   // keeps the (unused) per-problem bound reads of an idle lane in range
                    } else {
                        active = true;
                        k = 0;
                        res_px = res_dx = res_pu = res_du = 0;
                        rho = rho_lc = rho0; dlt = dlt_lc = 0;
                        // x0
                        load_chunk<NX>(prm.x0 + (size_t)prob * NX, [&](int i, float v) { X0.set(i, static_cast<T>(v)); });
                        // Xref -> XRQ = Xref .* Q (work->Q = diag(Q)+rho, admm.cpp:218) and the terminal
                        // term PT = -(xref_N' Pinf)' (admm.cpp:238)
                        T xr_last[NX];
#pragma unroll
                        for (int r = 0; r < NX; ++r) xr_last[r] = 0;
                        if (prm.Xref) {
                            load_chunk<SXL>(prm.Xref + (size_t)prob * SXL, [&](int e, float v) {
                                XRQ.set(e, static_cast<T>(v) * pack[SP::Qd + e % NX]);
                                if (e >= SXL - NX) xr_last[e - (SXL - NX)] = static_cast<T>(v);
                            });
                        } else {
#pragma unroll
                            for (int e = 0; e < SXL; ++e) XRQ.set(e, T(0));
                        }
#pragma unroll
                        for (int c = 0; c < NX; ++c) {
                            T acc = 0, acc1 = 0;
#pragma unroll
                            for (int r = 0; r < NX; ++r) {
                                acc = N::fma(xr_last[r], cP[r * NX + c], acc);
                                if constexpr (C::ADAPT) acc1 = N::fma(xr_last[r], cdP[r * NX + c], acc1);
                            }
                            PT.set(c, -acc);
                            if constexpr (C::ADAPT) PT.set(NX + c, -acc1);
                        }
                        if (prm.Uref) {
                            load_chunk<SUL>(prm.Uref + (size_t)prob * SUL,
                                            [&](int e, float v) { URR.set(e, static_cast<T>(v) * pack[SP::Rd + e % NU]); });
                        } else {
#pragma unroll
                            for (int e = 0; e < SUL; ++e) URR.set(e, T(0));
                        }
                        // cold workspace (tiny_api.cpp:68-105): duals and slacks zero, d = d0
#pragma unroll
                        for (int e = 0; e < SXL; ++e) { G.set(e, T(0)); V.set(e, T(0)); }
#pragma unroll
                        for (int e = 0; e < SUL; ++e) { Y.set(e, T(0)); Z.set(e, T(0)); D.set(e, pack[SP::d0 + e]); }
                        if constexpr (C::CONSTR) {
#pragma unroll
                            for (int e = 0; e < SXL; ++e) { GC.set(e, T(0)); GL.set(e, T(0)); SXT.set(e, T(0)); }
#pragma unroll
                            for (int e = 0; e < SUL; ++e) { YC.set(e, T(0)); YL.set(e, T(0)); SUT.set(e, T(0)); }
                        }
                    }
                }
            }
            if (!__any_sync(FULL, active)) break;
        }

        // ------------------------------------------------- forward rollout + slack + dual + residuals
        T rpx = 0, rdx = 0, rpu = 0, rdu = 0;
        // adaptive-rho accumulators (rho_benchmark.cpp:146-173); only evaluated on sweeps where some
        // lane of the warp is at an adaptation iteration (i > 0 && i % 5 == 0, admm.cpp:339)
        const bool do_adapt = C::ADAPT && prm.adaptive_rho && __any_sync(FULL, active && k > 0 && k % 5 == 0);
        T a_pri = 0, a_prin = 0, a_dua = 0, a_duan = 0;
        T x[NX];
        T xprev[NX], gprev[NX], uprev[NU], yprev[NU];   // ADAPT: lagged column for the A'g terms
#pragma unroll
        for (int r = 0; r < NX; ++r) { x[r] = X0.get(r); xprev[r] = 0; gprev[r] = 0; }
#pragma unroll
        for (int a = 0; a < NU; ++a) { uprev[a] = 0; yprev[a] = 0; }
        const T* pxmin = pack + SP::xmin;
        const T* pxmax = pack + SP::xmax;
        const T* pumin = pack + SP::umin;
        const T* pumax = pack + SP::umax;
        const size_t pbx = (size_t)(prob < 0 ? 0 : prob) * SXL, pbu = (size_t)(prob < 0 ? 0 : prob) * SUL;
        (void)pbx; (void)pbu;

#pragma unroll(C::TU)
        for (int i = 0; i < NH; ++i) {
            // ---- state column i: vnew = clamp(x + g), g += x - vnew (admm.cpp:85,92,184)
            T gnew[NX];
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                const int e = i * NX + r;
                const T g = G.get(e), vo = V.get(e);
                T lo, hi;
                if constexpr (C::PPB) {
                    lo = prm.en_state_bound ? static_cast<T>(__ldg(prm.x_min + pbx + e)) : -CUDART_INF_F;
                    hi = prm.en_state_bound ? static_cast<T>(__ldg(prm.x_max + pbx + e)) : CUDART_INF_F;
                }
                else { lo = pxmin[e]; hi = pxmax[e]; }
                T vn = x[r] + g;
                vn = N::min(hi, N::max(lo, vn));
                const T gn = (g + x[r]) - vn;
                rpx = N::max(rpx, N::abs(x[r] - vn));
                rdx = N::max(rdx, N::abs(vo - vn));
                V.set(e, vn);
                G.set(e, gn);
                gnew[r] = gn;
                if constexpr (C::ADAPT) {
                    if (do_adapt && i > 0) {   // dynamics rows of A_matrix: (A x + B u - x_next) - vnew_next = -f - vnew
                        a_pri = N::max(a_pri, N::abs(cf[r] + vn));
                        a_prin = N::max(a_prin, N::max(N::abs(vn), N::abs(cf[r])));
                    }
                }
            }
            if constexpr (C::CONSTR) {
                T extra[NX];
#pragma unroll
                for (int r = 0; r < NX; ++r) extra[r] = 0;
                if (soc_x) {   // admm.cpp:103,112-122,191
#pragma unroll
                    for (int r = 0; r < NX; ++r) scr[r * BLOCK] = x[r] + GC.get(i * NX + r);
                    for (int c = 0; c < prm.n_state_cones; ++c) project_soc_col<T, BLOCK>(scr, prm.Acx[c], prm.qcx[c], prm.cx[c]);
#pragma unroll
                    for (int r = 0; r < NX; ++r) {
                        const T vc = scr[r * BLOCK];
                        const T gcn = (GC.get(i * NX + r) + x[r]) - vc;
                        GC.set(i * NX + r, gcn);
                        extra[r] += vc - gcn;
                    }
                }
                if (lin_x) {   // admm.cpp:139,148-159,201
                    T vl[NX];
#pragma unroll
                    for (int r = 0; r < NX; ++r) vl[r] = x[r] + GL.get(i * NX + r);
                    for (int c = 0; c < nsl; ++c) {
                        T val = 0;
#pragma unroll
                        for (int r = 0; r < NX; ++r) val = N::fma(cAlx[c * NX + r], vl[r], val);
                        if (val > cblx[c]) {
                            const T dist = (val - cblx[c]) / cnrx[c];
#pragma unroll
                            for (int r = 0; r < NX; ++r) vl[r] = vl[r] - dist * cAlx[c * NX + r];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < NX; ++r) {
                        const T gln = (GL.get(i * NX + r) + x[r]) - vl[r];
                        GL.set(i * NX + r, gln);
                        extra[r] += vl[r] - gln;
                    }
                }
#pragma unroll
                for (int r = 0; r < NX; ++r) SXT.set(i * NX + r, extra[r]);
            }
            if constexpr (C::ADAPT) {
                // dual residual blocks of column i-1 need g_i (post update): x-block A'g_i - g_{i-1}, u-block y_{i-1} + B'g_i
                if (do_adapt && i > 0) {
#pragma unroll
                    for (int c = 0; c < NX; ++c) {
                        T aty = 0;
#pragma unroll
                        for (int r = 0; r < NX; ++r) aty = N::fma(cA[r * NX + c], gnew[r], aty);
                        if (i > 1) aty -= gprev[c];
                        const T qx = pack[SP::Qd + c] * xprev[c];   // Px = qv = Q .* x for columns < N-1
                        a_dua = N::max(a_dua, N::abs(qx + qx + aty));
                        a_duan = N::max(a_duan, N::max(N::abs(qx), N::abs(aty)));
                    }
#pragma unroll
                    for (int a = 0; a < NU; ++a) {
                        T aty = yprev[a];
#pragma unroll
                        for (int r = 0; r < NX; ++r) aty = N::fma(cB[r * NU + a], gnew[r], aty);
                        const T ru = pack[SP::Rd + a] * uprev[a];
                        a_dua = N::max(a_dua, N::abs(ru + ru + aty));
                        a_duan = N::max(a_duan, N::max(N::abs(ru), N::abs(aty)));
                    }
                }
                if (do_adapt && i == NH - 1) {   // last state block: Px = Pinf x_N, qv = Q .* x_N, ATy = -g_N
#pragma unroll
                    for (int r = 0; r < NX; ++r) {
                        T px = 0;
#pragma unroll
                        for (int c = 0; c < NX; ++c) px = N::fma(N::fma(dlt, cdP[r * NX + c], cP[r * NX + c]), x[c], px);
                        const T qx = pack[SP::Qd + r] * x[r];
                        const T aty = -gnew[r];
                        a_dua = N::max(a_dua, N::abs(px + qx + aty));
                        a_duan = N::max(a_duan, N::max(N::max(N::abs(px), N::abs(qx)), N::abs(aty)));
                    }
                }
            }
            if (i < NH - 1) {
                // ---- u_i = -Kinf x_i - d_i (admm.cpp:29); Kinf = Kinf0 + dlt * dKinf under adaptive rho
                T u[NU];
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    T acc = 0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc = N::fma(cK[a * NX + c], x[c], acc);
                    if constexpr (C::ADAPT) {
                        T acc1 = 0;
#pragma unroll
                        for (int c = 0; c < NX; ++c) acc1 = N::fma(cdK[a * NX + c], x[c], acc1);
                        acc = N::fma(dlt, acc1, acc);
                    }
                    u[a] = -acc - D.get(i * NU + a);
                }
                // ---- input column i: znew = clamp(u + y), y += u - znew (admm.cpp:88,97,187)
                T ynew[NU];
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    const int e = i * NU + a;
                    const T yv = Y.get(e), zo = Z.get(e);
                    T lo, hi;
                    if constexpr (C::PPB) {
                        lo = prm.en_input_bound ? static_cast<T>(__ldg(prm.u_min + pbu + e)) : -CUDART_INF_F;
                        hi = prm.en_input_bound ? static_cast<T>(__ldg(prm.u_max + pbu + e)) : CUDART_INF_F;
                    }
                    else { lo = pumin[e]; hi = pumax[e]; }
                    T zn = u[a] + yv;
                    zn = N::min(hi, N::max(lo, zn));
                    const T yn = (yv + u[a]) - zn;
                    rpu = N::max(rpu, N::abs(u[a] - zn));
                    rdu = N::max(rdu, N::abs(zo - zn));
                    Z.set(e, zn);
                    Y.set(e, yn);
                    ynew[a] = yn;
                    if constexpr (C::ADAPT) { if (do_adapt) a_prin = N::max(a_prin, N::max(N::abs(u[a]), N::abs(zn))); }
                }
                if constexpr (C::CONSTR) {
                    T extra[NU];
#pragma unroll
                    for (int a = 0; a < NU; ++a) extra[a] = 0;
                    if (soc_u) {
#pragma unroll
                        for (int a = 0; a < NU; ++a) scr[a * BLOCK] = u[a] + YC.get(i * NU + a);
                        for (int c = 0; c < prm.n_input_cones; ++c) project_soc_col<T, BLOCK>(scr, prm.Acu[c], prm.qcu[c], prm.cu[c]);
#pragma unroll
                        for (int a = 0; a < NU; ++a) {
                            const T zc = scr[a * BLOCK];
                            const T ycn = (YC.get(i * NU + a) + u[a]) - zc;
                            YC.set(i * NU + a, ycn);
                            extra[a] += zc - ycn;
                        }
                    }
                    if (lin_u) {
                        T zl[NU];
#pragma unroll
                        for (int a = 0; a < NU; ++a) zl[a] = u[a] + YL.get(i * NU + a);
                        for (int c = 0; c < nil; ++c) {
                            T val = 0;
#pragma unroll
                            for (int a = 0; a < NU; ++a) val = N::fma(cAlu[c * NU + a], zl[a], val);
                            if (val > cblu[c]) {
                                const T dist = (val - cblu[c]) / cnru[c];
#pragma unroll
                                for (int a = 0; a < NU; ++a) zl[a] = zl[a] - dist * cAlu[c * NU + a];
                            }
                        }
#pragma unroll
                        for (int a = 0; a < NU; ++a) {
                            const T yln = (YL.get(i * NU + a) + u[a]) - zl[a];
                            YL.set(i * NU + a, yln);
                            extra[a] += zl[a] - yln;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < NU; ++a) SUT.set(i * NU + a, extra[a]);
                }
                // ---- x_{i+1} = A x_i + B u_i + f (admm.cpp:30)
                T xn[NX];
#pragma unroll
                for (int r = 0; r < NX; ++r) {
                    T acc = 0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc = N::fma(cA[r * NX + c], x[c], acc);
#pragma unroll
                    for (int a = 0; a < NU; ++a) acc = N::fma(cB[r * NU + a], u[a], acc);
                    xn[r] = acc + cf[r];
                }
                if constexpr (C::ADAPT) {
#pragma unroll
                    for (int r = 0; r < NX; ++r) { xprev[r] = x[r]; gprev[r] = gnew[r]; }
#pragma unroll
                    for (int a = 0; a < NU; ++a) { uprev[a] = u[a]; yprev[a] = ynew[a]; }
                }
#pragma unroll
                for (int r = 0; r < NX; ++r) x[r] = xn[r];
            }
        }
        k += 1;   // work->iter += 1 (admm.cpp:328)

        // ------------------------------------------------- adaptive rho (admm.cpp:331-357), i = k-1
        rho_lc = rho; dlt_lc = dlt;   // update_linear_cost of this iteration ran with the pre-adaptation cache
        if constexpr (C::ADAPT) {
            if (do_adapt && (k - 1) > 0 && (k - 1) % 5 == 0) {
                a_pri = N::max(a_pri, rpu);                       // input rows: u - znew
                const T eps = T(1e-10);
                const T npri = a_pri / (a_prin + eps), ndua = a_dua / (a_duan + eps);
                T nr = rho * N::sqrt(npri / (ndua + eps));
                if (prm.rho_clip) nr = N::min(N::max(nr, static_cast<T>(prm.rho_min)), static_cast<T>(prm.rho_max));
                dlt += nr - rho;      // Kinf, Pinf += (rho' - rho) * d/drho (rho_benchmark.cpp:199-212)
                rho = nr;
            }
        }

        // ------------------------------------------------- termination (admm.cpp:253-271, 364-388)
        bool finish = false;
        int st = 11;
        if (k % check_every == 0) {
            res_px = rpx; res_dx = rdx * rho; res_pu = rpu; res_du = rdu * rho;
            if (res_px < tol_pri && res_pu < tol_pri && res_dx < tol_dua && res_du < tol_dua) { finish = true; st = 1; }
        }
        if (k >= max_iter) finish = true;
        if (active && finish) {
            store_chunk<SXL>(prm.x + (size_t)prob * SXL, [&](int e) { return static_cast<float>(V.get(e)); });
            store_chunk<SUL>(prm.u + (size_t)prob * SUL, [&](int e) { return static_cast<float>(Z.get(e)); });
            prm.iter[prob] = k;
            prm.status[prob] = st;
            if (prm.residuals) {
                float4 rr = make_float4(static_cast<float>(res_px), static_cast<float>(res_dx), static_cast<float>(res_pu), static_cast<float>(res_du));
                *reinterpret_cast<float4*>(prm.residuals + 4 * (size_t)prob) = rr;
            }
            if (prm.rho_out) prm.rho_out[prob] = static_cast<float>(rho);
            active = false;
        }
        if (!__any_sync(FULL, active)) continue;   // whole warp idle: go refill (or exit) without a backward sweep

        // ------------------------------------------------- backward Riccati sweep for the next iteration
        // q, r, p_N of update_linear_cost (admm.cpp:214-247) are formed on the fly with the rho / Pinf
        // that update_linear_cost saw; Kinf' uses the current (possibly adapted) Kinf.
        T p[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            const int e = (NH - 1) * NX + c;
            T w = V.get(e) - G.get(e);
            if constexpr (C::CONSTR) w += SXT.get(e);
            T pt = PT.get(c);
            if constexpr (C::ADAPT) pt = N::fma(dlt_lc, PT.get(NX + c), pt);
            p[c] = pt - rho_lc * w;
        }
#pragma unroll(C::TU)
        for (int i = NH - 2; i >= 0; --i) {
            T rr[NU], t[NU];
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                const int e = i * NU + a;
                T w = Z.get(e) - Y.get(e);
                if constexpr (C::CONSTR) w += SUT.get(e);
                rr[a] = -URR.get(e) - rho_lc * w;
                T acc = 0;
#pragma unroll
                for (int r = 0; r < NX; ++r) acc = N::fma(cB[r * NU + a], p[r], acc);
                t[a] = acc + rr[a] + cBPf[a];
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                T acc = 0;
#pragma unroll
                for (int b = 0; b < NU; ++b) acc = N::fma(cQuu[a * NU + b], t[b], acc);
                D.set(i * NU + a, acc);
            }
            T pn[NX];
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                const int e = i * NX + c;
                T w = V.get(e) - G.get(e);
                if constexpr (C::CONSTR) w += SXT.get(e);
                const T q = -XRQ.get(e) - rho_lc * w;
                T acc = 0;
#pragma unroll
                for (int r = 0; r < NX; ++r) acc = N::fma(cAK[c * NX + r], p[r], acc);
                T kr = 0;
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    T kc = cK[a * NX + c];
                    if constexpr (C::ADAPT) kc = N::fma(dlt, cdK[a * NX + c], kc);
                    kr = N::fma(kc, rr[a], kr);
                }
                pn[c] = (q + acc - kr) + cAPf[c];
            }
#pragma unroll
            for (int c = 0; c < NX; ++c) p[c] = pn[c];
        }
    }
}

template <class C>
inline size_t tpp_smem_bytes(int pack_elems) {
    return ((size_t)((pack_elems + 31) & ~31) + (size_t)C::COLS * C::BLOCK) * sizeof(typename C::T);
}

}  // namespace tmpc
