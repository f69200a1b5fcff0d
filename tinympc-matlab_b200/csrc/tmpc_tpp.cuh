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
// Mapping.  One CUDA thread owns one MPC problem for all of its ADMM iterations; a mat-vec is then a
// pure FMA stream with no shuffles.  Operand delivery decides the speed of that stream
// (profiles/microbench/RESULTS.md): a shared-memory broadcast returns at most one coefficient per
// clock per SM (<= 25 % of the FP32 pipe), so the family matrices A, B, Kinf, AmBKt, Quu_inv and the
// shared bounds travel in the kernel-parameter constant bank and reach the FMA pipe through uniform
// registers (LDCU), paired two columns at a time into packed FFMA2 (fma.rn.f32x2).  The cold tables
// (Pinf, d0, cost diagonals, linear rows) are staged once per CTA into shared memory with a TMA bulk
// copy (cp.async.bulk + mbarrier).  Per-problem state lives in conflict-free shared-memory columns
// ([element][thread]) for the whole solve; HBM is touched only to read x0/Xref/Uref once and to write
// the solution once.  Problems need 1..max_iter iterations, so lanes are refilled: a lane that
// finishes claims the next problem from a global counter (warp-aggregated atomic) while its
// neighbours keep iterating.
//
// State compression.  For the box constraint the reference keeps four arrays per variable: slack
// vnew, previous slack v, dual g (and the trajectory x).  Since vnew = clamp(x + g) and
// g_new = (g + x) - vnew (admm.cpp:85,92,184), the single pre-clamp value t = x + g determines both:
// vnew = clamp(t), g_new = t - vnew.  Only t is stored (TV / TZ), which halves shared-memory traffic
// and doubles the number of resident problems per SM.
//
// Loop rotation.  The reference runs backward -> forward -> slack -> dual -> linear cost -> check.
// On a cold workspace (q = r = p = 0) the first backward pass is problem independent, so its result
// d0 is precomputed on the host; each iteration here is forward+slack+dual -> check -> backward (for
// the next iteration): the same sequence of values, without the backward pass of the last iteration.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tmpc_common.h"

namespace tmpc {

enum : int { FEAT_BOX = 0, FEAT_CONSTR = 1, FEAT_ADAPT = 2 };

// ----------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
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
// family tables that ride in the kernel-parameter constant bank (row-major, plus transposed copies
// so that every dot product reads its coefficients as adjacent pairs)
// ----------------------------------------------------------------------------------------------
template <typename T, int NX, int NU, int NH, bool ADAPT>
struct ConstPack {
    T A[NX * NX];      // x_next rows
    T B[NX * NU];
    T BT[NU * NX];     // B'  (backward: B' p)
    T K[NU * NX];      // Kinf (forward)
    T KT[NX * NU];     // -Kinf' (backward: - Kinf' r, negated so that it chains into the AmBKt accumulation)
    T AK[NX * NX];     // AmBKt
    T Quu[NU * NU];
    T f[NX], APf[NX], BPf[NU];
    T Qd[NX], Rd[NU];
    T xmin[NX * NH], xmax[NX * NH], umin[NU * (NH - 1)], umax[NU * (NH - 1)];
    T dK[ADAPT ? NU * NX : 2], dKT[ADAPT ? NX * NU : 2];
    T AT[ADAPT ? NX * NX : 2];   // A' (adaptive rho: A' g)
};

template <typename T, int NX, int NU, int NH, bool ADAPT>
inline void fill_const_pack(ConstPack<T, NX, NU, NH, ADAPT>& c, const double* pk, const PackLayout& L) {
    auto cp = [&](T* dst, int at, int n) { for (int i = 0; i < n; ++i) dst[i] = static_cast<T>(pk[at + i]); };
    auto tr = [&](T* dst, int at, int rows, int cols) {   // dst (cols x rows) = transpose of row-major (rows x cols)
        for (int r = 0; r < rows; ++r) for (int k = 0; k < cols; ++k) dst[k * rows + r] = static_cast<T>(pk[at + r * cols + k]);
    };
    cp(c.A, L.A, NX * NX); cp(c.B, L.B, NX * NU); tr(c.BT, L.B, NX, NU);
    cp(c.K, L.Kinf, NU * NX); tr(c.KT, L.Kinf, NU, NX);
    for (int i = 0; i < NX * NU; ++i) c.KT[i] = -c.KT[i];
    cp(c.AK, L.AmBKt, NX * NX); cp(c.Quu, L.Quu_inv, NU * NU);
    cp(c.f, L.f, NX); cp(c.APf, L.APf, NX); cp(c.BPf, L.BPf, NU); cp(c.Qd, L.Qd, NX); cp(c.Rd, L.Rd, NU);
    cp(c.xmin, L.xmin, NX * NH); cp(c.xmax, L.xmax, NX * NH); cp(c.umin, L.umin, NU * (NH - 1)); cp(c.umax, L.umax, NU * (NH - 1));
    if (ADAPT) { cp(c.dK, L.dKinf, NU * NX); tr(c.dKT, L.dKinf, NU, NX); tr(c.AT, L.A, NX, NX); for (int i = 0; i < NX * NU; ++i) c.dKT[i] = -c.dKT[i]; }
}

// ----------------------------------------------------------------------------------------------
// per-thread state: a shared-memory column [elem][thread]
// ----------------------------------------------------------------------------------------------
template <typename T, int OFF, int BLOCK>
struct Col {
    T* base;
    __device__ __forceinline__ explicit Col(T* colbase) : base(colbase + OFF * BLOCK) {}
    __device__ __forceinline__ T get(int i) const { return base[i * BLOCK]; }
    __device__ __forceinline__ void set(int i, T v) { base[i * BLOCK] = v; }
};

enum : int { REFS_NONE = 0, REFS_SMEM = 1, REFS_L2 = 2 };

template <typename T_, int NX_, int NU_, int NH_, int FEAT_, int BLOCK_, int REFS_, bool PPB_, int MINB_ = 1, bool FB_ = false>
struct TppCfg {
    using T = T_;
    static constexpr int NX = NX_, NU = NU_, NH = NH_, FEAT = FEAT_, BLOCK = BLOCK_, MINB = MINB_;
    static constexpr int REFMODE = REFS_;                   // where the per-problem reference terms live
    static constexpr bool REFS = REFS_ != REFS_NONE;        // per-problem Xref/Uref present
    static constexpr bool REFS_SM = REFS_ == REFS_SMEM;     // ... in shared-memory columns
    static constexpr bool REFS_G = REFS_ == REFS_L2;        // ... in an L2-resident, lane-interleaved global scratch
    static constexpr bool PPB = PPB_;                       // per-problem bounds read from global memory
    static constexpr bool FB = FB_ && !PPB_;                // "fast box": shared bounds are time-invariant and contain 0
    static constexpr bool CONSTR = FEAT_ == FEAT_CONSTR;
    static constexpr bool ADAPT = FEAT_ == FEAT_ADAPT;
    static constexpr int SX = NX * NH, SU = NU * (NH - 1);
    using CPack = ConstPack<T_, NX_, NU_, NH_, ADAPT>;
    static constexpr int VX = (NX_ % 4 == 0) ? 4 : ((NX_ % 2 == 0) ? 2 : 1);   // REFS_L2: elements per vector load
    static constexpr int VU = (NU_ % 4 == 0) ? 4 : ((NU_ % 2 == 0) ? 2 : 1);
    // element offsets of the shared-memory columns
    static constexpr int oTV = 0;
    static constexpr int oTZ = oTV + SX;
    static constexpr int oD = oTZ + SU;
    static constexpr int oXRQ = oD + SU;
    static constexpr int oURR = oXRQ + (REFS_SM ? SX : 0);
    static constexpr int oGC = oURR + (REFS_SM ? SU : 0);
    static constexpr int oGL = oGC + (CONSTR ? SX : 0);
    static constexpr int oSX = oGL + (CONSTR ? SX : 0);
    static constexpr int oYC = oSX + (CONSTR ? SX : 0);
    static constexpr int oYL = oYC + (CONSTR ? SU : 0);
    static constexpr int oSU = oYL + (CONSTR ? SU : 0);
    static constexpr int oSCR = oSU + (CONSTR ? SU : 0);
    static constexpr int COLS = oSCR + (CONSTR ? (NX > NU ? NX : NU) : 0);   // + cone scratch column
};

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return sqrtf(a); }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double min(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};

// init + sum_c m(c) * x[c]; fp32 pairs adjacent columns into packed FFMA2 (two partial sums)
template <int C, typename FM>
__device__ __forceinline__ float dot(FM&& m, const float (&x)[C], float init) {
    if constexpr (C >= 2) {
        float2 acc = make_float2(init, 0.f);
#pragma unroll
        for (int c = 0; c + 1 < C; c += 2) acc = __ffma2_rn(make_float2(m(c), m(c + 1)), make_float2(x[c], x[c + 1]), acc);
        float r = acc.x + acc.y;
        if constexpr (C & 1) r = fmaf(m(C - 1), x[C - 1], r);
        return r;
    } else {
        return fmaf(m(0), x[0], init);
    }
}
template <int C, typename FM>
__device__ __forceinline__ double dot(FM&& m, const double (&x)[C], double init) {
    double acc = init;
#pragma unroll
    for (int c = 0; c < C; ++c) acc = ::fma(m(c), x[c], acc);
    return acc;
}

// init + sum_c m1(c) x1[c] + sum_c m2(c) x2[c] in ONE accumulation chain (one pair-sum instead of two)
template <int C1, int C2, typename F1, typename F2>
__device__ __forceinline__ float dot2(F1&& m1, const float (&x1)[C1], F2&& m2, const float (&x2)[C2], float init) {
    float2 acc = make_float2(init, 0.f);
#pragma unroll
    for (int c = 0; c + 1 < C1; c += 2) acc = __ffma2_rn(make_float2(m1(c), m1(c + 1)), make_float2(x1[c], x1[c + 1]), acc);
#pragma unroll
    for (int c = 0; c + 1 < C2; c += 2) acc = __ffma2_rn(make_float2(m2(c), m2(c + 1)), make_float2(x2[c], x2[c + 1]), acc);
    if constexpr (C1 & 1) acc.x = fmaf(m1(C1 - 1), x1[C1 - 1], acc.x);
    if constexpr (C2 & 1) acc.y = fmaf(m2(C2 - 1), x2[C2 - 1], acc.y);
    return acc.x + acc.y;
}
template <int C1, int C2, typename F1, typename F2>
__device__ __forceinline__ double dot2(F1&& m1, const double (&x1)[C1], F2&& m2, const double (&x2)[C2], double init) {
    double acc = init;
#pragma unroll
    for (int c = 0; c < C1; ++c) acc = ::fma(m1(c), x1[c], acc);
#pragma unroll
    for (int c = 0; c < C2; ++c) acc = ::fma(m2(c), x2[c], acc);
    return acc;
}

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

// vector type of V elements of T for the lane-interleaved scratch
template <typename T, int V> struct VecOf;
template <> struct VecOf<float, 4> { using type = float4; static __device__ __forceinline__ void unpack(const float4& v, float* d) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; } };
template <> struct VecOf<float, 2> { using type = float2; static __device__ __forceinline__ void unpack(const float2& v, float* d) { d[0] = v.x; d[1] = v.y; } };
template <> struct VecOf<float, 1> { using type = float; static __device__ __forceinline__ void unpack(const float& v, float* d) { d[0] = v; } };
template <> struct VecOf<double, 4> { using type = double4; static __device__ __forceinline__ void unpack(const double4& v, double* d) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; } };
template <> struct VecOf<double, 2> { using type = double2; static __device__ __forceinline__ void unpack(const double2& v, double* d) { d[0] = v.x; d[1] = v.y; } };
template <> struct VecOf<double, 1> { using type = double; static __device__ __forceinline__ void unpack(const double& v, double* d) { d[0] = v; } };

template <class C>
__global__ void __launch_bounds__(C::BLOCK, C::MINB)
tpp_kernel(const __grid_constant__ SolveParams prm, const __grid_constant__ typename C::CPack cp) {
    using T = typename C::T;
    using N = Num<T>;
    using SP = StaticPack<C::NX, C::NU, C::NH>;
    constexpr int NX = C::NX, NU = C::NU, NH = C::NH, BLOCK = C::BLOCK, SXL = C::SX, SUL = C::SU;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t pack_bar;
    T* pack = reinterpret_cast<T*>(smem_raw);
    const uint32_t pack_bytes = static_cast<uint32_t>(prm.pack_elems) * sizeof(T);

    // ---- stage the cold family tables into shared memory: one TMA bulk copy per CTA ----
    if (threadIdx.x == 0) {
        mbar_init(&pack_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&pack_bar, pack_bytes);
        tma_bulk_g2s(pack, prm.pack, pack_bytes, &pack_bar);
    }
    __syncthreads();
    mbar_wait(&pack_bar, 0);

    T* colbase = pack + ((prm.pack_elems + 31) & ~31) + threadIdx.x;
    Col<T, C::oTV, BLOCK> TV(colbase);      // t = x + g   (pre-clamp state slack)
    Col<T, C::oTZ, BLOCK> TZ(colbase);      // t = u + y   (pre-clamp input slack)
    Col<T, C::oD, BLOCK> D(colbase);        // d of the backward pass
    Col<T, C::oXRQ, BLOCK> XRQ(colbase);    // Xref .* Q
    Col<T, C::oURR, BLOCK> URR(colbase);    // Uref .* R
    Col<T, C::oGC, BLOCK> GC(colbase);      // cone duals (state)
    Col<T, C::oGL, BLOCK> GL(colbase);      // linear duals (state)
    Col<T, C::oSX, BLOCK> SXT(colbase);     // (vc - gc) + (vl - gl)
    Col<T, C::oYC, BLOCK> YC(colbase);
    Col<T, C::oYL, BLOCK> YL(colbase);
    Col<T, C::oSU, BLOCK> SUT(colbase);
    T* scr = colbase + C::oSCR * BLOCK;     // cone scratch column (CONSTR only)

    // REFS_L2: reference terms in a global scratch laid out [element / V][slot][V] (V = 4/2/1 elements per lane and
    // load): coalesced across lanes, one vector load per V elements, re-read every iteration out of L2
    const size_t slots = (size_t)gridDim.x * BLOCK;
    const size_t slot = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    T* const gxr = static_cast<T*>(prm.ref_scratch) + slot * C::VX;
    T* const gur = static_cast<T*>(prm.ref_scratch) + (size_t)SXL * slots + slot * C::VU;
    auto xrq_set = [&](int e, T v) { if constexpr (C::REFS_G) gxr[(size_t)(e / C::VX) * slots * C::VX + (e % C::VX)] = v; else XRQ.set(e, v); };
    auto urr_set = [&](int e, T v) { if constexpr (C::REFS_G) gur[(size_t)(e / C::VU) * slots * C::VU + (e % C::VU)] = v; else URR.set(e, v); };
    // fetch the NX (NU) reference terms of time step i
    auto xrq_step = [&](int i, T (&dst)[NX]) {
        if constexpr (C::REFS_G) {
            using V = typename VecOf<T, C::VX>::type;
#pragma unroll
            for (int k = 0; k < NX / C::VX; ++k) {
                const V v = *reinterpret_cast<const V*>(gxr + (size_t)(i * (NX / C::VX) + k) * slots * C::VX);
                VecOf<T, C::VX>::unpack(v, &dst[k * C::VX]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NX; ++c) dst[c] = XRQ.get(i * NX + c);
        }
    };
    auto urr_step = [&](int i, T (&dst)[NU]) {
        if constexpr (C::REFS_G) {
            using V = typename VecOf<T, C::VU>::type;
#pragma unroll
            for (int k = 0; k < NU / C::VU; ++k) {
                const V v = *reinterpret_cast<const V*>(gur + (size_t)(i * (NU / C::VU) + k) * slots * C::VU);
                VecOf<T, C::VU>::unpack(v, &dst[k * C::VU]);
            }
        } else {
#pragma unroll
            for (int a = 0; a < NU; ++a) dst[a] = URR.get(i * NU + a);
        }
    };
    (void)slots; (void)slot;

    const T* cP = pack + SP::Pinf;
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
    const bool en_sb = prm.en_state_bound != 0, en_ib = prm.en_input_bound != 0;
    (void)en_sb; (void)en_ib;

    const int lane = threadIdx.x & 31;
    int prob = 0;           // problem owned by this lane
    bool active = false;    // lane holds an unfinished problem
    bool exhausted = false; // the work counter ran past the batch
    int k = 0;              // ADMM iterations done on the current problem
    T res_px = 0, res_dx = 0, res_pu = 0, res_du = 0;   // last evaluated residuals (admm.cpp:257-260)
    // adaptive rho state (cache->rho and the Taylor offset of Kinf/Pinf); *_lc are the values
    // update_linear_cost saw (it runs BEFORE the adaptation inside an iteration)
    T rho = rho0, rho_lc = rho0, dlt = 0, dlt_lc = 0;
    T x0r[NX], ptr_[C::ADAPT ? 2 * NX : NX];            // x0 and the terminal term -(xref_N' Pinf)'
#pragma unroll
    for (int r = 0; r < NX; ++r) { x0r[r] = 0; ptr_[r] = 0; if constexpr (C::ADAPT) ptr_[NX + r] = 0; }

    // box bounds of trajectory element e (state) / (input)
    auto xbounds = [&](int i, int r, size_t pb, T& lo, T& hi) {
        const int e = i * NX + r; (void)e;
        if constexpr (C::PPB) {
            lo = en_sb ? static_cast<T>(__ldg(prm.x_min + pb + e)) : -N::inf();
            hi = en_sb ? static_cast<T>(__ldg(prm.x_max + pb + e)) : N::inf();
        } else if constexpr (C::FB) { lo = cp.xmin[r]; hi = cp.xmax[r]; }
        else { lo = cp.xmin[e]; hi = cp.xmax[e]; }
    };
    auto ubounds = [&](int i, int a, size_t pb, T& lo, T& hi) {
        const int e = i * NU + a; (void)e;
        if constexpr (C::PPB) {
            lo = en_ib ? static_cast<T>(__ldg(prm.u_min + pb + e)) : -N::inf();
            hi = en_ib ? static_cast<T>(__ldg(prm.u_max + pb + e)) : N::inf();
        } else if constexpr (C::FB) { lo = cp.umin[a]; hi = cp.umax[a]; }
        else { lo = cp.umin[e]; hi = cp.umax[e]; }
    };

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
                        prob = 0;   // keeps the (unused) per-problem reads of an idle lane in range
                    } else {
                        active = true;
                        k = 0;
                        res_px = res_dx = res_pu = res_du = 0;
                        rho = rho_lc = rho0; dlt = dlt_lc = 0;
                        load_chunk<NX>(prm.x0 + (size_t)prob * NX, [&](int i, float v) { x0r[i] = static_cast<T>(v); });
                        // Xref -> XRQ = Xref .* Q (work->Q = diag(Q)+rho, admm.cpp:218) and the terminal
                        // term PT = -(xref_N' Pinf)' (admm.cpp:238)
                        if constexpr (C::REFS) {
                            // one time step per (rolled) trip keeps the number of loads in flight -- and registers -- small
                            T xr_last[NX];
#pragma unroll
                            for (int r = 0; r < NX; ++r) xr_last[r] = 0;
                            if (prm.Xref) {
                                const float* src = prm.Xref + (size_t)prob * SXL;
#pragma unroll 1
                                for (int i = 0; i < NH; ++i)
                                    load_chunk<NX>(src + i * NX, [&](int r, float v) {
                                        xrq_set(i * NX + r, static_cast<T>(v) * cp.Qd[r]);
                                        xr_last[r] = static_cast<T>(v);      // after the last trip: xref_N
                                    });
                            } else {
#pragma unroll 4
                                for (int e = 0; e < SXL; ++e) xrq_set(e, T(0));
                            }
#pragma unroll
                            for (int c = 0; c < NX; ++c) {
                                T acc = 0, acc1 = 0;
#pragma unroll
                                for (int r = 0; r < NX; ++r) {
                                    acc = N::fma(xr_last[r], cP[r * NX + c], acc);
                                    if constexpr (C::ADAPT) acc1 = N::fma(xr_last[r], cdP[r * NX + c], acc1);
                                }
                                ptr_[c] = -acc;
                                if constexpr (C::ADAPT) ptr_[NX + c] = -acc1;
                            }
                            if (prm.Uref) {
                                const float* src = prm.Uref + (size_t)prob * SUL;
#pragma unroll 1
                                for (int i = 0; i < NH - 1; ++i)
                                    load_chunk<NU>(src + i * NU, [&](int a, float v) { urr_set(i * NU + a, static_cast<T>(v) * cp.Rd[a]); });
                            } else {
#pragma unroll 4
                                for (int e = 0; e < SUL; ++e) urr_set(e, T(0));
                            }
                        }
                        // cold workspace (tiny_api.cpp:68-105): duals and slacks zero, d = d0
#pragma unroll 4
                        for (int e = 0; e < SXL; ++e) TV.set(e, T(0));
#pragma unroll 4
                        for (int e = 0; e < SUL; ++e) { TZ.set(e, T(0)); D.set(e, pack[SP::d0 + e]); }
                        if constexpr (C::CONSTR) {
#pragma unroll 4
                            for (int e = 0; e < SXL; ++e) { GC.set(e, T(0)); GL.set(e, T(0)); SXT.set(e, T(0)); }
#pragma unroll 4
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
        const bool first = (k == 0);   // cold start: v = 0, g = 0 whatever the bounds are
        T x[NX];
        T xprev[NX], gprev[NX], uprev[NU], yprev[NU];   // ADAPT: lagged column for the A'g terms
#pragma unroll
        for (int r = 0; r < NX; ++r) { x[r] = x0r[r]; xprev[r] = 0; gprev[r] = 0; }
#pragma unroll
        for (int a = 0; a < NU; ++a) { uprev[a] = 0; yprev[a] = 0; }
        const size_t pbx = (size_t)prob * SXL, pbu = (size_t)prob * SUL;

#pragma unroll 1
        for (int i = 0; i < NH; ++i) {
            // ---- state column i: vnew = clamp(x + g), g += x - vnew (admm.cpp:85,92,184)
            T gnew[NX];
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                const int e = i * NX + r;
                T lo, hi;
                xbounds(i, r, pbx, lo, hi);
                const T tvo = TV.get(e);
                T vo = N::min(hi, N::max(lo, tvo));
                T g = tvo - vo;
                if constexpr (!C::FB) { if (first) { vo = 0; g = 0; } }   // FB: 0 is inside the box, clamp(0) = 0 already
                const T tvn = x[r] + g;
                const T vn = N::min(hi, N::max(lo, tvn));
                rpx = N::max(rpx, N::abs(x[r] - vn));
                rdx = N::max(rdx, N::abs(vo - vn));
                TV.set(e, tvn);
                gnew[r] = tvn - vn;
                if constexpr (C::ADAPT) {
                    if (do_adapt && i > 0) {   // dynamics rows of A_matrix: (A x + B u - x_next) - vnew_next = -f - vnew
                        a_pri = N::max(a_pri, N::abs(cp.f[r] + vn));
                        a_prin = N::max(a_prin, N::max(N::abs(vn), N::abs(cp.f[r])));
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
                        T aty = dot<NX>([&](int r) { return cp.AT[c * NX + r]; }, gnew, T(0));
                        if (i > 1) aty -= gprev[c];
                        const T qx = cp.Qd[c] * xprev[c];   // Px = qv = Q .* x for columns < N-1
                        a_dua = N::max(a_dua, N::abs(qx + qx + aty));
                        a_duan = N::max(a_duan, N::max(N::abs(qx), N::abs(aty)));
                    }
#pragma unroll
                    for (int a = 0; a < NU; ++a) {
                        const T aty = dot<NX>([&](int r) { return cp.BT[a * NX + r]; }, gnew, yprev[a]);
                        const T ru = cp.Rd[a] * uprev[a];
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
                        const T qx = cp.Qd[r] * x[r];
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
                    T acc = dot<NX>([&](int c) { return cp.K[a * NX + c]; }, x, T(0));
                    if constexpr (C::ADAPT) acc = N::fma(dlt, dot<NX>([&](int c) { return cp.dK[a * NX + c]; }, x, T(0)), acc);
                    u[a] = -acc - D.get(i * NU + a);
                }
                // ---- input column i: znew = clamp(u + y), y += u - znew (admm.cpp:88,97,187)
                T ynew[NU];
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    const int e = i * NU + a;
                    T lo, hi;
                    ubounds(i, a, pbu, lo, hi);
                    const T tzo = TZ.get(e);
                    T zo = N::min(hi, N::max(lo, tzo));
                    T yv = tzo - zo;
                    if constexpr (!C::FB) { if (first) { zo = 0; yv = 0; } }
                    const T tzn = u[a] + yv;
                    const T zn = N::min(hi, N::max(lo, tzn));
                    rpu = N::max(rpu, N::abs(u[a] - zn));
                    rdu = N::max(rdu, N::abs(zo - zn));
                    TZ.set(e, tzn);
                    ynew[a] = tzn - zn;
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
                    xn[r] = dot2<NX, NU>([&](int c) { return cp.A[r * NX + c]; }, x, [&](int a) { return cp.B[r * NU + a]; }, u, cp.f[r]);
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
            // solution = (vnew, znew) = clamp of the stored pre-clamp values
#pragma unroll 1
            for (int i = 0; i < NH; ++i)
                store_chunk<NX>(prm.x + pbx + i * NX, [&](int r) {
                    T lo, hi; xbounds(i, r, pbx, lo, hi);
                    return static_cast<float>(N::min(hi, N::max(lo, TV.get(i * NX + r))));
                });
#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i)
                store_chunk<NU>(prm.u + pbu + i * NU, [&](int a) {
                    T lo, hi; ubounds(i, a, pbu, lo, hi);
                    return static_cast<float>(N::min(hi, N::max(lo, TZ.get(i * NU + a))));
                });
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
        // v - g = 2 clamp(t) - t for the box slack/dual pair.
        T p[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            const int e = (NH - 1) * NX + c;
            T lo, hi;
            xbounds(NH - 1, c, pbx, lo, hi);
            const T tv = TV.get(e);
            const T v = N::min(hi, N::max(lo, tv));
            T w = N::fma(T(2), v, -tv);
            if constexpr (C::CONSTR) w += SXT.get(e);
            T pt = ptr_[c];
            if constexpr (C::ADAPT) pt = N::fma(dlt_lc, ptr_[NX + c], pt);
            p[c] = pt - rho_lc * w;
        }
        // reference terms of step i are fetched one step ahead (they may live in L2)
        T xq_nx[NX], ur_nx[NU];
#pragma unroll
        for (int c = 0; c < NX; ++c) xq_nx[c] = T(0);
#pragma unroll
        for (int a = 0; a < NU; ++a) ur_nx[a] = T(0);
        if constexpr (C::REFS) { xrq_step(NH - 2, xq_nx); urr_step(NH - 2, ur_nx); }
#pragma unroll 1
        for (int i = NH - 2; i >= 0; --i) {
            T rr[NU], t[NU];
            T xq_cur[NX], ur_cur[NU];
#pragma unroll
            for (int c = 0; c < NX; ++c) xq_cur[c] = xq_nx[c];
#pragma unroll
            for (int a = 0; a < NU; ++a) ur_cur[a] = ur_nx[a];
            if constexpr (C::REFS) {
                if (i > 0) { xrq_step(i - 1, xq_nx); urr_step(i - 1, ur_nx); }
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                const int e = i * NU + a;
                T lo, hi;
                ubounds(i, a, pbu, lo, hi);
                const T tz = TZ.get(e);
                const T z = N::min(hi, N::max(lo, tz));
                T w = N::fma(T(2), z, -tz);
                if constexpr (C::CONSTR) w += SUT.get(e);
                rr[a] = -ur_cur[a] - rho_lc * w;
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) t[a] = dot<NX>([&](int r) { return cp.BT[a * NX + r]; }, p, rr[a] + cp.BPf[a]);
#pragma unroll
            for (int a = 0; a < NU; ++a) D.set(i * NU + a, dot<NU>([&](int b) { return cp.Quu[a * NU + b]; }, t, T(0)));
            T pn[NX];
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                const int e = i * NX + c;
                T lo, hi;
                xbounds(i, c, pbx, lo, hi);
                const T tv = TV.get(e);
                const T v = N::min(hi, N::max(lo, tv));
                T w = N::fma(T(2), v, -tv);
                if constexpr (C::CONSTR) w += SXT.get(e);
                const T q = -xq_cur[c] - rho_lc * w;
                // q + APf + AmBKt p - Kinf' r in one chain (the pack holds -Kinf')
                T acc = dot2<NX, NU>([&](int r) { return cp.AK[c * NX + r]; }, p, [&](int a) { return cp.KT[c * NU + a]; }, rr, q + cp.APf[c]);
                if constexpr (C::ADAPT) acc = N::fma(dlt, dot<NU>([&](int a) { return cp.dKT[c * NU + a]; }, rr, T(0)), acc);
                pn[c] = acc;
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
