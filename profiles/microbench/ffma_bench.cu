// ffma_bench.cu -- FP32 CUDA-core roofline and operand-delivery micro-benchmarks for sm_100a.
// Measures, per SM and clock, how many FP32 FMAs the B200 sustains when the second multiplicand is
//   reg    : a register                      (pure pipe peak; this is the roofline denominator)
//   const  : a __constant__ bank operand with a working set of W floats (constant-cache capacity)
//   lds    : a shared-memory broadcast (LDS.128 feeding 4 FFMAs)
//   ffma2  : packed fma.rn.f32x2, register operands
// plus a mix test: FFMA stream + thread-private shared-memory column traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ffma_bench ffma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__constant__ float cpack[4096];

constexpr int NACC = 12;

template <int W>
__global__ void k_const(float* out, int iters, float x0) {
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = x0 + j + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < W; ++j) acc[j % NACC] = fmaf(acc[j % NACC], cpack[j], x0);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W>
__global__ void k_reg(float* out, int iters, float x0, float c0) {
    float acc[NACC];
    float c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = c0 + 0.001f * j;
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = x0 + j + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < W; ++j) acc[j % NACC] = fmaf(acc[j % NACC], c[j % 8], x0);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W>
__global__ void k_ffma2(float* out, int iters, float x0, float c0) {
    float2 acc[NACC];
    float2 c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = make_float2(c0 + 0.001f * j, c0 - 0.001f * j);
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = make_float2(x0 + j + threadIdx.x, x0 - j);
    const float2 xx = make_float2(x0, x0 * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < W; ++j) acc[j % NACC] = __ffma2_rn(acc[j % NACC], c[j % 8], xx);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory broadcast operands: one LDS.128 per 4 FFMAs, working set W floats
template <int W>
__global__ void k_lds(float* out, int iters, float x0) {
    __shared__ __align__(16) float sp[W];
    for (int j = threadIdx.x; j < W; j += blockDim.x) sp[j] = 0.5f + 1e-4f * j;
    __syncthreads();
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = x0 + j + threadIdx.x;
    const float4* s4 = reinterpret_cast<const float4*>(sp);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
            float4 c = s4[j];
            acc[(4 * j + 0) % NACC] = fmaf(acc[(4 * j + 0) % NACC], c.x, x0);
            acc[(4 * j + 1) % NACC] = fmaf(acc[(4 * j + 1) % NACC], c.y, x0);
            acc[(4 * j + 2) % NACC] = fmaf(acc[(4 * j + 2) % NACC], c.z, x0);
            acc[(4 * j + 3) % NACC] = fmaf(acc[(4 * j + 3) % NACC], c.w, x0);
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix: constant-operand FFMA stream + thread-private smem column traffic (1 LDS + 1 STS per R FFMAs)
template <int W, int R>
__global__ void k_mix(float* out, int iters, float x0) {
    extern __shared__ float col[];
    float* mine = col + threadIdx.x;
    const int BS = blockDim.x;
    for (int j = 0; j < 64; ++j) mine[j * BS] = x0 + j;
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = x0 + j + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            acc[j % NACC] = fmaf(acc[j % NACC], cpack[j], x0);
            if (j % R == 0) {
                float v = mine[((j / R) % 64) * BS];
                mine[((j / R + 7) % 64) * BS] = v + acc[j % NACC];
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int g_sms = 148;
static double g_clock_mhz = 0;

template <typename F>
double time_ms(F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

void report(const char* name, int W, int threads, int ctas_per_sm, double ms, double fma_per_thread, int mult = 1) {
    double total = fma_per_thread * threads * ctas_per_sm * g_sms * mult;
    double tflops = 2.0 * total / (ms * 1e-3) / 1e12;
    printf("%-10s W=%5d thr=%4d cta/sm=%d  %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM(@%.0fMHz)\n", name, W, threads, ctas_per_sm, ms,
           tflops, total / (ms * 1e-3) / g_sms / (g_clock_mhz * 1e6), g_clock_mhz);
}

template <int W> void run_const(float* out, int threads, int cps) {
    int iters = (1 << 22) / W;
    double ms = time_ms([&] { k_const<W><<<g_sms * cps, threads>>>(out, iters, 0.999f); });
    report("const", W, threads, cps, ms, (double)iters * W);
}
template <int W> void run_lds(float* out, int threads, int cps) {
    int iters = (1 << 22) / W;
    double ms = time_ms([&] { k_lds<W><<<g_sms * cps, threads>>>(out, iters, 0.999f); });
    report("lds128", W, threads, cps, ms, (double)iters * W);
}
template <int W, int R> void run_mix(float* out, int threads, int cps) {
    int iters = (1 << 21) / W;
    CK(cudaFuncSetAttribute(k_mix<W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    double ms = time_ms([&] { k_mix<W, R><<<g_sms * cps, threads, 64 * threads * sizeof(float)>>>(out, iters, 0.999f); });
    char nm[32]; snprintf(nm, 32, "mix1:%d", R);
    report(nm, W, threads, cps, ms, (double)iters * W);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    g_clock_mhz = clk / 1000.0;
    printf("device %s, %d SMs, max clock %.0f MHz, nominal FP32 peak %.1f TFLOP/s\n", p.name, g_sms, g_clock_mhz, g_sms * 128 * 2 * g_clock_mhz * 1e6 / 1e12);
    std::vector<float> h(4096);
    for (int i = 0; i < 4096; ++i) h[i] = 0.5f + 1e-4f * i;
    CK(cudaMemcpyToSymbol(cpack, h.data(), sizeof(float) * 4096));
    float* out; CK(cudaMalloc(&out, sizeof(float) * g_sms * 8 * 1024));

    for (int threads : {128, 256, 512, 1024}) {
        int iters = (1 << 22) / 512;
        double ms = time_ms([&] { k_reg<512><<<g_sms, threads>>>(out, iters, 0.999f, 0.5f); });
        report("reg", 512, threads, 1, ms, (double)iters * 512);
        ms = time_ms([&] { k_ffma2<512><<<g_sms, threads>>>(out, iters, 0.999f, 0.5f); });
        report("ffma2", 512, threads, 1, ms, (double)iters * 512, 2);
    }
    for (int threads : {128, 256, 512}) {
        run_const<64>(out, threads, 1);
        run_const<256>(out, threads, 1);
        run_const<512>(out, threads, 1);
        run_const<1024>(out, threads, 1);
        run_const<2048>(out, threads, 1);
        run_const<4096>(out, threads, 1);
    }
    for (int threads : {128, 256, 512}) {
        run_lds<512>(out, threads, 1);
        run_lds<2048>(out, threads, 1);
    }
    for (int threads : {128, 256}) {
        run_mix<512, 4>(out, threads, 1);
        run_mix<512, 8>(out, threads, 1);
        run_mix<512, 16>(out, threads, 1);
    }
    return 0;
}
