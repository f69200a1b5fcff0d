// fp64_bench.cu -- FP64-pipe micro-benchmarks for sm_100a (B200), behind the round-2 mixed-precision design question:
// can the per-element slack / dual / residual work of the ADMM iteration run in NATIVE double precision next to the fp32
// FFMA2 mat-vecs?  Measures per SM and clock the sustained rate of
//   dfma, dadd, dmnmx (fmin/fmax on doubles), f2f up (float -> double), f2f down (double -> float),
//   and ffma2 alone vs ffma2 + dadd interleaved in the same warps (do the two pipes overlap?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NACC = 8;

enum Op { DFMA, DADD, DMNMX, F2F_UP, F2F_DOWN, FFMA2, FFMA2_DADD, FFMA2_DADD_1TO4 };

template <int OP>
__global__ void kern(double* out, int iters, double x0, float f0) {
    double acc[NACC];
    float2 facc[NACC];
    float fl[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) { acc[j] = x0 + j + threadIdx.x; facc[j] = make_float2(f0 + j, f0 - j); fl[j] = f0 * (j + 1) + threadIdx.x; }
    const double c = x0 * 0.999, lo = -1e300 + x0, hi = 1e300 - x0;
    const float2 fc = make_float2(f0 * 0.999f, f0 * 1.001f), fx = make_float2(f0, 0.5f * f0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                if (OP == DFMA) acc[j] = fma(acc[j], c, x0);
                if (OP == DADD) acc[j] = acc[j] + c;
                if (OP == DMNMX) acc[j] = fmin(fmax(acc[j], lo + r), hi - j);          // 2 min/max per statement
                if (OP == F2F_UP) { acc[j] = acc[j] + (double)fl[j]; fl[j] = __int_as_float(__float_as_int(fl[j]) ^ (r + 1)); }   // 1 DADD + 1 F2F + 1 LOP
                if (OP == F2F_DOWN) { fl[j] = (float)acc[j]; acc[j] = __hiloint2double(__double2hiint(acc[j]) ^ (r + 1), __float_as_int(fl[j])); }
                if (OP == FFMA2 || OP == FFMA2_DADD || OP == FFMA2_DADD_1TO4) facc[j] = __ffma2_rn(facc[j], fc, fx);
                if (OP == FFMA2_DADD) acc[j] = acc[j] + c;
                if (OP == FFMA2_DADD_1TO4 && (j & 3) == 0) acc[j] = acc[j] + c;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j] + facc[j].x + facc[j].y + fl[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double ops_per_inner, int sms, double clock_ghz) {
    const int threads = 512, blocks = sms * 2, iters = 4000;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<OP><<<blocks, threads>>>(out, 10, 1.0, 1.0f);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<OP><<<blocks, threads>>>(out, iters, 1.0, 1.0f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double n = (double)threads * blocks * iters * 8 * NACC * ops_per_inner;
    printf("%-22s %8.3f ms  %8.2f G thread-ops/s  %7.2f thread-ops / clk / SM (at %.3f GHz)\n", name, ms, n / ms / 1e6, n / (ms * 1e-3) / sms / (clock_ghz * 1e9), clock_ghz);
    CK(cudaFree(out));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz / 1e6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<DFMA>("dfma", 1, p.multiProcessorCount, ghz);
    run<DADD>("dadd", 1, p.multiProcessorCount, ghz);
    run<DMNMX>("dmnmx (x2 per stmt)", 2, p.multiProcessorCount, ghz);
    run<F2F_UP>("f2f up + dadd", 1, p.multiProcessorCount, ghz);
    run<F2F_DOWN>("f2f down", 1, p.multiProcessorCount, ghz);
    run<FFMA2>("ffma2 (pairs)", 1, p.multiProcessorCount, ghz);
    run<FFMA2_DADD>("ffma2 + dadd 1:1", 1, p.multiProcessorCount, ghz);
    run<FFMA2_DADD_1TO4>("ffma2 + dadd 4:1", 1, p.multiProcessorCount, ghz);
    return 0;
}
