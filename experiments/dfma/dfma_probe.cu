// dfma_probe.cu -- what one warp / one SM sub-partition can do with DFMA on B200 (sm_100a).
//   chains = independent accumulators per thread (ILP), warps = warps per SM sub-partition (block = 128*warps threads, 1 block/SM)
//   pattern 0: acc = fma(acc, a, b)            one varying register operand (a, b loop invariant)
//   pattern 1: acc = fma(c_k, x_k, acc)        three distinct register operands per DFMA (x_k, c_k from register arrays)
//   pattern 2: S = fma(c,x_k,S); g_k = fma(c,P,g_k) pairs, c from a broadcast LDS.128 (the tape engines' inner pattern)
// Prints cycles per DFMA per warp and DFMA/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int PAT>
__global__ void __launch_bounds__(1024, 1) k(double* out, long long* cyc, int iters, double a, double b)
{
    __shared__ __align__(16) double tape[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tape[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[CH], x[8], g[8];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a + i * 1e-7; g[i] = i; }
    const double P = b * 1.0000001;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (PAT == 0) {
#pragma unroll
            for (int u = 0; u < 32; ++u)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
        } else if (PAT == 1) {
#pragma unroll
            for (int u = 0; u < 32; ++u)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = fma(x[(u + i) & 7], g[(u + 3 * i) & 7], acc[i]);
        } else {
            const unsigned base = (unsigned)__cvta_generic_to_shared(tape) + ((it & 7) << 9);
#pragma unroll
            for (int u = 0; u < 16; ++u) {      // 16 LDS.128 -> 32 coefficients -> 64 DFMA
                double c0, c1;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(base + 16 * u));
                acc[(2 * u) % CH] = fma(c0, x[(2 * u) & 7], acc[(2 * u) % CH]);
                g[(2 * u) & 7] = fma(c0, P, g[(2 * u) & 7]);
                acc[(2 * u + 1) % CH] = fma(c1, x[(2 * u + 1) & 7], acc[(2 * u + 1) % CH]);
                g[(2 * u + 1) & 7] = fma(c1, P, g[(2 * u + 1) & 7]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += g[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CH, int PAT>
void run(int warps_per_smsp, int sms, double* out, long long* cyc)
{
    const int iters = 2000;
    const int threads = 128 * warps_per_smsp;
    k<CH, PAT><<<sms, threads>>>(out, cyc, 10, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CH, PAT><<<sms, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter = PAT == 2 ? 64.0 : 32.0 * CH;
    const double dfma_warp = per_iter * iters;
    printf("pat %d chains %d warps/SMSP %d: %.2f clk per DFMA per warp, %.1f DFMA lanes/clk/SM (by clock64), %.2f TFLOP/s (by events)\n", PAT, CH, warps_per_smsp,
           (double)c / dfma_warp, dfma_warp * 32.0 * 4 * warps_per_smsp / (double)c, 2.0 * dfma_warp * threads * sms / (ms * 1e-3) * 1e-12);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024); cudaMalloc(&cyc, 8);
    for (int w : {1, 2, 3, 4}) { run<1, 0>(w, sms, out, cyc); run<2, 0>(w, sms, out, cyc); run<4, 0>(w, sms, out, cyc); run<8, 0>(w, sms, out, cyc); }
    for (int w : {1, 2, 3, 4}) { run<1, 1>(w, sms, out, cyc); run<2, 1>(w, sms, out, cyc); run<4, 1>(w, sms, out, cyc); run<8, 1>(w, sms, out, cyc); }
    for (int w : {1, 2, 3, 4}) { run<2, 2>(w, sms, out, cyc); run<4, 2>(w, sms, out, cyc); run<8, 2>(w, sms, out, cyc); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
