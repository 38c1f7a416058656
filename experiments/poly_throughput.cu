// Micro-benchmarks that size the FP64 roofline for the MB-pol polynomial kernels (run under gpurun).
//   1. DFMA peak: independent register-resident FMA chains -> the measured FP64 "peak" used by bench.py
//   2. raw throughput of the generated 3B polynomial device functions (no geometry, no scatter)
// Build: make -C dlpoly2-mbpol-heat_b200/csrc exp     Run: build/poly_throughput [nthreads_per_block]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include POLY_GEN_H

__constant__ double mbpol_coef2b[1153];
__constant__ double mbpol_coef3b[1163];

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_peak(double* out, int iters, double a, double b)
{
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int B>
__global__ void __launch_bounds__(256, 1) poly3b_raw(const double* __restrict__ xin, double* __restrict__ out, int n)
{
    // thread t of the CTA owns clusters  (blockIdx*B + b)*blockDim + t,  b = 0..B-1
    double buf[B * POLY3B_RECORD_DOUBLES];
    for (int b = 0; b < B; ++b) {
        size_t t = ((size_t)blockIdx.x * B + b) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 36; ++i) buf[b * POLY3B_RECORD_DOUBLES + i] = xin[(size_t)i * n + t];
    }
    POLY3B_CALL_ALL(buf, B);
    for (int b = 0; b < B; ++b) {
        size_t t = ((size_t)blockIdx.x * B + b) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 37; ++i) out[(size_t)i * n + t] = buf[b * POLY3B_RECORD_DOUBLES + 36 + i];
    }
}

template <int B>
__global__ void __launch_bounds__(256, 1) poly2b_raw(const double* __restrict__ xin, double* __restrict__ out, int n)
{
    double buf[B * POLY2B_RECORD_DOUBLES];
    for (int b = 0; b < B; ++b) {
        size_t t = ((size_t)blockIdx.x * B + b) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 31; ++i) buf[b * POLY2B_RECORD_DOUBLES + i] = xin[(size_t)i * n + t];
    }
    POLY2B_CALL_ALL(buf, B);
    for (int b = 0; b < B; ++b) {
        size_t t = ((size_t)blockIdx.x * B + b) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 32; ++i) out[(size_t)i * n + t] = buf[b * POLY2B_RECORD_DOUBLES + 31 + i];
    }
}

template <int B>
static void run_poly(int sms, int bs, cudaEvent_t e0, cudaEvent_t e1)
{
    float ms;
    int n = sms * bs * B * 4;      // 4 CTAs per SM, B clusters per thread
    std::vector<double> hx((size_t)36 * n);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = 0.3 + 0.5 * ((i * 2654435761u) % 1000) / 1000.0;
    double *dx, *dout; CK(cudaMalloc(&dx, hx.size() * 8)); CK(cudaMalloc(&dout, (size_t)37 * n * 8));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        poly3b_raw<B><<<n / bs / B, bs>>>(dx, dout, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) printf("poly3b_raw B=%d parts=%d bs=%d n=%d: %.3f ms  %.2f Mtrimer/s  %.2f T fp64-instr-lanes/s  (ref-flop %.2f TFLOP/s)\n", B, POLY3B_NPARTS, bs, n, ms,
               n / ms * 1e-3, (double)n * POLY3B_FP64_INSTR / ms * 1e-9, (double)n * 111818.0 / ms * 1e-9);
    }
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        poly2b_raw<B><<<n / bs / B, bs>>>(dx, dout, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) printf("poly2b_raw B=%d parts=%d bs=%d n=%d: %.3f ms  %.2f Mdimer/s  %.2f T fp64-instr-lanes/s  (ref-flop %.2f TFLOP/s)\n", B, POLY2B_NPARTS, bs, n, ms,
               n / ms * 1e-3, (double)n * POLY2B_FP64_INSTR / ms * 1e-9, (double)n * 30601.0 / ms * 1e-9);
    }
    double ho; CK(cudaMemcpy(&ho, dout, 8, cudaMemcpyDeviceToHost));
    printf("  sample out[0]=%.15g\n", ho);
    CK(cudaFree(dx)); CK(cudaFree(dout));
}

int main(int argc, char** argv)
{
    int bs = argc > 1 ? atoi(argv[1]) : 256;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s  SMs %d  clock %d kHz\n", prop.name, sms, prop.clockRate);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    {   // 1. DFMA peak
        double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 256));
        int iters = 20000;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            dfma_peak<8><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double fl = 2.0 * 8 * 8 * (double)iters * sms * 8 * 256;
            printf("dfma_peak<8 chains> blocks/SM=8: %.3f ms  %.2f TFLOP/s (%.1f DFMA/clk/SM at %d MHz nominal)\n", ms, fl / ms * 1e-9,
                   fl / 2 / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1000);
        }
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            dfma_peak<4><<<sms, 256>>>(out, iters, 1.0000001, 1e-9);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double fl = 2.0 * 4 * 8 * (double)iters * sms * 256;
            printf("dfma_peak<4 chains> 1 block(256)/SM: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
        CK(cudaFree(out));
    }
    // 2. polynomial throughput
    std::vector<double> c3(1163), c2(1153);
    for (size_t i = 0; i < c3.size(); ++i) c3[i] = 1e-3 * ((i * 7919) % 1000 - 500) / 500.0;
    for (size_t i = 0; i < c2.size(); ++i) c2[i] = 1e-3 * ((i * 7919) % 1000 - 500) / 500.0;
    CK(cudaMemcpyToSymbol(mbpol_coef3b, c3.data(), c3.size() * 8));
    CK(cudaMemcpyToSymbol(mbpol_coef2b, c2.data(), c2.size() * 8));
    run_poly<1>(sms, bs, e0, e1);
    run_poly<2>(sms, bs, e0, e1);
    run_poly<4>(sms, bs, e0, e1);
    run_poly<8>(sms, bs, e0, e1);
    return 0;
}
