// tmem_probe.cu -- can tensor memory (TMEM) serve as a per-thread scratchpad next to shared memory?
// One "row" = 5 broadcast LDS.128 (10 tape slots) feeding 20 DFMAs, two dynamic-index x reads and two dynamic-index
// gradient read-modify-writes -- the shape of the tape engine's inner loop (csrc/mbpol_tape.cuh).
//   MODE 0: x / g columns in shared memory (LDS.64 / STS.64)      -- what the engine does
//   MODE 1: x / g columns in TMEM (tcgen05.ld/st 32x32b.x2, warp-uniform dynamic column)
//   MODE 2: no x / g traffic at all (lower bound)
//   MODE 3: TMEM traffic only (no tape loads, no FMAs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cuda_runtime.h>
#define BS 256
#define NV 37
__device__ __forceinline__ double lds(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double tld(unsigned ta)
{
    unsigned lo, hi;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tld2(unsigned ta, unsigned tb, double& a, double& b)   // two loads, one wait
{
    unsigned l0, h0, l1, h1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(l0), "=r"(h0) : "r"(ta));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(l1), "=r"(h1) : "r"(tb));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __hiloint2double(h0, l0);
    b = __hiloint2double(h1, l1);
}
__device__ __forceinline__ void tst(unsigned ta, double v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(ta), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void twait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Smem {
    double tape[2048];
    double xs[NV][BS];
    double gs[NV][BS];
    unsigned tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(BS, 1) k_probe(int nrows, double* out, long long* clk)
{
    extern __shared__ __align__(16) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const unsigned tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2048; i += BS) sm.tape[i] = 1.0 / (1 + i);
    for (int c = 0; c < NV; ++c) { sm.xs[c][tid] = 1.0 + 1e-3 * c + 1e-6 * tid; sm.gs[c][tid] = 0.0; }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((unsigned long long)__cvta_generic_to_shared(&sm.tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // this warp's TMEM window: lanes 32*(warp%4).., columns 256*(warp/4)..; x at column 2c, g at column 74 + 2c
    const unsigned tbase = sm.tmem_base + ((32u * (warp & 3u)) << 16) + 256u * (warp >> 2);
    for (int c = 0; c < NV; ++c) { tst(tbase + 2 * c, 1.0 + 1e-3 * c + 1e-6 * tid); tst(tbase + 74 + 2 * c, 0.0); }
    twait_st();
    const unsigned tape0 = (unsigned)__cvta_generic_to_shared(&sm.tape[0]);
    const unsigned xs_t = (unsigned)__cvta_generic_to_shared(&sm.xs[0][tid]);
    const unsigned gs_delta = sizeof(sm.xs);
    double g[10], x[10], S0 = 0, S1 = 0, e = 0;
    for (int c = 0; c < 10; ++c) { g[c] = 0; x[c] = 1.0 + 0.01 * c + 1e-7 * tid; }
    __syncthreads();
    const long long t0 = clock64();
    unsigned cp = tape0;
    for (int r = 0; r < nrows; ++r) {
        const unsigned w = (r * 7u) % NV, v = (r * 11u + 3u) % NV;
        double xw = 1.0, xv = 1.0;
        if (MODE == 0) { xw = lds(xs_t + w * (BS * 8)); xv = lds(xs_t + v * (BS * 8)); }
        if (MODE == 1 || MODE == 3) tld2(tbase + 2 * w, tbase + 2 * v, xw, xv);
        const double P = xw * xv;
        if (MODE != 3) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                double a, b;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(a), "=d"(b) : "r"(cp), "n"(0));
                cp += 16;
                S0 = fma(a, x[2 * k], S0); g[2 * k] = fma(a, P, g[2 * k]);
                S1 = fma(b, x[2 * k + 1], S1); g[2 * k + 1] = fma(b, P, g[2 * k + 1]);
            }
            if (cp >= tape0 + 2000 * 8) cp = tape0;
        } else {
            S0 += P;
        }
        const double S = S0 + S1;
        e = fma(S, P, e);
        if (MODE == 0) {
            sts(xs_t + gs_delta + w * (BS * 8), fma(S, xv, lds(xs_t + gs_delta + w * (BS * 8))));
            sts(xs_t + gs_delta + v * (BS * 8), fma(S, xw, lds(xs_t + gs_delta + v * (BS * 8))));
        }
        if (MODE == 1 || MODE == 3) {
            double gw, gv;
            tld2(tbase + 74 + 2 * w, tbase + 74 + 2 * v, gw, gv);
            gw = fma(S, xv, gw);
            gv = (w == v) ? fma(S, xw, gw) : fma(S, xw, gv);
            tst(tbase + 74 + 2 * w, gw);
            tst(tbase + 74 + 2 * v, gv);
            twait_st();
        }
        S0 *= 1e-3; S1 *= 1e-3;
    }
    const long long t1 = clock64();
    double acc = e;
    for (int c = 0; c < 10; ++c) acc += g[c];
    if (MODE == 0) for (int c = 0; c < NV; ++c) acc += sm.gs[c][tid];
    if (MODE == 1 || MODE == 3) for (int c = 0; c < NV; ++c) acc += tld(tbase + 74 + 2 * c);
    out[blockIdx.x * BS + tid] = acc;
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(sm.tmem_base));
}

template <int MODE>
void run(int nrows, double* out, long long* clk, const char* what)
{
    cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_probe<MODE><<<148, BS, sizeof(Smem)>>>(nrows, out, clk);
    cudaEventRecord(a);
    k_probe<MODE><<<148, BS, sizeof(Smem)>>>(nrows, out, clk);
    cudaEventRecord(b);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double o[4]; cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i];
    printf("%-44s %8.3f ms  %7.1f clk/row/SM (8 warps)  chk %.6e  %s\n", what, ms, avg / 148 / nrows, o[0] + o[1], cudaGetErrorString(err));
}

int main()
{
    double* out; long long* clk;
    cudaMalloc(&out, 148 * BS * sizeof(double));
    cudaMalloc(&clk, 148 * sizeof(long long));
    const int n = 200000;
    run<2>(n, out, clk, "tape LDS.128 + DFMA only");
    run<0>(n, out, clk, "x/g columns in shared memory");
    run<1>(n, out, clk, "x/g columns in TMEM");
    run<3>(n, out, clk, "TMEM traffic only");
    return 0;
}
