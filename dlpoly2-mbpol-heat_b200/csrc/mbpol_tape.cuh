// mbpol_tape.cuh -- the "coefficient tape" polynomial engine (see tools/gen_tape.py for the format and the why).
//
// One thread owns one cluster; all 256 threads of the CTA walk the SAME op list in lock step, so the op words are
// warp-uniform (constant memory), the coefficient tape is staged global -> shared once per CTA with cp.async
// (double buffered, 8 KB chunks) and read with broadcast LDS.128, and the only per-thread state is
// x[NV], g[NV] (registers, static names) + a handful of trie-walk scalars.  Dynamic-index accesses to x / g
// (once per op) go through per-thread columns in shared memory, addressed with 32-bit shared addresses.
#pragma once
#include <cuda_runtime.h>

#define TAPE_BS 256   /* threads per CTA: fixed (shared-memory column stride) */

__device__ __forceinline__ double tape_lds(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void tape_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
#define TAPE_LDS_V2(vx, vy, addr, imm) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(vx), "=d"(vy) : "r"(addr), "n"(imm))

#include "gen/tape_2b.h"
#include "gen/tape_3b.h"

extern __constant__ unsigned c_tape_ops2b[TAPE2B_NOPS];
extern __constant__ unsigned c_tape_ops3b[TAPE3B_NOPS];
extern __constant__ int c_tape_chunks2b[TAPE2B_NCHUNKS + 1][2];
extern __constant__ int c_tape_chunks3b[TAPE3B_NCHUNKS + 1][2];

template <int NV>
struct TapeSmem {   // dynamic shared memory layout
    double buf[2][TAPE_CHUNK_DOUBLES];   // staged tape chunks (16-byte aligned)
    double xs[NV + 1][TAPE_BS];          // x, per-thread columns (dynamic-index reads); row NV = the unit variable
    double gs[NV + 1][TAPE_BS];          // gradient contributions with a dynamic index; row NV = dump slot
};

__device__ __forceinline__ void tape_cp_async16(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tape_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tape_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// x -> E = poly(x), g = dE/dx.  Must be called by ALL threads of a TAPE_BS-wide CTA (it contains barriers).
template <int WHICH, int NV>
__device__ __forceinline__ void tape_eval(const double* __restrict__ tape, double (&x)[NV], double (&g)[NV], double& E, TapeSmem<NV>& sm)
{
    constexpr int NTYPES = (WHICH == 3) ? TAPE3B_NTYPES : TAPE2B_NTYPES;
    constexpr int nchunks = (WHICH == 3) ? TAPE3B_NCHUNKS : TAPE2B_NCHUNKS;
    const unsigned* ops = (WHICH == 3) ? c_tape_ops3b : c_tape_ops2b;
    const int(*chunks)[2] = (WHICH == 3) ? c_tape_chunks3b : c_tape_chunks2b;
    const unsigned tid = threadIdx.x;
    const unsigned buf0 = (unsigned)__cvta_generic_to_shared(&sm.buf[0][0]);
    const unsigned xs_t = (unsigned)__cvta_generic_to_shared(&sm.xs[0][tid]);   // + var * (TAPE_BS*8)
    const unsigned gs_t = (unsigned)__cvta_generic_to_shared(&sm.gs[0][tid]);
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        tape_sts(xs_t + c * (TAPE_BS * 8), x[c]);
        tape_sts(gs_t + c * (TAPE_BS * 8), 0.0);
        g[c] = 0.0;
    }
    tape_sts(xs_t + NV * (TAPE_BS * 8), 1.0);
    tape_sts(gs_t + NV * (TAPE_BS * 8), 0.0);
    auto stage = [&](int k) {
        const int t0 = chunks[k][1], n = chunks[k + 1][1] - t0;      // both even: 16-byte granules
        const unsigned dst = buf0 + (k & 1) * (TAPE_CHUNK_DOUBLES * 8);
        for (int i = 2 * tid; i < n; i += 2 * TAPE_BS) tape_cp_async16(dst + 8 * i, tape + t0 + i);
        tape_cp_commit();
    };
    stage(0);
    double e = 0.0, s1 = 0.0, s2 = 0.0, xv = 0.0, xw = 0.0, p2 = 0.0;
    for (int k = 0; k < nchunks; ++k) {
        if (k + 1 < nchunks) {
            stage(k + 1);
            tape_cp_wait<1>();
        } else {
            tape_cp_wait<0>();
        }
        __syncthreads();                                   // chunk k visible to every thread
        unsigned cp = buf0 + (k & 1) * (TAPE_CHUNK_DOUBLES * 8);
        const int o1 = chunks[k + 1][0];
        for (int oi = chunks[k][0]; oi < o1; ++oi) {
            const unsigned op = ops[oi];
            const unsigned h = op & 255u, voff = (op >> 20) * (TAPE_BS * 8);
            if (h < NTYPES) {                              // NODE3: one row of leaves under the product P = p2 * x[u]
                const double xu = tape_lds(xs_t + voff);
                double q0x, q0y;
                TAPE_LDS_V2(q0x, q0y, cp, 0);
                const double P = p2 * xu;
                double S0 = q0x, S1 = 0.0;
                if (WHICH == 3) {
                    TAPE3B_ROW_SWITCH(h, cp, S0, S1, P, q0x, q0y)
                } else {
                    TAPE2B_ROW_SWITCH(h, cp, S0, S1, P, q0x, q0y)
                }
                S0 += S1;
                s2 = fma(S0, xu, s2);
                tape_sts(gs_t + voff, fma(S0, p2, tape_lds(gs_t + voff)));
            } else if (h == NTYPES + 1) {                  // ENTER2
                xw = tape_lds(xs_t + voff);
                p2 = xv * xw;
                s2 = tape_lds(cp);
            } else if (h == NTYPES + 2) {                  // EXIT2
                s1 = fma(s2, xw, s1);
                tape_sts(gs_t + voff, fma(s2, xv, tape_lds(gs_t + voff)));
            } else if (h == NTYPES + 0) {                  // ENTER1
                xv = tape_lds(xs_t + voff);
                s1 = tape_lds(cp);
            } else if (h == NTYPES + 3) {                  // EXIT1
                e = fma(s1, xv, e);
                tape_sts(gs_t + voff, tape_lds(gs_t + voff) + s1);
            } else {                                       // LIN
                const double a = tape_lds(cp);
                e = fma(a, tape_lds(xs_t + voff), e);
                tape_sts(gs_t + voff, tape_lds(gs_t + voff) + a);
            }
            cp += (op >> 8) & 0xfffu;
        }
        __syncthreads();                                   // everyone is done with this buffer before it is refilled
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) g[c] += tape_lds(gs_t + c * (TAPE_BS * 8));
    E = e;
}
