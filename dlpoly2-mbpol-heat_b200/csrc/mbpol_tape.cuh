// mbpol_tape.cuh -- the "coefficient tape" polynomial engine (see tools/gen_tape.py for the format and the why).
//
// The polynomial is a list of independent ROWS sorted by (type = leaf-variable set, u, w, v):
//     P = x_u x_w x_v,  S = own + sum_c a_c x_c,  E += S P,  g_c += a_c P,  g_u += S x_w x_v,  g_w += S x_u x_v,  g_v += S x_u x_w.
// One thread owns one cluster; all 256 threads of the CTA walk the SAME rows in lock step.  Engine 1 (this part of the
// file; default for the 2B polynomial): the coefficient tape is staged global -> shared once per CTA with cp.async (double
// buffered 16 KB chunks) and read with broadcast LDS.128; one small counted loop per (chunk, type, u) segment with x_u and
// the segment's g_u sum in registers, static register names for x[], g[]; the (w,v) word of a row comes from constant
// memory; dynamic-index reads of x and updates of g go through per-thread columns in shared memory (32-bit shared
// addresses).  Engine 3 (mbpol_fat.cuh): the same machinery with all rows of one (w,v) prefix merged into a fat row.
// (Two earlier engines -- straight-line PTX parts and a constant-bank tape -- were 1.6-2x slower and left the product
// after round 1; experiments/README.md sections 1, 2 and 4 keep what was learnt from them.)
#pragma once
#include <cuda_runtime.h>

// threads per CTA = stride of the per-thread shared-memory columns.  The 2B and the 3B kernels may use different sizes
// (MBPOL_BS2 / MBPOL_BS3, csrc/Makefile): every translation unit instantiates kernels of ONE family and defines TAPE_BS
// accordingly before including this file (mbpol_inst.cu); the host code launches with MBPOL_BS2 / MBPOL_BS3.
#ifndef MBPOL_BS2
#define MBPOL_BS2 256
#endif
#ifndef MBPOL_BS3
#define MBPOL_BS3 256
#endif
#ifndef TAPE_BS
#define TAPE_BS MBPOL_BS3
#endif
#ifndef TAPE_STAGGER_NS
#define TAPE_STAGGER_NS 1000      /* k_dimers 2.37 -> 2.33 ms */
#endif

__device__ __forceinline__ double tape_lds(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void tape_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
#define TAPE_LDS_V2(vx, vy, addr, imm) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(vx), "=d"(vy) : "r"(addr), "n"(imm))
#define TAPE_LD2(vx, vy, addr, imm) TAPE_LDS_V2(vx, vy, addr, imm)      /* engine 1: tape pairs come from shared memory */

// ---- per-row prologue / epilogue shared by every row type (used by the generated TAPE*_SEGMENT_SWITCH) ----
#define TAPE_COL (TAPE_BS * 8)       /* bytes between two variables' columns */
#ifndef TAPE_UNROLL
#define TAPE_UNROLL 2
#endif
#define TAPE_PRAGMA_(x) _Pragma(#x)
#define TAPE_PRAGMA(x) TAPE_PRAGMA_(x)
#define TAPE_ROW_LOOP(nrows) TAPE_PRAGMA(unroll TAPE_UNROLL) for (unsigned r_ = 0; r_ < (nrows); ++r_)
// Segment = run of rows with the same (type, u): x_u is a register for the whole segment, sum_rows S x_w x_v is
// collected in Gu (E += x_u Gu and g_u += Gu once per segment).  Only w and v are dynamic per row: two x reads and two
// gradient read-modify-writes in the per-thread shared-memory columns.
#define TAPE_SEG_BEGIN(seg)                                                                              \
    const unsigned type = (seg) & 255u, nrows = (seg) >> 16;                                             \
    const unsigned au_ = xs_t + (((seg) >> 8) & 255u) * TAPE_COL;                                        \
    const double xu = tape_lds(au_);                                                                     \
    double Gu = 0.0;
#define TAPE_SEG_END                                                                                     \
    e = fma(Gu, xu, e);                                                                                  \
    tape_sts(au_ + gs_delta, tape_lds(au_ + gs_delta) + Gu);
// The row header -- (w,v) word, the two x loads and the first TAPE_NPF tape pairs -- is software pipelined: the header
// of the NEXT row (nw_/nv_, nxw/nxv, nq*; loaded by TAPE_ROW_FETCH) is issued before the FMAs of this row.
#define TAPE_ROW_HEADER(ROWIDX)                                                                          \
        const unsigned vw_ = rowvars[ROWIDX];                                                            \
        nw_ = xs_t + (vw_ & 255u) * TAPE_COL;                                                            \
        nv_ = xs_t + ((vw_ >> 8) & 255u) * TAPE_COL;                                                     \
        nxv = tape_lds(nv_);                                                                             \
        nxw = tape_lds(nw_);
#define TAPE_ROW_FETCH(ROWIDX, CP)                                                                       \
    {                                                                                                    \
        TAPE_ROW_HEADER(ROWIDX)                                                                          \
        TAPE_PF_LOADS(CP)                                                                                \
    }
// rows are evaluated in pairs (suffix S = a, b; K = 0, 1 = position in the pair): see tools/gen_tape.py
#define TAPE_PAIR_LOOP(nrows) for (; r_ + 2 <= (nrows); r_ += 2)
#define TAPE_ROW_PRE(ROWBYTES, S, K)                                                                     \
    const unsigned aw_##S = nw_, av_##S = nv_;                                                           \
    const double xw##S = nxw, xv##S = nxv;                                                               \
    TAPE_PF_TAKE(S)                                                                                      \
    const double p2##S = xv##S * xw##S, P##S = p2##S * xu;                                               \
    double S0##S = q0x##S, S1##S = 0.0, S2##S = 0.0, S3##S = 0.0;                                        \
    TAPE_ROW_FETCH(min(row + (K) + 1, row_last), cp + ((K) + 1) * (ROWBYTES))
#define TAPE_ROW_TAIL(S)                                                                                 \
    Gu = fma(S0##S, p2##S, Gu);                                                                          \
    {                                                                                                    \
        const double R_ = S0##S * xu;                                                                    \
        tape_sts(av_##S + gs_delta, fma(R_, xw##S, tape_lds(av_##S + gs_delta)));                        \
        tape_sts(aw_##S + gs_delta, fma(R_, xv##S, tape_lds(aw_##S + gs_delta)));                        \
    }
#define TAPE_ROW_POST(S)                                                                                 \
    S0##S = TAPE_SSUM(S);                                                                                \
    TAPE_ROW_TAIL(S)
#define TAPE_ROWS_ADVANCE(BYTES, N)                                                                      \
    cp += (BYTES);                                                                                       \
    row += (N);

#include "gen/tape_2b.h"
#include "gen/tape_3b.h"
#if TAPE_NSUM == 2      /* sum of the partial Horner accumulators in use (x + 0.0 is not folded away in FP64) */
#define TAPE_SSUM(S) S0##S + S1##S
#else
#define TAPE_SSUM(S) (S0##S + S1##S) + (S2##S + S3##S)
#endif

extern __constant__ unsigned c_tape_rowvars2b[TAPE2B_NROWS];
extern __constant__ unsigned c_tape_rowvars3b[TAPE3B_NROWS];
extern __constant__ unsigned c_tape_segs2b[TAPE2B_NSEGS];
extern __constant__ unsigned c_tape_segs3b[TAPE3B_NSEGS];
extern __constant__ int c_tape_chunks2b[TAPE2B_NCHUNKS + 1][3];
extern __constant__ int c_tape_chunks3b[TAPE3B_NCHUNKS + 1][3];

constexpr size_t tape_smem_bytes(int nv, int bs, int chunk_doubles) { return (size_t)16 * chunk_doubles + (size_t)16 * (nv + 1) * bs; }

template <int NV>
struct TapeSmem {   // dynamic shared memory layout
    double buf[2][TAPE_CHUNK_DOUBLES];   // staged tape chunks (16-byte aligned); prefetches may run <= 64 B past a chunk,
    double xs[NV + 1][TAPE_BS];          // i.e. into buf[1] / xs: harmless reads.  x, per-thread columns; row NV = unit variable
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
    constexpr int nchunks = (WHICH == 3) ? TAPE3B_NCHUNKS : TAPE2B_NCHUNKS;
    const unsigned* rowvars = (WHICH == 3) ? c_tape_rowvars3b : c_tape_rowvars2b;
    const unsigned* segs = (WHICH == 3) ? c_tape_segs3b : c_tape_segs2b;
    const int(*chunks)[3] = (WHICH == 3) ? c_tape_chunks3b : c_tape_chunks2b;
    const unsigned tid = threadIdx.x;
    const unsigned buf0 = (unsigned)__cvta_generic_to_shared(&sm.buf[0][0]);
    const unsigned xs_t = (unsigned)__cvta_generic_to_shared(&sm.xs[0][tid]);
    const unsigned gs_delta = (unsigned)sizeof(sm.xs);                       // gs column = xs column + gs_delta
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        tape_sts(xs_t + c * TAPE_COL, x[c]);
        tape_sts(xs_t + gs_delta + c * TAPE_COL, 0.0);
        g[c] = 0.0;
    }
    tape_sts(xs_t + NV * TAPE_COL, 1.0);
    tape_sts(xs_t + gs_delta + NV * TAPE_COL, 0.0);
    auto stage = [&](int k) {
        const int t0 = chunks[k][2], n = chunks[k + 1][2] - t0;      // both even: 16-byte granules
        const unsigned dst = buf0 + (k & 1) * (TAPE_CHUNK_DOUBLES * 8);
        for (int i = 2 * tid; i < n; i += 2 * TAPE_BS) tape_cp_async16(dst + 8 * i, tape + t0 + i);
        tape_cp_commit();
    };
    stage(0);
    double e = 0.0;
    for (int k = 0; k < nchunks; ++k) {
        if (k + 1 < nchunks) {
            stage(k + 1);
            tape_cp_wait<1>();
        } else {
            tape_cp_wait<0>();
        }
        __syncthreads();                                   // chunk k visible to every thread
#if TAPE_STAGGER_NS > 0
        if (threadIdx.x >= TAPE_BS / 2) __nanosleep(TAPE_STAGGER_NS);      // de-phase the two warps of a sub-partition (see mbpol_sym.cuh)
#endif
        unsigned cp = buf0 + (k & 1) * (TAPE_CHUNK_DOUBLES * 8);
        unsigned row = chunks[k][1];
        const unsigned row_last = chunks[k + 1][1] - 1;    // last row of this chunk (prefetch clamp)
        const int sg1 = chunks[k + 1][0];
        unsigned nw_, nv_;
        double nxw, nxv;
        TAPE_PF_DECL                                       // the TAPE_NPF prefetched tape pairs of the next row
        TAPE_ROW_FETCH(row, cp)
        for (int sg = chunks[k][0]; sg < sg1; ++sg) {
            TAPE_SEG_BEGIN(segs[sg])
            if (WHICH == 3) {
                TAPE3B_SEGMENT_SWITCH(type, nrows)
            } else {
                TAPE2B_SEGMENT_SWITCH(type, nrows)
            }
            TAPE_SEG_END
        }
        __syncthreads();                                   // everyone is done with this buffer before it is refilled
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) g[c] += tape_lds(xs_t + gs_delta + c * TAPE_COL);
    E = e;
}
