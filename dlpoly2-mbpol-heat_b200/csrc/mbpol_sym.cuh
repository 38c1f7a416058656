// mbpol_sym.cuh -- engine 4 ("sym"): fat rows batched over a subgroup K of the molecule permutations of the 3B
// polynomial (tools/gen_sym.py has the why and the format).
//
// A super row = the K-orbit of a fat row: |K| dynamic (w_j, v_j) pairs (row words, x reads and gradient read-modify-writes
// in the per-thread shared-memory columns) and ONE static body in which every coefficient of the tape is loaded once and
// feeds 2|K| DFMAs on |K| independent chains -- image j works on the statically renamed registers x[sigma_j(u)],
// x[sigma_j(c)], g[...].  Staging as in engines 1 and 3: cp.async double-buffered 16 KB chunks, broadcast LDS.128, one loop over
// the super rows of a chunk with the shape dispatch inside.  The generated header also carries the row epilogue (inside the
// shape bodies), the header prefetch macros and the pair-sum variables; this file is the loop around them.
//
// The kernel's 420 KB of straight-line bodies pass through the instruction cache once per tile, so HOW the eight warps of the CTA
// walk the rows matters as much as the instruction count (experiments/README.md section 9): all eight in lock step share every
// fetched line (hit rate 7/8) but also wait for every miss together; SYM_SPLIT_ORDER 1 (default) lets warps 4..7 walk the rows of
// each chunk in the cyclic order that starts in the middle, so that the two warps of an SM sub-partition (w and w + 4) miss at
// different moments: 15.13 -> 14.98 ms.  One start point per warp (8) thrashes the cache: 32.5 ms.
#pragma once
#include "mbpol_tape.cuh"
// acc += a * b with the operand order fixed in PTX (tools/gen_sym.py --asmfma 1)
#define SYM_FMA(acc, a, b) asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b))
#ifndef SYM_STAGGER_NS
#define SYM_STAGGER_NS 0         /* extra start delay of warps 4..7 per chunk: without SYM_SPLIT_ORDER 0 / 1000 / 3000 ns -> 15.38 / 15.27 / 15.13 ms; with it 0 is best (15.02 vs 15.31) */
#endif
#ifndef SYM_SPLIT_ORDER
#define SYM_SPLIT_ORDER 1
#endif
#ifndef MBPOL_SYM_HEADER
#define MBPOL_SYM_HEADER "gen/sym_3b.h"
#endif
#include MBPOL_SYM_HEADER

// own staging buffers: the chunk size of this engine is independent of engines 1 and 3 (2048 doubles; 1024 / 1536 / 4096 measured
// 17.4 / 15.3 / 16.3 ms against 14.8 ms: the chunk also sets how far apart the two halves of the CTA walk)
extern __constant__ unsigned c_sym_rowvars3b[SYM3B_NROWS * SYM3B_K + SYM_ROWVARS_SPARE];
extern __constant__ unsigned c_sym_segs3b[SYM3B_NSEGS];
extern __constant__ int c_sym_chunks3b[SYM3B_NCHUNKS + 1][3];

struct SymSmem {
    double buf[2][SYM_CHUNK_DOUBLES];
    double xs[37][TAPE_BS];
    double gs[37][TAPE_BS];
};

// x -> E = poly(x), g = dE/dx.  Must be called by ALL threads of a TAPE_BS-wide CTA (it contains barriers).
// XSTAGED: the x columns were already filled by the caller (sym_stage_x in mbpol_device.cuh); x[] is read back from them.
template <bool XSTAGED = false>
__device__ __forceinline__ void sym_eval3b(const double* __restrict__ tape, double (&x)[36], double (&g)[36], double& E, SymSmem& sm)
{
    constexpr int NV = 36;
    const unsigned* rowvars = c_sym_rowvars3b;
    const unsigned tid = threadIdx.x;
    const unsigned buf0 = (unsigned)__cvta_generic_to_shared(&sm.buf[0][0]);
    const unsigned xs_t = (unsigned)__cvta_generic_to_shared(&sm.xs[0][tid]);
    const unsigned gs_delta = (unsigned)sizeof(sm.xs);
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        if (XSTAGED) x[c] = tape_lds(xs_t + c * TAPE_COL);
        else tape_sts(xs_t + c * TAPE_COL, x[c]);
        tape_sts(xs_t + gs_delta + c * TAPE_COL, 0.0);
        g[c] = 0.0;
    }
    SYM_PAIR_INIT                                          // pair sums y = x_c + x_c' and their gradient accumulators (tools/gen_sym.py, PAIRS)
    tape_sts(xs_t + NV * TAPE_COL, 1.0);
    tape_sts(xs_t + gs_delta + NV * TAPE_COL, 0.0);
    auto stage = [&](int k) {
        const int t0 = c_sym_chunks3b[k][2], n = c_sym_chunks3b[k + 1][2] - t0;      // both even: 16-byte granules
        const unsigned dst = buf0 + (k & 1) * (SYM_CHUNK_DOUBLES * 8);
        for (int i = 2 * tid; i < n; i += 2 * TAPE_BS) tape_cp_async16(dst + 8 * i, tape + t0 + i);
        tape_cp_commit();
    };
    stage(0);
    double e = 0.0;
    for (int k = 0; k < SYM3B_NCHUNKS; ++k) {
        if (k + 1 < SYM3B_NCHUNKS) {
            stage(k + 1);
            tape_cp_wait<1>();
        } else {
            tape_cp_wait<0>();
        }
        __syncthreads();                                   // chunk k visible to every thread
#if SYM_STAGGER_NS > 0
        // The two warps of an SM sub-partition (w and w + 4) walk the same code: in lock step both wait for the same instruction-cache
        // miss.  Warps 4..7 start every chunk a little later: the leaders take the misses while the followers still have cached
        // code to run, and the followers find the lines in place.
        if (threadIdx.x >= TAPE_BS / 2) __nanosleep(SYM_STAGGER_NS);
#endif
        const unsigned cbase = buf0 + (k & 1) * (SYM_CHUNK_DOUBLES * 8);
        const unsigned row0 = c_sym_chunks3b[k][1], row1 = c_sym_chunks3b[k + 1][1];
        // The rows of a chunk are independent, so a warp may walk them in any cyclic order: warps 4..7 start in the middle of the
        // chunk (see the head of this file).
        unsigned row = row0;
#if SYM_SPLIT_ORDER == 1
        if (threadIdx.x >= TAPE_BS / 2) row += (row1 - row0) / 2;
#elif SYM_SPLIT_ORDER == 3
        if (threadIdx.x >= TAPE_BS / 2) row += (row1 - row0) / 3;
#elif SYM_SPLIT_ORDER == 8
        row += ((threadIdx.x >> 5) * (row1 - row0)) / (TAPE_BS / 32);
#endif
        unsigned row_next = row + 1 == row1 ? row0 : row + 1;
        unsigned cp;
        SYM_ROW_DECL
        SYM_ROW_WORDS(row)
        SYM_ROW_FETCH(row, row_next)
#pragma unroll 1
        for (unsigned it = row0; it < row1; ++it) {
            const unsigned row_next2 = row_next + 1 == row1 ? row0 : row_next + 1;
            SYM_ROW_PRE
            SYM3B_ROW_SWITCH(shape)
            SYM_ROW_POST
            row = row_next;
            row_next = row_next2;
        }
        __syncthreads();                                   // everyone is done with this buffer before it is refilled
    }
    SYM_PAIR_FINI
#pragma unroll
    for (int c = 0; c < NV; ++c) g[c] += tape_lds(xs_t + gs_delta + c * TAPE_COL);
    E = e;
}
