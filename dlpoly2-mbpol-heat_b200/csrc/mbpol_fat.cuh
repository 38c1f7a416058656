// mbpol_fat.cuh -- engine 3 ("fat"): the coefficient-tape engine with FAT rows (tools/gen_fat.py has the why and the format).
//
// A fat row = all rows of the row engine that share the (w, v) prefix: p2 = x_w x_v is dynamic (one row word, two
// shared-memory reads, two gradient read-modify-writes per fat row), its children (u_k, leaf set) are static code working
// on registers only: P = p2 x[u_k], S_k = own + sum a_c x[c], g[c] += a_c P, g[u_k] += S_k p2, T += S_k x[u_k];
// E += T p2, g_w += T x_v, g_v += T x_w.  Same staging as engine 1 (cp.async double-buffered chunks, broadcast LDS.128,
// per-thread shared-memory columns for the dynamic x / g accesses); one loop over the fat rows of a chunk, the row's shape
// (its static body) is selected by an if-chain inside the loop.
#pragma once
#include "mbpol_tape.cuh"
#include "gen/fat_2b.h"
#include "gen/fat_3b.h"

static_assert(FAT_CHUNK_DOUBLES == TAPE_CHUNK_DOUBLES, "engines 1 and 3 share the staging buffers");
extern __constant__ unsigned c_fat_rowvars2b[FAT2B_NROWS + 1];
extern __constant__ unsigned c_fat_rowvars3b[FAT3B_NROWS + 1];
extern __constant__ unsigned c_fat_segs2b[FAT2B_NSEGS];
extern __constant__ unsigned c_fat_segs3b[FAT3B_NSEGS];
extern __constant__ int c_fat_chunks2b[FAT2B_NCHUNKS + 1][3];
extern __constant__ int c_fat_chunks3b[FAT3B_NCHUNKS + 1][3];

// header of the NEXT fat row (row word = w | v << 8 | shape << 16, then x_w, x_v) is fetched while this one is computed; the
// table has one spare entry
#define FAT_ROW_FETCH(ROWIDX)                                                                            \
    {                                                                                                    \
        const unsigned vw_ = rowvars[ROWIDX];                                                            \
        nshape = vw_ >> 16;                                                                              \
        nw_ = xs_t + (vw_ & 255u) * TAPE_COL;                                                            \
        nv_ = xs_t + ((vw_ >> 8) & 255u) * TAPE_COL;                                                     \
        nxw = tape_lds(nw_);                                                                             \
        nxv = tape_lds(nv_);                                                                             \
    }
#define FAT_ROW_PRE                                                                                      \
    const unsigned aw_ = nw_, av_ = nv_, shape = nshape;                                                 \
    const double xw = nxw, xv = nxv;                                                                     \
    const double p2 = xv * xw;                                                                           \
    double Tw = 0.0;                                                                                     \
    FAT_ROW_FETCH(row + 1)
#define FAT_ROW_POST                                                                                     \
    e = fma(Tw, p2, e);                                                                                  \
    tape_sts(aw_ + gs_delta, fma(Tw, xv, tape_lds(aw_ + gs_delta)));                                     \
    tape_sts(av_ + gs_delta, fma(Tw, xw, tape_lds(av_ + gs_delta)));

// x -> E = poly(x), g = dE/dx.  Must be called by ALL threads of a TAPE_BS-wide CTA (it contains barriers).
template <int WHICH, int NV>
__device__ __forceinline__ void fat_eval(const double* __restrict__ tape, double (&x)[NV], double (&g)[NV], double& E, TapeSmem<NV>& sm)
{
    static_assert(NV == (WHICH == 3 ? FAT3B_NV : FAT2B_NV), "polynomial / variable count mismatch");
    constexpr int nchunks = (WHICH == 3) ? FAT3B_NCHUNKS : FAT2B_NCHUNKS;
    const unsigned* rowvars = (WHICH == 3) ? c_fat_rowvars3b : c_fat_rowvars2b;
    const int(*chunks)[3] = (WHICH == 3) ? c_fat_chunks3b : c_fat_chunks2b;
    const unsigned tid = threadIdx.x;
    const unsigned buf0 = (unsigned)__cvta_generic_to_shared(&sm.buf[0][0]);
    const unsigned xs_t = (unsigned)__cvta_generic_to_shared(&sm.xs[0][tid]);
    const unsigned gs_delta = (unsigned)sizeof(sm.xs);
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
        unsigned cp = buf0 + (k & 1) * (TAPE_CHUNK_DOUBLES * 8);
        const unsigned row1 = chunks[k + 1][1];
        unsigned nw_, nv_, nshape;
        double nxw, nxv;
        FAT_ROW_FETCH(chunks[k][1])
#pragma unroll 1
        for (unsigned row = chunks[k][1]; row < row1; ++row) {
            FAT_ROW_PRE
            if constexpr (WHICH == 3) {
                FAT3B_ROW_SWITCH(shape)
            } else {
                FAT2B_ROW_SWITCH(shape)
            }
            FAT_ROW_POST
        }
        __syncthreads();                                   // everyone is done with this buffer before it is refilled
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) g[c] += tape_lds(xs_t + gs_delta + c * TAPE_COL);
    E = e;
}
