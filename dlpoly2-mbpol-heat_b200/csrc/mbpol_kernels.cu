// mbpol_kernels.cu -- libmbpol_b200.so: kernels + C ABI (include/mbpol_b200.h).  sm_100a only, FP64.
//
// Pipeline of one mbpol_forces-equivalent step (reference: mbpol.f90:98-449), all on one stream:
//   K0  k_gather_image      xxx/yyy/zzz + listttm2 -> imaged monomers, SoA [9][nw]            (mbpol.f90:184-196, 1275-1301)
//   K1a k_cell_index/k_cell_fill   counting sort of waters into cells (oxygen positions)
//   K1b k_neighbours        27-cell search with the reference's exact predicates -> per-water neighbour rows
//   K1c k_fill_dimers, k_count_trimers, k_fill_trimers (+ cub exclusive scans) -> compact (i,j) / (i,j,k) lists
//   K2  k_dimers            2B energy/gradient (+ pair matrix for HEAT_STRESS), scatter     (mbpol.f90:453-790)
//   K3  k_trimers           3B energy/gradient, scatter                                      (mbpol.f90:794-1271)
//   K5  k_flux              sum_i e_i v_i, sum_i S_i v_i, sum_i W_i v_i                       (heatcurrent.f90:464-483, 703-728)
// Scatter: FP64 red.global.add into the packed output buffer; scalars are block-reduced first.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nothing without a profiler attached
#include <nccl.h>      // types and prototypes only: the library is bound at run time (nccl_api below)

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "mbpol_b200.h"
#define MBPOL_GATHER_KERNELS
#include "mbpol_cluster.cuh"
#include "gen/mbpol_model_data.h"
#include <cstdlib>

__constant__ MbpolParams mbpol_par;
__constant__ unsigned c_tape_rowvars2b[TAPE2B_NROWS];
__constant__ unsigned c_tape_rowvars3b[TAPE3B_NROWS];
__constant__ unsigned c_tape_segs2b[TAPE2B_NSEGS];
__constant__ unsigned c_tape_segs3b[TAPE3B_NSEGS];
__constant__ int c_tape_chunks2b[TAPE2B_NCHUNKS + 1][3];
__constant__ int c_tape_chunks3b[TAPE3B_NCHUNKS + 1][3];
__constant__ unsigned c_fat_rowvars2b[FAT2B_NROWS + 1];
__constant__ unsigned c_fat_rowvars3b[FAT3B_NROWS + 1];
__constant__ unsigned c_fat_segs2b[FAT2B_NSEGS];
__constant__ unsigned c_fat_segs3b[FAT3B_NSEGS];
__constant__ int c_fat_chunks2b[FAT2B_NCHUNKS + 1][3];
__constant__ int c_fat_chunks3b[FAT3B_NCHUNKS + 1][3];
__constant__ unsigned c_v3word[36] = {      // compact prologue of k_trimers<4> (sym_stage_x): 3 * V3A[n] | 3 * V3B[n] << 8 | V3K[n] << 16
#define W3(a, b, k) (3u * a | (3u * b) << 8 | (unsigned)k << 16)
    W3(1, 2, 0), W3(4, 5, 0), W3(7, 8, 0), W3(0, 1, 1), W3(0, 2, 1), W3(3, 4, 1), W3(3, 5, 1), W3(6, 7, 1), W3(6, 8, 1),
    W3(1, 4, 2), W3(1, 5, 2), W3(1, 7, 2), W3(1, 8, 2), W3(2, 4, 2), W3(2, 5, 2), W3(2, 7, 2), W3(2, 8, 2), W3(4, 7, 2), W3(4, 8, 2), W3(5, 7, 2), W3(5, 8, 2),
    W3(0, 4, 3), W3(0, 5, 3), W3(0, 7, 3), W3(0, 8, 3), W3(3, 1, 3), W3(3, 2, 3), W3(3, 7, 3), W3(3, 8, 3), W3(6, 1, 3), W3(6, 2, 3), W3(6, 4, 3), W3(6, 5, 3),
    W3(0, 3, 4), W3(0, 6, 4), W3(3, 6, 4)
#undef W3
};
__constant__ unsigned c_sym_rowvars3b[SYM3B_NROWS * SYM3B_K + SYM_ROWVARS_SPARE];
__constant__ unsigned c_sym_segs3b[SYM3B_NSEGS];
__constant__ int c_sym_chunks3b[SYM3B_NCHUNKS + 1][3];
__constant__ int c_gather_vp[3][36];
__constant__ int c_gather_vp2[31];

using namespace mbpol_k;

// the (kernel, engine) instantiations live in mbpol_inst.cu
namespace mbpol_k {
extern template __global__ void k_dimers<1>(const long long*, const int2*, int, const double*, CellInfo, const double*, double*);
extern template __global__ void k_dimers<3>(const long long*, const int2*, int, const double*, CellInfo, const double*, double*);
extern template __global__ void k_trimers<1>(const long long*, const int*, int, const double*, CellInfo, const double*, double*);
extern template __global__ void k_trimers<3>(const long long*, const int*, int, const double*, CellInfo, const double*, double*);
extern template __global__ void k_trimers<4>(const long long*, const int*, int, const double*, CellInfo, const double*, double*);
extern template __global__ void k_eval_dimers<1>(int, const double*, double*, double*, double*, const double*);
extern template __global__ void k_eval_dimers<3>(int, const double*, double*, double*, double*, const double*);
extern template __global__ void k_eval_trimers<1>(int, const double*, double*, double*, double*, const double*);
extern template __global__ void k_eval_trimers<3>(int, const double*, double*, double*, double*, const double*);
extern template __global__ void k_eval_trimers<4>(int, const double*, double*, double*, double*, const double*);
}

namespace {

// ---------------------------------------------------------------------------------------------------------
// K0: gather + image_monomers.  mono[c*nw + m], c = 0..8 (O xyz, H1 xyz, H2 xyz)
// ---------------------------------------------------------------------------------------------------------
// monoA[m][12]: the same nine coordinates water-major plus the RAW coordinate sum of the water's three atoms (o, o+1, o+2 as
// do_dimer / do_trimer address them) -- 96 bytes = three sectors per water for the gather kernels, which visit waters at random
__global__ void k_gather_image(int nw, const int* __restrict__ list, const double* __restrict__ x, const double* __restrict__ y,
                               const double* __restrict__ z, CellInfo C, double* __restrict__ mono, double* __restrict__ monoA)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nw) return;
    double w[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int i = list[3 * m + a] - 1;
        w[3 * a] = x[i]; w[3 * a + 1] = y[i]; w[3 * a + 2] = z[i];
    }
#pragma unroll
    for (int h = 1; h < 3; ++h) {
        double dx = __dsub_rn(w[3 * h], w[0]), dy = __dsub_rn(w[3 * h + 1], w[1]), dz = __dsub_rn(w[3 * h + 2], w[2]);
        image_vec(C, dx, dy, dz);
        w[3 * h] = __dadd_rn(w[0], dx); w[3 * h + 1] = __dadd_rn(w[1], dy); w[3 * h + 2] = __dadd_rn(w[2], dz);
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) mono[(size_t)c * nw + m] = w[c];
    const int o = list[3 * m] - 1;
    double2* A = reinterpret_cast<double2*>(monoA + (size_t)m * 12);
    A[0] = make_double2(w[0], w[1]); A[1] = make_double2(w[2], w[3]); A[2] = make_double2(w[4], w[5]); A[3] = make_double2(w[6], w[7]);
    A[4] = make_double2(w[8], (x[o] + x[o + 1]) + x[o + 2]);
    A[5] = make_double2((y[o] + y[o + 1]) + y[o + 2], (z[o] + z[o + 1]) + z[o + 2]);
}

// ---------------------------------------------------------------------------------------------------------
// K1a: cells.  Fractional coordinates s = rcell * r wrapped to [0,1); cell edge >= r2f in every direction, so all
// partners within either cutoff sit in the 27 surrounding cells.  The grid is only a candidate filter: membership is
// decided by the exact predicates in k_neighbours.
// ---------------------------------------------------------------------------------------------------------
struct Grid { int nc[3]; double rc[9]; int nwb, ncb; };   // nwb waters and ncb cells per bead: beads never see each other

__device__ __forceinline__ int cell_of(const Grid& G, double x, double y, double z, int* cxyz)
{
    double s[3] = {G.rc[0] * x + G.rc[3] * y + G.rc[6] * z, G.rc[1] * x + G.rc[4] * y + G.rc[7] * z, G.rc[2] * x + G.rc[5] * y + G.rc[8] * z};
    int c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s[k] -= floor(s[k]);
        c[k] = min(max((int)(s[k] * G.nc[k]), 0), G.nc[k] - 1);
        if (cxyz) cxyz[k] = c[k];
    }
    return (c[2] * G.nc[1] + c[1]) * G.nc[0] + c[0];
}

__global__ void k_cell_index(int nw, const double* __restrict__ mono, Grid G, int* __restrict__ cell_id, int* __restrict__ cell_count)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nw) return;
    const int c = (m / G.nwb) * G.ncb + cell_of(G, mono[m], mono[(size_t)nw + m], mono[2 * (size_t)nw + m], nullptr);
    cell_id[m] = c;
    atomicAdd(&cell_count[c], 1);
}

__global__ void k_cell_fill(int nw, const int* __restrict__ cell_id, const int* __restrict__ cell_start, int* __restrict__ cell_fill,
                            int* __restrict__ sorted)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nw) return;
    const int c = cell_id[m];
    sorted[cell_start[c] + atomicAdd(&cell_fill[c], 1)] = m;
}

// ---------------------------------------------------------------------------------------------------------
// K1b: neighbour rows of the owned waters.
//   nb2[i][*]: j > i with  r^2 <= r2f^2     (reject: r^2 .gt. r2f**2, mbpol.f90:502)
//   nb3[i][*]: j != i with sqrt(r^2) <= r3f (reject: r .gt. r3f,      mbpol.f90:865-869)
// d = O_min(i,j) - O_max(i,j), then the reference's minimum image.
// ---------------------------------------------------------------------------------------------------------
// The cell fill order comes from atomics; sorting every cell's (short) member list makes the neighbour rows -- and
// therefore the lists and the per-rank shards cut out of them -- identical on every GPU and from call to call.
__global__ void k_sort_cells(int nw, const int* __restrict__ cell_id, const int* __restrict__ cell_start, const int* __restrict__ unsorted,
                             int* __restrict__ sorted)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per water: its rank = members of its cell with a smaller index
    if (m >= nw) return;
    const int c = cell_id[m];
    const int p0 = cell_start[c], p1 = cell_start[c + 1];
    int rank = 0;
    for (int p = p0; p < p1; ++p) rank += unsorted[p] < m;
    sorted[p0 + rank] = m;
}

// One WARP per water of this rank's range [w0, w0 + nown): the lanes share the scan of the 27 surrounding cells; accepted
// partners are appended with a ballot/prefix so the row order is (cell order, member order) -- deterministic.  Rows are
// indexed by the LOCAL water number i - w0: every rank builds only the rows of the waters it owns (mbpol.f90:218-226, 311-314
// split the pair / triple loops over the ranks of a bead in the same spirit).
__global__ void k_neighbours(int w0, int nown, int nw, const double* __restrict__ mono, CellInfo C, Grid G, const int* __restrict__ cell_start,
                             const int* __restrict__ sorted, int* __restrict__ nb2, int* __restrict__ n2, int* __restrict__ nb3,
                             int* __restrict__ n3, int* __restrict__ overflow)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    if (t >= nown) return;
    const int i = w0 + t;
    const double xi = mono[i], yi = mono[(size_t)nw + i], zi = mono[2 * (size_t)nw + i];
    int ci[3];
    cell_of(G, xi, yi, zi, ci);
    const double r2f2 = __dmul_rn(mbpol_par.r2f, mbpol_par.r2f);
    int c2 = 0, c3 = 0;
    int lo[3], cnt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // unique neighbour cells even when the grid is narrower than 3
        cnt[k] = min(G.nc[k], 3);
        lo[k] = (G.nc[k] >= 3) ? ci[k] - 1 : 0;
    }
    for (int dz = 0; dz < cnt[2]; ++dz)
        for (int dy = 0; dy < cnt[1]; ++dy)
            for (int dx = 0; dx < cnt[0]; ++dx) {
                const int cx = (lo[0] + dx + G.nc[0]) % G.nc[0], cy = (lo[1] + dy + G.nc[1]) % G.nc[1], cz = (lo[2] + dz + G.nc[2]) % G.nc[2];
                const int c = (i / G.nwb) * G.ncb + (cz * G.nc[1] + cy) * G.nc[0] + cx;
                const int p1 = cell_start[c + 1];
                for (int p0 = cell_start[c]; p0 < p1; p0 += 32) {
                    const int p = p0 + (int)lane;
                    bool in2 = false, in3 = false;
                    int j = -1;
                    if (p < p1) {
                        j = sorted[p];
                        if (j != i) {
                            const double xj = mono[j], yj = mono[(size_t)nw + j], zj = mono[2 * (size_t)nw + j];
                            double ddx, ddy, ddz;
                            if (i < j) { ddx = __dsub_rn(xi, xj); ddy = __dsub_rn(yi, yj); ddz = __dsub_rn(zi, zj); }
                            else       { ddx = __dsub_rn(xj, xi); ddy = __dsub_rn(yj, yi); ddz = __dsub_rn(zj, zi); }
                            image_vec(C, ddx, ddy, ddz);
                            const double rsq = norm2_exact(ddx, ddy, ddz);
                            in2 = j > i && !(rsq > r2f2);
                            in3 = !(__dsqrt_rn(rsq) > mbpol_par.r3f);
                        }
                    }
                    const unsigned b2 = __ballot_sync(0xffffffffu, in2), b3 = __ballot_sync(0xffffffffu, in3);
                    const int q2 = c2 + __popc(b2 & lt), q3 = c3 + __popc(b3 & lt);
                    if (in2 && q2 < NB2_CAP) nb2[(size_t)t * NB2_CAP + q2] = j;
                    if (in3 && q3 < NB3_CAP) nb3[(size_t)t * NB3_CAP + q3] = j;
                    c2 += __popc(b2);
                    c3 += __popc(b3);
                }
            }
    if (lane == 0) {
        if (c2 > NB2_CAP || c3 > NB3_CAP) atomicExch(overflow, 1);
        n2[t] = min(c2, NB2_CAP);
        n3[t] = min(c3, NB3_CAP);
    }
}

__global__ void k_fill_dimers(int w0, int nown, const int* __restrict__ nb2, const int* __restrict__ n2, const long long* __restrict__ off,
                              int2* __restrict__ dimers, long long cap, int* __restrict__ overflow)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nown) return;
    const long long o = off[t];
    if (o + n2[t] > cap) { atomicExch(overflow, 2); return; }
    for (int q = 0; q < n2[t]; ++q) dimers[o + q] = make_int2(w0 + t, nb2[(size_t)t * NB2_CAP + q]);
}

// the 3B list predicate for one pair, with the reference's arithmetic (d = O_min - O_max, minimum image, r .gt. r3f rejects:
// mbpol.f90:843-869) -- the same bits k_neighbours computes, so a rank needs no neighbour rows of waters it does not own
__device__ __forceinline__ bool within_r3f(int nw, const double* __restrict__ mono, const CellInfo& C, int a, int b)
{
    const int i = min(a, b), j = max(a, b);
    double dx = __dsub_rn(mono[i], mono[j]), dy = __dsub_rn(mono[(size_t)nw + i], mono[(size_t)nw + j]),
           dz = __dsub_rn(mono[2 * (size_t)nw + i], mono[2 * (size_t)nw + j]);
    image_vec(C, dx, dy, dz);
    return !(__dsqrt_rn(norm2_exact(dx, dy, dz)) > mbpol_par.r3f);
}

// A triple is accepted iff at least two of its three O-O distances are <= r3f (mbpol.f90:865-869).  Enumerate it from
// a centre c that is within r3f of both others; when all three pairs are close every water is such a centre and the
// triple is emitted from the smallest index only.  One warp per owned centre.  The count pass also leaves the centre's
// load-balancing weight (4 x trimers + dimers) for the next step's shard boundaries.
template <bool FILL>
__global__ void k_trimers_enum(int w0, int nown, int nw, const double* __restrict__ mono, CellInfo C, const int* __restrict__ nb3,
                               const int* __restrict__ n3, const int* __restrict__ n2, long long* __restrict__ cnt, double* __restrict__ weight,
                               const long long* __restrict__ off, int* __restrict__ trimers, long long cap, int* __restrict__ overflow)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // one warp per centre
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    if (t >= nown) return;
    const int c = w0 + t;
    const int* row = nb3 + (size_t)t * NB3_CAP;
    const int n = n3[t];
    const int npairs = n * (n - 1) / 2;
    long long k = 0;
    const long long base = FILL ? off[t] : 0;
    if (FILL && base + (off[t + 1] - base) > cap) {
        if (lane == 0) atomicExch(overflow, 3);
        return;
    }
    for (int e0 = 0; e0 < npairs; e0 += 32) {
        const int e = e0 + (int)lane;
        bool emit = false;
        int a = 0, b = 0;
        if (e < npairs) {
            // pair index e -> (p, q), p < q, in the order (0,1), (0,2), (1,2), (0,3), ...: q = row of the triangle
            int q = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
            while (q * (q - 1) / 2 > e) --q;
            while ((q + 1) * q / 2 <= e) ++q;
            const int pp = e - q * (q - 1) / 2;
            a = row[pp];
            b = row[q];
            const bool closed = within_r3f(nw, mono, C, a, b);
            emit = !(closed && !(c < a && c < b));
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (FILL && emit) {
            int t0 = c, t1 = a, t2 = b, s_;
            if (t0 > t1) { s_ = t0; t0 = t1; t1 = s_; }
            if (t1 > t2) { s_ = t1; t1 = t2; t2 = s_; }
            if (t0 > t1) { s_ = t0; t0 = t1; t1 = s_; }
            int* o = trimers + 3 * (base + k + __popc(m & lt));
            o[0] = t0; o[1] = t1; o[2] = t2;
        }
        k += __popc(m);
    }
    if (!FILL && lane == 0) {
        cnt[t] = k;
        weight[c] = (double)(4 * k + n2[t]);
    }
}

// shard totals -> the scalar block (slots S_ND / S_NT of bead 0: they ride through the reduction, so every rank learns the
// box totals) and the device-side counts the evaluation kernels loop over
__global__ void k_publish_counts(const long long* __restrict__ off2_end, const long long* __restrict__ off3_end, long long cap2, long long cap3,
                                 long long grid2, long long grid3, long long* __restrict__ counts, double* __restrict__ scal)
{
    const long long nd = min(*off2_end, cap2), nt = min(*off3_end, cap3);
    counts[0] = nd; counts[1] = nt;
    counts[2] = *off2_end; counts[3] = *off3_end;
    scal[S_ND] += (double)nd;
    scal[S_NT] += (double)nt;
    // a list that does not fit its estimated capacity (or has no launch at all) makes EVERY rank repeat the step: the mark
    // travels through the all-reduce of the scalar block
    if (*off2_end > cap2 || *off3_end > cap3 || (*off2_end > 0 && grid2 <= 0) || (*off3_end > 0 && grid3 <= 0)) scal[S_OVF] += 1.0;
}

// ---------------------------------------------------------------------------------------------------------
// K5: heat-flux contractions over the atoms of THIS shard's partial sums (linear, so partials add up)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_flux(OutPtrs O, const double* __restrict__ vx, const double* __restrict__ vy,
                       const double* __restrict__ vz, int flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double j[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    // the O/H atoms of every bead; compute_heat_flux2 loops 1..3*mxatms/4 (heatcurrent.f90:710)
    const bool in_range = i < O.mxatms;
    const int bead = in_range ? i / O.mxb : -1;
    if (in_range && (i % O.mxb) < 3 * O.nwb) {
        const double v[3] = {vx[i], vy[i], vz[i]};
        const size_t n = O.mxb;
        if (flags & MBPOL_B200_WANT_ATOMIC_ENERGY) {
            const double e = O.aeng[i];
            j[0] = e * v[0]; j[1] = e * v[1]; j[2] = e * v[2];
        }
        if (flags & MBPOL_B200_WANT_ATOMIC_STRESS) {   // compute_heat_flux reads the upper triangle only (heatcurrent.f90:475-480)
            const double* s = O.astress + tensor_base(O, i);
            const double s11 = s[0], s12 = s[n * 3], s13 = s[n * 6], s22 = s[n * 4], s23 = s[n * 7], s33 = s[n * 8];
            j[3] = s11 * v[0] + s12 * v[1] + s13 * v[2];
            j[4] = s12 * v[0] + s22 * v[1] + s23 * v[2];
            j[5] = s13 * v[0] + s23 * v[1] + s33 * v[2];
        }
        if (flags & MBPOL_B200_WANT_FDECOMP) {
            const double* w = O.afd + tensor_base(O, i);
#pragma unroll
            for (int a = 0; a < 3; ++a) j[6 + a] = w[n * (a + 0)] * v[0] + w[n * (a + 3)] * v[1] + w[n * (a + 6)] * v[2];
        }
    }
    block_reduce_add_bead<9>(j, O.scal, S_JFLUX, bead);
}

// ---------------------------------------------------------------------------------------------------------
// FP64 roofline probe: 8 independent register-resident DFMA chains per thread (the denominator bench.py quotes)
// ---------------------------------------------------------------------------------------------------------
// out[i] += acc[i]: the caller's "+=" done on the device (fused end of mbpol_b200_forces)
__global__ void k_accumulate(size_t n, const double* __restrict__ acc, double* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] += acc[i];
}

__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =========================================================================================================
// host side
// =========================================================================================================
std::once_flag g_const_once[16];
double* g_tape2b[16] = {nullptr};     // engine 1: coefficients in row order
double* g_tape3b[16] = {nullptr};
double* g_fat3b[16] = {nullptr};      // engine 3: the coefficients in fat-row order
double* g_fat2b[16] = {nullptr};
double* g_sym3b[16] = {nullptr};      // engine 4: the coefficients of the super rows, weighted |orbit| / |K|

// polynomial engine: 1 = coefficient tape staged in shared memory (rows), 3 = the same with FAT rows, 4 = fat rows batched
// over molecule permutations (3B only).  MBPOL_B200_ENGINE=tape|fat|sym overrides the 3B default, MBPOL_B200_ENGINE_2B=tape|fat
// the 2B one.
#ifndef MBPOL_DEFAULT_ENGINE
#define MBPOL_DEFAULT_ENGINE 4
#endif
int default_engine()
{
    const char* e = std::getenv("MBPOL_B200_ENGINE");
    if (e && std::string(e) == "tape") return 1;
    if (e && std::string(e) == "fat") return 3;
    if (e && std::string(e) == "sym") return 4;
    return MBPOL_DEFAULT_ENGINE;
}
#ifndef MBPOL_DEFAULT_ENGINE_2B
#define MBPOL_DEFAULT_ENGINE_2B 1   /* measured: rows 3.6-3.8 ms, fat rows 4.05 ms for the 2B polynomial (521 fat rows of 27 shapes merge only 2 rows each) */
#endif
int default_engine_2b()
{
    const char* e = std::getenv("MBPOL_B200_ENGINE_2B");
    if (e && std::string(e) == "tape") return 1;
    if (e && std::string(e) == "fat") return 3;
    return MBPOL_DEFAULT_ENGINE_2B;
}
constexpr size_t SMEM2B = tape_smem_bytes(31, MBPOL_BS2, TAPE_CHUNK_DOUBLES), SMEM3B = tape_smem_bytes(36, MBPOL_BS3, TAPE_CHUNK_DOUBLES),
                 SMEM3S = tape_smem_bytes(36, MBPOL_BS3, SYM_CHUNK_DOUBLES);

template <class K>
cudaError_t allow_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define CUDA_TRY(h, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
            return MBPOL_B200_ECUDA;                                                                        \
        }                                                                                                   \
    } while (0)

void invert3(const double* a, double* b, double& d)   // invert.f
{
    b[0] = a[4] * a[8] - a[5] * a[7]; b[1] = a[2] * a[7] - a[1] * a[8]; b[2] = a[1] * a[5] - a[2] * a[4];
    b[3] = a[5] * a[6] - a[3] * a[8]; b[4] = a[0] * a[8] - a[2] * a[6]; b[5] = a[2] * a[3] - a[0] * a[5];
    b[6] = a[3] * a[7] - a[4] * a[6]; b[7] = a[1] * a[6] - a[0] * a[7]; b[8] = a[0] * a[4] - a[1] * a[3];
    d = a[0] * b[0] + a[3] * b[1] + a[6] * b[2];
    double r = 0.0;
    if (std::fabs(d) > 0.0) r = 1.0 / d;
    for (int i = 0; i < 9; ++i) b[i] = r * b[i];
}

int upload_constants(int device)
{
    MbpolParams P;
    const double k2[8] = {MBPOL_P2b_k_HH_intra, MBPOL_P2b_k_OH_intra, MBPOL_P2b_k_HH_coul, MBPOL_P2b_k_OH_coul,
                          MBPOL_P2b_k_OO_coul, MBPOL_P2b_k_XH_main, MBPOL_P2b_k_XO_main, MBPOL_P2b_k_XX_main};
    std::copy(k2, k2 + 8, P.k2);
    P.gamma_in = MBPOL_P2b_in_plane_gamma; P.gamma_out = MBPOL_P2b_out_of_plane_gamma;
    P.r2i = MBPOL_P2b_r2i; P.r2f = MBPOL_P2b_r2f;
    const double k3[5] = {MBPOL_P3b_kHH_intra, MBPOL_P3b_kOH_intra, MBPOL_P3b_kHH, MBPOL_P3b_kOH, MBPOL_P3b_kOO};
    const double d3[5] = {MBPOL_P3b_dHH_intra, MBPOL_P3b_dOH_intra, MBPOL_P3b_dHH, MBPOL_P3b_dOH, MBPOL_P3b_dOO};
    std::copy(k3, k3 + 5, P.k3); std::copy(d3, d3 + 5, P.d3);
    P.r3i = MBPOL_P3b_r3i; P.r3f = MBPOL_P3b_r3f;
    if (cudaMemcpyToSymbol(c_tape_rowvars2b, TAPE2B_ROWVARS_HOST, sizeof(TAPE2B_ROWVARS_HOST)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(c_tape_rowvars3b, TAPE3B_ROWVARS_HOST, sizeof(TAPE3B_ROWVARS_HOST)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(c_tape_segs2b, TAPE2B_SEGS_HOST, sizeof(TAPE2B_SEGS_HOST)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(c_tape_segs3b, TAPE3B_SEGS_HOST, sizeof(TAPE3B_SEGS_HOST)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(c_tape_chunks2b, TAPE2B_CHUNKS_HOST, sizeof(TAPE2B_CHUNKS_HOST)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(c_tape_chunks3b, TAPE3B_CHUNKS_HOST, sizeof(TAPE3B_CHUNKS_HOST)) != cudaSuccess) return 1;
    {   // coefficient tapes: one double per coefficient USE, in walk order (tools/gen_tape.py)
        std::vector<double> t2(TAPE2B_NTAPE), t3(TAPE3B_NTAPE);
        for (int i = 0; i < TAPE2B_NTAPE; ++i) t2[i] = TAPE2B_COEFIDX_HOST[i] < 0 ? 0.0 : MBPOL_COEF2B_HOST[TAPE2B_COEFIDX_HOST[i]];
        for (int i = 0; i < TAPE3B_NTAPE; ++i) t3[i] = TAPE3B_COEFIDX_HOST[i] < 0 ? 0.0 : MBPOL_COEF3B_HOST[TAPE3B_COEFIDX_HOST[i]];
        double *d2 = nullptr, *d3 = nullptr;
        if (cudaMalloc(&d2, t2.size() * 8) != cudaSuccess || cudaMalloc(&d3, t3.size() * 8) != cudaSuccess) return 1;
        if (cudaMemcpy(d2, t2.data(), t2.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        if (cudaMemcpy(d3, t3.data(), t3.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        g_tape2b[device] = d2; g_tape3b[device] = d3;
        std::vector<double> tf(FAT3B_NTAPE);               // engine 3: the same coefficients in fat-row order
        for (int i = 0; i < FAT3B_NTAPE; ++i) tf[i] = FAT3B_COEFIDX_HOST[i] < 0 ? 0.0 : MBPOL_COEF3B_HOST[FAT3B_COEFIDX_HOST[i]];
        double* df = nullptr;
        if (cudaMalloc(&df, tf.size() * 8) != cudaSuccess || cudaMemcpy(df, tf.data(), tf.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        g_fat3b[device] = df;
        std::vector<double> tf2(FAT2B_NTAPE);
        for (int i = 0; i < FAT2B_NTAPE; ++i) tf2[i] = FAT2B_COEFIDX_HOST[i] < 0 ? 0.0 : MBPOL_COEF2B_HOST[FAT2B_COEFIDX_HOST[i]];
        double* df2 = nullptr;
        if (cudaMalloc(&df2, tf2.size() * 8) != cudaSuccess || cudaMemcpy(df2, tf2.data(), tf2.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        g_fat2b[device] = df2;
        if (cudaMemcpyToSymbol(c_fat_rowvars2b, FAT2B_ROWVARS_HOST, sizeof(FAT2B_ROWVARS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_fat_segs2b, FAT2B_SEGS_HOST, sizeof(FAT2B_SEGS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_fat_chunks2b, FAT2B_CHUNKS_HOST, sizeof(FAT2B_CHUNKS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_fat_rowvars3b, FAT3B_ROWVARS_HOST, sizeof(FAT3B_ROWVARS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_fat_segs3b, FAT3B_SEGS_HOST, sizeof(FAT3B_SEGS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_fat_chunks3b, FAT3B_CHUNKS_HOST, sizeof(FAT3B_CHUNKS_HOST)) != cudaSuccess) return 1;
        std::vector<double> ts(SYM3B_NTAPE);                // engine 4: value = a * |orbit| / |K|
        for (int i = 0; i < SYM3B_NTAPE; ++i)
            ts[i] = SYM3B_COEFIDX_HOST[i] < 0 ? 0.0 : MBPOL_COEF3B_HOST[SYM3B_COEFIDX_HOST[i]] * (double)SYM3B_COEFNUM_HOST[i] / (double)SYM3B_K;
        double* ds = nullptr;
        if (cudaMalloc(&ds, ts.size() * 8) != cudaSuccess || cudaMemcpy(ds, ts.data(), ts.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        g_sym3b[device] = ds;
        if (cudaMemcpyToSymbol(c_sym_rowvars3b, SYM3B_ROWVARS_HOST, sizeof(SYM3B_ROWVARS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_sym_segs3b, SYM3B_SEGS_HOST, sizeof(SYM3B_SEGS_HOST)) != cudaSuccess) return 1;
        if (cudaMemcpyToSymbol(c_sym_chunks3b, SYM3B_CHUNKS_HOST, sizeof(SYM3B_CHUNKS_HOST)) != cudaSuccess) return 1;
    }
    if (cudaMemcpyToSymbol(mbpol_par, &P, sizeof(P)) != cudaSuccess) return 1;
    {   // gather kernel: variable relabelling when molecules 0 and s of a trimer are swapped
        static const int va[36] = {1, 4, 7, 0, 0, 3, 3, 6, 6, 1, 1, 1, 1, 2, 2, 2, 2, 4, 4, 5, 5, 0, 0, 0, 0, 3, 3, 3, 3, 6, 6, 6, 6, 0, 0, 3};
        static const int vb[36] = {2, 5, 8, 1, 2, 4, 5, 7, 8, 4, 5, 7, 8, 4, 5, 7, 8, 7, 8, 7, 8, 4, 5, 7, 8, 1, 2, 7, 8, 1, 2, 4, 5, 3, 6, 6};
        int vp[3][36];
        for (int s = 0; s < 3; ++s)
            for (int n = 0; n < 36; ++n) {
                auto q = [&](int a) { const int m = a / 3, mm = (m == 0) ? s : ((m == s) ? 0 : m); return 3 * mm + a % 3; };
                const int a = q(va[n]), b = q(vb[n]);
                vp[s][n] = -1;
                for (int k = 0; k < 36; ++k)
                    if ((va[k] == a && vb[k] == b) || (va[k] == b && vb[k] == a)) vp[s][n] = k;
                if (vp[s][n] < 0) return 1;
            }
        if (cudaMemcpyToSymbol(c_gather_vp, vp, sizeof(vp)) != cudaSuccess) return 1;
        // 2B: sites 0..9 = Oa Ha1 Ha2 Ob Hb1 Hb2 Xa1 Xa2 Xb1 Xb2; swapping the molecules maps site s -> sw[s]
        static const int a2[31] = {1, 4, 0, 0, 3, 3, 1, 1, 2, 2, 0, 0, 3, 3, 0, 6, 6, 7, 7, 8, 8, 9, 9, 0, 0, 3, 3, 6, 6, 7, 7};
        static const int b2[31] = {2, 5, 1, 2, 4, 5, 4, 5, 4, 5, 4, 5, 1, 2, 3, 4, 5, 4, 5, 1, 2, 1, 2, 8, 9, 6, 7, 8, 9, 8, 9};
        static const int sw[10] = {3, 4, 5, 0, 1, 2, 8, 9, 6, 7};
        int vp2[31];
        for (int n = 0; n < 31; ++n) {
            const int a = sw[a2[n]], b = sw[b2[n]];
            vp2[n] = -1;
            for (int k = 0; k < 31; ++k)
                if ((a2[k] == a && b2[k] == b) || (a2[k] == b && b2[k] == a)) vp2[n] = k;
            if (vp2[n] < 0) return 1;
        }
        if (cudaMemcpyToSymbol(c_gather_vp2, vp2, sizeof(vp2)) != cudaSuccess) return 1;
    }
    return 0;
}

int ensure_constants(int device)
{
    static int status[16] = {0};
    if (device < 0 || device >= 16) return 1;
    std::call_once(g_const_once[device], [&]() { status[device] = upload_constants(device); });
    return status[device];
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

struct mbpol_b200_handle {
    int device = 0, nw = 0, flags = 0, rank = 0, nranks = 1, engine = 1, engine2b = 1;
    int nbeads = 1, nwb = 0, mxb = 0;      // bead-batched calls: nw = nbeads*nwb waters, mxatms = nbeads*mxb atoms
    int mxatms = 0, imcon = 0;
    bool have_vel = false, uploaded = false;
    std::string err;
    cudaStream_t stream = nullptr, vstream = nullptr, s2b = nullptr;      // s2b: the 2B chain of a step, concurrent with the 3B chain on `stream`      // vstream: velocity upload (needed only by k_flux at the end of the step)
    cudaEvent_t vel_ev = nullptr;
    cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    CellInfo C;
    Grid G;
    int ncells = 0;
    DevBuf<double> xyz, vel, mono, monoA, out, rec, rec2;         // rec: one 40-double record per trimer of the shard (k_trimers -> k_trimers_gather)
    DevBuf<int> list, cell_id, cell_cnt, cell_fill, sorted, unsorted, nb2, n2, nb3, n3, trimers, flag, inc_off, inc_cur, inc_off2, inc_cur2;
    DevBuf<unsigned> inc, inc2;                            // incidence lists water -> (trimer << 2 | slot) of the shard
    DevBuf<int2> dimers;
    DevBuf<long long> off2, off3, cnt3, n2ll, counts;
    DevBuf<char> scan_tmp, scan_tmp2;
    std::vector<int> host_list, list_cache;     // list_cache: the caller's listttm2 of the last upload
    int list_mxb = -1;
    double* pinned = nullptr;
    size_t pinned_cap = 0;
    std::vector<cudaEvent_t> dl_ev;            // one event behind every piece of the pipelined download
    // shard of this rank: the waters [w0, w0 + nown) are its centres (trimers) / first members (dimers)
    int w0 = 0, nown = 0;
    long long est_nd = -1, est_nt = -1;        // counts of the previous step: sizes and grids of this one (no host round trip)
    struct HostBack { long long counts[4]; double tail[3]; int flag; int pad; } *back = nullptr;     // pinned: read back once per step
    double* host_weight_d = nullptr;           // pinned: per-water load-balancing weights of the previous step (all ranks)
    std::vector<int> host_weight;
    bool have_weight = false, weights_pending = false, two_streams = false, dbg_times = false;
    cudaEvent_t dbg_ev = nullptr;
    double dbg_sum[3] = {0, 0, 0};
    long long dbg_n = 0, dbg_seen = 0;
    // fused end of mbpol_b200_forces (caller arrays in pinned memory): the caller's current values travel up on vstream while the
    // step runs, the device adds them (k_accumulate) and the sums go straight back into the caller's arrays -- no host "+=" pass
    struct FusedJob { double* dst; size_t off, len; };
    std::vector<FusedJob> fused;
    DevBuf<double> acc;
    unsigned acc_mask = ~0u;      // which per-atom arrays the zero-filled parts of acc were prepared for
    cudaEvent_t acc_ev = nullptr;
    ncclComm_t comm = nullptr;
    int reduce_mode = 0;                       // 0 none, 1 all-reduce, 2 reduce-scatter of the per-atom part (scalars: all-reduce)
    long long nd_total = 0, nt_total = 0, nd_shard = 0, nt_shard = 0, launches = 0;
    bool ms_stale = false;
    float ms[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // 0..7: device phases (events), 8..10: host wall time of upload / compute / download in the last mbpol_b200_forces
    // packed output buffer: [22 per-atom arrays of mxatms | pad to nranks equal chunks | NSCAL scalars per bead]
    size_t atom_count() const { return (size_t)22 * mxatms; }
    size_t chunk() const { const size_t c = (atom_count() + nranks - 1) / nranks; return (c + 1) & ~(size_t)1; }
    size_t scal_offset() const { return chunk() * nranks; }
    size_t out_count() const { return scal_offset() + (size_t)NSCAL * nbeads; }
    size_t weight_offset() const { return out_count(); }        // per-water weights of this step, as doubles, right behind the scalars
    size_t out_alloc() const { return out_count() + (size_t)nw + 2; }
    OutPtrs views()
    {
        OutPtrs O;
        const size_t n = mxatms;
        O.fx = out.p; O.fy = out.p + n; O.fz = out.p + 2 * n; O.aeng = out.p + 3 * n;
        O.astress = out.p + 4 * n; O.afd = out.p + 13 * n; O.scal = out.p + scal_offset(); O.mxatms = mxatms; O.nwb = nwb; O.mxb = mxb;
        return O;
    }
};

namespace {
__global__ void k_widen(int n, const int* __restrict__ a, long long* __restrict__ b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}

int scan_ll(mbpol_b200_handle* h, const long long* in, long long* out, int n)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, h->stream);
    CUDA_TRY(h, h->scan_tmp.reserve(bytes));
    CUDA_TRY(h, cub::DeviceScan::ExclusiveSum(h->scan_tmp.p, bytes, in, out, n, h->stream));
    return 0;
}
int scan_int(mbpol_b200_handle* h, const int* in, int* out, int n, bool second_stream = false)
{
    size_t bytes = 0;
    cudaStream_t s = second_stream ? h->s2b : h->stream;
    auto& tmp = second_stream ? h->scan_tmp2 : h->scan_tmp;      // the two chains of a step scan concurrently
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s);
    CUDA_TRY(h, tmp.reserve(bytes));
    CUDA_TRY(h, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, s));
    return 0;
}

// NCCL is bound lazily with dlopen / dlsym at the first communicator call, not at load time: a process that imports PyTorch AFTER
// this library must still get PyTorch's own (newer) libnccl.so.2 -- two copies under one SONAME cannot coexist.  An NCCL that is
// already in the process (RTLD_NOLOAD) is reused, otherwise the system library is opened.
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduceScatter) ReduceScatter = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
const NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(dlsym(lib, "ncclReduceScatter"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(lib, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.ReduceScatter && api.AllReduce && api.GetErrorString && api.GroupStart && api.GroupEnd;
    });
    return api;
}

// NVTX ranges of the step (the user-visible trace of the reference's three -DDEBUG timers, mbpol.f90:209-212, 280-284, 383-387)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define NCCL_TRY(h, call)                                                                                   \
    do {                                                                                                    \
        ncclResult_t r_ = (call);                                                                           \
        if (r_ != ncclSuccess) {                                                                            \
            (h)->err = std::string(#call) + ": " + nccl_api().GetErrorString(r_);                                  \
            return MBPOL_B200_ECUDA;                                                                        \
        }                                                                                                   \
    } while (0)
}  // namespace

// Shard boundaries: the waters are cut into nranks contiguous ranges of (nearly) equal WEIGHT, weight = 4 x trimers + dimers a
// water owned in the previous step; without weights (first step, one rank) the cut is uniform.
// Pure host arithmetic on data every rank holds identically -- exposed for the CPU tests.
extern "C" void mbpol_b200_shard_bounds(int nw, int nranks, const int* weight /* nw or NULL */, int* bounds /* nranks + 1 */)
{
    bounds[0] = 0;
    long long total = 0;
    if (weight)
        for (int i = 0; i < nw; ++i) total += weight[i] > 0 ? weight[i] : 0;
    if (!weight || total <= 0) {
        for (int r = 1; r <= nranks; ++r) bounds[r] = (int)((long long)nw * r / nranks);
        return;
    }
    long long acc = 0;
    int r = 1;
    for (int i = 0; i < nw && r < nranks; ++i) {
        acc += weight[i] > 0 ? weight[i] : 0;
        while (r < nranks && acc * nranks >= total * r) bounds[r++] = i + 1;
    }
    for (; r <= nranks; ++r) bounds[r] = nw;
    bounds[nranks] = nw;
}

extern "C" {

int mbpol_b200_init(int nttm2, int device, int flags, mbpol_b200_handle** handle)
{
    return mbpol_b200_init_batched(nttm2, 1, device, flags, handle);
}

int mbpol_b200_init_batched(int nttm2, int nbeads, int device, int flags, mbpol_b200_handle** handle)
{
    if (!handle || nttm2 < 0 || nttm2 % 3 || nbeads < 1 || (long long)nbeads * nttm2 > 0x3fffffffLL) return MBPOL_B200_EINVAL;
    *handle = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MBPOL_B200_ENODEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return MBPOL_B200_ENODEVICE;
    if (device >= ndev) return MBPOL_B200_EINVAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return MBPOL_B200_ENODEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return MBPOL_B200_ECUDA;
    if (ensure_constants(device)) return MBPOL_B200_ECUDA;
    auto* h = new mbpol_b200_handle;
    h->device = device; h->nbeads = nbeads; h->nwb = nttm2 / 3; h->nw = nbeads * h->nwb; h->flags = flags; h->engine = default_engine(); h->engine2b = default_engine_2b();
    {   // dynamic shared memory of the selected engines (a build with more warps per SM may leave no room for the engines it does not use)
        cudaError_t e2 = h->engine2b == 3 ? allow_smem(k_dimers<3>, SMEM2B) : allow_smem(k_dimers<1>, SMEM2B);
        cudaError_t e3 = h->engine == 4 ? allow_smem(k_trimers<4>, SMEM3S) : (h->engine == 3 ? allow_smem(k_trimers<3>, SMEM3B) : allow_smem(k_trimers<1>, SMEM3B));
        if (e2 != cudaSuccess || e3 != cudaSuccess) {
            cudaGetLastError();
            delete h;
            return MBPOL_B200_ECUDA;
        }
    }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MBPOL_B200_ECUDA; }
    if (const char* e = std::getenv("MBPOL_B200_TWO_STREAMS")) h->two_streams = std::atoi(e) != 0;
    if (const char* e = std::getenv("MBPOL_B200_DEBUG_TIMES")) h->dbg_times = std::atoi(e) != 0;
    if (h->dbg_times) cudaEventCreate(&h->dbg_ev);
    {
        int lo_pri = 0, hi_pri = 0;
        cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
        if (cudaStreamCreateWithPriority(&h->s2b, cudaStreamNonBlocking, hi_pri) != cudaSuccess) { delete h; return MBPOL_B200_ECUDA; }
    }
    if (cudaStreamCreateWithFlags(&h->vstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->vel_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->acc_ev, cudaEventDisableTiming) != cudaSuccess) { delete h; return MBPOL_B200_ECUDA; }
    for (auto& e : h->ev) cudaEventCreate(&e);
    if (cudaMallocHost(&h->back, sizeof(*h->back)) != cudaSuccess) { delete h; return MBPOL_B200_ECUDA; }
    *handle = h;
    return MBPOL_B200_OK;
}

int mbpol_b200_fini(mbpol_b200_handle* h)
{
    if (!h) return MBPOL_B200_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->dbg_times && h->dbg_n)
        fprintf(stderr, "mbpol_b200 rank %d: %lld steps, host launch phase %.3f ms, final wait %.3f ms, device time behind the reduction %.3f ms\n",
                h->rank, h->dbg_n, h->dbg_sum[0] / h->dbg_n, h->dbg_sum[1] / h->dbg_n, h->dbg_sum[2] / h->dbg_n);
    if (h->dbg_ev) cudaEventDestroy(h->dbg_ev);
    if (h->comm) nccl_api().CommDestroy(h->comm);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->back) cudaFreeHost(h->back);
    if (h->host_weight_d) cudaFreeHost(h->host_weight_d);
    for (cudaEvent_t e : h->dl_ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
    if (h->vstream) cudaStreamDestroy(h->vstream);
    if (h->s2b) cudaStreamDestroy(h->s2b);
    if (h->vel_ev) cudaEventDestroy(h->vel_ev);
    if (h->acc_ev) cudaEventDestroy(h->acc_ev);
    delete h;
    return MBPOL_B200_OK;
}

const char* mbpol_b200_last_error(const mbpol_b200_handle* h) { return h ? h->err.c_str() : "null handle"; }

void mbpol_b200_cutoffs(double* r2f, double* r3f) { *r2f = MBPOL_P2b_r2f; *r3f = MBPOL_P3b_r3f; }

int mbpol_b200_set_shard(mbpol_b200_handle* h, int rank, int nranks)
{
    if (!h || nranks < 1 || rank < 0 || rank >= nranks) return MBPOL_B200_EINVAL;
    if (h->comm && (rank != h->rank || nranks != h->nranks)) { h->err = "the shard of a handle with a communicator is fixed by mbpol_b200_comm_init"; return MBPOL_B200_EINVAL; }
    if (rank != h->rank || nranks != h->nranks) { h->est_nd = h->est_nt = -1; h->have_weight = h->weights_pending = false; }
    h->rank = rank; h->nranks = nranks;
    return MBPOL_B200_OK;
}

// ---- NCCL inside the boundary: the handle owns the communicator, the reduction runs on the handle's stream --------------
int mbpol_b200_comm_unique_id(void* id)
{
    static_assert(sizeof(ncclUniqueId) == MBPOL_B200_UNIQUE_ID_BYTES, "NCCL unique id size");
    if (!id) return MBPOL_B200_EINVAL;
    if (!nccl_api().ok) return MBPOL_B200_ECUDA;
    return nccl_api().GetUniqueId(reinterpret_cast<ncclUniqueId*>(id)) == ncclSuccess ? MBPOL_B200_OK : MBPOL_B200_ECUDA;
}

int mbpol_b200_comm_init(mbpol_b200_handle* h, int rank, int nranks, const void* id)
{
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks || h->comm) return MBPOL_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    if (!nccl_api().ok) { h->err = "libnccl.so.2 not found (or too old): the in-library reduction needs NCCL"; return MBPOL_B200_ECUDA; }
    NCCL_TRY(h, nccl_api().CommInitRank(&h->comm, nranks, uid, rank));
    h->rank = rank; h->nranks = nranks;
    h->est_nd = h->est_nt = -1; h->have_weight = h->weights_pending = false;
    h->reduce_mode = nranks > 1 ? 2 : 0;
    return MBPOL_B200_OK;
}

int mbpol_b200_comm_mode(mbpol_b200_handle* h, int mode)
{
    if (!h || mode < 0 || mode > 2 || (mode && !h->comm)) return MBPOL_B200_EINVAL;
    h->reduce_mode = h->nranks > 1 ? mode : 0;
    return MBPOL_B200_OK;
}

int mbpol_b200_host_register(void* p, size_t bytes) { return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? MBPOL_B200_OK : MBPOL_B200_ECUDA; }
int mbpol_b200_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? MBPOL_B200_OK : MBPOL_B200_ECUDA; }

void* mbpol_b200_stream(mbpol_b200_handle* h) { return h ? (void*)h->stream : nullptr; }

int mbpol_b200_upload(mbpol_b200_handle* h, int imcon, const double* cell, int nttm2, const int* listttm2, int mxatms,
                      const double* xxx, const double* yyy, const double* zzz, const double* vxx, const double* vyy, const double* vzz)
{
    if (!h) return MBPOL_B200_EINVAL;
    if (imcon < 1 || imcon > 3 || !cell || nttm2 != 3 * h->nwb || !listttm2 || mxatms < nttm2 || !xxx || !yyy || !zzz) {
        h->err = "mbpol_b200_upload: invalid argument (imcon must be 1, 2 or 3; nttm2 must match init)";
        return MBPOL_B200_EINVAL;
    }
    // the topology rarely changes between MD steps: it is validated and uploaded only when it differs from the cached copy
    const bool same_list = h->list_mxb == mxatms && h->list_cache.size() == (size_t)nttm2 &&
                           std::memcmp(h->list_cache.data(), listttm2, (size_t)nttm2 * 4) == 0;
    if (!same_list) {
        for (int t = 0; t < nttm2; t += 3)      // H atoms are addressed as o+1, o+2 (mbpol.f90:526-527, 592-594): the whole water must be in range
            if (listttm2[t] < 1 || listttm2[t] + 2 > mxatms || listttm2[t + 1] < 1 || listttm2[t + 1] > mxatms || listttm2[t + 2] < 1 || listttm2[t + 2] > mxatms) {
                h->err = "mbpol_b200_upload: listttm2 entry out of range [1, mxatms]";
                return MBPOL_B200_EINVAL;
            }
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    // bead-batched handles: the arrays hold nbeads consecutive systems of mxatms atoms (one bead otherwise)
    h->mxb = mxatms;
    mxatms *= h->nbeads;
    h->mxatms = mxatms; h->imcon = imcon;
    const size_t n = mxatms;
    CellInfo& C = h->C;
    C.imcon = imcon;
    std::copy(cell, cell + 9, C.cell);
    double det;
    invert3(cell, C.rcell, det);
    C.a1 = 1.0 / cell[0]; C.b1 = 1.0 / cell[4]; C.c1 = 1.0 / cell[8];
    // candidate grid: fractional coordinates through the inverse cell; for imcon 1/2 the cell is diagonal
    double gc[9] = {0};
    if (imcon == 1) { gc[0] = gc[4] = gc[8] = cell[0]; }
    else if (imcon == 2) { gc[0] = cell[0]; gc[4] = cell[4]; gc[8] = cell[8]; }
    else std::copy(cell, cell + 9, gc);
    double grc[9], gdet;
    invert3(gc, grc, gdet);
    if (!(std::fabs(gdet) > 0.0)) { h->err = "singular cell"; return MBPOL_B200_EINVAL; }
    std::copy(grc, grc + 9, h->G.rc);
    const double rcut = MBPOL_P2b_r2f * 1.0001;
    h->ncells = 1;
    for (int k = 0; k < 3; ++k) {   // perpendicular width of the cell along reciprocal vector k = 1/|row k of inverse|
        const double len = std::sqrt(grc[k] * grc[k] + grc[3 + k] * grc[3 + k] + grc[6 + k] * grc[6 + k]);
        int nc = (int)std::floor((1.0 / len) / rcut);
        nc = std::max(1, std::min(nc, 1024));
        h->G.nc[k] = nc;
        h->ncells *= nc;
    }
    h->G.nwb = h->nwb;
    h->G.ncb = h->ncells;
    h->ncells *= h->nbeads;
    CUDA_TRY(h, h->xyz.reserve(3 * n));
    CUDA_TRY(h, h->list.reserve((size_t)nttm2 * h->nbeads));
    CUDA_TRY(h, cudaMemcpyAsync(h->xyz.p, xxx, n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->xyz.p + n, yyy, n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->xyz.p + 2 * n, zzz, n * 8, cudaMemcpyHostToDevice, h->stream));
    if (same_list) {
        // device copy is current
    } else if (h->nbeads == 1) {
        h->list_cache.assign(listttm2, listttm2 + nttm2);
        h->list_mxb = mxatms / h->nbeads;
        CUDA_TRY(h, cudaMemcpyAsync(h->list.p, h->list_cache.data(), (size_t)nttm2 * 4, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else {
        h->list_cache.assign(listttm2, listttm2 + nttm2);
        h->list_mxb = mxatms / h->nbeads;   // the same topology in every bead: atom indices shifted by the bead's offset into the concatenated arrays
        h->host_list.resize((size_t)nttm2 * h->nbeads);
        for (int b = 0; b < h->nbeads; ++b)
            for (int t = 0; t < nttm2; ++t) h->host_list[(size_t)b * nttm2 + t] = listttm2[t] + b * h->mxb;
        CUDA_TRY(h, cudaMemcpyAsync(h->list.p, h->host_list.data(), h->host_list.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    h->have_vel = vxx && vyy && vzz;
    if (h->have_vel) {
        CUDA_TRY(h, h->vel.reserve(3 * n));
        // on their own stream: the step only waits for them in front of k_flux, so the copy overlaps the list build and the 2B / 3B kernels
        CUDA_TRY(h, cudaMemcpyAsync(h->vel.p, vxx, n * 8, cudaMemcpyHostToDevice, h->vstream));
        CUDA_TRY(h, cudaMemcpyAsync(h->vel.p + n, vyy, n * 8, cudaMemcpyHostToDevice, h->vstream));
        CUDA_TRY(h, cudaMemcpyAsync(h->vel.p + 2 * n, vzz, n * 8, cudaMemcpyHostToDevice, h->vstream));
        CUDA_TRY(h, cudaEventRecord(h->vel_ev, h->vstream));
    }
    h->uploaded = true;
    return MBPOL_B200_OK;
}

// One step on the handle's stream.  With counts from the previous step at hand (est_nd / est_nt) nothing between the first
// and the last launch waits for the host: list capacities and grids come from the estimates, the kernels read the true counts
// from device memory, and an overflow flag (checked after the single synchronisation at the end) triggers a sized retry.
static int compute_once(mbpol_b200_handle* h, bool sized_by_estimate)
{
    using dclk = std::chrono::steady_clock;
    const dclk::time_point dbg_t0 = dclk::now();
    dclk::time_point dbg_t1 = dbg_t0, dbg_t2 = dbg_t0;
    const int nw = h->nw;
    const size_t n = h->mxatms;
    cudaStream_t s = h->stream;
    const int TB = 128;
    h->launches = 0;
    const int gb = (nw + TB - 1) / TB;
    CUDA_TRY(h, h->mono.reserve(9 * (size_t)nw));
    CUDA_TRY(h, h->monoA.reserve(12 * (size_t)nw));
    CUDA_TRY(h, h->out.reserve(h->out_alloc()));
    CUDA_TRY(h, h->cell_id.reserve(nw)); CUDA_TRY(h, h->sorted.reserve(nw)); CUDA_TRY(h, h->unsorted.reserve(nw));
    CUDA_TRY(h, h->cell_cnt.reserve(h->ncells + 1)); CUDA_TRY(h, h->cell_fill.reserve(h->ncells + 1));
    CUDA_TRY(h, h->flag.reserve(1)); CUDA_TRY(h, h->counts.reserve(4));
    CUDA_TRY(h, h->inc_off.reserve(nw + 1)); CUDA_TRY(h, h->inc_cur.reserve(nw + 1));
    CUDA_TRY(h, h->inc_off2.reserve(nw + 1)); CUDA_TRY(h, h->inc_cur2.reserve(nw + 1));

    CUDA_TRY(h, cudaEventRecord(h->ev[0], s));
    CUDA_TRY(h, cudaMemsetAsync(h->out.p, 0, h->out_alloc() * 8, s));
    OutPtrs O = h->views();
    h->back->flag = 0;
    for (long long& c : h->back->counts) c = 0;
    if (nw == 0) {   // no water at all (the reference's loops simply do not execute): zero contributions, nothing to launch
        h->nd_total = h->nt_total = h->nd_shard = h->nt_shard = 0;
        for (int k = 1; k < 10; ++k) CUDA_TRY(h, cudaEventRecord(h->ev[k], s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
        for (float& m : h->ms) m = 0.f;
        return MBPOL_B200_OK;
    }
    CUDA_TRY(h, cudaMemsetAsync(h->cell_cnt.p, 0, (size_t)(h->ncells + 1) * 4, s));
    CUDA_TRY(h, cudaMemsetAsync(h->cell_fill.p, 0, (size_t)(h->ncells + 1) * 4, s));
    CUDA_TRY(h, cudaMemsetAsync(h->flag.p, 0, 4, s));
    k_gather_image<<<gb, TB, 0, s>>>(nw, h->list.p, h->xyz.p, h->xyz.p + n, h->xyz.p + 2 * n, h->C, h->mono.p, h->monoA.p);
    k_cell_index<<<gb, TB, 0, s>>>(nw, h->mono.p, h->G, h->cell_id.p, h->cell_cnt.p);
    if (int rc = scan_int(h, h->cell_cnt.p, h->cell_cnt.p, h->ncells + 1)) return rc;   // in-place exclusive scan -> cell_start
    k_cell_fill<<<gb, TB, 0, s>>>(nw, h->cell_id.p, h->cell_cnt.p, h->cell_fill.p, h->unsorted.p);
    k_sort_cells<<<gb, TB, 0, s>>>(nw, h->cell_id.p, h->cell_cnt.p, h->unsorted.p, h->sorted.p);
    h->launches += 6;
    // shard of this rank.  The weights of the previous step (all-reduced, read back at its end) are turned into the water ranges
    // HERE, while the shard-independent kernels above are running: no host work sits between two steps with the GPU idle.
    {
        if (h->weights_pending) {
            h->host_weight.resize(nw);
            for (int i = 0; i < nw; ++i) h->host_weight[i] = (int)h->host_weight_d[i];
            h->weights_pending = false;
        }
        std::vector<int> bounds(h->nranks + 1);
        mbpol_b200_shard_bounds(nw, h->nranks, (h->nranks > 1 && h->have_weight) ? h->host_weight.data() : nullptr, bounds.data());
        h->w0 = bounds[h->rank];
        h->nown = bounds[h->rank + 1] - bounds[h->rank];
    }
    const int w0 = h->w0, nown = h->nown;
    const int go = (nown + TB - 1) / TB;
    CUDA_TRY(h, h->nb2.reserve((size_t)std::max(nown, 1) * NB2_CAP)); CUDA_TRY(h, h->nb3.reserve((size_t)std::max(nown, 1) * NB3_CAP));
    CUDA_TRY(h, h->n2.reserve(nown + 1)); CUDA_TRY(h, h->n3.reserve(nown + 1));
    CUDA_TRY(h, h->off2.reserve(nown + 1)); CUDA_TRY(h, h->off3.reserve(nown + 1)); CUDA_TRY(h, h->cnt3.reserve(nown + 1));
    CUDA_TRY(h, h->n2ll.reserve(nown + 1));
    nvtxRangePushA("mbpol_b200: lists");
    const int gw = (int)(((size_t)std::max(nown, 1) * 32 + TB - 1) / TB);      // one warp per owned water
    if (nown > 0) {
        k_neighbours<<<gw, TB, 0, s>>>(w0, nown, nw, h->mono.p, h->C, h->G, h->cell_cnt.p, h->sorted.p, h->nb2.p, h->n2.p, h->nb3.p, h->n3.p, h->flag.p);
        k_widen<<<go, TB, 0, s>>>(nown, h->n2.p, h->n2ll.p);
        k_trimers_enum<false><<<gw, TB, 0, s>>>(w0, nown, nw, h->mono.p, h->C, h->nb3.p, h->n3.p, h->n2.p, h->cnt3.p, h->out.p + h->weight_offset(), nullptr, nullptr, 0, h->flag.p);
        h->launches += 3;
    }
    CUDA_TRY(h, cudaMemsetAsync(h->n2ll.p + nown, 0, 8, s));
    CUDA_TRY(h, cudaMemsetAsync(h->cnt3.p + nown, 0, 8, s));
    if (int rc = scan_ll(h, h->n2ll.p, h->off2.p, nown + 1)) return rc;
    if (int rc = scan_ll(h, h->cnt3.p, h->off3.p, nown + 1)) return rc;
    h->launches += 4;
    long long cap2, cap3;
    if (sized_by_estimate) {
        cap2 = h->est_nd + h->est_nd / 4 + 4096;
        cap3 = h->est_nt + h->est_nt / 4 + 4096;
    } else {   // first step (or after an overflow): read the totals back once
        long long tot[2];
        int ovf = 0;
        CUDA_TRY(h, cudaMemcpyAsync(&tot[0], h->off2.p + nown, 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaMemcpyAsync(&tot[1], h->off3.p + nown, 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaMemcpyAsync(&ovf, h->flag.p, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
        if (ovf == 1) { h->err = "neighbour capacity exceeded (NB2_CAP/NB3_CAP): unphysical density"; return MBPOL_B200_ECAPACITY; }
        cap2 = tot[0]; cap3 = tot[1];
        h->est_nd = tot[0]; h->est_nt = tot[1];
    }
    CUDA_TRY(h, h->dimers.reserve((size_t)std::max<long long>(cap2, 1)));
    CUDA_TRY(h, h->trimers.reserve((size_t)std::max<long long>(3 * cap3, 3)));
    CUDA_TRY(h, h->rec.reserve((size_t)std::max<long long>(cap3, 1) * MBPOL_REC3));
    CUDA_TRY(h, h->inc.reserve((size_t)std::max<long long>(3 * cap3, 3)));
    CUDA_TRY(h, h->rec2.reserve((size_t)std::max<long long>(cap2, 1) * MBPOL_REC2));
    CUDA_TRY(h, h->inc2.reserve((size_t)std::max<long long>(2 * cap2, 2)));
    if (nown > 0) {
        k_fill_dimers<<<go, TB, 0, s>>>(w0, nown, h->nb2.p, h->n2.p, h->off2.p, h->dimers.p, cap2, h->flag.p);
        k_trimers_enum<true><<<gw, TB, 0, s>>>(w0, nown, nw, h->mono.p, h->C, h->nb3.p, h->n3.p, h->n2.p, nullptr, nullptr, h->off3.p, h->trimers.p, cap3, h->flag.p);
        h->launches += 2;
    }
    k_publish_counts<<<1, 1, 0, s>>>(h->off2.p + nown, h->off3.p + nown, cap2, cap3, h->est_nd, h->est_nt, h->counts.p, O.scal);
    ++h->launches;
    CUDA_TRY(h, cudaEventRecord(h->ev[1], s));
    nvtxRangePop();

    // ---- 2B.  Optionally (MBPOL_B200_TWO_STREAMS=1) on its own high-priority stream, concurrent with the 3B chain below: both
    // polynomial kernels fill whole SMs, so nothing overlaps inside an SM, but the CTAs of the two kernels pack the SMs back to back.
    // Measured: one GPU 20.12 -> 20.06 ms, two 10.64 -> 10.56, but four 5.77 -> 5.94 and eight 3.22 -> 3.33 ms per step (the
    // emulated 1/8 shard on one GPU had gained 0.08 ms): off by default.  The two gathers add into the same per-atom arrays:
    // k_trimers_gather waits for ev[2].
    const double* xr = h->xyz.p; const double* yr = h->xyz.p + n; const double* zr = h->xyz.p + 2 * n;
    nvtxRangePushA("mbpol_b200: 2B");
    if (h->est_nd > 0) {
        cudaStream_t s2 = h->two_streams ? h->s2b : s;
        if (h->two_streams) CUDA_TRY(h, cudaStreamWaitEvent(s2, h->ev[1], 0));
        const unsigned gd = (unsigned)((h->est_nd + MBPOL_BS2 - 1) / MBPOL_BS2);
        const unsigned gi = (unsigned)std::min<long long>((2 * h->est_nd + 255) / 256, 1 << 20);
        CUDA_TRY(h, cudaMemsetAsync(h->inc_off2.p, 0, (size_t)(nw + 1) * 4, s2));
        CUDA_TRY(h, cudaMemsetAsync(h->inc_cur2.p, 0, (size_t)(nw + 1) * 4, s2));
        k_inc_count<2><<<gi, 256, 0, s2>>>(h->counts.p, reinterpret_cast<const int*>(h->dimers.p), h->inc_off2.p);
        if (int rc = scan_int(h, h->inc_off2.p, h->inc_off2.p, nw + 1, h->two_streams)) return rc;
        k_inc_fill<2><<<gi, 256, 0, s2>>>(h->counts.p, reinterpret_cast<const int*>(h->dimers.p), h->inc_off2.p, h->inc_cur2.p, h->inc2.p);
        if (h->engine2b == 3)
            k_dimers<3><<<gd, MBPOL_BS2, SMEM2B, s2>>>(h->counts.p, h->dimers.p, nw, h->mono.p, h->C, g_fat2b[h->device], h->rec2.p);
        else
            k_dimers<1><<<gd, MBPOL_BS2, SMEM2B, s2>>>(h->counts.p, h->dimers.p, nw, h->mono.p, h->C, g_tape2b[h->device], h->rec2.p);
        k_dimers_gather<<<(unsigned)((nw + 3) / 4), 128, 0, s2>>>(nw, h->inc_off2.p, h->inc2.p, h->dimers.p, h->rec2.p, h->monoA.p, h->list.p, xr, yr, zr, h->C, O, h->flags);
        h->launches += 6;
        CUDA_TRY(h, cudaEventRecord(h->ev[2], s2));
    } else {
        CUDA_TRY(h, cudaEventRecord(h->ev[2], s));
    }
    nvtxRangePop();
    nvtxRangePushA("mbpol_b200: 3B");
    // ---- 3B: incidence lists of the shard (count -> exclusive scan -> fill), records, gather -----------------------------
    if (h->est_nt > 0) {
        const unsigned gt = (unsigned)((h->est_nt + MBPOL_BS3 - 1) / MBPOL_BS3);
        const unsigned gi = (unsigned)std::min<long long>((3 * h->est_nt + 255) / 256, 1 << 20);
        CUDA_TRY(h, cudaEventRecord(h->ev[9], s));
        CUDA_TRY(h, cudaMemsetAsync(h->inc_off.p, 0, (size_t)(nw + 1) * 4, s));
        CUDA_TRY(h, cudaMemsetAsync(h->inc_cur.p, 0, (size_t)(nw + 1) * 4, s));
        k_inc_count<3><<<gi, 256, 0, s>>>(h->counts.p + 1, h->trimers.p, h->inc_off.p);
        if (int rc = scan_int(h, h->inc_off.p, h->inc_off.p, nw + 1)) return rc;
        k_inc_fill<3><<<gi, 256, 0, s>>>(h->counts.p + 1, h->trimers.p, h->inc_off.p, h->inc_cur.p, h->inc.p);
        CUDA_TRY(h, cudaEventRecord(h->ev[5], s));
        if (h->engine == 4)
            k_trimers<4><<<gt, MBPOL_BS3, SMEM3S, s>>>(h->counts.p + 1, h->trimers.p, nw, h->mono.p, h->C, g_sym3b[h->device], h->rec.p);
        else if (h->engine == 3)
            k_trimers<3><<<gt, MBPOL_BS3, SMEM3B, s>>>(h->counts.p + 1, h->trimers.p, nw, h->mono.p, h->C, g_fat3b[h->device], h->rec.p);
        else
            k_trimers<1><<<gt, MBPOL_BS3, SMEM3B, s>>>(h->counts.p + 1, h->trimers.p, nw, h->mono.p, h->C, g_tape3b[h->device], h->rec.p);
        CUDA_TRY(h, cudaEventRecord(h->ev[6], s));
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev[2], 0));      // the 2B gather has finished with the per-atom arrays
        k_trimers_gather<<<(unsigned)((nw + 3) / 4), 128, 0, s>>>(nw, h->inc_off.p, h->inc.p, h->trimers.p, h->rec.p, h->monoA.p, h->list.p, xr, yr, zr, h->C, O, h->flags);
        CUDA_TRY(h, cudaEventRecord(h->ev[7], s));
        h->launches += 6;
    } else {
        CUDA_TRY(h, cudaEventRecord(h->ev[9], s));
        for (int k = 5; k < 8; ++k) CUDA_TRY(h, cudaEventRecord(h->ev[k], s));
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev[2], 0));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[3], s));
    nvtxRangePop();
    NvtxRange tail_range("mbpol_b200: flux + reduction + read-back");
    if (h->have_vel) {
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->vel_ev, 0));
        k_flux<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(O, h->vel.p, h->vel.p + n, h->vel.p + 2 * n, h->flags);
        ++h->launches;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[8], s));
    // ---- reduction over the ranks (NCCL over NVLink, on this stream): forces, per-atom arrays, scalars, weights -------------
    if (h->reduce_mode && h->comm) {
        if (h->reduce_mode == 2) {
            NCCL_TRY(h, nccl_api().GroupStart());      // one fused launch for the two collectives
            NCCL_TRY(h, nccl_api().ReduceScatter(h->out.p, h->out.p + (size_t)h->rank * h->chunk(), h->chunk(), ncclDouble, ncclSum, h->comm, s));
            NCCL_TRY(h, nccl_api().AllReduce(O.scal, O.scal, (size_t)NSCAL * h->nbeads + nw, ncclDouble, ncclSum, h->comm, s));      // scalars + weights
            NCCL_TRY(h, nccl_api().GroupEnd());
        } else {
            NCCL_TRY(h, nccl_api().AllReduce(h->out.p, h->out.p, h->out_count() + nw, ncclDouble, ncclSum, h->comm, s));
        }
        h->launches += 2;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[4], s));
    if (!h->fused.empty()) {      // fused end of mbpol_b200_forces: caller values (already on the device) + this step, straight into the caller's arrays
        const size_t a = h->fused.front().off, b = h->fused.back().off + h->fused.back().len;
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->acc_ev, 0));
        k_accumulate<<<(unsigned)std::min<size_t>((b - a + 255) / 256, 148 * 8), 256, 0, s>>>(b - a, h->acc.p + a, h->out.p + a);
        ++h->launches;
        for (const auto& j : h->fused) CUDA_TRY(h, cudaMemcpyAsync(j.dst, h->out.p + j.off, j.len * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaMemcpyAsync(h->pinned + h->scal_offset(), h->out.p + h->scal_offset(), (size_t)NSCAL * h->nbeads * 8, cudaMemcpyDeviceToHost, s));
    }
    // ---- the one read-back of the step: counts, overflow flag, weights ---------------------------------------------------
    CUDA_TRY(h, cudaMemcpyAsync(h->back->counts, h->counts.p, 32, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&h->back->flag, h->flag.p, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->back->tail, O.scal + S_ND, 24, cudaMemcpyDeviceToHost, s));      // S_ND, S_NT, S_OVF
    if (h->nranks > 1) {
        if (!h->host_weight_d) CUDA_TRY(h, cudaMallocHost(&h->host_weight_d, (size_t)(nw + 1) * 8));
        CUDA_TRY(h, cudaMemcpyAsync(h->host_weight_d, h->out.p + h->weight_offset(), (size_t)nw * 8, cudaMemcpyDeviceToHost, s));
    }
    if (h->dbg_times) CUDA_TRY(h, cudaEventRecord(h->dbg_ev, s));
    CUDA_TRY(h, cudaGetLastError());
    dbg_t1 = dclk::now();
    CUDA_TRY(h, cudaStreamSynchronize(s));
    dbg_t2 = dclk::now();
    if (h->dbg_times && h->fused.empty() && ++h->dbg_seen > 4) {      // MBPOL_B200_DEBUG_TIMES=1 (compute-only steps, after the first four): host wall time of the launch phase and of the final wait, GPU time behind the reduction
        float tail_ms = 0.f;
        cudaEventElapsedTime(&tail_ms, h->ev[4], h->dbg_ev);
        h->dbg_sum[0] += std::chrono::duration<double, std::milli>(dbg_t1 - dbg_t0).count();
        h->dbg_sum[1] += std::chrono::duration<double, std::milli>(dbg_t2 - dbg_t1).count();
        h->dbg_sum[2] += tail_ms;
        ++h->dbg_n;
    }
    if (h->back->flag == 1) { h->err = "neighbour capacity exceeded (NB2_CAP/NB3_CAP): unphysical density"; return MBPOL_B200_ECAPACITY; }
    if (h->back->tail[2] > 0.0) {      // (all-reduced when a communicator is attached: every rank takes this branch together)
        h->est_nd = h->back->counts[2]; h->est_nt = h->back->counts[3];
        return -1;       // an estimate was too small somewhere: the caller repeats the step with exact sizes
    }
    h->nd_shard = h->back->counts[0]; h->nt_shard = h->back->counts[1];
    h->est_nd = h->nd_shard; h->est_nt = h->nt_shard;
    // box totals: the counts rode through the reduction in the scalar block (bead 0); without one, the shard is all this rank knows
    h->nd_total = (long long)(h->back->tail[0] + 0.5); h->nt_total = (long long)(h->back->tail[1] + 0.5);
    h->have_weight = h->nranks > 1 && h->reduce_mode && h->comm;      // without a communicator the weights of the other ranks are unknown
    h->weights_pending = h->have_weight;      // converted at the start of the next step, under its first kernels
    h->ms_stale = true;      // the phase times are read from the events only when somebody asks (refresh_phase_times)
    return MBPOL_B200_OK;
}

static void refresh_phase_times(mbpol_b200_handle* h)
{
    if (!h->ms_stale) return;
    h->ms_stale = false;
    if (h->nw == 0) return;
    cudaEventElapsedTime(&h->ms[0], h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->ms[1], h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&h->ms[2], h->ev[9], h->ev[3]);      // 3B chain (overlaps the 2B chain: ms[1] + ms[2] > the step)
    cudaEventElapsedTime(&h->ms[3], h->ev[0], h->ev[4]);
    cudaEventElapsedTime(&h->ms[4], h->ev[5], h->ev[6]);      // k_trimers alone (the roofline kernel)
    cudaEventElapsedTime(&h->ms[5], h->ev[6], h->ev[7]);      // k_trimers_gather
    cudaEventElapsedTime(&h->ms[6], h->ev[8], h->ev[4]);      // reduction
    cudaEventElapsedTime(&h->ms[7], h->ev[9], h->ev[5]);      // incidence lists
}

int mbpol_b200_compute(mbpol_b200_handle* h)
{
    if (!h || !h->uploaded) return MBPOL_B200_EINVAL;
    NvtxRange range("mbpol_b200_compute");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = compute_once(h, h->est_nd >= 0 && h->est_nt >= 0);
    if (rc == -1) {      // rare: the lists outgrew the previous step's counts by > 25 % (collective decision, see k_publish_counts)
        rc = compute_once(h, false);
        if (rc == -1) { h->err = "work lists changed size during the step"; return MBPOL_B200_ECUDA; }
    }
    return rc;
}

int mbpol_b200_download(mbpol_b200_handle* h, double* fxx, double* fyy, double* fzz, double* eng2, double* vir2, double* eng3,
                        double* vir3, double* stress, double* atomic_energy, double* atomic_stress, double* atomic_fdecomp, double* jflux)
{
    if (!h || !h->uploaded || !fxx || !fyy || !fzz) return MBPOL_B200_EINVAL;
    NvtxRange range("mbpol_b200_download");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = h->mxatms, cnt = h->out_count();
    if (h->pinned_cap < cnt) {
        if (h->pinned) cudaFreeHost(h->pinned);
        h->pinned = nullptr; h->pinned_cap = 0;
        CUDA_TRY(h, cudaMallocHost(&h->pinned, cnt * 8));
        h->pinned_cap = cnt;
    }
    double* P = h->pinned;
    // After a reduce-scatter this rank holds (and returns) only its slice [lo, hi) of the flat per-atom part; the caller's sum over the
    // ranks (gdsum, dlpoly_cmd.f:2551, 2568; heatcurrent.f90:1129-1259) completes the arrays exactly as it does for the reference.
    const bool slice = h->reduce_mode == 2 && h->comm && h->nranks > 1;
    const size_t lo = slice ? std::min((size_t)h->rank * h->chunk(), h->atom_count()) : 0;
    const size_t hi = slice ? std::min((size_t)(h->rank + 1) * h->chunk(), h->atom_count()) : h->atom_count();
    // Copy only what the caller asked for, in pieces with an event behind each piece: the host threads accumulate piece k
    // into the caller's arrays (the reference's "+=" semantics) while piece k+1 is still crossing PCIe.
    struct Job { double* dst; size_t off, len; };
    std::vector<Job> jobs;
    auto add = [&](double* dst, size_t a, size_t b) {      // caller array dst covers flat [a, b)
        if (!dst) return;
        const size_t x = std::max(a, lo), y = std::min(b, hi);
        if (x >= y) return;
        const size_t piece = std::max<size_t>((y - x + 1) / 2, 1 << 16);
        for (size_t q = x; q < y; q += piece) jobs.push_back({dst + (q - a), q, std::min(piece, y - q)});
    };
    add(fxx, 0, n); add(fyy, n, 2 * n); add(fzz, 2 * n, 3 * n);
    add(atomic_energy, 3 * n, 4 * n);
    add(atomic_stress, 4 * n, 13 * n);
    add(atomic_fdecomp, 13 * n, 22 * n);
    while (h->dl_ev.size() < jobs.size() + 1) {
        cudaEvent_t e;
        CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->dl_ev.push_back(e);
    }
    size_t total = 0;
    for (const Job& j : jobs) total += j.len;
    const size_t so = h->scal_offset();
    const bool one_copy = !jobs.empty() && total < (1u << 20);      // a small slice (multi-GPU): one copy, one wait -- latency, not bandwidth, is what counts
    if (one_copy) {
        const size_t a = jobs.front().off, b = jobs.back().off + jobs.back().len;
        CUDA_TRY(h, cudaMemcpyAsync(P + a, h->out.p + a, (b - a) * 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(P + so, h->out.p + so, (size_t)NSCAL * h->nbeads * 8, cudaMemcpyDeviceToHost, h->stream));
        for (size_t k = 0; k <= jobs.size(); ++k) CUDA_TRY(h, cudaEventRecord(h->dl_ev[k], h->stream));
    } else {
        for (size_t k = 0; k < jobs.size(); ++k) {
            CUDA_TRY(h, cudaMemcpyAsync(P + jobs[k].off, h->out.p + jobs[k].off, jobs[k].len * 8, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaEventRecord(h->dl_ev[k], h->stream));
        }
        CUDA_TRY(h, cudaMemcpyAsync(P + so, h->out.p + so, (size_t)NSCAL * h->nbeads * 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->dl_ev[jobs.size()], h->stream));
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nth = total < (1u << 20) ? 1u : std::min(16u, std::max(1u, hw / (unsigned)std::max(1, h->nranks)));
    auto work = [&](unsigned t) {
        cudaSetDevice(h->device);
        for (size_t k = 0; k < jobs.size(); ++k) {
            cudaEventSynchronize(h->dl_ev[k]);
            const Job& j = jobs[k];
            const double* src = P + j.off;
            const size_t a = j.len * t / nth, b = j.len * (t + 1) / nth;
            for (size_t i = a; i < b; ++i) j.dst[i] += src[i];
        }
    };
    if (nth == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nth; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const bool reduced = h->reduce_mode && h->comm && h->nranks > 1;
    for (int b = 0; b < h->nbeads; ++b) {     // scalars are per bead: eng2[b], ..., stress[9*b + q], jflux[9*b + q]
        const double* sc = P + so + (size_t)b * NSCAL;
        // after an in-library reduction every rank holds the TOTALS: only rank 0 adds them, so that the caller's gdsum of the
        // four scalars (mbpol.f90:395-431) and of stress (dlpoly_cmd.f:2568) still gives the right sums
        const bool mine = !reduced || h->rank == 0;
        if (mine) {
            if (eng2) eng2[b] += sc[S_ENG2];
            if (vir2) vir2[b] += sc[S_VIR2];
            if (eng3) eng3[b] += sc[S_ENG3];
            if (vir3) vir3[b] += sc[S_VIR3];
            if (stress) for (int q = 0; q < 9; ++q) stress[9 * b + q] += sc[S_STRESS + q];
        }
        if (jflux) for (int q = 0; q < 9; ++q) jflux[9 * b + q] = mine ? sc[S_JFLUX + q] : 0.0;
    }
    return MBPOL_B200_OK;
}

int mbpol_b200_forces(mbpol_b200_handle* h, int imcon, const double* cell, int nttm2, const int* listttm2, int mxatms,
                      const double* xxx, const double* yyy, const double* zzz, double* fxx, double* fyy, double* fzz, double* eng2,
                      double* vir2, double* eng3, double* vir3, double* stress, double* atomic_energy, double* atomic_stress,
                      double* atomic_fdecomp, const double* vxx, const double* vyy, const double* vzz, double* jflux)
{
    if (!h) return MBPOL_B200_EINVAL;
    if (atomic_stress && !(h->flags & MBPOL_B200_WANT_ATOMIC_STRESS)) {
        h->err = "atomic_stress requested but the handle was created without MBPOL_B200_WANT_ATOMIC_STRESS";
        return MBPOL_B200_EINVAL;
    }
    if (atomic_energy && !(h->flags & MBPOL_B200_WANT_ATOMIC_ENERGY)) {
        h->err = "atomic_energy requested but the handle was created without MBPOL_B200_WANT_ATOMIC_ENERGY";
        return MBPOL_B200_EINVAL;
    }
    if (atomic_fdecomp && !(h->flags & MBPOL_B200_WANT_FDECOMP)) {
        h->err = "atomic_fdecomp requested but the handle was created without MBPOL_B200_WANT_FDECOMP";
        return MBPOL_B200_EINVAL;
    }
    if (jflux && !(vxx && vyy && vzz)) {
        h->err = "jflux requested without velocities (vxx, vyy, vzz)";
        return MBPOL_B200_EINVAL;
    }
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point t) { return std::chrono::duration<float, std::milli>(clk::now() - t).count(); };
    clk::time_point t0 = clk::now();
    if (int rc = mbpol_b200_upload(h, imcon, cell, nttm2, listttm2, mxatms, xxx, yyy, zzz, vxx, vyy, vzz)) return rc;
    // Caller arrays in pinned (page-locked) memory: fused end.  Their current values go up on the side stream now, the device adds
    // them after the reduction and copies the sums back inside the step -- the reference's "+=" without a host pass over 23 MB.
    h->fused.clear();
    {
        double* arr[6] = {fxx, fyy, fzz, atomic_energy, atomic_stress, atomic_fdecomp};
        const size_t n = h->mxatms;
        const size_t a0[6] = {0, n, 2 * n, 3 * n, 4 * n, 13 * n}, a1[6] = {n, 2 * n, 3 * n, 4 * n, 13 * n, 22 * n};
        bool pinned = fxx && fyy && fzz && !std::getenv("MBPOL_B200_NO_FUSED_DOWNLOAD");
        unsigned mask = 0;
        for (int k = 0; k < 6 && pinned; ++k)
            if (arr[k]) {
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, arr[k]) != cudaSuccess || at.type != cudaMemoryTypeHost) { cudaGetLastError(); pinned = false; }
                mask |= 1u << k;
            }
        if (pinned) {
            const bool slice = h->reduce_mode == 2 && h->comm && h->nranks > 1;
            const size_t lo = slice ? std::min((size_t)h->rank * h->chunk(), h->atom_count()) : 0;
            const size_t hi = slice ? std::min((size_t)(h->rank + 1) * h->chunk(), h->atom_count()) : h->atom_count();
            CUDA_TRY(h, h->acc.reserve(h->out_alloc()));
            if (h->pinned_cap < h->out_count()) {
                if (h->pinned) cudaFreeHost(h->pinned);
                h->pinned = nullptr; h->pinned_cap = 0;
                CUDA_TRY(h, cudaMallocHost(&h->pinned, h->out_count() * 8));
                h->pinned_cap = h->out_count();
            }
            if (h->acc_mask != mask) {      // arrays that are not requested stay zero in acc
                CUDA_TRY(h, cudaMemsetAsync(h->acc.p, 0, h->atom_count() * 8, h->vstream));
                h->acc_mask = mask;
            }
            for (int k = 0; k < 6; ++k) {
                if (!arr[k]) continue;
                const size_t x = std::max(a0[k], lo), y = std::min(a1[k], hi);
                if (x >= y) continue;
                h->fused.push_back({arr[k] + (x - a0[k]), x, y - x});
                CUDA_TRY(h, cudaMemcpyAsync(h->acc.p + x, arr[k] + (x - a0[k]), (y - x) * 8, cudaMemcpyHostToDevice, h->vstream));
            }
            CUDA_TRY(h, cudaEventRecord(h->acc_ev, h->vstream));
        }
    }
    h->ms[8] = ms_since(t0); t0 = clk::now();
    const int rcc = mbpol_b200_compute(h);
    const bool fused = !h->fused.empty();
    h->fused.clear();
    if (rcc) return rcc;
    h->ms[9] = ms_since(t0); t0 = clk::now();
    int rc = MBPOL_B200_OK;
    if (fused) {      // the per-atom arrays are already home; the scalars wait in the pinned block
        const double* P = h->pinned;
        const size_t so = h->scal_offset();
        const bool reduced = h->reduce_mode && h->comm && h->nranks > 1;
        for (int b = 0; b < h->nbeads; ++b) {
            const double* sc = P + so + (size_t)b * NSCAL;
            const bool mine = !reduced || h->rank == 0;
            if (mine) {
                if (eng2) eng2[b] += sc[S_ENG2];
                if (vir2) vir2[b] += sc[S_VIR2];
                if (eng3) eng3[b] += sc[S_ENG3];
                if (vir3) vir3[b] += sc[S_VIR3];
                if (stress) for (int q = 0; q < 9; ++q) stress[9 * b + q] += sc[S_STRESS + q];
            }
            if (jflux) for (int q = 0; q < 9; ++q) jflux[9 * b + q] = mine ? sc[S_JFLUX + q] : 0.0;
        }
    } else {
        rc = mbpol_b200_download(h, fxx, fyy, fzz, eng2, vir2, eng3, vir3, stress, atomic_energy, atomic_stress, atomic_fdecomp, jflux);
    }
    h->ms[10] = ms_since(t0);
    return rc;
}

int mbpol_b200_forces_batched(mbpol_b200_handle* h, int nbeads, int imcon, const double* cell, int nttm2, const int* listttm2, int mxatms,
                              const double* xxx, const double* yyy, const double* zzz, double* fxx, double* fyy, double* fzz, double* eng2,
                              double* vir2, double* eng3, double* vir3, double* stress, double* atomic_energy, double* atomic_stress,
                              double* atomic_fdecomp, const double* vxx, const double* vyy, const double* vzz, double* jflux)
{
    if (!h || nbeads != h->nbeads) return MBPOL_B200_EINVAL;
    return mbpol_b200_forces(h, imcon, cell, nttm2, listttm2, mxatms, xxx, yyy, zzz, fxx, fyy, fzz, eng2, vir2, eng3, vir3, stress,
                             atomic_energy, atomic_stress, atomic_fdecomp, vxx, vyy, vzz, jflux);
}

int mbpol_b200_device_outputs(mbpol_b200_handle* h, void** dev_ptr, long long* count)
{
    if (!h || !h->out.p) return MBPOL_B200_EINVAL;
    *dev_ptr = h->out.p;
    *count = (long long)h->out_count();
    return MBPOL_B200_OK;
}

int mbpol_b200_counts(mbpol_b200_handle* h, long long* ndt, long long* ntt, long long* nds, long long* nts)
{
    if (!h) return MBPOL_B200_EINVAL;
    if (ndt) *ndt = h->nd_total;
    if (ntt) *ntt = h->nt_total;
    if (nds) *nds = h->nd_shard;
    if (nts) *nts = h->nt_shard;
    return MBPOL_B200_OK;
}

int mbpol_b200_get_lists(mbpol_b200_handle* h, int* dimers, long long dimer_cap, int* trimers, long long trimer_cap)
{
    if (!h) return MBPOL_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (dimers && dimer_cap >= h->nd_shard && h->nd_shard > 0)
        CUDA_TRY(h, cudaMemcpy(dimers, h->dimers.p, (size_t)h->nd_shard * 8, cudaMemcpyDeviceToHost));
    if (trimers && trimer_cap >= h->nt_shard && h->nt_shard > 0)
        CUDA_TRY(h, cudaMemcpy(trimers, h->trimers.p, (size_t)h->nt_shard * 12, cudaMemcpyDeviceToHost));
    return MBPOL_B200_OK;
}

int mbpol_b200_kernel_times(mbpol_b200_handle* h, float* ms, long long* launches)
{
    if (!h) return MBPOL_B200_EINVAL;
    refresh_phase_times(h);
    if (ms) std::copy(h->ms, h->ms + 4, ms);
    if (launches) *launches = h->launches;
    return MBPOL_B200_OK;
}

int mbpol_b200_phase_times(mbpol_b200_handle* h, float* ms, int n)
{
    if (!h || !ms || n < 0) return MBPOL_B200_EINVAL;
    refresh_phase_times(h);
    for (int k = 0; k < n; ++k) ms[k] = k < 11 ? h->ms[k] : 0.f;
    return MBPOL_B200_OK;
}

int mbpol_b200_shard(mbpol_b200_handle* h, int* first_water, int* n_waters, long long* scalar_offset)
{
    if (!h) return MBPOL_B200_EINVAL;
    if (first_water) *first_water = h->w0;
    if (n_waters) *n_waters = h->nown;
    if (scalar_offset) *scalar_offset = (long long)h->scal_offset();
    return MBPOL_B200_OK;
}

int mbpol_b200_measure_fp64_peak(double* tflops, double* ms_out)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return MBPOL_B200_ENODEVICE;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * sms * 8 * 256) != cudaSuccess) return MBPOL_B200_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return MBPOL_B200_ECUDA;
    *tflops = 2.0 * 64.0 * iters * (double)sms * 8 * 256 / (best * 1e-3) * 1e-12;
    if (ms_out) *ms_out = best;
    return MBPOL_B200_OK;
}

// ---- batched per-cluster API -----------------------------------------------------------------------------
static int eval_clusters(int n, int nat3, const double* w, double* E, double* g, double* fm, int fmlen)
{
    if (n <= 0 || !w || !E || !g) return MBPOL_B200_EINVAL;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return MBPOL_B200_ENODEVICE;
    if (ensure_constants(device)) return MBPOL_B200_ECUDA;
    double *dw = nullptr, *dE = nullptr, *dg = nullptr, *dfm = nullptr;
    int rc = MBPOL_B200_OK;
    auto ok = [&](cudaError_t e) { if (e != cudaSuccess) rc = MBPOL_B200_ECUDA; return e == cudaSuccess; };
    if (ok(cudaMalloc(&dw, (size_t)n * nat3 * 8)) && ok(cudaMalloc(&dE, (size_t)n * 8)) && ok(cudaMalloc(&dg, (size_t)n * nat3 * 8)) &&
        (!fm || ok(cudaMalloc(&dfm, (size_t)n * fmlen * 8))) && ok(cudaMemcpy(dw, w, (size_t)n * nat3 * 8, cudaMemcpyHostToDevice))) {
        const int eng = default_engine(), eng2 = default_engine_2b();
        if (nat3 == 18) ok(eng2 == 3 ? allow_smem(k_eval_dimers<3>, SMEM2B) : allow_smem(k_eval_dimers<1>, SMEM2B));
        else ok(eng == 4 ? allow_smem(k_eval_trimers<4>, SMEM3S) : (eng == 3 ? allow_smem(k_eval_trimers<3>, SMEM3B) : allow_smem(k_eval_trimers<1>, SMEM3B)));
        if (nat3 == 18) {
            if (eng2 == 3) k_eval_dimers<3><<<(n + MBPOL_BS2 - 1) / MBPOL_BS2, MBPOL_BS2, SMEM2B>>>(n, dw, dE, dg, dfm, g_fat2b[device]);
            else k_eval_dimers<1><<<(n + MBPOL_BS2 - 1) / MBPOL_BS2, MBPOL_BS2, SMEM2B>>>(n, dw, dE, dg, dfm, g_tape2b[device]);
        } else {
            if (eng == 4) k_eval_trimers<4><<<(n + MBPOL_BS3 - 1) / MBPOL_BS3, MBPOL_BS3, SMEM3S>>>(n, dw, dE, dg, dfm, g_sym3b[device]);
            else if (eng == 3) k_eval_trimers<3><<<(n + MBPOL_BS3 - 1) / MBPOL_BS3, MBPOL_BS3, SMEM3B>>>(n, dw, dE, dg, dfm, g_fat3b[device]);
            else k_eval_trimers<1><<<(n + MBPOL_BS3 - 1) / MBPOL_BS3, MBPOL_BS3, SMEM3B>>>(n, dw, dE, dg, dfm, g_tape3b[device]);
        }
        ok(cudaGetLastError());
        ok(cudaDeviceSynchronize());
        ok(cudaMemcpy(E, dE, (size_t)n * 8, cudaMemcpyDeviceToHost));
        ok(cudaMemcpy(g, dg, (size_t)n * nat3 * 8, cudaMemcpyDeviceToHost));
        if (fm) ok(cudaMemcpy(fm, dfm, (size_t)n * fmlen * 8, cudaMemcpyDeviceToHost));
    }
    cudaFree(dw); cudaFree(dE); cudaFree(dg); cudaFree(dfm);
    return rc;
}

int mbpol_b200_eval_dimers(int n, const double* w, double* E, double* g, double* fm) { return eval_clusters(n, 18, w, E, g, fm, 108); }
int mbpol_b200_eval_trimers(int n, const double* w, double* E, double* g, double* fm) { return eval_clusters(n, 27, w, E, g, fm, 243); }

// ---- the reference's per-cluster symbols -----------------------------------------------------------------
void mbpol_2b_poly_(const double* w1, const double* w2, double* E, double* g1, double* g2, double* force_matrix)
{
    double w[18], g[18];
    std::copy(w1, w1 + 9, w); std::copy(w2, w2 + 9, w + 9);
    if (mbpol_b200_eval_dimers(1, w, E, g, force_matrix) != MBPOL_B200_OK) { *E = std::nan(""); return; }
    std::copy(g, g + 9, g1); std::copy(g + 9, g + 18, g2);      // assigned, like mbpol-2b-poly.cpp:660-663
}

void mbpol_3b_poly_(const double* w1, const double* w2, const double* w3, double* E, double* g1, double* g2, double* g3,
                    double* force_matrix)
{
    double w[27], g[27];
    std::copy(w1, w1 + 9, w); std::copy(w2, w2 + 9, w + 9); std::copy(w3, w3 + 9, w + 18);
    if (mbpol_b200_eval_trimers(1, w, E, g, force_matrix) != MBPOL_B200_OK) { *E = std::nan(""); return; }
    for (int i = 0; i < 9; ++i) { g1[i] += g[i]; g2[i] += g[9 + i]; g3[i] += g[18 + i]; }   // accumulated, mbpol-3b-poly.cpp:39-40
}

void mbpol_2b_cutoff_(double* r) { *r = MBPOL_P2b_r2f; }
void mbpol_3b_cutoff_(double* r) { *r = MBPOL_P3b_r3f; }

}  // extern "C"
