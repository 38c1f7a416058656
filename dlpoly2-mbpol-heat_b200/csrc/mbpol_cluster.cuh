// mbpol_cluster.cuh -- the per-cluster kernels (K2 dimers, K3 trimers, batched per-cluster evaluation) as templates on the
// polynomial engine.  Every (kernel, engine) pair is instantiated in its own translation unit (mbpol_inst.cu with
// -DMBPOL_INST=...) so that the heavy ptxas runs go in parallel; mbpol_kernels.cu only sees `extern template`s.
#pragma once
#include <cuda_runtime.h>

#include "mbpol_b200.h"
#include "mbpol_device.cuh"

extern __constant__ int c_gather_vp2[31];     // 2B: variable n with the two molecules swapped -> original variable
extern __constant__ int c_gather_vp[3][36];  // slot s: variable n of the relabelled frame (molecules 0 <-> s swapped) -> original variable

namespace mbpol_k {

constexpr int NB2_CAP = 96;    // neighbours j>i within r2f (6.5 A): ~19 on average in liquid water
constexpr int NB3_CAP = 40;    // neighbours within r3f (4.5 A): ~13 on average
constexpr int NSCAL = 32;
enum { S_ENG2 = 0, S_VIR2 = 1, S_ENG3 = 2, S_VIR3 = 3, S_STRESS = 4, S_JFLUX = 13, S_ND = 22, S_NT = 23, S_OVF = 24 };

struct OutPtrs {      // views into the packed output buffer
    double *fx, *fy, *fz, *aeng, *astress, *afd, *scal;
    int mxatms;       // all beads: nbeads * mxb
    int nwb, mxb;     // waters / atoms per bead (bead-batched calls; one bead otherwise); scalars of bead b at scal + b*NSCAL
};

// ---------------------------------------------------------------------------------------------------------
// block reduction of NV per-thread doubles -> atomicAdd into dst[0..NV)
// ---------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_add(double (&v)[NV], double* dst)
{
    __shared__ double red[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = lane < nwarp ? red[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) atomicAdd(dst + k, x);
        }
    }
}

// scalars are kept per bead: reduce over the block when all its clusters belong to one bead (always, except at the
// few blocks that straddle a bead boundary of a bead-batched call), else one atomic per thread
template <int NV>
__device__ __forceinline__ void block_reduce_add_bead(double (&v)[NV], double* scal, int slot, int bead /* < 0: idle lane */)
{
    __shared__ int bead0;
    if (threadIdx.x == 0) bead0 = bead;          // thread 0 of a launched block always owns a cluster
    __syncthreads();
    const int b0 = bead0;
    if (__syncthreads_and(bead < 0 || bead == b0)) {
        block_reduce_add<NV>(v, scal + (size_t)b0 * NSCAL + slot);
    } else if (bead >= 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) atomicAdd(scal + (size_t)bead * NSCAL + slot + k, v[k]);
    }
    __syncthreads();
}

// per-atom 3x3 tensors are stored bead-major, [nbeads][9][mxb]: component q of (0-based, all-bead) atom `atom` sits at
// base + tensor_base(atom) + mxb*q -- for one bead this is the Fortran layout atomic_stress(mxatms,3,3)
__device__ __forceinline__ size_t tensor_base(const OutPtrs& O, int atom) { return (size_t)atom + (size_t)(atom / O.mxb) * 8 * O.mxb; }

// per-atom force-decomposition tensor W (heatcurrent.f90:529-533, 703-728):
//   W_i(a,b) += (n*x_i - sum_j x_j)_a^raw * f_b ,  f = -g_i*engunit/n
__device__ __forceinline__ void scatter_fdecomp(const OutPtrs& O, int atom, const double* xraw_i, const double* xsum, int n,
                                                const double* f)
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double ra = n * xraw_i[a] - xsum[a];
#pragma unroll
        for (int b = 0; b < 3; ++b) atomicAdd(O.afd + tensor_base(O, atom) + (size_t)O.mxb * (a + 3 * b), ra * f[b]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: trimers.  One thread per accepted trimer of this shard.
// ---------------------------------------------------------------------------------------------------------
// imaging exactly as do_trimer (mbpol.f90:843-882): waters i and k next to water j; origin moved to O_j
__device__ __forceinline__ double ld_again(const double* p)     // a load the compiler cannot merge with an earlier one
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <bool AGAIN = false>
__device__ __forceinline__ void trimer_geometry(int nw, const double* __restrict__ mono, const CellInfo& C, int mi, int mj, int mk, double (&r)[27])
{
    double wj[9], wi[9], wk[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        wi[c] = AGAIN ? ld_again(mono + (size_t)c * nw + mi) : mono[(size_t)c * nw + mi];
        wj[c] = AGAIN ? ld_again(mono + (size_t)c * nw + mj) : mono[(size_t)c * nw + mj];
        wk[c] = AGAIN ? ld_again(mono + (size_t)c * nw + mk) : mono[(size_t)c * nw + mk];
    }
    double a0 = wi[0] - wj[0], a1 = wi[1] - wj[1], a2 = wi[2] - wj[2];
    double b0 = wj[0] - wk[0], b1 = wj[1] - wk[1], b2 = wj[2] - wk[2];
    image_vec(C, a0, a1, a2);
    image_vec(C, b0, b1, b2);
    const double ra[3] = {a0, a1, a2}, rb[3] = {b0, b1, b2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = ra[c];                               // O_i' - O_j
        r[3 + c] = (wi[3 + c] - wi[c]) + ra[c];
        r[6 + c] = (wi[6 + c] - wi[c]) + ra[c];
        r[9 + c] = 0.0;
        r[12 + c] = wj[3 + c] - wj[c];
        r[15 + c] = wj[6 + c] - wj[c];
        r[18 + c] = -rb[c];                         // O_k' - O_j
        r[21 + c] = (wk[3 + c] - wk[c]) - rb[c];
        r[24 + c] = (wk[6 + c] - wk[c]) - rb[c];
    }
}

// water-major copy (k_gather_image): nine coordinates + the raw coordinate sum of the water, six 16-byte loads
__device__ __forceinline__ void load_water_aos(const double* __restrict__ monoA, int m, double (&w)[9], double (&X)[3])
{
    const double2* A = reinterpret_cast<const double2*>(monoA + (size_t)m * 12);
    const double2 a0 = __ldg(A), a1 = __ldg(A + 1), a2 = __ldg(A + 2), a3 = __ldg(A + 3), a4 = __ldg(A + 4), a5 = __ldg(A + 5);
    w[0] = a0.x; w[1] = a0.y; w[2] = a1.x; w[3] = a1.y; w[4] = a2.x; w[5] = a2.y; w[6] = a3.x; w[7] = a3.y; w[8] = a4.x;
    X[0] += a4.y; X[1] += a5.x; X[2] += a5.y;
}
// trimer_geometry from the water-major copy; X += raw coordinate sum of the nine atoms
__device__ __forceinline__ void trimer_geometry_aos(const double* __restrict__ monoA, const CellInfo& C, int mi, int mj, int mk, double (&r)[27], double (&X)[3])
{
    double wj[9], wi[9], wk[9];
    load_water_aos(monoA, mi, wi, X);
    load_water_aos(monoA, mj, wj, X);
    load_water_aos(monoA, mk, wk, X);
    double a0 = wi[0] - wj[0], a1 = wi[1] - wj[1], a2 = wi[2] - wj[2];
    double b0 = wj[0] - wk[0], b1 = wj[1] - wk[1], b2 = wj[2] - wk[2];
    image_vec(C, a0, a1, a2);
    image_vec(C, b0, b1, b2);
    const double ra[3] = {a0, a1, a2}, rb[3] = {b0, b1, b2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = ra[c];
        r[3 + c] = (wi[3 + c] - wi[c]) + ra[c];
        r[6 + c] = (wi[6 + c] - wi[c]) + ra[c];
        r[9 + c] = 0.0;
        r[12 + c] = wj[3 + c] - wj[c];
        r[15 + c] = wj[6 + c] - wj[c];
        r[18 + c] = -rb[c];
        r[21 + c] = (wk[3 + c] - wk[c]) - rb[c];
        r[24 + c] = (wk[6 + c] - wk[c]) - rb[c];
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3 phase 1: one thread per accepted trimer of this shard -- geometry, variables, switch, polynomial (the tape walk) and
// a 40-double RECORD per trimer (trimer_record); nothing is scattered here.  The records of a CTA are staged through the
// (now idle) column area of the shared memory and leave as coalesced 16-byte stores.
// ---------------------------------------------------------------------------------------------------------
#define MBPOL_REC3_STRIDE 42      /* doubles between two staged records: 336 B, 16-byte aligned, spreads the banks */

template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, (256 / TAPE_BS) > 0 ? (256 / TAPE_BS) : 1)
k_trimers(const long long* __restrict__ pcount, const int* __restrict__ tri, int nw, const double* __restrict__ mono, CellInfo C,
          const double* __restrict__ tape, double* __restrict__ rec)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const long long ntri = *pcount;          // written on the device by the list build: no host round trip before the launch
    const PolyCtx ctx{dyn_smem, tape};
    for (long long first = (long long)blockIdx.x * blockDim.x; first < ntri; first += (long long)gridDim.x * blockDim.x) {
    const long long t0_ = first + threadIdx.x;
    const long long t = t0_ < ntri ? t0_ : ntri - 1;     // idle lanes shadow the last trimer: the tape walk has CTA barriers
    double out[MBPOL_REC3];
    {
        const int mi = tri[3 * t], mj = tri[3 * t + 1], mk = tri[3 * t + 2];
        double r[27], xv[36], gp[36], sw[4], Epoly;
        trimer_geometry(nw, mono, C, mi, mj, mk, r);
#ifndef MBPOL_UNROLLED_PROLOGUE
        if constexpr (ENGINE == 4) {
            SymSmem& sm = *reinterpret_cast<SymSmem*>(dyn_smem);
            sym_stage_x(r, sw, sm);
            sym_eval3b<true>(tape, xv, gp, Epoly, sm);
        } else
#endif
        {
            trimer_x_switch(r, xv, sw);
            poly3b_eval<ENGINE>(xv, gp, Epoly, ctx);
        }
        trimer_record(xv, gp, Epoly, sw, out);
    }
    __syncthreads();      // every thread has read its last gradient column: the column area is free
    double* stage = ENGINE == 4 ? reinterpret_cast<SymSmem*>(dyn_smem)->xs[0] : reinterpret_cast<TapeSmem<36>*>(dyn_smem)->xs[0];
    static_assert(sizeof(double) * MBPOL_REC3_STRIDE * TAPE_BS <= 2 * sizeof(double) * 37 * TAPE_BS, "staging area");
#pragma unroll
    for (int n = 0; n < MBPOL_REC3; ++n) stage[threadIdx.x * MBPOL_REC3_STRIDE + n] = out[n];
    __syncthreads();
    const long long left = ntri - first;
    const int nvalid = left < (long long)blockDim.x ? (int)left : (int)blockDim.x;
    double2* dst = reinterpret_cast<double2*>(rec + (size_t)first * MBPOL_REC3);
    for (int i = threadIdx.x; i < nvalid * (MBPOL_REC3 / 2); i += blockDim.x) {
        const int tr = i / (MBPOL_REC3 / 2), pr = i % (MBPOL_REC3 / 2);
        dst[i] = *reinterpret_cast<const double2*>(stage + tr * MBPOL_REC3_STRIDE + 2 * pr);
    }
    __syncthreads();      // the staged records are out before the next tile overwrites the columns
    }
}

#ifdef MBPOL_GATHER_KERNELS      /* non-template kernels: compiled in mbpol_kernels.cu only */
// ---------------------------------------------------------------------------------------------------------
// K3 phase 2: GATHER.  One warp per water w; the lanes walk the incidence list of w (every trimer of this shard that contains
// w, with w's position in the sorted triple), rebuild the trimer's geometry exactly as phase 1 did, relabel the molecules
// so that w comes first, turn the record into the force coefficients of the 21 variables that touch an atom of w and
// accumulate w's share of everything do_trimer adds up (mbpol.f90:896-1267): forces, E/9 per atom, the force-decomposition
// tensor W, the per-atom stress S, and the virial / stress / energy scalars.  One warp reduction per WATER, plain stores:
// no FP64 atomics on per-atom data at all.
// ---------------------------------------------------------------------------------------------------------
__device__ constexpr int G3N = 21;          // variables with an atom of molecule 0: HH_a, OH_a (2), Ha-H (8), Oa-H (4), O-Ha (4), Oa-O (2)
__device__ constexpr int G3VAR[G3N] = {0, 3, 4, 9, 10, 11, 12, 13, 14, 15, 16, 21, 22, 23, 24, 25, 26, 29, 30, 33, 34};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifndef MBPOL_GATHER_MINBLOCKS
#define MBPOL_GATHER_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128, MBPOL_GATHER_MINBLOCKS)
k_trimers_gather(int nw, const int* __restrict__ inc_off, const unsigned* __restrict__ inc, const int* __restrict__ tri,
                 const double* __restrict__ rec, const double* __restrict__ mono, const int* __restrict__ list,
                 const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr, CellInfo C, OutPtrs O, int flags)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool active = w < nw;
    const int beg = active ? inc_off[w] : 0, end = active ? inc_off[w + 1] : 0;
    if (!__syncthreads_or(end > beg)) return;      // no water of this block occurs in the shard's lists (most blocks of a multi-GPU run)
    const bool want_e = flags & MBPOL_B200_WANT_ATOMIC_ENERGY, want_w = flags & MBPOL_B200_WANT_FDECOMP, want_s = flags & MBPOL_B200_WANT_ATOMIC_STRESS;
    const double eu = MBPOL_ENGUNIT;
    double G[9], A[27], S[18], sc[8], esum = 0.0;     // G = sum g (kcal/mol/A); A = sum X_c (x) g; S; scalars E, vir, stress(6)
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 27; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 18; ++i) S[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i] = 0.0;
    // the incidence code and the triple of the NEXT incidence of this lane are fetched while the current one is evaluated: one level
    // of dependent loads (coordinates, record) per iteration instead of three
    unsigned ncode = 0;
    int nm[3] = {0, 0, 0};
    if (beg + lane < end) {
        ncode = inc[beg + lane];
        const int t = (int)(ncode >> 2);
        nm[0] = tri[3 * t]; nm[1] = tri[3 * t + 1]; nm[2] = tri[3 * t + 2];
    }
    for (int i = beg + lane; i < end; i += 32) {
        const unsigned code = ncode;
        const int t = (int)(code >> 2), s = (int)(code & 3u);
        const int m3[3] = {nm[0], nm[1], nm[2]};
        if (i + 32 < end) {
            ncode = inc[i + 32];
            const int tn = (int)(ncode >> 2);
            nm[0] = tri[3 * tn]; nm[1] = tri[3 * tn + 1]; nm[2] = tri[3 * tn + 2];
        }
        double r0[27], r[27], X[3] = {0.0, 0.0, 0.0};
        trimer_geometry_aos(mono, C, m3[0], m3[1], m3[2], r0, X);
        // relabel: molecule 0 <-> molecule s
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const double v0 = r0[c], vm = r0[9 * m + c], vs = (s == 1) ? r0[9 + c] : ((s == 2) ? r0[18 + c] : r0[c]);
                r[9 * m + c] = (m == 0) ? vs : ((m == s) ? v0 : vm);
            }
        const double* R = rec + (size_t)t * MBPOL_REC3;
        const int* vp = c_gather_vp[s];
        double g[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) g[q] = 0.0;
#pragma unroll
        for (int k = 0; k < G3N; ++k) {
            const int n = G3VAR[k], a = V3A[n], b = V3B[n];
            const double d0 = r[3 * a] - r[3 * b], d1 = r[3 * a + 1] - r[3 * b + 1], d2 = r[3 * a + 2] - r[3 * b + 2];
            const double rinv = rsqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const int no = vp[n];
            double cv = R[no];
            if (n >= 33) cv += R[37 + no - 33];
            cv *= rinv;
            const double f0 = cv * d0, f1 = cv * d1, f2 = cv * d2;
            if (a < 3) { g[3 * a] += f0; g[3 * a + 1] += f1; g[3 * a + 2] += f2; }
            if (b < 3) { g[3 * b] -= f0; g[3 * b + 1] -= f1; g[3 * b + 2] -= f2; }
            if (want_s) {   // S_i(a,b) += 1/2*eu*cv_n d (x) d for BOTH atoms of variable n (mbpol.f90:945-966)
                const double h = 0.5 * eu * cv;
                const double T[6] = {h * d0 * d0, h * d0 * d1, h * d0 * d2, h * d1 * d1, h * d1 * d2, h * d2 * d2};
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    if (a < 3) S[6 * a + q] += T[q];
                    if (b < 3) S[6 * b + q] += T[q];
                }
            }
        }
        const double E = R[36];
        if (s == 0) sc[0] += E * eu;                   // every trimer has exactly one slot-0 incidence
        esum += E;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double gx = g[3 * a] * eu, gy = g[3 * a + 1] * eu, gz = g[3 * a + 2] * eu;
            G[3 * a] += g[3 * a]; G[3 * a + 1] += g[3 * a + 1]; G[3 * a + 2] += g[3 * a + 2];
            sc[1] += r[3 * a] * gx + r[3 * a + 1] * gy + r[3 * a + 2] * gz;
            sc[2] -= r[3 * a] * gx; sc[3] -= r[3 * a] * gy; sc[4] -= r[3 * a] * gz;
            sc[5] -= r[3 * a + 1] * gy; sc[6] -= r[3 * a + 1] * gz; sc[7] -= r[3 * a + 2] * gz;
        }
        if (want_w) {   // X_c = sum of the RAW coordinates of the 9 atoms of the cluster (carried by the water-major copy)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) A[9 * a + i + 3 * j] += X[i] * g[3 * a + j];
        }
    }
    // one reduction per water
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = warp_sum(G[i]);
    esum = warp_sum(esum);
    if (want_w) {
#pragma unroll
        for (int i = 0; i < 27; ++i) A[i] = warp_sum(A[i]);
    }
    if (want_s) {
#pragma unroll
        for (int i = 0; i < 18; ++i) S[i] = warp_sum(S[i]);
    }
    if (active && lane == 0 && end > beg) {
        const int o = list[3 * w] - 1;                  // 0-based O index; H at o+1, o+2 (mbpol.f90:592-594)
        const size_t n = O.mxb;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int atom = o + a;
            O.fx[atom] -= G[3 * a] * eu; O.fy[atom] -= G[3 * a + 1] * eu; O.fz[atom] -= G[3 * a + 2] * eu;
            if (want_e) O.aeng[atom] += (1.0 / 9.0) * esum * eu;
            if (want_w) {   // W_i = sum_c (9 x_i - X_c) (x) (-g_i eu / 9) = -x_i (x) G_i eu + A_i eu / 9
                const double x3[3] = {xr[atom], yr[atom], zr[atom]};
                double* p = O.afd + tensor_base(O, atom);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) p[n * (i + 3 * j)] += (1.0 / 9.0) * eu * A[9 * a + i + 3 * j] - x3[i] * G[3 * a + j] * eu;
            }
            if (want_s) {
                double* p = O.astress + tensor_base(O, atom);
                const double* q = S + 6 * a;
                p[n * 0] += q[0]; p[n * 1] += q[1]; p[n * 2] += q[2];
                p[n * 3] += q[1]; p[n * 4] += q[3]; p[n * 5] += q[4];
                p[n * 6] += q[2]; p[n * 7] += q[4]; p[n * 8] += q[5];
            }
        }
    }
    // scalars: E, virial, stress of this shard (per bead)
    const int bead = active ? w / O.nwb : -1;          // (thread 0 of every launched block owns a water)
    double v2[2] = {sc[0], sc[1]};
    block_reduce_add_bead<2>(v2, O.scal, S_ENG3, bead);
    double st[9] = {sc[2], sc[3], sc[4], sc[3], sc[5], sc[6], sc[4], sc[6], sc[7]};
    block_reduce_add_bead<9>(st, O.scal, S_STRESS, bead);
}

// incidence lists: water -> (trimer << 2 | slot) for the trimers [0, ntri) of this shard.  Consecutive list entries share their
// centre water (the enumeration emits a water's clusters together), so the lanes of a warp that hit the same water are
// aggregated first (__match_any_sync): one atomic per distinct water and warp instead of one per entry.
template <int M>      // M = 3: trimers (int[3]), M = 2: dimers (int2)
__global__ void k_inc_count(const long long* __restrict__ pcount, const int* __restrict__ tri, int* __restrict__ cnt)
{
    const long long n3 = M * *pcount;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n3; i0 += stride) {      // (block-uniform trip count)
        const long long i = i0 + threadIdx.x;
        const bool live = i < n3;
        const int w = live ? tri[i] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, w);
        if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[w], __popc(peers));
    }
}
template <int M>
__global__ void k_inc_fill(const long long* __restrict__ pcount, const int* __restrict__ tri, const int* __restrict__ off, int* __restrict__ cur,
                           unsigned* __restrict__ inc)
{
    const long long n3 = M * *pcount;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n3; i0 += stride) {
        const long long i = i0 + threadIdx.x;
        const bool live = i < n3;
        const int w = live ? tri[i] : -1;
        const unsigned lane = threadIdx.x & 31;
        const unsigned peers = __match_any_sync(0xffffffffu, w);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (live && lane == (unsigned)leader) base = atomicAdd(&cur[w], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live)
            inc[off[w] + base + __popc(peers & ((1u << lane) - 1u))] =
                M == 3 ? ((unsigned)((i / 3) << 2) | (unsigned)(i % 3)) : ((unsigned)((i / 2) << 1) | (unsigned)(i % 2));
    }
}

#endif   // MBPOL_GATHER_KERNELS

// ---------------------------------------------------------------------------------------------------------
// K2 phase 1: one thread per accepted dimer of this shard -- geometry (do_dimer's imaging, mbpol.f90:496-509; origin moved to
// O_j), lone pairs, 31 variables, polynomial, switch -> a 34-double record (dimer_record).  Staged and stored like the 3B records.
// ---------------------------------------------------------------------------------------------------------
#define MBPOL_REC2_STRIDE 36

__device__ __forceinline__ void dimer_geometry(int nw, const double* __restrict__ mono, const CellInfo& C, int mi, int mj, double (&w)[18])
{
    double wi[9], wj[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        wi[c] = mono[(size_t)c * nw + mi];
        wj[c] = mono[(size_t)c * nw + mj];
    }
    double a0 = wi[0] - wj[0], a1 = wi[1] - wj[1], a2 = wi[2] - wj[2];
    image_vec(C, a0, a1, a2);
    const double ra[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        w[c] = ra[c];
        w[3 + c] = (wi[3 + c] - wi[c]) + ra[c];
        w[6 + c] = (wi[6 + c] - wi[c]) + ra[c];
        w[9 + c] = 0.0;
        w[12 + c] = wj[3 + c] - wj[c];
        w[15 + c] = wj[6 + c] - wj[c];
    }
}

// dimer_geometry from the water-major copy (k_gather_image); X += raw coordinate sum of the six atoms
__device__ __forceinline__ void dimer_geometry_aos(const double* __restrict__ monoA, const CellInfo& C, int mi, int mj, double (&w)[18], double (&X)[3])
{
    double wi[9], wj[9];
    load_water_aos(monoA, mi, wi, X);
    load_water_aos(monoA, mj, wj, X);
    double a0 = wi[0] - wj[0], a1 = wi[1] - wj[1], a2 = wi[2] - wj[2];
    image_vec(C, a0, a1, a2);
    const double ra[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        w[c] = ra[c];
        w[3 + c] = (wi[3 + c] - wi[c]) + ra[c];
        w[6 + c] = (wi[6 + c] - wi[c]) + ra[c];
        w[9 + c] = 0.0;
        w[12 + c] = wj[3 + c] - wj[c];
        w[15 + c] = wj[6 + c] - wj[c];
    }
}

template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, (256 / TAPE_BS) > 0 ? (256 / TAPE_BS) : 1)
k_dimers(const long long* __restrict__ pcount, const int2* __restrict__ dim, int nw, const double* __restrict__ mono, CellInfo C,
         const double* __restrict__ tape, double* __restrict__ rec)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const long long ndim = *pcount;
    const PolyCtx ctx{dyn_smem, tape};
    for (long long first = (long long)blockIdx.x * blockDim.x; first < ndim; first += (long long)gridDim.x * blockDim.x) {
        const long long t0_ = first + threadIdx.x;
        const long long t = t0_ < ndim ? t0_ : ndim - 1;     // idle lanes shadow the last dimer: the tape walk has CTA barriers
        double out[MBPOL_REC2];
        {
            double w[18];
            dimer_geometry(nw, mono, C, dim[t].x, dim[t].y, w);
            dimer_record<ENGINE>(w, out, ctx);
        }
        __syncthreads();      // every thread has read its last gradient column: the column area is free
        double* stage = reinterpret_cast<TapeSmem<31>*>(dyn_smem)->xs[0];
        static_assert(MBPOL_REC2_STRIDE <= 2 * 32, "staging area");
#pragma unroll
        for (int n = 0; n < MBPOL_REC2; ++n) stage[threadIdx.x * MBPOL_REC2_STRIDE + n] = out[n];
        __syncthreads();
        const long long left = ndim - first;
        const int nvalid = left < (long long)blockDim.x ? (int)left : (int)blockDim.x;
        double2* dst = reinterpret_cast<double2*>(rec + (size_t)first * MBPOL_REC2);
        for (int i = threadIdx.x; i < nvalid * (MBPOL_REC2 / 2); i += blockDim.x) {
            const int tr = i / (MBPOL_REC2 / 2), pr = i % (MBPOL_REC2 / 2);
            dst[i] = *reinterpret_cast<const double2*>(stage + tr * MBPOL_REC2_STRIDE + 2 * pr);
        }
        __syncthreads();      // the staged records are out before the next tile overwrites the columns
    }
}

#ifdef MBPOL_GATHER_KERNELS
// ---------------------------------------------------------------------------------------------------------
// K2 phase 2: GATHER, one warp per water w over its dimer incidences (dimer << 1 | slot).  The dimer is rebuilt as phase 1 saw it,
// the two molecules are swapped when w is the second one, and the rows of the folded pair matrix that belong to w's atoms
// (dimer_rows_a) give everything do_dimer accumulates for them (mbpol.f90:519-654, 732-786): force = -row sum, E/6 per atom, W,
// the (non-symmetric) per-atom stress S_i(a,b) = 1/2 eu sum_j (r_i - r_j)_a FM[i][j]_b, virial / stress / energy scalars.
// ---------------------------------------------------------------------------------------------------------
#ifndef MBPOL_GATHER2_MINBLOCKS
#define MBPOL_GATHER2_MINBLOCKS 3      /* 2 / 3 / 4 blocks per SM: 2B phase 3.02 / 2.95 / 3.04 ms */
#endif
__global__ void __launch_bounds__(128, MBPOL_GATHER2_MINBLOCKS)
k_dimers_gather(int nw, const int* __restrict__ inc_off, const unsigned* __restrict__ inc, const int2* __restrict__ dim,
                const double* __restrict__ rec, const double* __restrict__ mono, const int* __restrict__ list,
                const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr, CellInfo C, OutPtrs O, int flags)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool active = w < nw;
    const int beg = active ? inc_off[w] : 0, end = active ? inc_off[w + 1] : 0;
    if (!__syncthreads_or(end > beg)) return;      // no water of this block occurs in the shard's lists (most blocks of a multi-GPU run)
    const bool want_e = flags & MBPOL_B200_WANT_ATOMIC_ENERGY, want_w = flags & MBPOL_B200_WANT_FDECOMP, want_s = flags & MBPOL_B200_WANT_ATOMIC_STRESS;
    const double eu = MBPOL_ENGUNIT;
    double G[9], A[27], S[27], sc[8], esum = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 27; ++i) { A[i] = 0.0; S[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i] = 0.0;
    unsigned ncode = 0;      // next incidence of this lane, fetched one iteration ahead (see k_trimers_gather)
    int2 nm = make_int2(0, 0);
    if (beg + lane < end) {
        ncode = inc[beg + lane];
        nm = dim[ncode >> 1];
    }
    for (int i = beg + lane; i < end; i += 32) {
        const unsigned code = ncode;
        const int t = (int)(code >> 1), s = (int)(code & 1u);
        const int2 m2 = nm;
        if (i + 32 < end) {
            ncode = inc[i + 32];
            nm = dim[ncode >> 1];
        }
        double w0[18], wv[18], X[3] = {0.0, 0.0, 0.0};
        dimer_geometry_aos(mono, C, m2.x, m2.y, w0, X);
#pragma unroll
        for (int c = 0; c < 9; ++c) {       // relabel: the gathering water becomes molecule a
            wv[c] = s ? w0[9 + c] : w0[c];
            wv[9 + c] = s ? w0[c] : w0[9 + c];
        }
        double xc[30];
        dimer_sites(wv, xc);
        const double* R = rec + (size_t)t * MBPOL_REC2;
        double c[31];
#pragma unroll
        for (int n = 0; n < 31; ++n) c[n] = R[s ? c_gather_vp2[n] : n];
        double fm[3][6][3];
        dimer_rows_a(xc, c, R[32], fm);
        const double E = R[31];
        if (s == 0) sc[0] += E * eu;
        esum += E;
        double g[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    if (j != a) v += fm[a][j][k];
                g[3 * a + k] = v;
                G[3 * a + k] += v;
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double gx = g[3 * a] * eu, gy = g[3 * a + 1] * eu, gz = g[3 * a + 2] * eu;
            sc[1] += wv[3 * a] * gx + wv[3 * a + 1] * gy + wv[3 * a + 2] * gz;
            sc[2] -= wv[3 * a] * gx; sc[3] -= wv[3 * a] * gy; sc[4] -= wv[3 * a] * gz;
            sc[5] -= wv[3 * a + 1] * gy; sc[6] -= wv[3 * a + 1] * gz; sc[7] -= wv[3 * a + 2] * gz;
        }
        if (want_s) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    if (j == a) continue;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const double ra = wv[3 * a + p] - wv[3 * j + p];
#pragma unroll
                        for (int q = 0; q < 3; ++q) S[9 * a + p + 3 * q] += ra * fm[a][j][q];
                    }
                }
        }
        if (want_w) {      // X_c: raw coordinate sum of the six atoms (carried by the water-major copy)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int q = 0; q < 3; ++q) A[9 * a + p + 3 * q] += X[p] * g[3 * a + q];
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = warp_sum(G[i]);
    esum = warp_sum(esum);
    if (want_w) {
#pragma unroll
        for (int i = 0; i < 27; ++i) A[i] = warp_sum(A[i]);
    }
    if (want_s) {
#pragma unroll
        for (int i = 0; i < 27; ++i) S[i] = warp_sum(S[i]);
    }
    if (active && lane == 0 && end > beg) {
        const int o = list[3 * w] - 1;
        const size_t n = O.mxb;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int atom = o + a;
            O.fx[atom] -= G[3 * a] * eu; O.fy[atom] -= G[3 * a + 1] * eu; O.fz[atom] -= G[3 * a + 2] * eu;
            if (want_e) O.aeng[atom] += (1.0 / 6.0) * esum * eu;
            if (want_w) {   // W_i = sum_c (6 x_i - X_c) (x) (-g_i eu / 6)
                const double x3[3] = {xr[atom], yr[atom], zr[atom]};
                double* p = O.afd + tensor_base(O, atom);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) p[n * (i + 3 * j)] += (1.0 / 6.0) * eu * A[9 * a + i + 3 * j] - x3[i] * G[3 * a + j] * eu;
            }
            if (want_s) {
                double* p = O.astress + tensor_base(O, atom);
#pragma unroll
                for (int q = 0; q < 9; ++q) p[n * q] += 0.5 * eu * S[9 * a + q];
            }
        }
    }
    const int bead = active ? w / O.nwb : -1;
    double v2[2] = {sc[0], sc[1]};
    block_reduce_add_bead<2>(v2, O.scal, S_ENG2, bead);
    double st[9] = {sc[2], sc[3], sc[4], sc[3], sc[5], sc[6], sc[4], sc[6], sc[7]};
    block_reduce_add_bead<9>(st, O.scal, S_STRESS, bead);
}
#endif   // MBPOL_GATHER_KERNELS

// ---------------------------------------------------------------------------------------------------------
// batched per-cluster kernels (legacy ABI / KAT tests)
// ---------------------------------------------------------------------------------------------------------
template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, (256 / TAPE_BS) > 0 ? (256 / TAPE_BS) : 1) k_eval_dimers(int n, const double* __restrict__ w, double* E, double* g, double* fm, const double* __restrict__ tape)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const PolyCtx ctx{dyn_smem, tape};
    const int t0_ = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = min(t0_, n - 1);
    double ww[18], gg[18], e, f[108];
    for (int i = 0; i < 18; ++i) ww[i] = w[(size_t)18 * t + i];
    dimer_eval<ENGINE>(ww, e, gg, fm ? f : nullptr, ctx);
    if (t0_ >= n) return;
    E[t] = e;
    for (int i = 0; i < 18; ++i) g[(size_t)18 * t + i] = gg[i];
    if (fm)
        for (int i = 0; i < 108; ++i) fm[(size_t)108 * t + i] = f[i];
}

template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, (256 / TAPE_BS) > 0 ? (256 / TAPE_BS) : 1) k_eval_trimers(int n, const double* __restrict__ w, double* E, double* g, double* fm, const double* __restrict__ tape)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const PolyCtx ctx{dyn_smem, tape};
    const int t0_ = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = min(t0_, n - 1);
    double r[27], cv[36], gg[27], e;
#pragma unroll
    for (int i = 0; i < 27; ++i) r[i] = w[(size_t)27 * t + i];
    trimer_eval<ENGINE>(r, e, cv, ctx);
    if (t0_ >= n) return;
    trimer_gradient(r, cv, gg);
    E[t] = e;
#pragma unroll
    for (int i = 0; i < 27; ++i) g[(size_t)27 * t + i] = gg[i];
    if (fm) {
        double* f = fm + (size_t)243 * t;
        for (int i = 0; i < 243; ++i) f[i] = 0.0;
#pragma unroll
        for (int n_ = 0; n_ < 36; ++n_) {
            const int a = V3A[n_], b = V3B[n_];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double d = cv[n_] * (r[3 * a + c] - r[3 * b + c]);
                f[27 * a + 3 * b + c] = d;
                f[27 * b + 3 * a + c] = -d;
            }
        }
    }
}


}  // namespace mbpol_k
