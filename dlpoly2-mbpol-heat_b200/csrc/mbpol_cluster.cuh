// mbpol_cluster.cuh -- the per-cluster kernels (K2 dimers, K3 trimers, batched per-cluster evaluation) as templates on the
// polynomial engine.  Every (kernel, engine) pair is instantiated in its own translation unit (mbpol_inst.cu with
// -DMBPOL_INST=...) so that the heavy ptxas runs go in parallel; mbpol_kernels.cu only sees `extern template`s.
#pragma once
#include <cuda_runtime.h>

#include "mbpol_b200.h"
#include "mbpol_device.cuh"

namespace mbpol_k {

constexpr int NB2_CAP = 96;    // neighbours j>i within r2f (6.5 A): ~19 on average in liquid water
constexpr int NB3_CAP = 40;    // neighbours within r3f (4.5 A): ~13 on average
constexpr int NSCAL = 32;
enum { S_ENG2 = 0, S_VIR2 = 1, S_ENG3 = 2, S_VIR3 = 3, S_STRESS = 4, S_JFLUX = 13, S_ND = 22, S_NT = 23 };

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

// accumulation of one evaluated trimer (mbpol.f90:896-1267): forces, per-atom energy, W, S; returns the block scalars
__device__ __forceinline__ void trimer_scatter(const double (&r)[27], double E, const double (&cv)[36], const int (&o)[3],
                                               const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr,
                                               const OutPtrs& O, int flags, double (&sc)[8])
{
    double g[27];
    trimer_gradient(r, cv, g);
    const double eu = MBPOL_ENGUNIT;
    sc[0] = E * eu;
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        const int atom = o[a / 3] - 1 + (a % 3);
        const double gx = g[3 * a] * eu, gy = g[3 * a + 1] * eu, gz = g[3 * a + 2] * eu;
        atomicAdd(O.fx + atom, -gx);
        atomicAdd(O.fy + atom, -gy);
        atomicAdd(O.fz + atom, -gz);
        sc[1] += r[3 * a] * gx + r[3 * a + 1] * gy + r[3 * a + 2] * gz;
        sc[2] -= r[3 * a] * gx; sc[3] -= r[3 * a] * gy; sc[4] -= r[3 * a] * gz;
        sc[5] -= r[3 * a + 1] * gy; sc[6] -= r[3 * a + 1] * gz; sc[7] -= r[3 * a + 2] * gz;
    }
    if (flags & MBPOL_B200_WANT_ATOMIC_ENERGY) {
        const double e9 = (1.0 / 9.0) * E * eu;
#pragma unroll
        for (int a = 0; a < 9; ++a) atomicAdd(O.aeng + o[a / 3] - 1 + (a % 3), e9);
    }
    if (flags & MBPOL_B200_WANT_FDECOMP) {
        double xs[3] = {0, 0, 0}, xraw[27];
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const int atom = o[a / 3] - 1 + (a % 3);
            xraw[3 * a] = xr[atom]; xraw[3 * a + 1] = yr[atom]; xraw[3 * a + 2] = zr[atom];
            xs[0] += xraw[3 * a]; xs[1] += xraw[3 * a + 1]; xs[2] += xraw[3 * a + 2];
        }
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const double f[3] = {-(1.0 / 9.0) * g[3 * a] * eu, -(1.0 / 9.0) * g[3 * a + 1] * eu, -(1.0 / 9.0) * g[3 * a + 2] * eu};
            scatter_fdecomp(O, o[a / 3] - 1 + (a % 3), xraw + 3 * a, xs, 9, f);
        }
    }
    if (flags & MBPOL_B200_WANT_ATOMIC_STRESS) {
        // S_i(a,b) = 1/2*eu*sum_{j != i} (r_i-r_j)_a FM[i][j]_b with FM[i][j] = cv_n (r_i - r_j): a symmetric
        // tensor cv_n d (x) d that goes to BOTH atoms of variable n (mbpol.f90:945-966)
        double S[9][6];
#pragma unroll
        for (int a = 0; a < 9; ++a)
#pragma unroll
            for (int q = 0; q < 6; ++q) S[a][q] = 0.0;
#pragma unroll
        for (int n = 0; n < 36; ++n) {
            const int a = V3A[n], b = V3B[n];
            const double d0 = r[3 * a] - r[3 * b], d1 = r[3 * a + 1] - r[3 * b + 1], d2 = r[3 * a + 2] - r[3 * b + 2];
            const double h = 0.5 * eu * cv[n];
            const double T[6] = {h * d0 * d0, h * d0 * d1, h * d0 * d2, h * d1 * d1, h * d1 * d2, h * d2 * d2};
#pragma unroll
            for (int q = 0; q < 6; ++q) { S[a][q] += T[q]; S[b][q] += T[q]; }
        }
        const size_t n = O.mxb;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            double* p = O.astress + tensor_base(O, o[a / 3] - 1 + (a % 3));
            atomicAdd(p + n * 0, S[a][0]); atomicAdd(p + n * 1, S[a][1]); atomicAdd(p + n * 2, S[a][2]);
            atomicAdd(p + n * 3, S[a][1]); atomicAdd(p + n * 4, S[a][3]); atomicAdd(p + n * 5, S[a][4]);
            atomicAdd(p + n * 6, S[a][2]); atomicAdd(p + n * 7, S[a][4]); atomicAdd(p + n * 8, S[a][5]);
        }
    }
}

__device__ __forceinline__ void trimer_block_scalars(double (&sc)[8], const OutPtrs& O, int bead)
{
    double v2[2] = {sc[0], sc[1]};
    block_reduce_add_bead<2>(v2, O.scal, S_ENG3, bead);
    double st[9] = {sc[2], sc[3], sc[4], sc[3], sc[5], sc[6], sc[4], sc[6], sc[7]};
    block_reduce_add_bead<9>(st, O.scal, S_STRESS, bead);
}

template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, 256 / TAPE_BS)
k_trimers(long long ntri, const int* __restrict__ tri, int nw, const double* __restrict__ mono, const int* __restrict__ list,
          const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr, CellInfo C, OutPtrs O, int flags,
          const double* __restrict__ tape)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const long long t0_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t0_ < ntri;
    const long long t = valid ? t0_ : ntri - 1;     // idle lanes shadow the last trimer: the tape walk has CTA barriers
    const PolyCtx ctx{dyn_smem, tape};
    double sc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // E, virial, stress xx xy xz yy yz zz
    int bead = -1;
    {
        const int mi = tri[3 * t], mj = tri[3 * t + 1], mk = tri[3 * t + 2];
        if (valid) bead = mi / O.nwb;
        double r[27];
        double E, cv[36];
        auto geom = [&](double (&rr)[27], bool again) {
            if (again) trimer_geometry<true>(nw, mono, C, mi, mj, mk, rr);
            else trimer_geometry<false>(nw, mono, C, mi, mj, mk, rr);
        };
        trimer_eval_regeom<ENGINE>(geom, r, E, cv, ctx);
        const int o[3] = {list[3 * mi], list[3 * mj], list[3 * mk]};   // 1-based O indices; H at o+1, o+2 (mbpol.f90:592-594)
        if (valid) trimer_scatter(r, E, cv, o, xr, yr, zr, O, flags, sc);
    }
    trimer_block_scalars(sc, O, bead);
}

// ---------------------------------------------------------------------------------------------------------
// K2: dimers
// ---------------------------------------------------------------------------------------------------------
template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, 256 / TAPE_BS)
k_dimers(long long ndim, const int2* __restrict__ dim, int nw, const double* __restrict__ mono, const int* __restrict__ list,
         const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr, CellInfo C, OutPtrs O, int flags,
         const double* __restrict__ tape)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const long long t0_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t0_ < ndim;
    const long long t = valid ? t0_ : ndim - 1;
    const PolyCtx ctx{dyn_smem, tape};
    double sc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int bead = -1;
    {
        const int mi = dim[t].x, mj = dim[t].y;
        if (valid) bead = mi / O.nwb;
        double wi[9], wj[9], w[18];
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            wi[c] = mono[(size_t)c * nw + mi];
            wj[c] = mono[(size_t)c * nw + mj];
        }
        double a0 = wi[0] - wj[0], a1 = wi[1] - wj[1], a2 = wi[2] - wj[2];
        image_vec(C, a0, a1, a2);                       // do_dimer, mbpol.f90:496-509; origin moved to O_j
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
        double E, g[18];
        double fm[108];
        const bool want_s = (flags & MBPOL_B200_WANT_ATOMIC_STRESS) != 0;
        dimer_eval<ENGINE>(w, E, g, want_s ? fm : nullptr, ctx);
        const double eu = MBPOL_ENGUNIT;
        const int o[2] = {list[3 * mi], list[3 * mj]};
        if (valid) {
        sc[0] = E * eu;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int atom = o[a / 3] - 1 + (a % 3);
            const double gx = g[3 * a] * eu, gy = g[3 * a + 1] * eu, gz = g[3 * a + 2] * eu;
            atomicAdd(O.fx + atom, -gx);
            atomicAdd(O.fy + atom, -gy);
            atomicAdd(O.fz + atom, -gz);
            sc[1] += w[3 * a] * gx + w[3 * a + 1] * gy + w[3 * a + 2] * gz;
            sc[2] -= w[3 * a] * gx; sc[3] -= w[3 * a] * gy; sc[4] -= w[3 * a] * gz;
            sc[5] -= w[3 * a + 1] * gy; sc[6] -= w[3 * a + 1] * gz; sc[7] -= w[3 * a + 2] * gz;
        }
        if (flags & MBPOL_B200_WANT_ATOMIC_ENERGY) {
            const double e6 = (1.0 / 6.0) * E * eu;
#pragma unroll
            for (int a = 0; a < 6; ++a) atomicAdd(O.aeng + o[a / 3] - 1 + (a % 3), e6);
        }
        if (flags & MBPOL_B200_WANT_FDECOMP) {
            double xs[3] = {0, 0, 0}, xraw[18];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const int atom = o[a / 3] - 1 + (a % 3);
                xraw[3 * a] = xr[atom]; xraw[3 * a + 1] = yr[atom]; xraw[3 * a + 2] = zr[atom];
                xs[0] += xraw[3 * a]; xs[1] += xraw[3 * a + 1]; xs[2] += xraw[3 * a + 2];
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double f[3] = {-(1.0 / 6.0) * g[3 * a] * eu, -(1.0 / 6.0) * g[3 * a + 1] * eu, -(1.0 / 6.0) * g[3 * a + 2] * eu};
                scatter_fdecomp(O, o[a / 3] - 1 + (a % 3), xraw + 3 * a, xs, 6, f);
            }
        }
        if (want_s) {   // mbpol.f90:558-579: S_i(a,b) = 1/2*eu*sum_{j != i} (r_i - r_j)_a * FM[i][j]_b  (FM is NOT antisymmetric)
            const size_t n = O.mxb;
            for (int i = 0; i < 6; ++i) {
                double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                for (int j = 0; j < 6; ++j) {
                    if (i == j) continue;
                    for (int a = 0; a < 3; ++a) {
                        const double ra_ = w[3 * i + a] - w[3 * j + a];
                        for (int b = 0; b < 3; ++b) S[a + 3 * b] += ra_ * fm[18 * i + 3 * j + b];
                    }
                }
                double* p = O.astress + tensor_base(O, o[i / 3] - 1 + (i % 3));
                for (int q = 0; q < 9; ++q) atomicAdd(p + n * q, 0.5 * eu * S[q]);
            }
        }
        }   // valid
    }
    double v2[2] = {sc[0], sc[1]};
    block_reduce_add_bead<2>(v2, O.scal, S_ENG2, bead);
    double st[9] = {sc[2], sc[3], sc[4], sc[3], sc[5], sc[6], sc[4], sc[6], sc[7]};
    block_reduce_add_bead<9>(st, O.scal, S_STRESS, bead);
}

// ---------------------------------------------------------------------------------------------------------
// batched per-cluster kernels (legacy ABI / KAT tests)
// ---------------------------------------------------------------------------------------------------------
template <int ENGINE>
__global__ void __launch_bounds__(TAPE_BS, 256 / TAPE_BS) k_eval_dimers(int n, const double* __restrict__ w, double* E, double* g, double* fm, const double* __restrict__ tape)
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
__global__ void __launch_bounds__(TAPE_BS, 256 / TAPE_BS) k_eval_trimers(int n, const double* __restrict__ w, double* E, double* g, double* fm, const double* __restrict__ tape)
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
