// mbnrg_kernels.cu -- MB-nrg ion-water / ion-ion 2-body path on the GPU (include/mbnrg_b200.h).
//
//   N0 k_nrg_gather   xxx/yyy/zzz + listttm2 -> imaged water monomers, SoA [9][nw]              (mbnrg.f90:214-243, image_water_monomers)
//   N1 k_nrg_pairs    every (ion, water of this shard) candidate: the reference's exact 10 A predicate (mbnrg.f90:519-526)
//                     and the evaluator's own cutoff test (mbnrg-2b-h2o-na-poly.cpp:224-234); ballot-compacted list of the
//                     pairs that actually contribute
//   N2 k_nrg_h2o_ion  one thread per listed pair: 8 variables, the 429-coefficient degree-5 polynomial as OUR shared-prefix
//                     Horner schedule (gen/polyhi_body.inc, tools/gen_mbnrg.py; 2 411 FP64 instructions, 2 128 of them FMAs, against the 3 888
//                     multiplies and adds of the Maple-generated reference body), chain rule, switch, scatter             (mbnrg-2b-h2o-na-poly.cpp:220-329,
//                     mbnrg.f90:706-775)
//   N3 k_nrg_ion_ion  one thread per ion pair j < k                                                 (mbnrg-2b-na-cl-poly.cpp:129-186)
// Scalars (energy, virial, stress) are block-reduced before one FP64 atomic per block; forces use FP64 red.global.add.
// The workload is small next to MB-pol (a few 1e3 pairs of ~2.4k FP64 instructions): these kernels are latency-, not
// throughput-critical, so there is no tape engine here -- the straight-line body is 39 KB of code and runs from the I-cache.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "mbnrg_b200.h"
#include "mbpol_device.cuh"
#include "gen/mbnrg_model_data.h"

namespace {

constexpr double NRG_R2F_LIST = 10.0;   // mbnrg.f90:130
constexpr int NSC = 16;                  // scalars: 0 eng2, 1 vir2, 2..10 stress, 11 unused

struct NrgIon {   // parameters of one ion species, order of MBNRG_PAR_HI
    double kHH, kOH, kXH, kXO, kXL, dHH, dOH, dXH, dXO, dXL, gin, gout, r2i, r2f;
};

__device__ __forceinline__ double v_exp(double r0, double k, const double* a, const double* b, double (&g)[3])
{
    g[0] = a[0] - b[0]; g[1] = a[1] - b[1]; g[2] = a[2] - b[2];
    const double r = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const double e = exp(k * (r0 - r));
    const double gg = -k * e / r;
    g[0] *= gg; g[1] *= gg; g[2] *= gg;
    return e;
}
__device__ __forceinline__ double v_coul(double r0, double k, const double* a, const double* b, double (&g)[3])
{
    g[0] = a[0] - b[0]; g[1] = a[1] - b[1]; g[2] = a[2] - b[2];
    const double r = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const double e = exp(k * (r0 - r));
    const double rinv = 1.0 / r;
    const double val = e * rinv;
    const double gg = -(k + rinv) * val * rinv;
    g[0] *= gg; g[1] *= gg; g[2] *= gg;
    return val;
}

// water w[9] (O,H1,H2) + ion x[3] -> E, g1[9], g2[3] (all in kcal/mol, NOT yet scaled by engunit); returns false outside r2f
__device__ __forceinline__ bool h2o_ion_eval(const NrgIon& P, const double* __restrict__ a, const double (&w)[9], const double (&x)[3],
                                             double& E, double (&g1)[9], double (&g2)[3])
{
    const double dXO[3] = {w[0] - x[0], w[1] - x[1], w[2] - x[2]};
    const double rXO = sqrt(norm2_exact(dXO[0], dXO[1], dXO[2]));
    if (rXO > P.r2f) return false;
    double oh1[3], oh2[3], l1[3], l2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { oh1[i] = w[3 + i] - w[i]; oh2[i] = w[6 + i] - w[i]; }
    const double vv[3] = {oh1[1] * oh2[2] - oh1[2] * oh2[1], oh1[2] * oh2[0] - oh1[0] * oh2[2], oh1[0] * oh2[1] - oh1[1] * oh2[0]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double inp = w[i] + 0.5 * P.gin * (oh1[i] + oh2[i]);
        const double outp = P.gout * vv[i];
        l1[i] = inp + outp; l2[i] = inp - outp;
    }
    double xv[8], g[8], c[8][3];
    xv[0] = v_exp(P.dHH, P.kHH, w + 3, w + 6, c[0]);
    xv[1] = v_exp(P.dOH, P.kOH, w, w + 3, c[1]);
    xv[2] = v_exp(P.dOH, P.kOH, w, w + 6, c[2]);
    xv[3] = v_coul(P.dXH, P.kXH, x, w + 3, c[3]);
    xv[4] = v_coul(P.dXH, P.kXH, x, w + 6, c[4]);
    xv[5] = v_coul(P.dXO, P.kXO, x, w, c[5]);
    xv[6] = v_exp(P.dXL, P.kXL, x, l1, c[6]);
    xv[7] = v_exp(P.dXL, P.kXL, x, l2, c[7]);
    double Ep;
    {
        const double* x = xv;   // the generated body names its inputs x[], g[], E
        double E;
#define PFMA(p, q, r) fma(p, q, r)
#define PMUL(p, q) ((p) * (q))
#define PADD(p, q) ((p) + (q))
#define PCOEF(k) __ldg(a + (k))
#include "gen/polyhi_body.inc"
#undef PFMA
#undef PMUL
#undef PADD
#undef PCOEF
        Ep = E;
    }
    // chain rule: xg = [O, H1, H2 | X | L1, L2]
    double gO[3], gH1[3], gH2[3], gX[3], gL1[3], gL2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double d0 = g[0] * c[0][i], d1 = g[1] * c[1][i], d2 = g[2] * c[2][i], d3 = g[3] * c[3][i], d4 = g[4] * c[4][i],
                     d5 = g[5] * c[5][i], d6 = g[6] * c[6][i], d7 = g[7] * c[7][i];
        gH1[i] = d0 - d1 - d3;
        gH2[i] = -d0 - d2 - d4;
        gO[i] = d1 + d2 - d5;
        gX[i] = d3 + d4 + d5 + d6 + d7;
        gL1[i] = -d6;
        gL2[i] = -d7;
    }
    {   // lone pairs back onto the atoms (monomer::grads)
        const double gm[3] = {gL1[0] - gL2[0], gL1[1] - gL2[1], gL1[2] - gL2[2]};
        const double t1[3] = {oh2[1] * gm[2] - oh2[2] * gm[1], oh2[2] * gm[0] - oh2[0] * gm[2], oh2[0] * gm[1] - oh2[1] * gm[0]};
        const double t2[3] = {oh1[1] * gm[2] - oh1[2] * gm[1], oh1[2] * gm[0] - oh1[0] * gm[2], oh1[0] * gm[1] - oh1[1] * gm[0]};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double gsum = gL1[i] + gL2[i];
            const double inp = 0.5 * P.gin * gsum;
            const double gh1 = inp + P.gout * t1[i];
            const double gh2 = inp - P.gout * t2[i];
            gO[i] += gsum - (gh1 + gh2);
            gH1[i] += gh1;
            gH2[i] += gh2;
        }
    }
    double gsw;
    const double sw = switch_fn(rXO, P.r2i, P.r2f, gsw);
    gsw *= Ep / rXO;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double d = gsw * dXO[i];
        g1[i] = sw * gO[i] + d;
        g1[3 + i] = sw * gH1[i];
        g1[6 + i] = sw * gH2[i];
        g2[i] = sw * gX[i] - d;
    }
    E = sw * Ep;
    return true;
}

// ion pair: E = sum_n a[n-1] v^n with v = exp(k (d - r)); Horner, the derivative along the same chain
__device__ __forceinline__ bool ion_ion_eval(const double* __restrict__ A, const double* __restrict__ P, const double (&w)[3],
                                             const double (&x)[3], double& E, double (&g1)[3], double (&g2)[3])
{
    const double d3[3] = {w[0] - x[0], w[1] - x[1], w[2] - x[2]};
    const double r = sqrt(norm2_exact(d3[0], d3[1], d3[2]));
    if (r > P[3]) return false;
    double c[3];
    const double v = v_exp(P[1], P[0], w, x, c);
    double T = A[12] * v;
    double D = 2.0 * T + A[11];
    T = (T + A[11]) * v;
#pragma unroll
    for (int n = 10; n >= 0; --n) {
        D = D * v + A[n] + T;
        T = (A[n] + T) * v;
    }
    double gsw;
    const double sw = switch_fn(r, P[2], P[3], gsw);
    gsw *= T / r;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double d = gsw * d3[i];
        g1[i] = sw * D * c[i] + d;
        g2[i] = -sw * D * c[i] - d;
    }
    E = sw * T;
    return true;
}

__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

// block reduction of NV scalars, one atomic per scalar and block
template <int NVAL>
__device__ __forceinline__ void block_add(double (&v)[NVAL], double* __restrict__ dst)
{
    __shared__ double part[32][NVAL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NVAL; ++k) {
        double s = v[k];
        for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) part[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NVAL) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += part[w][threadIdx.x];
        if (s != 0.0) red_add(dst + threadIdx.x, s);
    }
}

__global__ void k_nrg_gather(int nw, const int* __restrict__ list, const double* __restrict__ x, const double* __restrict__ y,
                             const double* __restrict__ z, CellInfo C, double* __restrict__ mono)
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
}

// the water of pair (ion j, water m) placed next to the ion exactly as do_dimer does (mbnrg.f90:519-541)
__device__ __forceinline__ bool place_water(int nw, const double* __restrict__ mono, const CellInfo& C, int m, const double (&x)[3],
                                            double (&w)[9])
{
#pragma unroll
    for (int c = 0; c < 9; ++c) w[c] = mono[(size_t)c * nw + m];
    double dx = __dsub_rn(w[0], x[0]), dy = __dsub_rn(w[1], x[1]), dz = __dsub_rn(w[2], x[2]);
    image_vec(C, dx, dy, dz);
    if (norm2_exact(dx, dy, dz) > NRG_R2F_LIST * NRG_R2F_LIST) return false;
    const double o[3] = {__dadd_rn(x[0], dx), __dadd_rn(x[1], dy), __dadd_rn(x[2], dz)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        w[3 + k] = __dadd_rn(__dsub_rn(w[3 + k], w[k]), o[k]);
        w[6 + k] = __dadd_rn(__dsub_rn(w[6 + k], w[k]), o[k]);
        w[k] = o[k];
    }
    return true;
}

__global__ void k_nrg_pairs(int nw, int lo, int nwr, int nions, const double* __restrict__ mono, const double* __restrict__ xyz, int n,
                            const int* __restrict__ code, const NrgIon* __restrict__ par, CellInfo C, int2* __restrict__ pairs,
                            unsigned long long* __restrict__ counters)
{
    const long long total = (long long)nions * nwr;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
        const long long c = base + threadIdx.x;
        bool in_list = false, active = false;
        int j = 0, m = 0;
        if (c < total) {
            j = (int)(c / nwr);
            m = lo + (int)(c % nwr);
            const double x[3] = {xyz[j], xyz[n + j], xyz[2 * (size_t)n + j]};
            double w[9];
            in_list = place_water(nw, mono, C, m, x, w);
            if (in_list) {
                const double dx = __dsub_rn(w[0], x[0]), dy = __dsub_rn(w[1], x[1]), dz = __dsub_rn(w[2], x[2]);
                active = !(sqrt(norm2_exact(dx, dy, dz)) > par[code[j] - 1].r2f);
            }
        }
        const unsigned bl = __ballot_sync(0xffffffffu, in_list), ba = __ballot_sync(0xffffffffu, active);
        const int lane = threadIdx.x & 31;
        unsigned long long pos = 0;
        if (lane == 0) {
            if (bl) atomicAdd(counters + 0, (unsigned long long)__popc(bl));
            if (ba) pos = atomicAdd(counters + 1, (unsigned long long)__popc(ba));
        }
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (active) pairs[pos + __popc(ba & ((1u << lane) - 1u))] = make_int2(j, m);
    }
}

__global__ void __launch_bounds__(128) k_nrg_h2o_ion(int nw, const double* __restrict__ mono, const double* __restrict__ xyz, int n,
                                                      const int* __restrict__ list, const int* __restrict__ listions,
                                                      const int* __restrict__ code, const NrgIon* __restrict__ par,
                                                      const double* __restrict__ coef, CellInfo C, const int2* __restrict__ pairs,
                                                      const unsigned long long* __restrict__ counters, double* __restrict__ out)
{
    const unsigned long long np = counters[1];
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    double sc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) sc[k] = 0.0;
    if (t < np) {
        const int2 p = pairs[t];
        const int j = p.x, m = p.y, ion = code[j] - 1;
        const double x[3] = {xyz[j], xyz[n + j], xyz[2 * (size_t)n + j]};
        double w[9], E, g1[9], g2[3];
        place_water(nw, mono, C, m, x, w);
        if (h2o_ion_eval(par[ion], coef + 429 * ion, w, x, E, g1, g2)) {
            sc[0] = E * MBPOL_ENGUNIT;
#pragma unroll
            for (int k = 0; k < 9; ++k) { g1[k] *= MBPOL_ENGUNIT; sc[1] += w[k] * g1[k]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { g2[k] *= MBPOL_ENGUNIT; sc[1] += x[k] * g2[k]; }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    sc[2 + 3 * a + b] = -(w[a] * g1[b] + w[3 + a] * g1[3 + b] + w[6 + a] * g1[6 + b] + x[a] * g2[b]);
            double* fx = out;
            double* fy = out + n;
            double* fz = out + 2 * (size_t)n;
            const int o1 = list[3 * m] - 1, o2 = listions[j] - 1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                red_add(fx + o1 + k, -g1[3 * k]);
                red_add(fy + o1 + k, -g1[3 * k + 1]);
                red_add(fz + o1 + k, -g1[3 * k + 2]);
            }
            red_add(fx + o2, -g2[0]);
            red_add(fy + o2, -g2[1]);
            red_add(fz + o2, -g2[2]);
        }
    }
    // the reference computes strs(2,3,6) and mirrors them into (4,7,8) (mbnrg.f90:745-771)
    sc[2 + 3] = sc[2 + 1]; sc[2 + 6] = sc[2 + 2]; sc[2 + 7] = sc[2 + 5];
    block_add<11>(sc, out + 3 * (size_t)n);
}

__global__ void __launch_bounds__(128) k_nrg_ion_ion(int nions, const double* __restrict__ xyz, int n, const int* __restrict__ listions,
                                                      const int* __restrict__ code, const double* __restrict__ coef_ii,
                                                      const double* __restrict__ par_ii, const unsigned char* __restrict__ have_ii,
                                                      CellInfo C, unsigned long long* __restrict__ counters, double* __restrict__ out)
{
    const long long total = (long long)nions * (nions - 1) / 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double sc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) sc[k] = 0.0;
    if (t < total) {
        // t -> (j, k), j < k
        int j = (int)((2.0 * nions - 1.0 - sqrt((2.0 * nions - 1.0) * (2.0 * nions - 1.0) - 8.0 * (double)t)) * 0.5);
        while ((long long)j * (2 * nions - j - 1) / 2 > t) --j;
        while ((long long)(j + 1) * (2 * nions - j - 2) / 2 <= t) ++j;
        const int k = j + 1 + (int)(t - (long long)j * (2 * nions - j - 1) / 2);
        const int a = code[j] - 1, b = code[k] - 1;
        const double y[3] = {xyz[k], xyz[n + k], xyz[2 * (size_t)n + k]};
        double dx = __dsub_rn(xyz[j], y[0]), dy = __dsub_rn(xyz[n + j], y[1]), dz = __dsub_rn(xyz[2 * (size_t)n + j], y[2]);
        image_vec(C, dx, dy, dz);
        if (have_ii[9 * a + b] && !(norm2_exact(dx, dy, dz) > NRG_R2F_LIST * NRG_R2F_LIST)) {
            const double x[3] = {__dadd_rn(y[0], dx), __dadd_rn(y[1], dy), __dadd_rn(y[2], dz)};
            double E, g1[3], g2[3];
            atomicAdd(counters + 2, 1ull);
            if (ion_ion_eval(coef_ii + 13 * (9 * a + b), par_ii + 4 * (9 * a + b), x, y, E, g1, g2)) {
                sc[0] = E * MBPOL_ENGUNIT;
#pragma unroll
                for (int q = 0; q < 3; ++q) { g1[q] *= MBPOL_ENGUNIT; g2[q] *= MBPOL_ENGUNIT; sc[1] += x[q] * g1[q] + y[q] * g2[q]; }
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int q = 0; q < 3; ++q) sc[2 + 3 * p + q] = -(x[p] * g1[q] + y[p] * g2[q]);
                const int o1 = listions[j] - 1, o2 = listions[k] - 1;
                red_add(out + o1, -g1[0]); red_add(out + n + o1, -g1[1]); red_add(out + 2 * (size_t)n + o1, -g1[2]);
                red_add(out + o2, -g2[0]); red_add(out + n + o2, -g2[1]); red_add(out + 2 * (size_t)n + o2, -g2[2]);
            }
        }
    }
    sc[2 + 3] = sc[2 + 1]; sc[2 + 6] = sc[2 + 2]; sc[2 + 7] = sc[2 + 5];
    block_add<11>(sc, out + 3 * (size_t)n);
}

__global__ void k_nrg_eval_hi(int n, int ion, const NrgIon* __restrict__ par, const double* __restrict__ coef, const double* __restrict__ w,
                              const double* __restrict__ x, double* __restrict__ E, double* __restrict__ g1, double* __restrict__ g2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double ww[9], xx[3], e = 0.0, a[9], b[3];
    for (int k = 0; k < 9; ++k) { ww[k] = w[9 * t + k]; a[k] = 0.0; }
    for (int k = 0; k < 3; ++k) { xx[k] = x[3 * t + k]; b[k] = 0.0; }
    h2o_ion_eval(par[ion - 1], coef + 429 * (ion - 1), ww, xx, e, a, b);
    E[t] = e;
    for (int k = 0; k < 9; ++k) g1[9 * t + k] = a[k];
    for (int k = 0; k < 3; ++k) g2[3 * t + k] = b[k];
}
__global__ void k_nrg_eval_ii(int n, int pa, int pb, const double* __restrict__ coef_ii, const double* __restrict__ par_ii,
                              const double* __restrict__ w, const double* __restrict__ x, double* __restrict__ E, double* __restrict__ g1,
                              double* __restrict__ g2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double ww[3], xx[3], e = 0.0, a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) { ww[k] = w[3 * t + k]; xx[k] = x[3 * t + k]; }
    const int q = 9 * (pa - 1) + (pb - 1);
    ion_ion_eval(coef_ii + 13 * q, par_ii + 4 * q, ww, xx, e, a, b);
    E[t] = e;
    for (int k = 0; k < 3; ++k) { g1[3 * t + k] = a[k]; g2[3 * t + k] = b[k]; }
}

// =========================================================================================================
// 3B ion-water-water (mbnrg-3b-h2o-h2o-<ion>-poly.cpp, poly-3b-h2o-ion-v1x.cpp; driver mbnrg.f90:354-427, do_trimer :781-993 --
// compiled out in the reference build by `!#define ENABLE_X3B`, mbnrg.f90:4, and therefore a separate, optional entry point here)
// =========================================================================================================
constexpr double NRG_R3F_LIST = 7.0;    // mbnrg.f90:135
constexpr int NRG3_NEAR_CAP = 256;      // waters within 7 A of one ion: ~48 in liquid water

struct Nrg3 {   // parameters of one ion species, order of tools/extract_mbnrg.py PAR_3B
    double r3i, r3f, k[7], d[7];        // HH_intra OH_intra HH OH OO XH XO
};

#include "gen/poly3_parts.inc"

// sites 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 X; variable order of mbnrg-3b-h2o-h2o-na-poly.cpp:168-189
__device__ constexpr int N3A[21] = {1, 4, 0, 0, 3, 3, 1, 1, 2, 2, 0, 0, 3, 3, 0, 6, 6, 6, 6, 6, 6};
__device__ constexpr int N3B[21] = {2, 5, 1, 2, 4, 5, 4, 5, 4, 5, 4, 5, 1, 2, 3, 1, 2, 4, 5, 0, 3};
__device__ constexpr int N3K[21] = {0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 5, 5, 5, 5, 6, 6};

// c[21] = w1 (9), w2 (9), x (3) -> E and the gradient on the 7 sites g[21] (kcal/mol, assigned).  The polynomial is OUR shared-prefix
// Horner schedule (tools/gen_mbnrg.py: 25.5 k FP64 instructions in 24 device functions through a per-thread record).
__device__ __forceinline__ void h2o_h2o_ion_eval(const Nrg3& P, const double* __restrict__ a, const double (&c)[21], double& E, double (&g)[21])
{
    double rec[NRG3_RECORD_DOUBLES];
    double fac[21];      // -k x / d of every variable
#pragma unroll
    for (int n = 0; n < 21; ++n) {
        const int p = 3 * N3A[n], q = 3 * N3B[n];
        const double dx = c[p] - c[q], dy = c[p + 1] - c[q + 1], dz = c[p + 2] - c[q + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        const double xv = exp(-P.k[N3K[n]] * (d - P.d[N3K[n]]));
        rec[n] = xv;
        fac[n] = -P.k[N3K[n]] * xv / d;
    }
    NRG3_CALL_ALL(rec, a)
    const double Ep = rec[21];
    double rab[3], rac[3], rbc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { rab[i] = c[i] - c[9 + i]; rac[i] = c[i] - c[18 + i]; rbc[i] = c[9 + i] - c[18 + i]; }
    const double drab = sqrt(rab[0] * rab[0] + rab[1] * rab[1] + rab[2] * rab[2]);
    const double drac = sqrt(rac[0] * rac[0] + rac[1] * rac[1] + rac[2] * rac[2]);
    const double drbc = sqrt(rbc[0] * rbc[0] + rbc[1] * rbc[1] + rbc[2] * rbc[2]);
    double gab, gac, gbc;
    const double sab = switch_fn(drab, P.r3i, P.r3f, gab), sac = switch_fn(drac, P.r3i, P.r3f, gac), sbc = switch_fn(drbc, P.r3i, P.r3f, gbc);
    const double sw = sab * sac + sab * sbc + sac * sbc;
#pragma unroll
    for (int i = 0; i < 21; ++i) g[i] = 0.0;
#pragma unroll
    for (int n = 0; n < 21; ++n) {
        const int p = 3 * N3A[n], q = 3 * N3B[n];
        const double gg = (rec[22 + n] * sw) * fac[n];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double f = gg * (c[p + i] - c[q + i]);
            g[p + i] += f;
            g[q + i] -= f;
        }
    }
    gab *= (sac + sbc) * Ep / drab;
    gac *= (sab + sbc) * Ep / drac;
    gbc *= (sab + sac) * Ep / drbc;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g[i] += gab * rab[i] + gac * rac[i];
        g[9 + i] += gbc * rbc[i] - gab * rab[i];
        g[18 + i] -= gac * rac[i] + gbc * rbc[i];
    }
    E = sw * Ep;
}

// one warp per ion: the waters whose imaged oxygen-ion distance is < r3f (strict, mbnrg.f90:840-842), in index order
__global__ void k_nrg3_near(int nw, int nions, const double* __restrict__ mono, const double* __restrict__ xyz, int n, CellInfo C,
                            int* __restrict__ near, int* __restrict__ nnear, int* __restrict__ overflow)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (k >= nions) return;
    const double x[3] = {xyz[k], xyz[n + k], xyz[2 * (size_t)n + k]};
    int cnt = 0;
    for (int i0 = 0; i0 < nw; i0 += 32) {
        const int i = i0 + (int)lane;
        bool in = false;
        if (i < nw) {
            double dx = __dsub_rn(mono[i], x[0]), dy = __dsub_rn(mono[(size_t)nw + i], x[1]), dz = __dsub_rn(mono[2 * (size_t)nw + i], x[2]);
            image_vec(C, dx, dy, dz);
            in = __dsqrt_rn(norm2_exact(dx, dy, dz)) < NRG_R3F_LIST;
        }
        const unsigned b = __ballot_sync(0xffffffffu, in);
        const int pos = cnt + __popc(b & ((1u << lane) - 1u));
        if (in && pos < NRG3_NEAR_CAP) near[(size_t)k * NRG3_NEAR_CAP + pos] = i;
        cnt += __popc(b);
    }
    if (lane == 0) {
        if (cnt > NRG3_NEAR_CAP) atomicExch(overflow, 1);
        nnear[k] = min(cnt, NRG3_NEAR_CAP);
    }
}

// exclusive prefix of the trimer counts n (n - 1) / 2 per ion (a few hundred ions at most: one thread)
__global__ void k_nrg3_offsets(int nions, const int* __restrict__ nnear, long long* __restrict__ off)
{
    if (blockIdx.x || threadIdx.x) return;
    long long t = 0;
    for (int k = 0; k < nions; ++k) {
        off[k] = t;
        t += (long long)nnear[k] * (nnear[k] - 1) / 2;
    }
    off[nions] = t;
}

// one thread per accepted trimer of this rank's contiguous range of the (ion, i, j) list
__global__ void __launch_bounds__(128) k_nrg3_trimers(int rank, int nranks, int nw, int nions, const double* __restrict__ mono,
                                                       const double* __restrict__ xyz, int n, const int* __restrict__ list,
                                                       const int* __restrict__ listions, const int* __restrict__ code, const Nrg3* __restrict__ par,
                                                       const double* __restrict__ coef, CellInfo C, const int* __restrict__ near,
                                                       const int* __restrict__ nnear, const long long* __restrict__ off, double* __restrict__ out)
{
    const long long T = off[nions], lo = T * rank / nranks, hi = T * (rank + 1) / nranks;
    for (long long base = lo + (long long)blockIdx.x * blockDim.x; base < hi; base += (long long)gridDim.x * blockDim.x) {
        const long long t = base + threadIdx.x;
        double sc[11];
#pragma unroll
        for (int q = 0; q < 11; ++q) sc[q] = 0.0;
        if (t < hi) {
            int k = 0;
            while (k + 1 < nions && off[k + 1] <= t) ++k;
            const long long e = t - off[k];      // pair (p, q), p < q, in the order (0,1), (0,2), ..., (0,n-1), (1,2), ...
            const int nn = nnear[k];
            int p = (int)((2.0 * nn - 1.0 - sqrt((2.0 * nn - 1.0) * (2.0 * nn - 1.0) - 8.0 * (double)e)) * 0.5);
            while ((long long)p * (2 * nn - p - 1) / 2 > e) --p;
            while ((long long)(p + 1) * (2 * nn - p - 2) / 2 <= e) ++p;
            const int q = p + 1 + (int)(e - (long long)p * (2 * nn - p - 1) / 2);
            const int i = near[(size_t)k * NRG3_NEAR_CAP + p], j = near[(size_t)k * NRG3_NEAR_CAP + q];
            double wi[9], c[21];
#pragma unroll
            for (int a = 0; a < 9; ++a) { wi[a] = mono[(size_t)a * nw + i]; c[9 + a] = mono[(size_t)a * nw + j]; }
            // do_trimer's placement (mbnrg.f90:849-858): water i and the ion next to water j
            double ax = __dsub_rn(wi[0], c[9]), ay = __dsub_rn(wi[1], c[10]), az = __dsub_rn(wi[2], c[11]);
            image_vec(C, ax, ay, az);
            double bx = __dsub_rn(c[9], xyz[k]), by = __dsub_rn(c[10], xyz[n + k]), bz = __dsub_rn(c[11], xyz[2 * (size_t)n + k]);
            image_vec(C, bx, by, bz);
            const double ra[3] = {ax, ay, az}, rb[3] = {bx, by, bz};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double o = c[9 + a] + ra[a];
                c[3 + a] = (wi[3 + a] - wi[a]) + o;
                c[6 + a] = (wi[6 + a] - wi[a]) + o;
                c[a] = o;
                c[18 + a] = c[9 + a] - rb[a];
            }
            const int ion = code[k] - 1;
            double E, g[21];
            h2o_h2o_ion_eval(par[ion], coef + 1831 * ion, c, E, g);
            sc[0] = E * MBPOL_ENGUNIT;
#pragma unroll
            for (int a = 0; a < 21; ++a) { g[a] *= MBPOL_ENGUNIT; sc[1] += c[a] * g[a]; }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    double s_ = 0.0;
#pragma unroll
                    for (int m = 0; m < 7; ++m) s_ += c[3 * m + a] * g[3 * m + b];
                    sc[2 + 3 * a + b] = -s_;
                }
            double* fx = out;
            double* fy = out + n;
            double* fz = out + 2 * (size_t)n;
            const int o1 = list[3 * i] - 1, o2 = list[3 * j] - 1, o3 = listions[k] - 1;
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                red_add(fx + o1 + m, -g[3 * m]); red_add(fy + o1 + m, -g[3 * m + 1]); red_add(fz + o1 + m, -g[3 * m + 2]);
                red_add(fx + o2 + m, -g[9 + 3 * m]); red_add(fy + o2 + m, -g[9 + 3 * m + 1]); red_add(fz + o2 + m, -g[9 + 3 * m + 2]);
            }
            red_add(fx + o3, -g[18]); red_add(fy + o3, -g[19]); red_add(fz + o3, -g[20]);
        }
        // the reference computes strs(1,2,3,5,6,9) and mirrors (4,7,8) (mbnrg.f90:924-988)
        sc[2 + 3] = sc[2 + 1]; sc[2 + 6] = sc[2 + 2]; sc[2 + 7] = sc[2 + 5];
        block_add<11>(sc, out + 3 * (size_t)n);
        __syncthreads();
    }
}

__global__ void k_nrg_eval_3b(int n, int ion, const Nrg3* __restrict__ par, const double* __restrict__ coef, const double* __restrict__ w,
                              double* __restrict__ E, double* __restrict__ g)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double c[21], e, gg[21];
    for (int k = 0; k < 21; ++k) c[k] = w[21 * t + k];
    h2o_h2o_ion_eval(par[ion - 1], coef + 1831 * (ion - 1), c, e, gg);
    E[t] = e;
    for (int k = 0; k < 21; ++k) g[21 * t + k] = gg[k];
}

void host_invert3(const double* a, double* b, double& d)   // invert.f
{
    b[0] = a[4] * a[8] - a[5] * a[7]; b[1] = a[2] * a[7] - a[1] * a[8]; b[2] = a[1] * a[5] - a[2] * a[4];
    b[3] = a[5] * a[6] - a[3] * a[8]; b[4] = a[0] * a[8] - a[2] * a[6]; b[5] = a[2] * a[3] - a[0] * a[5];
    b[6] = a[3] * a[7] - a[4] * a[6]; b[7] = a[1] * a[6] - a[0] * a[7]; b[8] = a[0] * a[4] - a[1] * a[3];
    d = a[0] * b[0] + a[3] * b[1] + a[6] * b[2];
    double r = 0.0;
    if (std::fabs(d) > 0.0) r = 1.0 / d;
    for (int i = 0; i < 9; ++i) b[i] = r * b[i];
}

}  // namespace

struct mbnrg_b200_handle {
    int device = 0, nttm2 = 0, nw = 0, nions = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    std::vector<int> codes;
    // device buffers
    double *xyz = nullptr, *mono = nullptr, *out = nullptr, *coef = nullptr, *coef_ii = nullptr, *par_ii = nullptr, *scratch = nullptr;
    NrgIon* par = nullptr;
    int *list = nullptr, *listions = nullptr, *code = nullptr;
    int2* pairs = nullptr;
    unsigned char* have_ii = nullptr;
    unsigned long long* counters = nullptr;
    // 3B
    Nrg3* par3 = nullptr;
    double* coef3 = nullptr;
    int *near = nullptr, *nnear = nullptr, *ovf3 = nullptr;
    long long* off3 = nullptr;
    long long counts3[2] = {0, 0};
    float ms3[3] = {0, 0, 0};
    cudaEvent_t ev3[3] = {nullptr, nullptr, nullptr};
    size_t cap_atoms = 0, cap_scratch = 0;
    std::vector<double> host_out;
    long long counts[3] = {0, 0, 0};
    float ms[4] = {0, 0, 0, 0};
    int launches = 0;
};

#define NRG_TRY(h, call)                                                                                    \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
            return MBPOL_B200_ECUDA;                                                                        \
        }                                                                                                   \
    } while (0)

extern "C" {

int mbnrg_b200_init(int nttm2, int nions, const int* ion_code, int device, mbnrg_b200_handle** out)
{
    if (!out) return MBPOL_B200_EINVAL;
    *out = nullptr;
    if (nttm2 < 0 || nttm2 % 3 || nions < 0 || (nions > 0 && !ion_code)) return MBPOL_B200_EINVAL;
    for (int j = 0; j < nions; ++j)
        if (ion_code[j] < 1 || ion_code[j] > MBNRG_NIONS) return MBPOL_B200_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MBPOL_B200_ENODEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return MBPOL_B200_ENODEVICE;
    cudaDeviceProp prop;
    if (device >= ndev || cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return MBPOL_B200_ENODEVICE;
    auto* h = new mbnrg_b200_handle;
    *out = h;   // returned even on a CUDA error so that the caller can read last_error and must call fini
    h->device = device; h->nttm2 = nttm2; h->nw = nttm2 / 3; h->nions = nions;
    h->codes.assign(ion_code, ion_code + nions);
    NRG_TRY(h, cudaSetDevice(device));
    NRG_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) NRG_TRY(h, cudaEventCreate(&e));
    static_assert(sizeof(NrgIon) == 14 * sizeof(double), "parameter layout");
    NRG_TRY(h, cudaMalloc(&h->par, sizeof(MBNRG_PAR_HI)));
    NRG_TRY(h, cudaMalloc(&h->coef, sizeof(MBNRG_COEF_HI)));
    NRG_TRY(h, cudaMalloc(&h->coef_ii, sizeof(MBNRG_COEF_II)));
    NRG_TRY(h, cudaMalloc(&h->par_ii, sizeof(MBNRG_PAR_II)));
    NRG_TRY(h, cudaMalloc(&h->have_ii, sizeof(MBNRG_HAVE_II)));
    NRG_TRY(h, cudaMalloc(&h->counters, 4 * sizeof(unsigned long long)));
    NRG_TRY(h, cudaMemcpy(h->par, MBNRG_PAR_HI, sizeof(MBNRG_PAR_HI), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMemcpy(h->coef, MBNRG_COEF_HI, sizeof(MBNRG_COEF_HI), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMemcpy(h->coef_ii, MBNRG_COEF_II, sizeof(MBNRG_COEF_II), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMemcpy(h->par_ii, MBNRG_PAR_II, sizeof(MBNRG_PAR_II), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMemcpy(h->have_ii, MBNRG_HAVE_II, sizeof(MBNRG_HAVE_II), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMalloc(&h->list, std::max<size_t>(1, nttm2) * sizeof(int)));
    NRG_TRY(h, cudaMalloc(&h->mono, std::max<size_t>(1, 9 * (size_t)h->nw) * sizeof(double)));
    NRG_TRY(h, cudaMalloc(&h->listions, std::max<size_t>(1, nions) * sizeof(int)));
    NRG_TRY(h, cudaMalloc(&h->code, std::max<size_t>(1, nions) * sizeof(int)));
    NRG_TRY(h, cudaMalloc(&h->pairs, std::max<size_t>(1, (size_t)nions * h->nw) * sizeof(int2)));
    if (nions) NRG_TRY(h, cudaMemcpy(h->code, ion_code, nions * sizeof(int), cudaMemcpyHostToDevice));
    static_assert(sizeof(Nrg3) == 16 * sizeof(double), "3B parameter layout");
    NRG_TRY(h, cudaMalloc(&h->par3, sizeof(MBNRG_PAR_3B)));
    NRG_TRY(h, cudaMalloc(&h->coef3, sizeof(MBNRG_COEF_3B)));
    NRG_TRY(h, cudaMemcpy(h->par3, MBNRG_PAR_3B, sizeof(MBNRG_PAR_3B), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMemcpy(h->coef3, MBNRG_COEF_3B, sizeof(MBNRG_COEF_3B), cudaMemcpyHostToDevice));
    NRG_TRY(h, cudaMalloc(&h->near, std::max<size_t>(1, nions) * NRG3_NEAR_CAP * sizeof(int)));
    NRG_TRY(h, cudaMalloc(&h->nnear, std::max<size_t>(1, nions) * sizeof(int)));
    NRG_TRY(h, cudaMalloc(&h->off3, (std::max<size_t>(1, nions) + 1) * sizeof(long long)));
    NRG_TRY(h, cudaMalloc(&h->ovf3, sizeof(int)));
    for (auto& e : h->ev3) NRG_TRY(h, cudaEventCreate(&e));
    return MBPOL_B200_OK;
}

int mbnrg_b200_fini(mbnrg_b200_handle* h)
{
    if (!h) return MBPOL_B200_EINVAL;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : {(void*)h->xyz, (void*)h->mono, (void*)h->out, (void*)h->coef, (void*)h->coef_ii, (void*)h->par_ii, (void*)h->scratch,
                    (void*)h->par, (void*)h->list, (void*)h->listions, (void*)h->code, (void*)h->pairs, (void*)h->have_ii, (void*)h->counters,
                    (void*)h->par3, (void*)h->coef3, (void*)h->near, (void*)h->nnear, (void*)h->off3, (void*)h->ovf3})
        if (p) cudaFree(p);
    for (auto e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto e : h->ev3)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MBPOL_B200_OK;
}

const char* mbnrg_b200_last_error(const mbnrg_b200_handle* h) { return h ? h->err.c_str() : "null handle"; }

int mbnrg_b200_set_shard(mbnrg_b200_handle* h, int rank, int nranks)
{
    if (!h || nranks < 1 || rank < 0 || rank >= nranks) return MBPOL_B200_EINVAL;
    h->rank = rank; h->nranks = nranks;
    return MBPOL_B200_OK;
}

int mbnrg_b200_forces(mbnrg_b200_handle* h, int imcon, const double* cell, int nttm2, const int* listttm2, int nions, const int* listions,
                      int mxatms, const double* xxx, const double* yyy, const double* zzz, double* fxx, double* fyy, double* fzz,
                      double* eng2, double* vir2, double* stress)
{
    if (!h) return MBPOL_B200_EINVAL;
    if (imcon < 1 || imcon > 3 || !cell || nttm2 != h->nttm2 || nions != h->nions || mxatms < nions || !xxx || !yyy || !zzz || !fxx || !fyy ||
        !fzz || !eng2 || !vir2 || !stress || (nttm2 && !listttm2) || (nions && !listions)) {
        h->err = "mbnrg_b200_forces: invalid argument (imcon must be 1, 2 or 3; nttm2 / nions must match init)";
        return MBPOL_B200_EINVAL;
    }
    NRG_TRY(h, cudaSetDevice(h->device));
    const size_t n = mxatms;
    if (n > h->cap_atoms) {
        if (h->xyz) cudaFree(h->xyz);
        if (h->out) cudaFree(h->out);
        h->xyz = h->out = nullptr;
        NRG_TRY(h, cudaMalloc(&h->xyz, 3 * n * sizeof(double)));
        NRG_TRY(h, cudaMalloc(&h->out, (3 * n + NSC) * sizeof(double)));
        h->cap_atoms = n;
    }
    CellInfo C;
    C.imcon = imcon;
    std::copy(cell, cell + 9, C.cell);
    double det;
    host_invert3(cell, C.rcell, det);
    C.a1 = 1.0 / cell[0]; C.b1 = 1.0 / cell[4]; C.c1 = 1.0 / cell[8];
    cudaStream_t s = h->stream;
    const int nw = h->nw;
    const int lo = (int)((long long)nw * h->rank / h->nranks), hi = (int)((long long)nw * (h->rank + 1) / h->nranks), nwr = hi - lo;
    NRG_TRY(h, cudaMemcpyAsync(h->xyz, xxx, n * 8, cudaMemcpyHostToDevice, s));
    NRG_TRY(h, cudaMemcpyAsync(h->xyz + n, yyy, n * 8, cudaMemcpyHostToDevice, s));
    NRG_TRY(h, cudaMemcpyAsync(h->xyz + 2 * n, zzz, n * 8, cudaMemcpyHostToDevice, s));
    if (nttm2) NRG_TRY(h, cudaMemcpyAsync(h->list, listttm2, (size_t)nttm2 * 4, cudaMemcpyHostToDevice, s));
    if (nions) NRG_TRY(h, cudaMemcpyAsync(h->listions, listions, (size_t)nions * 4, cudaMemcpyHostToDevice, s));
    NRG_TRY(h, cudaMemsetAsync(h->out, 0, (3 * n + NSC) * sizeof(double), s));
    NRG_TRY(h, cudaMemsetAsync(h->counters, 0, 4 * sizeof(unsigned long long), s));
    h->launches = 0;
    NRG_TRY(h, cudaEventRecord(h->ev[0], s));
    if (nw && nions && nwr) {
        k_nrg_gather<<<(nw + 127) / 128, 128, 0, s>>>(nw, h->list, h->xyz, h->xyz + n, h->xyz + 2 * n, C, h->mono);
        const long long total = (long long)nions * nwr;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
        k_nrg_pairs<<<blocks, 256, 0, s>>>(nw, lo, nwr, nions, h->mono, h->xyz, (int)n, h->code, h->par, C, h->pairs, h->counters);
        h->launches += 2;
    }
    NRG_TRY(h, cudaEventRecord(h->ev[1], s));
    unsigned long long cnt[4] = {0, 0, 0, 0};
    if (nw && nions && nwr) {
        // the grid of the evaluation kernel is sized from the list length (one small D2H read, like the MB-pol path)
        NRG_TRY(h, cudaMemcpyAsync(cnt, h->counters, sizeof(cnt), cudaMemcpyDeviceToHost, s));
        NRG_TRY(h, cudaStreamSynchronize(s));
        if (cnt[1]) {
            k_nrg_h2o_ion<<<(unsigned)((cnt[1] + 127) / 128), 128, 0, s>>>(nw, h->mono, h->xyz, (int)n, h->list, h->listions, h->code, h->par,
                                                                           h->coef, C, h->pairs, h->counters, h->out);
            ++h->launches;
        }
    }
    NRG_TRY(h, cudaEventRecord(h->ev[2], s));
    if (h->rank == 0 && nions > 1) {
        const long long total = (long long)nions * (nions - 1) / 2;
        k_nrg_ion_ion<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(nions, h->xyz, (int)n, h->listions, h->code, h->coef_ii, h->par_ii,
                                                                      h->have_ii, C, h->counters, h->out);
        ++h->launches;
    }
    NRG_TRY(h, cudaEventRecord(h->ev[3], s));
    NRG_TRY(h, cudaGetLastError());
    h->host_out.resize(3 * n + NSC);
    NRG_TRY(h, cudaMemcpyAsync(h->host_out.data(), h->out, (3 * n + NSC) * sizeof(double), cudaMemcpyDeviceToHost, s));
    NRG_TRY(h, cudaMemcpyAsync(cnt, h->counters, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    NRG_TRY(h, cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k) {
        h->counts[k] = (long long)cnt[k];
        cudaEventElapsedTime(&h->ms[k], h->ev[k], h->ev[k + 1]);
    }
    cudaEventElapsedTime(&h->ms[3], h->ev[0], h->ev[3]);
    const double* o = h->host_out.data();
    for (size_t i = 0; i < n; ++i) {
        fxx[i] += o[i];
        fyy[i] += o[n + i];
        fzz[i] += o[2 * n + i];
    }
    *eng2 += o[3 * n];
    *vir2 += o[3 * n + 1];
    for (int k = 0; k < 9; ++k) stress[k] += o[3 * n + 2 + k];
    return MBPOL_B200_OK;
}

int mbnrg_b200_counts(const mbnrg_b200_handle* h, long long counts[3])
{
    if (!h || !counts) return MBPOL_B200_EINVAL;
    for (int k = 0; k < 3; ++k) counts[k] = h->counts[k];
    return MBPOL_B200_OK;
}

int mbnrg_b200_kernel_times(const mbnrg_b200_handle* h, float ms[4], int* launches)
{
    if (!h || !ms) return MBPOL_B200_EINVAL;
    for (int k = 0; k < 4; ++k) ms[k] = h->ms[k];
    if (launches) *launches = h->launches;
    return MBPOL_B200_OK;
}

// The 3B ion-water-water term (optional: the reference build has it compiled out, mbnrg.f90:4).  Same conventions as
// mbnrg_b200_forces: forces, *eng3, *vir3 and stress[9] are accumulated into; rank r of nranks evaluates the contiguous range
// [T r/N, T (r+1)/N) of the accepted (ion, water i < water j) trimers.
int mbnrg_b200_forces_3b(mbnrg_b200_handle* h, int imcon, const double* cell, int nttm2, const int* listttm2, int nions, const int* listions,
                         int mxatms, const double* xxx, const double* yyy, const double* zzz, double* fxx, double* fyy, double* fzz,
                         double* eng3, double* vir3, double* stress)
{
    if (!h) return MBPOL_B200_EINVAL;
    if (imcon < 1 || imcon > 3 || !cell || nttm2 != h->nttm2 || nions != h->nions || mxatms < nions || !xxx || !yyy || !zzz || !fxx || !fyy ||
        !fzz || !eng3 || !vir3 || !stress || (nttm2 && !listttm2) || (nions && !listions)) {
        h->err = "mbnrg_b200_forces_3b: invalid argument (imcon must be 1, 2 or 3; nttm2 / nions must match init)";
        return MBPOL_B200_EINVAL;
    }
    NRG_TRY(h, cudaSetDevice(h->device));
    const size_t n = mxatms;
    if (n > h->cap_atoms) {
        if (h->xyz) cudaFree(h->xyz);
        if (h->out) cudaFree(h->out);
        h->xyz = h->out = nullptr;
        NRG_TRY(h, cudaMalloc(&h->xyz, 3 * n * sizeof(double)));
        NRG_TRY(h, cudaMalloc(&h->out, (3 * n + NSC) * sizeof(double)));
        h->cap_atoms = n;
    }
    CellInfo C;
    C.imcon = imcon;
    std::copy(cell, cell + 9, C.cell);
    double det;
    host_invert3(cell, C.rcell, det);
    C.a1 = 1.0 / cell[0]; C.b1 = 1.0 / cell[4]; C.c1 = 1.0 / cell[8];
    cudaStream_t s = h->stream;
    const int nw = h->nw;
    NRG_TRY(h, cudaMemcpyAsync(h->xyz, xxx, n * 8, cudaMemcpyHostToDevice, s));
    NRG_TRY(h, cudaMemcpyAsync(h->xyz + n, yyy, n * 8, cudaMemcpyHostToDevice, s));
    NRG_TRY(h, cudaMemcpyAsync(h->xyz + 2 * n, zzz, n * 8, cudaMemcpyHostToDevice, s));
    if (nttm2) NRG_TRY(h, cudaMemcpyAsync(h->list, listttm2, (size_t)nttm2 * 4, cudaMemcpyHostToDevice, s));
    if (nions) NRG_TRY(h, cudaMemcpyAsync(h->listions, listions, (size_t)nions * 4, cudaMemcpyHostToDevice, s));
    NRG_TRY(h, cudaMemsetAsync(h->out, 0, (3 * n + NSC) * sizeof(double), s));
    NRG_TRY(h, cudaMemsetAsync(h->ovf3, 0, sizeof(int), s));
    NRG_TRY(h, cudaEventRecord(h->ev3[0], s));
    long long off_end = 0;
    int ovf = 0;
    if (nw > 1 && nions) {
        k_nrg_gather<<<(nw + 127) / 128, 128, 0, s>>>(nw, h->list, h->xyz, h->xyz + n, h->xyz + 2 * n, C, h->mono);
        k_nrg3_near<<<(nions * 32 + 127) / 128, 128, 0, s>>>(nw, nions, h->mono, h->xyz, (int)n, C, h->near, h->nnear, h->ovf3);
        k_nrg3_offsets<<<1, 32, 0, s>>>(nions, h->nnear, h->off3);
        NRG_TRY(h, cudaEventRecord(h->ev3[1], s));
        // latency-sized workload (a few 1e4 trimers of 25 k FP64 instructions): the grid is a fixed multiple of the SM count and the
        // kernel strides over its range, so no list length has to come back to the host first
        k_nrg3_trimers<<<148 * 4, 128, 0, s>>>(h->rank, h->nranks, nw, nions, h->mono, h->xyz, (int)n, h->list, h->listions, h->code, h->par3,
                                               h->coef3, C, h->near, h->nnear, h->off3, h->out);
        NRG_TRY(h, cudaMemcpyAsync(&off_end, h->off3 + nions, 8, cudaMemcpyDeviceToHost, s));
        NRG_TRY(h, cudaMemcpyAsync(&ovf, h->ovf3, 4, cudaMemcpyDeviceToHost, s));
    } else {
        NRG_TRY(h, cudaEventRecord(h->ev3[1], s));
    }
    NRG_TRY(h, cudaEventRecord(h->ev3[2], s));
    NRG_TRY(h, cudaGetLastError());
    h->host_out.resize(3 * n + NSC);
    NRG_TRY(h, cudaMemcpyAsync(h->host_out.data(), h->out, (3 * n + NSC) * sizeof(double), cudaMemcpyDeviceToHost, s));
    NRG_TRY(h, cudaStreamSynchronize(s));
    if (ovf) { h->err = "more than NRG3_NEAR_CAP waters within 7 A of an ion: unphysical density"; return MBPOL_B200_ECAPACITY; }
    h->counts3[0] = off_end;
    h->counts3[1] = off_end * (h->rank + 1) / h->nranks - off_end * h->rank / h->nranks;
    cudaEventElapsedTime(&h->ms3[0], h->ev3[0], h->ev3[1]);
    cudaEventElapsedTime(&h->ms3[1], h->ev3[1], h->ev3[2]);
    cudaEventElapsedTime(&h->ms3[2], h->ev3[0], h->ev3[2]);
    const double* o = h->host_out.data();
    for (size_t i = 0; i < n; ++i) {
        fxx[i] += o[i];
        fyy[i] += o[n + i];
        fzz[i] += o[2 * n + i];
    }
    *eng3 += o[3 * n];
    *vir3 += o[3 * n + 1];
    for (int k = 0; k < 9; ++k) stress[k] += o[3 * n + 2 + k];
    return MBPOL_B200_OK;
}

int mbnrg_b200_counts_3b(const mbnrg_b200_handle* h, long long counts[2], float ms[3])
{
    if (!h || !counts) return MBPOL_B200_EINVAL;
    counts[0] = h->counts3[0]; counts[1] = h->counts3[1];
    if (ms) for (int k = 0; k < 3; ++k) ms[k] = h->ms3[k];
    return MBPOL_B200_OK;
}

int mbnrg_b200_eval_h2o_h2o_ion(mbnrg_b200_handle* h, int ion, int n, const double* w /* [n][21]: w1, w2, x */, double* E, double* g /* [n][21] */)
{
    if (!h || ion < 1 || ion > MBNRG_NIONS || n < 0 || !w || !E || !g) return MBPOL_B200_EINVAL;
    if (n == 0) return MBPOL_B200_OK;
    NRG_TRY(h, cudaSetDevice(h->device));
    const size_t need = (size_t)n * 43;
    if (need > h->cap_scratch) {
        if (h->scratch) cudaFree(h->scratch);
        h->scratch = nullptr;
        NRG_TRY(h, cudaMalloc(&h->scratch, need * sizeof(double)));
        h->cap_scratch = need;
    }
    double *dw = h->scratch, *dE = dw + (size_t)n * 21, *dg = dE + n;
    NRG_TRY(h, cudaMemcpyAsync(dw, w, (size_t)n * 21 * 8, cudaMemcpyHostToDevice, h->stream));
    k_nrg_eval_3b<<<(n + 127) / 128, 128, 0, h->stream>>>(n, ion, h->par3, h->coef3, dw, dE, dg);
    NRG_TRY(h, cudaGetLastError());
    NRG_TRY(h, cudaMemcpyAsync(E, dE, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaMemcpyAsync(g, dg, (size_t)n * 21 * 8, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaStreamSynchronize(h->stream));
    return MBPOL_B200_OK;
}

static int eval_common(mbnrg_b200_handle* h, int n, int nwd, const double* w, const double* x, double*& dw, double*& dx, double*& dE,
                       double*& dg1, double*& dg2)
{
    const size_t need = (size_t)n * (nwd + 3 + 1 + nwd + 3);
    if (need > h->cap_scratch) {
        if (h->scratch) cudaFree(h->scratch);
        h->scratch = nullptr;
        NRG_TRY(h, cudaMalloc(&h->scratch, need * sizeof(double)));
        h->cap_scratch = need;
    }
    dw = h->scratch; dx = dw + (size_t)n * nwd; dE = dx + (size_t)n * 3; dg1 = dE + n; dg2 = dg1 + (size_t)n * nwd;
    NRG_TRY(h, cudaMemcpyAsync(dw, w, (size_t)n * nwd * 8, cudaMemcpyHostToDevice, h->stream));
    NRG_TRY(h, cudaMemcpyAsync(dx, x, (size_t)n * 3 * 8, cudaMemcpyHostToDevice, h->stream));
    return MBPOL_B200_OK;
}

int mbnrg_b200_eval_h2o_ion(mbnrg_b200_handle* h, int ion, int n, const double* w, const double* x, double* E, double* g1, double* g2)
{
    if (!h || ion < 1 || ion > MBNRG_NIONS || n < 0 || !w || !x || !E || !g1 || !g2) return MBPOL_B200_EINVAL;
    if (n == 0) return MBPOL_B200_OK;
    NRG_TRY(h, cudaSetDevice(h->device));
    double *dw, *dx, *dE, *dg1, *dg2;
    if (int rc = eval_common(h, n, 9, w, x, dw, dx, dE, dg1, dg2)) return rc;
    k_nrg_eval_hi<<<(n + 127) / 128, 128, 0, h->stream>>>(n, ion, h->par, h->coef, dw, dx, dE, dg1, dg2);
    NRG_TRY(h, cudaGetLastError());
    NRG_TRY(h, cudaMemcpyAsync(E, dE, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaMemcpyAsync(g1, dg1, (size_t)n * 72, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaMemcpyAsync(g2, dg2, (size_t)n * 24, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaStreamSynchronize(h->stream));
    return MBPOL_B200_OK;
}

int mbnrg_b200_eval_ion_ion(mbnrg_b200_handle* h, int ion_a, int ion_b, int n, const double* w, const double* x, double* E, double* g1,
                            double* g2)
{
    if (!h || ion_a < 1 || ion_a > MBNRG_NIONS || ion_b < 1 || ion_b > MBNRG_NIONS || n < 0 || !w || !x || !E || !g1 || !g2)
        return MBPOL_B200_EINVAL;
    if (!MBNRG_HAVE_II[9 * (ion_a - 1) + (ion_b - 1)]) {
        h->err = "mbnrg_b200_eval_ion_ion: the reference has no evaluator for this ion pair";
        return MBPOL_B200_EINVAL;
    }
    if (n == 0) return MBPOL_B200_OK;
    NRG_TRY(h, cudaSetDevice(h->device));
    double *dw, *dx, *dE, *dg1, *dg2;
    if (int rc = eval_common(h, n, 3, w, x, dw, dx, dE, dg1, dg2)) return rc;
    k_nrg_eval_ii<<<(n + 127) / 128, 128, 0, h->stream>>>(n, ion_a, ion_b, h->coef_ii, h->par_ii, dw, dx, dE, dg1, dg2);
    NRG_TRY(h, cudaGetLastError());
    NRG_TRY(h, cudaMemcpyAsync(E, dE, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaMemcpyAsync(g1, dg1, (size_t)n * 24, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaMemcpyAsync(g2, dg2, (size_t)n * 24, cudaMemcpyDeviceToHost, h->stream));
    NRG_TRY(h, cudaStreamSynchronize(h->stream));
    return MBPOL_B200_OK;
}

}  // extern "C"
