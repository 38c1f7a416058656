// mbpol_device.cuh -- per-cluster device code: geometry -> variables -> polynomial (coefficient-tape engines) -> chain rule.
//
// B200-native restatement of x2b_v6x_p::operator() (mbpol-2b-poly.cpp:358-682) and x3b_v2x_p::operator()
// (mbpol-3b-poly.cpp:148-312).  One thread owns one cluster.  The polynomial and its gradient are evaluated by one of the
// coefficient-tape engines: 1 = rows (mbpol_tape.cuh), 3 = fat rows (mbpol_fat.cuh), 4 = symmetry-batched fat rows
// (mbpol_sym.cuh, 3B only).
#pragma once
#include <cuda_runtime.h>
#include "mbpol_tape.cuh"
#include "mbpol_fat.cuh"
#include "mbpol_sym.cuh"

struct MbpolParams {   // non-linear parameters, mbpol-2b-poly.cpp:1846-1866, mbpol-3b-poly.cpp:1486-1503
    double k2[8];       // HH_intra, OH_intra, HH_coul, OH_coul, OO_coul, XH, XO, XX
    double gamma_in, gamma_out, r2i, r2f;
    double k3[5], d3[5];   // HH_intra, OH_intra, HH, OH, OO
    double r3i, r3f;
};
extern __constant__ MbpolParams mbpol_par;

#define MBPOL_ENGUNIT 418.4   /* mbpol.f90:37 */
#define MBPOL_PI 3.14159265358979323846

// ---------------------------------------------------------------------------------------------------------
// minimum image, bit-exact with mbpol.f90:1305-1361 (no FMA contraction: explicit _rn intrinsics; nint = round())
// ---------------------------------------------------------------------------------------------------------
struct CellInfo {
    int imcon;
    double cell[9];
    double rcell[9];     // invert.f result (imcon 3)
    double a1, b1, c1;   // 1/cell(1), 1/cell(5), 1/cell(9)
};

__device__ __forceinline__ void image_vec(const CellInfo& C, double& x, double& y, double& z)
{
    if (C.imcon == 1) {
        x = __dsub_rn(x, __dmul_rn(C.cell[0], round(__dmul_rn(x, C.a1))));
        y = __dsub_rn(y, __dmul_rn(C.cell[0], round(__dmul_rn(y, C.a1))));
        z = __dsub_rn(z, __dmul_rn(C.cell[0], round(__dmul_rn(z, C.a1))));
    } else if (C.imcon == 2) {
        x = __dsub_rn(x, __dmul_rn(C.cell[0], round(__dmul_rn(x, C.a1))));
        y = __dsub_rn(y, __dmul_rn(C.cell[4], round(__dmul_rn(y, C.b1))));
        z = __dsub_rn(z, __dmul_rn(C.cell[8], round(__dmul_rn(z, C.c1))));
    } else {
        const double* r = C.rcell;
        const double ssx = __dadd_rn(__dadd_rn(__dmul_rn(r[0], x), __dmul_rn(r[3], y)), __dmul_rn(r[6], z));
        const double ssy = __dadd_rn(__dadd_rn(__dmul_rn(r[1], x), __dmul_rn(r[4], y)), __dmul_rn(r[7], z));
        const double ssz = __dadd_rn(__dadd_rn(__dmul_rn(r[2], x), __dmul_rn(r[5], y)), __dmul_rn(r[8], z));
        const double xs = __dsub_rn(ssx, round(ssx)), ys = __dsub_rn(ssy, round(ssy)), zs = __dsub_rn(ssz, round(ssz));
        const double* c = C.cell;
        x = __dadd_rn(__dadd_rn(__dmul_rn(c[0], xs), __dmul_rn(c[3], ys)), __dmul_rn(c[6], zs));
        y = __dadd_rn(__dadd_rn(__dmul_rn(c[1], xs), __dmul_rn(c[4], ys)), __dmul_rn(c[7], zs));
        z = __dadd_rn(__dadd_rn(__dmul_rn(c[2], xs), __dmul_rn(c[5], ys)), __dmul_rn(c[8], zs));
    }
}

// r^2 exactly as the reference sums it: (x*x + y*y) + z*z without contraction
__device__ __forceinline__ double norm2_exact(double x, double y, double z)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

// cosine switch and its derivative w.r.t. r  (mbpol-2b-poly.cpp:340-354, mbpol-3b-poly.cpp:130-144)
__device__ __forceinline__ double switch_fn(double r, double ri, double rf, double& g)
{
    if (r > rf) { g = 0.0; return 0.0; }
    if (r > ri) {
        const double t1 = MBPOL_PI / (rf - ri);
        double sn, cs;
        sincos((r - ri) * t1, &sn, &cs);
        g = -sn * t1 / 2.0;
        return (1.0 + cs) / 2.0;
    }
    g = 0.0;
    return 1.0;
}

// ---------------------------------------------------------------------------------------------------------
// polynomial front-ends: every engine walks a coefficient tape staged in shared memory and needs the whole CTA
// ---------------------------------------------------------------------------------------------------------
struct PolyCtx {
    void* smem;              // the engine's dynamic shared memory (TapeSmem<NV> / SymSmem)
    const double* tape;      // device coefficient tape of the engine
};

// =========================================================================================================
// 3B
// =========================================================================================================
// atoms: 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 Oc 7 Hc1 8 Hc2 ; variable order of mbpol-3b-poly.cpp:170-206
__device__ constexpr int V3A[36] = {1, 4, 7, 0, 0, 3, 3, 6, 6, 1, 1, 1, 1, 2, 2, 2, 2, 4, 4, 5, 5, 0, 0, 0, 0, 3, 3, 3, 3, 6, 6, 6, 6, 0, 0, 3};
__device__ constexpr int V3B[36] = {2, 5, 8, 1, 2, 4, 5, 7, 8, 4, 5, 7, 8, 4, 5, 7, 8, 7, 8, 7, 8, 4, 5, 7, 8, 1, 2, 7, 8, 1, 2, 4, 5, 3, 6, 6};
__device__ constexpr int V3K[36] = {0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4};

// r[27]: imaged coordinates (any common origin).  Outputs in kcal/mol:
//   E, cv[n] with  dE/dr_a += cv[n]*(r_a - r_b),  dE/dr_b -= ...   for variable n = (a,b); the three switch-gradient
//   terms (mbpol-3b-poly.cpp:292-309) are folded into cv[33..35] because they act along the same O-O vectors.
// variables x_n = exp(-k (d - d0)) and 1/d of the 36 pair distances
__device__ __forceinline__ void trimer_variables(const double (&r)[27], double* xv, double (&rinv)[36])
{
#pragma unroll
    for (int n = 0; n < 36; ++n) {
        const int a = 3 * V3A[n], b = 3 * V3B[n];
        const double dx = r[a] - r[b], dy = r[a + 1] - r[b + 1], dz = r[a + 2] - r[b + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        rinv[n] = 1.0 / d;
        xv[n] = exp(-mbpol_par.k3[V3K[n]] * (d - mbpol_par.d3[V3K[n]]));
    }
}

// the same without 1/d (recomputed after the polynomial by trimer_rinv: 36 + 27 values less to keep alive across it)
__device__ __forceinline__ void trimer_x(const double (&r)[27], double* xv)
{
#pragma unroll
    for (int n = 0; n < 36; ++n) {
        const int a = 3 * V3A[n], b = 3 * V3B[n];
        const double dx = r[a] - r[b], dy = r[a + 1] - r[b + 1], dz = r[a + 2] - r[b + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        xv[n] = exp(-mbpol_par.k3[V3K[n]] * (d - mbpol_par.d3[V3K[n]]));
    }
}
__device__ __forceinline__ void trimer_rinv(const double (&r)[27], double (&rinv)[36])
{
#pragma unroll
    for (int n = 0; n < 36; ++n) {
        const int a = 3 * V3A[n], b = 3 * V3B[n];
        const double dx = r[a] - r[b], dy = r[a + 1] - r[b + 1], dz = r[a + 2] - r[b + 2];
        rinv[n] = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
    }
}

// Phase 1 of the box path (k_trimers): variables and the switch, everything the record of a trimer needs besides the
// polynomial.  sw[0] = s = sab sac + sab sbc + sac sbc, sw[1..3] = ds/dr_ab, ds/dr_ac, ds/dr_bc (mbpol-3b-poly.cpp:231-235, 292-301)
__device__ __forceinline__ void trimer_x_switch(const double (&r)[27], double* xv, double (&sw)[4])
{
    double doo[3];
#pragma unroll
    for (int n = 0; n < 36; ++n) {
        const int a = 3 * V3A[n], b = 3 * V3B[n];
        const double dx = r[a] - r[b], dy = r[a + 1] - r[b + 1], dz = r[a + 2] - r[b + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        if (n >= 33) doo[n - 33] = d;
        xv[n] = exp(-mbpol_par.k3[V3K[n]] * (d - mbpol_par.d3[V3K[n]]));
    }
    double gab, gac, gbc;
    const double sab = switch_fn(doo[0], mbpol_par.r3i, mbpol_par.r3f, gab);
    const double sac = switch_fn(doo[1], mbpol_par.r3i, mbpol_par.r3f, gac);
    const double sbc = switch_fn(doo[2], mbpol_par.r3i, mbpol_par.r3f, gbc);
    sw[0] = sab * sac + sab * sbc + sac * sbc;
    sw[1] = gab * (sac + sbc);
    sw[2] = gac * (sab + sbc);
    sw[3] = gbc * (sab + sac);
}

// The same as a COMPACT LOOP for the box kernel with engine 4 (k_trimers<4>): the 27 coordinates go to the thread's (still
// idle) gradient columns, a 9 x 4 loop computes distance -> exp and writes x_n straight into the thread's x column, where the
// tape walk expects it (sym_eval3b<true>).  The unrolled form above is 5 400 instructions that every warp executes once per
// trimer, i.e. always out of a cold instruction cache (ncu: 35 % of its stall samples are `no_instruction`).
extern __constant__ unsigned c_v3word[36];      // 3 * V3A[n] | 3 * V3B[n] << 8 | V3K[n] << 16
__device__ __forceinline__ void sym_stage_x(const double (&r)[27], double (&sw)[4], SymSmem& sm)
{
    const unsigned xs_t = (unsigned)__cvta_generic_to_shared(&sm.xs[0][threadIdx.x]);
    const unsigned rs_t = xs_t + (unsigned)sizeof(sm.xs);          // r[q] in gradient column q, d_OO in columns 27..29
#pragma unroll
    for (int q = 0; q < 27; ++q) tape_sts(rs_t + q * TAPE_COL, r[q]);
#pragma unroll 1
    for (int n0 = 0; n0 < 36; n0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + u;
            const unsigned word = c_v3word[n];
            const unsigned pa = rs_t + (word & 255u) * TAPE_COL, pb = rs_t + ((word >> 8) & 255u) * TAPE_COL;
            const double dx = tape_lds(pa) - tape_lds(pb), dy = tape_lds(pa + TAPE_COL) - tape_lds(pb + TAPE_COL),
                         dz = tape_lds(pa + 2 * TAPE_COL) - tape_lds(pb + 2 * TAPE_COL);
            const double d = sqrt(dx * dx + dy * dy + dz * dz);
            const unsigned kind = word >> 16;
            tape_sts(xs_t + n * TAPE_COL, exp(-mbpol_par.k3[kind] * (d - mbpol_par.d3[kind])));
            if (n >= 33) tape_sts(rs_t + (n - 6) * TAPE_COL, d);
        }
    }
    double gab, gac, gbc;
    const double sab = switch_fn(tape_lds(rs_t + 27 * TAPE_COL), mbpol_par.r3i, mbpol_par.r3f, gab);
    const double sac = switch_fn(tape_lds(rs_t + 28 * TAPE_COL), mbpol_par.r3i, mbpol_par.r3f, gac);
    const double sbc = switch_fn(tape_lds(rs_t + 29 * TAPE_COL), mbpol_par.r3i, mbpol_par.r3f, gbc);
    sw[0] = sab * sac + sab * sbc + sac * sbc;
    sw[1] = gab * (sac + sbc);
    sw[2] = gac * (sab + sbc);
    sw[3] = gbc * (sab + sac);
}

// The record phase 1 leaves per trimer (MBPOL_REC3 doubles): rec[n] = -k_n (dEpoly/dx_n s) x_n for the 36 variables,
// rec[36] = s Epoly (the 3B energy, kcal/mol), rec[37..39] = Epoly ds/dr_OO.  Phase 2 (k_trimers_gather) only multiplies by
// 1/d_n:  cv[n] = rec[n] / d_n (+ rec[37 + n - 33] / d_n for the three O-O variables) -- trimer_finish split in two.
#define MBPOL_REC3 40
__device__ __forceinline__ void trimer_record(const double* xv, const double* gp, double Epoly, const double (&sw)[4], double (&rec)[MBPOL_REC3])
{
#pragma unroll
    for (int n = 0; n < 36; ++n) rec[n] = -mbpol_par.k3[V3K[n]] * (gp[n] * sw[0]) * xv[n];
    rec[36] = sw[0] * Epoly;
    rec[37] = sw[1] * Epoly;
    rec[38] = sw[2] * Epoly;
    rec[39] = sw[3] * Epoly;
}

// switch + chain rule: (x, dEpoly/dx, Epoly) -> E and the per-variable force coefficients cv
__device__ __forceinline__ void trimer_finish(const double* xv, const double* gp, double Epoly, const double (&rinv)[36], double& E,
                                              double (&cv)[36])
{
    // switch on the three O-O distances: rinv[33..35] = 1/|Oa-Ob|, 1/|Oa-Oc|, 1/|Ob-Oc|
    const double drab = 1.0 / rinv[33], drac = 1.0 / rinv[34], drbc = 1.0 / rinv[35];
    double gab, gac, gbc;
    const double sab = switch_fn(drab, mbpol_par.r3i, mbpol_par.r3f, gab);
    const double sac = switch_fn(drac, mbpol_par.r3i, mbpol_par.r3f, gac);
    const double sbc = switch_fn(drbc, mbpol_par.r3i, mbpol_par.r3f, gbc);
    const double s = sab * sac + sab * sbc + sac * sbc;
#pragma unroll
    for (int n = 0; n < 36; ++n)
        cv[n] = -mbpol_par.k3[V3K[n]] * (gp[n] * s) * xv[n] * rinv[n];
    cv[33] += gab * (sac + sbc) * Epoly * rinv[33];
    cv[34] += gac * (sab + sbc) * Epoly * rinv[34];
    cv[35] += gbc * (sab + sac) * Epoly * rinv[35];
    E = s * Epoly;
}

// polynomial of one trimer with the engine's tape walk (all threads of the CTA must call it)
template <int ENGINE>
__device__ __forceinline__ void poly3b_eval(double (&xv)[36], double (&gp)[36], double& Epoly, const PolyCtx& ctx)
{
    static_assert(ENGINE == 1 || ENGINE == 3 || ENGINE == 4, "unknown polynomial engine");
    if constexpr (ENGINE == 4) sym_eval3b(ctx.tape, xv, gp, Epoly, *reinterpret_cast<SymSmem*>(ctx.smem));
    else if constexpr (ENGINE == 3) fat_eval<3, 36>(ctx.tape, xv, gp, Epoly, *reinterpret_cast<TapeSmem<36>*>(ctx.smem));
    else tape_eval<3, 36>(ctx.tape, xv, gp, Epoly, *reinterpret_cast<TapeSmem<36>*>(ctx.smem));
}

template <int ENGINE>
__device__ __forceinline__ void trimer_eval(const double (&r)[27], double& E, double (&cv)[36], const PolyCtx& ctx)
{
    double xv[36], gp[36], rinv[36], Epoly;
    trimer_variables(r, xv, rinv);
    poly3b_eval<ENGINE>(xv, gp, Epoly, ctx);
    trimer_finish(xv, gp, Epoly, rinv, E, cv);
}

// The engines keep x[36] + g[36] in registers while they walk the tape: instead of spilling the geometry (27 values) and
// the 36 inverse distances across the polynomial, both are recomputed afterwards (`geom` re-reads the monomers with loads
// the compiler cannot merge with the first pass).
template <int ENGINE, class Geom>
__device__ __forceinline__ void trimer_eval_regeom(Geom geom, double (&r)[27], double& E, double (&cv)[36], const PolyCtx& ctx)
{
    double xv[36], gp[36], Epoly;
    geom(r, false);
    trimer_x(r, xv);
    poly3b_eval<ENGINE>(xv, gp, Epoly, ctx);
    double rinv[36];
    geom(r, true);
    trimer_rinv(r, rinv);
    trimer_finish(xv, gp, Epoly, rinv, E, cv);
}

// gradient on the 9 atoms from cv:  g[3*atom + c]
__device__ __forceinline__ void trimer_gradient(const double (&r)[27], const double (&cv)[36], double (&g)[27])
{
#pragma unroll
    for (int i = 0; i < 27; ++i) g[i] = 0.0;
#pragma unroll
    for (int n = 0; n < 36; ++n) {
        const int a = 3 * V3A[n], b = 3 * V3B[n];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double f = cv[n] * (r[a + c] - r[b + c]);
            g[a + c] += f;
            g[b + c] -= f;
        }
    }
}

// =========================================================================================================
// 2B
// =========================================================================================================
// sites: 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 Xa1 7 Xa2 8 Xb1 9 Xb2 ; variable order of mbpol-2b-poly.cpp:415-445
__device__ constexpr int V2A[31] = {1, 4, 0, 0, 3, 3, 1, 1, 2, 2, 0, 0, 3, 3, 0, 6, 6, 7, 7, 8, 8, 9, 9, 0, 0, 3, 3, 6, 6, 7, 7};
__device__ constexpr int V2B[31] = {2, 5, 1, 2, 4, 5, 4, 5, 4, 5, 4, 5, 1, 2, 3, 4, 5, 4, 5, 1, 2, 1, 2, 8, 9, 6, 7, 8, 9, 8, 9};
__device__ constexpr int V2K[31] = {0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 5, 5, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7};

// lone-pair sites of one monomer (monomer::setup, mbpol-2b-poly.cpp:118-140); ohh -> x1, x2
__device__ __forceinline__ void lonepair_sites(const double* ohh, double* x1, double* x2)
{
    const double a0 = ohh[3] - ohh[0], a1 = ohh[4] - ohh[1], a2 = ohh[5] - ohh[2];
    const double b0 = ohh[6] - ohh[0], b1 = ohh[7] - ohh[1], b2 = ohh[8] - ohh[2];
    const double v[3] = {a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0};
    const double s[3] = {a0 + b0, a1 + b1, a2 + b2};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double in_plane = ohh[i] + 0.5 * mbpol_par.gamma_in * s[i];
        const double out_of_plane = mbpol_par.gamma_out * v[i];
        x1[i] = in_plane + out_of_plane;
        x2[i] = in_plane - out_of_plane;
    }
}

// fold a pair of lone-pair vectors (g1 on X1, g2 on X2) onto O,H1,H2 (+=) -- monomer::grads / distribute_X*
// (mbpol-2b-poly.cpp:144-177, 193-301).  dst triples are `stride` doubles apart.
__device__ __forceinline__ void lonepair_spread(const double* ohh, const double* g1, const double* g2, double* dst, int stride)
{
    const double oh1[3] = {ohh[3] - ohh[0], ohh[4] - ohh[1], ohh[5] - ohh[2]};
    const double oh2[3] = {ohh[6] - ohh[0], ohh[7] - ohh[1], ohh[8] - ohh[2]};
    const double gm[3] = {g1[0] - g2[0], g1[1] - g2[1], g1[2] - g2[2]};
    const double t1[3] = {oh2[1] * gm[2] - oh2[2] * gm[1], oh2[2] * gm[0] - oh2[0] * gm[2], oh2[0] * gm[1] - oh2[1] * gm[0]};
    const double t2[3] = {oh1[1] * gm[2] - oh1[2] * gm[1], oh1[2] * gm[0] - oh1[0] * gm[2], oh1[0] * gm[1] - oh1[1] * gm[0]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double gsum = g1[i] + g2[i];
        const double in_plane = 0.5 * mbpol_par.gamma_in * gsum;
        const double gh1 = in_plane + mbpol_par.gamma_out * t1[i];
        const double gh2 = in_plane - mbpol_par.gamma_out * t2[i];
        dst[i] += gsum - (gh1 + gh2);
        dst[i + stride] += gh1;
        dst[i + 2 * stride] += gh2;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Box path, phase 1 (k_dimers): everything of x2b_v6x_p::operator() up to the per-variable force coefficients.
// rec[n] = sw * dEpoly/dx_n * (dx_n/dr)/r  for the 31 variables: the pair vector of variable n = (p, q) is rec[n] (x_p - x_q);
// rec[31] = sw * Epoly (the 2B energy, kcal/mol), rec[32] = Epoly * dsw/dr / rOO (switch gradient along Oa - Ob), rec[33] = 0.
// Phase 2 (k_dimers_gather) folds these onto the atoms (dimer_rows_a).  A pair beyond r2f gives a zero record, like the
// reference's early return (mbpol-2b-poly.cpp:370-371).
// ---------------------------------------------------------------------------------------------------------
#define MBPOL_REC2 34
__device__ __forceinline__ void dimer_sites(const double* w, double (&xc)[30])
{
#pragma unroll
    for (int i = 0; i < 18; ++i) xc[i] = w[i];
    lonepair_sites(xc + 0, xc + 18, xc + 21);
    lonepair_sites(xc + 9, xc + 24, xc + 27);
}

template <int ENGINE>
__device__ __forceinline__ void dimer_record(const double* w, double (&rec)[MBPOL_REC2], const PolyCtx& ctx)
{
    double xc[30];
    dimer_sites(w, xc);
    const double dOO[3] = {w[0] - w[9], w[1] - w[10], w[2] - w[11]};
    const double rOO = sqrt(dOO[0] * dOO[0] + dOO[1] * dOO[1] + dOO[2] * dOO[2]);
    double xv[31], gp[31], gfac[31];
#pragma unroll
    for (int n = 0; n < 31; ++n) {
        const int a = 3 * V2A[n], b = 3 * V2B[n];
        const double dx = xc[a] - xc[b], dy = xc[a + 1] - xc[b + 1], dz = xc[a + 2] - xc[b + 2];
        const double rr = sqrt(dx * dx + dy * dy + dz * dz);
        const double k = mbpol_par.k2[V2K[n]];
        const double r0 = (V2K[n] <= 1) ? 1.0 : 4.0;
        const double e1 = exp(k * (r0 - rr));
        const double ri = 1.0 / rr;
        if (V2K[n] >= 2 && V2K[n] <= 4) {   // "coulomb" type: exp(k(r0-r))/r
            const double val = e1 * ri;
            xv[n] = val;
            gfac[n] = -(k + ri) * val * ri;
        } else {
            xv[n] = e1;
            gfac[n] = -k * e1 * ri;
        }
    }
    double gsw;
    const double sw = switch_fn(rOO, mbpol_par.r2i, mbpol_par.r2f, gsw);
    const double gsr = gsw / rOO;
    double Epoly;
    static_assert(ENGINE == 1 || ENGINE == 3, "unknown polynomial engine");
    if constexpr (ENGINE == 3) fat_eval<2, 31>(ctx.tape, xv, gp, Epoly, *reinterpret_cast<TapeSmem<31>*>(ctx.smem));
    else tape_eval<2, 31>(ctx.tape, xv, gp, Epoly, *reinterpret_cast<TapeSmem<31>*>(ctx.smem));
    const bool inside = !(rOO > mbpol_par.r2f);
#pragma unroll
    for (int n = 0; n < 31; ++n) rec[n] = inside ? sw * gp[n] * gfac[n] : 0.0;
    rec[31] = inside ? sw * Epoly : 0.0;
    rec[32] = inside ? gsr * Epoly : 0.0;
    rec[33] = 0.0;
}

// Phase 2: rows i = Oa, Ha1, Ha2 of the reference's folded pair force matrix FM[i][j], j = Oa Ha1 Ha2 Ob Hb1 Hb2
// (mbpol-2b-poly.cpp:503-653, 665-679), from the per-variable coefficients c[31] (already x switch) of a dimer whose molecule a
// is the gathering water.  Lone-pair contributions: the X sites of a are spread over a's atoms with the monomer::grads
// formulas (linear, so all pair vectors towards one column are summed first); the X sites of b are lumped onto Ob.  The matrix
// is NOT antisymmetric; the row sums are the gradient on the atoms of a.
__device__ __forceinline__ void dimer_rows_a(const double (&xc)[30], const double (&c)[31], double gswE, double (&fm)[3][6][3])
{
    double gx[2][3][3];      // pair vectors on Xa1 / Xa2 towards the columns Ob, Hb1, Hb2
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) fm[i][j][k] = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) gx[s][j][k] = 0.0;
#pragma unroll
    for (int n = 0; n < 31; ++n) {
        const int p = V2A[n], q = V2B[n];
        if (p >= 3 && p < 6 && q >= 3 && q < 6) continue;                 // intramolecular b: no row of a
        double d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = c[n] * (xc[3 * p + k] - xc[3 * q + k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (p < 6 && q < 6) {                                          // atom - atom
                if (p < 3) fm[p][q][k] += d[k];
                if (q < 3) fm[q][p][k] -= d[k];
            } else if (p >= 6 && p < 8 && q >= 3 && q < 6) gx[p - 6][q - 3][k] += d[k];      // Xa - (Hb): tmp[Xa][q] = d
            else if (p == 3 && q >= 6 && q < 8) gx[q - 6][0][k] -= d[k];                     // Ob - Xa: tmp[Xa][Ob] = -d
            else if (p >= 8 && q < 3) fm[q][3][k] -= d[k];                                   // Xb - Ha: tmp[Ha][Xb] = -d, lumped onto Ob
            else if (p < 3 && q >= 8) fm[p][3][k] += d[k];                                   // Oa - Xb: tmp[Oa][Xb] = d, lumped onto Ob
            else if (p >= 6 && p < 8 && q >= 8) gx[p - 6][0][k] += d[k];                     // Xa - Xb: spread over a, column Ob
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        lonepair_spread(&xc[0], gx[0][j], gx[1][j], acc, 3);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) fm[i][3 + j][k] += acc[3 * i + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) fm[0][3][k] += gswE * (xc[k] - xc[9 + k]);
}

// w[18]: imaged O,H,H of water a then water b (any common origin).  Outputs in kcal/mol: E, g[18];
// fm (optional, 108 doubles, index 18*i+3*j+c) = the reference's folded pair force matrix with zero diagonal blocks.
// Returns false (E = 0, g = 0) when rOO > r2f, like the reference's early return (mbpol-2b-poly.cpp:370-371).
template <int ENGINE>
__device__ __noinline__ bool dimer_eval(const double* w, double& E, double* g, double* fm, const PolyCtx& ctx)
{
    double xc[30];
#pragma unroll
    for (int i = 0; i < 18; ++i) xc[i] = w[i];
    const double dOO[3] = {w[0] - w[9], w[1] - w[10], w[2] - w[11]};
    const double rOO = sqrt(dOO[0] * dOO[0] + dOO[1] * dOO[1] + dOO[2] * dOO[2]);
    E = 0.0;
#pragma unroll
    for (int i = 0; i < 18; ++i) g[i] = 0.0;
    if (fm)
        for (int i = 0; i < 108; ++i) fm[i] = 0.0;
    const bool inside = !(rOO > mbpol_par.r2f);     // the tape walk has CTA barriers: every thread walks on regardless
    lonepair_sites(xc + 0, xc + 18, xc + 21);
    lonepair_sites(xc + 9, xc + 24, xc + 27);

    double xv[31], gp[31];
    double gfac[31];    // dv/dr / r : the variable's gradient is gfac * (r_a - r_b)
#pragma unroll
    for (int n = 0; n < 31; ++n) {
        const int a = 3 * V2A[n], b = 3 * V2B[n];
        const double dx = xc[a] - xc[b], dy = xc[a + 1] - xc[b + 1], dz = xc[a + 2] - xc[b + 2];
        const double rr = sqrt(dx * dx + dy * dy + dz * dz);
        const double k = mbpol_par.k2[V2K[n]];
        const double r0 = (V2K[n] <= 1) ? 1.0 : 4.0;
        const double e1 = exp(k * (r0 - rr));
        const double ri = 1.0 / rr;
        if (V2K[n] >= 2 && V2K[n] <= 4) {   // "coulomb" type: exp(k(r0-r))/r
            const double val = e1 * ri;
            xv[n] = val;
            gfac[n] = -(k + ri) * val * ri;
        } else {
            xv[n] = e1;
            gfac[n] = -k * e1 * ri;
        }
    }
    double Epoly;
    static_assert(ENGINE == 1 || ENGINE == 3, "unknown polynomial engine");
    if constexpr (ENGINE == 3) fat_eval<2, 31>(ctx.tape, xv, gp, Epoly, *reinterpret_cast<TapeSmem<31>*>(ctx.smem));
    else tape_eval<2, 31>(ctx.tape, xv, gp, Epoly, *reinterpret_cast<TapeSmem<31>*>(ctx.smem));
    if (!inside) return false;

    double xg[30];
#pragma unroll
    for (int i = 0; i < 30; ++i) xg[i] = 0.0;
    double tmp[300];   // 10 x 10 x 3 scratch (only when the pair matrix is requested)
    if (fm)
        for (int i = 0; i < 300; ++i) tmp[i] = 0.0;
#pragma unroll
    for (int n = 0; n < 31; ++n) {
        const int a = V2A[n], b = V2B[n];
        const double c = gp[n] * gfac[n];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double d = c * (xc[3 * a + i] - xc[3 * b + i]);
            xg[3 * a + i] += d;
            xg[3 * b + i] -= d;
            if (fm) {
                tmp[30 * a + 3 * b + i] = d;
                tmp[30 * b + 3 * a + i] = -d;
            }
        }
    }
    lonepair_spread(xc + 0, xg + 18, xg + 21, xg + 0, 3);
    lonepair_spread(xc + 9, xg + 24, xg + 27, xg + 9, 3);

    double gsw;
    const double sw = switch_fn(rOO, mbpol_par.r2i, mbpol_par.r2f, gsw);
    gsw *= Epoly / rOO;
#pragma unroll
    for (int i = 0; i < 18; ++i) g[i] = sw * xg[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g[i] += gsw * dOO[i];
        g[9 + i] -= gsw * dOO[i];
    }
    E = sw * Epoly;

    if (fm) {   // same folding order as mbpol-2b-poly.cpp:595-642
#define T_(i, j) (tmp + 30 * (i) + 3 * (j))
        lonepair_spread(xc + 0, T_(6, 3), T_(7, 3), T_(0, 3), 30);
        lonepair_spread(xc + 0, T_(6, 4), T_(7, 4), T_(0, 4), 30);
        lonepair_spread(xc + 0, T_(6, 5), T_(7, 5), T_(0, 5), 30);
        lonepair_spread(xc + 9, T_(8, 0), T_(9, 0), T_(3, 0), 30);
        lonepair_spread(xc + 9, T_(8, 1), T_(9, 1), T_(3, 1), 30);
        lonepair_spread(xc + 9, T_(8, 2), T_(9, 2), T_(3, 2), 30);
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) {
                T_(i, 3)[c] += T_(i, 8)[c];
                T_(i, 3)[c] += T_(i, 9)[c];
                T_(3 + i, 0)[c] += T_(3 + i, 6)[c];
                T_(3 + i, 0)[c] += T_(3 + i, 7)[c];
            }
        lonepair_spread(xc + 0, T_(6, 8), T_(7, 8), T_(0, 3), 30);
        lonepair_spread(xc + 0, T_(6, 9), T_(7, 9), T_(0, 3), 30);
        lonepair_spread(xc + 9, T_(8, 6), T_(9, 6), T_(3, 0), 30);
        lonepair_spread(xc + 9, T_(8, 7), T_(9, 7), T_(3, 0), 30);
#undef T_
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
                for (int c = 0; c < 3; ++c) fm[18 * i + 3 * j + c] = (i == j) ? 0.0 : sw * tmp[30 * i + 3 * j + c];
        for (int c = 0; c < 3; ++c) {
            fm[18 * 0 + 3 * 3 + c] += gsw * dOO[c];
            fm[18 * 3 + 3 * 0 + c] -= gsw * dOO[c];
        }
    }
    return true;
}
