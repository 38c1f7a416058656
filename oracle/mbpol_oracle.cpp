// mbpol_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the reference's MB-pol 2B+3B hot path, used exclusively as the
// checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Nothing in the product library (libmbpol_b200.so) includes, links or calls this file.
//
// Parity status: PINNED.  (1) The restated evaluators are compared with the reference's own
// objects (oracle/_ref/libmbpol_ref*.so, compiled unmodified from /root/reference/SOURCE by
// oracle/Makefile) and (2) with the known-answer vectors of SURVEY.md Appendix A, committed as
// tests/golden/kat_appendix_a.json, in tests/test_oracle.py.
//
// What is restated (file:line are relative to /root/reference/SOURCE):
//   oracle_poly{2b,3b}        E = sum_k a_k * orbit-sum(x), dE/dx   poly-2b-v6x.cpp:5-13843, poly-3b-v2x.cpp:5-28380
//                             (as a table-driven sum over the extracted monomials, not the Maple schedule)
//   oracle_mbpol_2b           x2b_v6x_p::operator()                 mbpol-2b-poly.cpp:358-682 (+ :24-85, :118-301, :340-354)
//   oracle_mbpol_3b           x3b_v2x_p::operator()                 mbpol-3b-poly.cpp:148-312 (+ :12-82, :130-144)
//   image_vectors             mbpol.f90:1305-1361  (imcon 1,2,3; invert.f for 3)
//   image_monomers            mbpol.f90:1275-1301
//   do_dimer / do_trimer      mbpol.f90:453-790 / 794-1271 (cutoff tests, in-place re-imaging, engunit,
//                             E/6 and E/9 per-atom energy, update_forces, HEAT_STRESS per-atom stress, strs, virial)
//   oracle_mbpol_forces       mbpol.f90:98-449 (gather, pair / triple loops strided by bead_rank/bead_size)
//   heat-flux contractions    heatcurrent.f90:464-483 (compute_heat_flux), :703-728 (compute_heat_flux2)
//
// Two driver modes:
//   faithful : the reference's O(N^2)/O(N^3) enumeration with in-place mutation of `monomers`
//   listed   : neighbour-list enumeration with the same predicates on un-mutated coordinates (canonical
//              semantics for the GPU path; identical up to +-1 ulp on O-O vectors)
//
// Arithmetic: no FMA contraction (build with -ffp-contract=off on x86-64 baseline).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <thread>
#include <vector>

#include "oracle_model.inc"

namespace {

const double engunit = 418.4;  // mbpol.f90:37

// ------------------------------------------------------------------------------------------------
// polynomials (table driven)
// ------------------------------------------------------------------------------------------------
template <int NV>
double poly_eval(const short (*mono)[5], int nmono, const double* a, const double* x, double* g)
{
    double E = 0.0;
    for (int i = 0; i < NV; ++i) g[i] = 0.0;
    for (int m = 0; m < nmono; ++m) {
        const short* r = mono[m];
        const double c = a[r[0]];
        int v[4], n = 0;
        for (int q = 1; q < 5; ++q)
            if (r[q] >= 0) v[n++] = r[q];
        double prod = c;
        for (int q = 0; q < n; ++q) prod *= x[v[q]];
        E += prod;
        for (int q = 0; q < n; ++q) {
            double d = c;
            for (int s = 0; s < n; ++s)
                if (s != q) d *= x[v[s]];
            g[v[q]] += d;
        }
    }
    return E;
}

// ------------------------------------------------------------------------------------------------
// optional: the reference's own evaluators, loaded from oracle/_ref (never part of the product)
// ------------------------------------------------------------------------------------------------
typedef void (*ref2b_t)(const double*, const double*, double*, double*, double*, double*);
typedef void (*ref3b_t)(const double*, const double*, const double*, double*, double*, double*, double*, double*);
ref2b_t ref_2b = nullptr;
ref3b_t ref_3b = nullptr;

// ------------------------------------------------------------------------------------------------
// 2B evaluator (mbpol-2b-poly.cpp)
// ------------------------------------------------------------------------------------------------
struct Fold {   // lone-pair geometry of one monomer, mbpol-2b-poly.cpp:87-177
    double oh1[3], oh2[3];
    void setup(const double* ohh, double* x1, double* x2)
    {
        for (int i = 0; i < 3; ++i) {
            oh1[i] = ohh[i + 3] - ohh[i];
            oh2[i] = ohh[i + 6] - ohh[i];
        }
        const double v[3] = {oh1[1] * oh2[2] - oh1[2] * oh2[1], oh1[2] * oh2[0] - oh1[0] * oh2[2],
                             oh1[0] * oh2[1] - oh1[1] * oh2[0]};
        for (int i = 0; i < 3; ++i) {
            const double in_plane = ohh[i] + 0.5 * ORC_P2b_in_plane_gamma * (oh1[i] + oh2[i]);
            const double out_of_plane = ORC_P2b_out_of_plane_gamma * v[i];
            x1[i] = in_plane + out_of_plane;
            x2[i] = in_plane - out_of_plane;
        }
    }
    // distribute a pair of lone-pair vectors onto O,H1,H2 (+=); strideO = distance between the three
    // destination triples in doubles (3 for gradients, 30 for rows of the 10x10x3 scratch matrix)
    void spread(const double* g1, const double* g2, double* dst, int stride) const
    {
        const double gm[3] = {g1[0] - g2[0], g1[1] - g2[1], g1[2] - g2[2]};
        const double t1[3] = {oh2[1] * gm[2] - oh2[2] * gm[1], oh2[2] * gm[0] - oh2[0] * gm[2],
                              oh2[0] * gm[1] - oh2[1] * gm[0]};
        const double t2[3] = {oh1[1] * gm[2] - oh1[2] * gm[1], oh1[2] * gm[0] - oh1[0] * gm[2],
                              oh1[0] * gm[1] - oh1[1] * gm[0]};
        for (int i = 0; i < 3; ++i) {
            const double gsum = g1[i] + g2[i];
            const double in_plane = 0.5 * ORC_P2b_in_plane_gamma * gsum;
            const double gh1 = in_plane + ORC_P2b_out_of_plane_gamma * t1[i];
            const double gh2 = in_plane - ORC_P2b_out_of_plane_gamma * t2[i];
            dst[i + 0 * stride] += gsum - (gh1 + gh2);
            dst[i + 1 * stride] += gh1;
            dst[i + 2 * stride] += gh2;
        }
    }
};

double switch_fn(double r, double ri, double rf, double& g)   // mbpol-2b-poly.cpp:340-354, mbpol-3b-poly.cpp:130-144
{
    if (r > rf) {
        g = 0.0;
        return 0.0;
    } else if (r > ri) {
        const double t1 = M_PI / (rf - ri);
        const double x = (r - ri) * t1;
        g = -std::sin(x) * t1 / 2.0;
        return (1.0 + std::cos(x)) / 2.0;
    }
    g = 0.0;
    return 1.0;
}

// sites: 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 Xa1 7 Xa2 8 Xb1 9 Xb2   (mbpol-2b-poly.cpp:374-387)
struct Var2 { int coul; int kind; int a, b; };
enum { K_HH_INTRA, K_OH_INTRA, K_HH_COUL, K_OH_COUL, K_OO_COUL, K_XH, K_XO, K_XX };
const Var2 VARS2[31] = {   // mbpol-2b-poly.cpp:415-445
    {0, K_HH_INTRA, 1, 2}, {0, K_HH_INTRA, 4, 5}, {0, K_OH_INTRA, 0, 1}, {0, K_OH_INTRA, 0, 2},
    {0, K_OH_INTRA, 3, 4}, {0, K_OH_INTRA, 3, 5},
    {1, K_HH_COUL, 1, 4}, {1, K_HH_COUL, 1, 5}, {1, K_HH_COUL, 2, 4}, {1, K_HH_COUL, 2, 5},
    {1, K_OH_COUL, 0, 4}, {1, K_OH_COUL, 0, 5}, {1, K_OH_COUL, 3, 1}, {1, K_OH_COUL, 3, 2},
    {1, K_OO_COUL, 0, 3},
    {0, K_XH, 6, 4}, {0, K_XH, 6, 5}, {0, K_XH, 7, 4}, {0, K_XH, 7, 5},
    {0, K_XH, 8, 1}, {0, K_XH, 8, 2}, {0, K_XH, 9, 1}, {0, K_XH, 9, 2},
    {0, K_XO, 0, 8}, {0, K_XO, 0, 9}, {0, K_XO, 3, 6}, {0, K_XO, 3, 7},
    {0, K_XX, 6, 8}, {0, K_XX, 6, 9}, {0, K_XX, 7, 8}, {0, K_XX, 7, 9}};

double k2_of(int kind)
{
    switch (kind) {
    case K_HH_INTRA: return ORC_P2b_k_HH_intra;
    case K_OH_INTRA: return ORC_P2b_k_OH_intra;
    case K_HH_COUL: return ORC_P2b_k_HH_coul;
    case K_OH_COUL: return ORC_P2b_k_OH_coul;
    case K_OO_COUL: return ORC_P2b_k_OO_coul;
    case K_XH: return ORC_P2b_k_XH_main;
    case K_XO: return ORC_P2b_k_XO_main;
    default: return ORC_P2b_k_XX_main;
    }
}

// returns 0 and leaves outputs zeroed when rOO > r2f (the reference returns 0.0 without touching
// g1,g2,force_matrix, mbpol-2b-poly.cpp:370-371; its caller never gets there, mbpol.f90:502-503)
double eval_2b(const double* w1, const double* w2, double* g1, double* g2, double* fm)
{
    for (int i = 0; i < 9; ++i) g1[i] = g2[i] = 0.0;
    for (int i = 0; i < 108; ++i) fm[i] = 0.0;
    const double dOO[3] = {w1[0] - w2[0], w1[1] - w2[1], w1[2] - w2[2]};
    const double rOOsq = dOO[0] * dOO[0] + dOO[1] * dOO[1] + dOO[2] * dOO[2];
    const double rOO = std::sqrt(rOOsq);
    if (rOO > ORC_P2b_r2f) return 0.0;

    double xc[30];
    std::copy(w1, w1 + 9, xc);
    std::copy(w2, w2 + 9, xc + 9);
    Fold ma, mb;
    ma.setup(xc + 0, xc + 18, xc + 21);
    mb.setup(xc + 9, xc + 24, xc + 27);

    double v[31], dv[31][3];
    for (int n = 0; n < 31; ++n) {          // variable::v_exp / v_coul, mbpol-2b-poly.cpp:24-66
        const Var2& q = VARS2[n];
        const double k = k2_of(q.kind);
        const double r0 = (q.kind == K_HH_INTRA || q.kind == K_OH_INTRA) ? 1.0 : 4.0;
        double* g = dv[n];
        for (int i = 0; i < 3; ++i) g[i] = xc[3 * q.a + i] - xc[3 * q.b + i];
        if (!q.coul) {
            const double r = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
            const double exp1 = std::exp(k * (r0 - r));
            const double gg = -k * exp1 / r;
            for (int i = 0; i < 3; ++i) g[i] *= gg;
            v[n] = exp1;
        } else {
            const double rsq = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
            const double r = std::sqrt(rsq);
            const double exp1 = std::exp(k * (r0 - r));
            const double rinv = 1.0 / r;
            const double val = exp1 * rinv;
            const double gg = -(k + rinv) * val * rinv;
            for (int i = 0; i < 3; ++i) g[i] *= gg;
            v[n] = val;
        }
    }
    double gp[31];
    const double E_poly = poly_eval<31>(ORC_MONO2B, ORC_NMONO2B, ORC_COEF2B, v, gp);

    double xg[30];
    std::fill(xg, xg + 30, 0.0);
    static thread_local double tmp[300];        // 10 x 10 x 3 scratch, index 30*i + 3*j + c  (mbpol-2b-poly.cpp:72-85, 454)
    std::fill(tmp, tmp + 300, 0.0);             // the reference leaves it uninitialised; only assigned entries are read
    for (int n = 0; n < 31; ++n) {              // variable::grads
        const Var2& q = VARS2[n];
        for (int i = 0; i < 3; ++i) {
            const double d = gp[n] * dv[n][i];
            xg[3 * q.a + i] += d;
            xg[3 * q.b + i] -= d;
            tmp[30 * q.a + 3 * q.b + i] = d;
            tmp[30 * q.b + 3 * q.a + i] = -d;
        }
    }
    ma.spread(xg + 18, xg + 21, xg + 0, 3);     // monomer::grads, mbpol-2b-poly.cpp:491-498
    mb.spread(xg + 24, xg + 27, xg + 9, 3);

    auto T = [&](int i, int j) { return tmp + 30 * i + 3 * j; };
    // row-side folding of lone-pair rows, mbpol-2b-poly.cpp:595-613
    ma.spread(T(6, 3), T(7, 3), T(0, 3), 30);
    ma.spread(T(6, 4), T(7, 4), T(0, 4), 30);
    ma.spread(T(6, 5), T(7, 5), T(0, 5), 30);
    mb.spread(T(8, 0), T(9, 0), T(3, 0), 30);
    mb.spread(T(8, 1), T(9, 1), T(3, 1), 30);
    mb.spread(T(8, 2), T(9, 2), T(3, 2), 30);
    // column-side: lone-pair columns lumped onto the owning O, mbpol-2b-poly.cpp:617-628
    auto AX = [&](int i, int jx, int jo) {
        for (int c = 0; c < 3; ++c) {
            T(i, jo)[c] += T(i, jx)[c];
            T(i, jx)[c] = 0;
        }
    };
    AX(0, 8, 3); AX(0, 9, 3); AX(1, 8, 3); AX(1, 9, 3); AX(2, 8, 3); AX(2, 9, 3);
    AX(3, 6, 0); AX(3, 7, 0); AX(4, 6, 0); AX(4, 7, 0); AX(5, 6, 0); AX(5, 7, 0);
    // X-X terms, mbpol-2b-poly.cpp:630-642
    ma.spread(T(6, 8), T(7, 8), T(0, 3), 30);
    ma.spread(T(6, 9), T(7, 9), T(0, 3), 30);
    mb.spread(T(8, 6), T(9, 6), T(3, 0), 30);
    mb.spread(T(8, 7), T(9, 7), T(3, 0), 30);

    double gsw;
    const double sw = switch_fn(rOO, ORC_P2b_r2i, ORC_P2b_r2f, gsw);
    for (int i = 0; i < 9; ++i) {
        g1[i] = sw * xg[i];
        g2[i] = sw * xg[i + 9];
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j)
            for (int c = 0; c < 3; ++c) fm[18 * i + 3 * j + c] = (i == j) ? 0.0 : sw * tmp[30 * i + 3 * j + c];
    gsw *= E_poly / rOO;
    for (int i = 0; i < 3; ++i) {
        const double d = gsw * dOO[i];
        g1[i] += d;
        g2[i] -= d;
        fm[i + 18 * 0 + 3 * 3] += d;
        fm[i + 18 * 3 + 3 * 0] -= d;
    }
    return sw * E_poly;
}

// ------------------------------------------------------------------------------------------------
// 3B evaluator (mbpol-3b-poly.cpp)
// ------------------------------------------------------------------------------------------------
// atoms: 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 Oc 7 Hc1 8 Hc2
struct Var3 { int kind; int a, b; };
enum { T_HH_INTRA, T_OH_INTRA, T_HH, T_OH, T_OO };
const Var3 VARS3[36] = {   // mbpol-3b-poly.cpp:170-206
    {T_HH_INTRA, 1, 2}, {T_HH_INTRA, 4, 5}, {T_HH_INTRA, 7, 8},
    {T_OH_INTRA, 0, 1}, {T_OH_INTRA, 0, 2}, {T_OH_INTRA, 3, 4}, {T_OH_INTRA, 3, 5}, {T_OH_INTRA, 6, 7}, {T_OH_INTRA, 6, 8},
    {T_HH, 1, 4}, {T_HH, 1, 5}, {T_HH, 1, 7}, {T_HH, 1, 8}, {T_HH, 2, 4}, {T_HH, 2, 5}, {T_HH, 2, 7}, {T_HH, 2, 8},
    {T_HH, 4, 7}, {T_HH, 4, 8}, {T_HH, 5, 7}, {T_HH, 5, 8},
    {T_OH, 0, 4}, {T_OH, 0, 5}, {T_OH, 0, 7}, {T_OH, 0, 8}, {T_OH, 3, 1}, {T_OH, 3, 2}, {T_OH, 3, 7}, {T_OH, 3, 8},
    {T_OH, 6, 1}, {T_OH, 6, 2}, {T_OH, 6, 4}, {T_OH, 6, 5},
    {T_OO, 0, 3}, {T_OO, 0, 6}, {T_OO, 3, 6}};

void kd3(int kind, double& k, double& d0)
{
    switch (kind) {
    case T_HH_INTRA: k = ORC_P3b_kHH_intra; d0 = ORC_P3b_dHH_intra; break;
    case T_OH_INTRA: k = ORC_P3b_kOH_intra; d0 = ORC_P3b_dOH_intra; break;
    case T_HH: k = ORC_P3b_kHH; d0 = ORC_P3b_dHH; break;
    case T_OH: k = ORC_P3b_kOH; d0 = ORC_P3b_dOH; break;
    default: k = ORC_P3b_kOO; d0 = ORC_P3b_dOO; break;
    }
}

// g1,g2,g3 are accumulated into (+=) exactly like the reference (mbpol-3b-poly.cpp:39-40, 77-78)
double eval_3b(const double* w1, const double* w2, const double* w3, double* g1, double* g2, double* g3, double* fm)
{
    double xyz[27];
    std::copy(w1, w1 + 9, xyz);
    std::copy(w2, w2 + 9, xyz + 9);
    std::copy(w3, w3 + 9, xyz + 18);
    double* grd[9] = {g1, g1 + 3, g1 + 6, g2, g2 + 3, g2 + 6, g3, g3 + 3, g3 + 6};
    for (int i = 0; i < 243; ++i) fm[i] = 0.0;

    double x[36];
    for (int n = 0; n < 36; ++n) {            // v_intra / v_main
        double k, r0;
        kd3(VARS3[n].kind, k, r0);
        const double* a1 = xyz + 3 * VARS3[n].a;
        const double* a2 = xyz + 3 * VARS3[n].b;
        const double dx[3] = {a1[0] - a2[0], a1[1] - a2[1], a1[2] - a2[2]};
        const double dsq = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        const double d = std::sqrt(dsq);
        x[n] = std::exp(-k * (d - r0));
    }
    double g[36];
    double retval = poly_eval<36>(ORC_MONO3B, ORC_NMONO3B, ORC_COEF3B, x, g);

    const double *Oa = xyz, *Ob = xyz + 9, *Oc = xyz + 18;
    double rab[3], rac[3], rbc[3];
    double drab = 0, drac = 0, drbc = 0;
    for (int n = 0; n < 3; ++n) {
        rab[n] = Oa[n] - Ob[n];
        drab += rab[n] * rab[n];
        rac[n] = Oa[n] - Oc[n];
        drac += rac[n] * rac[n];
        rbc[n] = Ob[n] - Oc[n];
        drbc += rbc[n] * rbc[n];
    }
    drab = std::sqrt(drab);
    drac = std::sqrt(drac);
    drbc = std::sqrt(drbc);
    double gab, gac, gbc;
    const double sab = switch_fn(drab, ORC_P3b_r3i, ORC_P3b_r3f, gab);
    const double sac = switch_fn(drac, ORC_P3b_r3i, ORC_P3b_r3f, gac);
    const double sbc = switch_fn(drbc, ORC_P3b_r3i, ORC_P3b_r3f, gbc);
    const double s = sab * sac + sab * sbc + sac * sbc;
    for (int n = 0; n < 36; ++n) g[n] *= s;

    for (int n = 0; n < 36; ++n) {            // g_intra / g_main, mbpol-3b-poly.cpp:26-82, 252-288
        double k, r0;
        kd3(VARS3[n].kind, k, r0);
        const int ia = VARS3[n].a, ib = VARS3[n].b;
        const double* a1 = xyz + 3 * ia;
        const double* a2 = xyz + 3 * ib;
        const double dx[3] = {a1[0] - a2[0], a1[1] - a2[1], a1[2] - a2[2]};
        const double dsq = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        const double d = std::sqrt(dsq);
        const double gg = -k * g[n] * std::exp(-k * (d - r0)) / d;
        for (int i = 0; i < 3; ++i) {
            grd[ia][i] += gg * dx[i];
            grd[ib][i] -= gg * dx[i];
            fm[27 * ia + 3 * ib + i] = gg * dx[i];
            fm[27 * ib + 3 * ia + i] = -gg * dx[i];
        }
    }
    gab *= (sac + sbc) * retval / drab;       // mbpol-3b-poly.cpp:292-309
    gac *= (sab + sbc) * retval / drac;
    gbc *= (sab + sac) * retval / drbc;
    retval *= s;
    for (int n = 0; n < 3; ++n) {
        grd[0][n] += gab * rab[n] + gac * rac[n];
        grd[3][n] += gbc * rbc[n] - gab * rab[n];
        grd[6][n] -= gac * rac[n] + gbc * rbc[n];
        fm[27 * 0 + 3 * 3 + n] += gab * rab[n];
        fm[27 * 3 + 3 * 0 + n] -= gab * rab[n];
        fm[27 * 0 + 3 * 6 + n] += gac * rac[n];
        fm[27 * 6 + 3 * 0 + n] -= gac * rac[n];
        fm[27 * 3 + 3 * 6 + n] += gbc * rbc[n];
        fm[27 * 6 + 3 * 3 + n] -= gbc * rbc[n];
    }
    return retval;
}

// ------------------------------------------------------------------------------------------------
// imaging (mbpol.f90:1275-1361, invert.f)
// ------------------------------------------------------------------------------------------------
inline double fnint(double v) { return std::round(v); }   // Fortran NINT: round half away from zero

void invert3(const double* a, double* b, double& d)       // invert.f:1-50 (adjugate / determinant)
{
    b[0] = a[4] * a[8] - a[5] * a[7];
    b[1] = a[2] * a[7] - a[1] * a[8];
    b[2] = a[1] * a[5] - a[2] * a[4];
    b[3] = a[5] * a[6] - a[3] * a[8];
    b[4] = a[0] * a[8] - a[2] * a[6];
    b[5] = a[2] * a[3] - a[0] * a[5];
    b[6] = a[3] * a[7] - a[4] * a[6];
    b[7] = a[1] * a[6] - a[0] * a[7];
    b[8] = a[0] * a[4] - a[1] * a[3];
    d = a[0] * b[0] + a[3] * b[1] + a[6] * b[2];
    double r = 0.0;
    if (std::fabs(d) > 0.0) r = 1.0 / d;
    for (int i = 0; i < 9; ++i) b[i] = r * b[i];
}

int image_vectors(int imcon, const double* cell, int n, double* dx)
{
    if (imcon == 1) {
        const double a1 = 1.0 / cell[0];
        for (int i = 0; i < 3 * n; ++i) dx[i] = dx[i] - cell[0] * fnint(dx[i] * a1);
    } else if (imcon == 2) {
        const double a1 = 1.0 / cell[0], b1 = 1.0 / cell[4], c1 = 1.0 / cell[8];
        for (int i = 0; i < n; ++i) {
            dx[3 * i + 0] = dx[3 * i + 0] - cell[0] * fnint(dx[3 * i + 0] * a1);
            dx[3 * i + 1] = dx[3 * i + 1] - cell[4] * fnint(dx[3 * i + 1] * b1);
            dx[3 * i + 2] = dx[3 * i + 2] - cell[8] * fnint(dx[3 * i + 2] * c1);
        }
    } else if (imcon == 3) {
        double rc[9], det;
        invert3(cell, rc, det);
        for (int i = 0; i < n; ++i) {
            double* d = dx + 3 * i;
            const double ssx = (rc[0] * d[0] + rc[3] * d[1] + rc[6] * d[2]);
            const double ssy = (rc[1] * d[0] + rc[4] * d[1] + rc[7] * d[2]);
            const double ssz = (rc[2] * d[0] + rc[5] * d[1] + rc[8] * d[2]);
            const double xss = ssx - fnint(ssx), yss = ssy - fnint(ssy), zss = ssz - fnint(ssz);
            d[0] = (cell[0] * xss + cell[3] * yss + cell[6] * zss);
            d[1] = (cell[1] * xss + cell[4] * yss + cell[7] * zss);
            d[2] = (cell[2] * xss + cell[5] * yss + cell[8] * zss);
        }
    } else {
        return 1;   // fixme() -> stop in the reference
    }
    return 0;
}

void image_monomers(int imcon, const double* cell, int nw, double* xyz)
{
    for (int i = 0; i < nw; ++i) {
        double* w = xyz + 9 * i;
        double oh[6];
        for (int k = 0; k < 3; ++k) {
            oh[k] = w[3 + k] - w[k];
            oh[k + 3] = w[6 + k] - w[k];
        }
        image_vectors(imcon, cell, 2, oh);
        for (int k = 0; k < 3; ++k) {
            w[3 + k] = w[k] + oh[k];
            w[6 + k] = w[k] + oh[k + 3];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// accumulators of one worker (the per-rank arrays that the caller later all-reduces)
// ------------------------------------------------------------------------------------------------
struct Acc {
    int mxatms = 0;
    std::vector<double> fx, fy, fz, aeng, astress, afd;
    double eng2 = 0, vir2 = 0, eng3 = 0, vir3 = 0, stress[9] = {0};
    double* fmat = nullptr;             // optional full force_matrix(mxatms,mxatms,3), column-major
    const double *xxx = nullptr, *yyy = nullptr, *zzz = nullptr;
    bool want_stress = true, want_fd = true;
    void init(int n)
    {
        mxatms = n;
        fx.assign(n, 0); fy.assign(n, 0); fz.assign(n, 0);
        aeng.assign(n, 0); astress.assign(9 * (size_t)n, 0); afd.assign(9 * (size_t)n, 0);
    }
};

// heatcurrent.f90:529-533 update_forces(iatom,jatom,f): force_matrix(iatom,jatom,:) += f.   The N x N x 3
// matrix is only materialised on request (tiny systems, to validate the compressed tensor W):
//   W_i(a,b) = sum_clusters sum_{j in cluster} (x_i - x_j)_a^raw * f_b   
inline void update_forces(Acc& A, int ia, int ja, const double* f)   // 1-based atom indices
{
    if (A.fmat) {
        const size_t n = A.mxatms;
        for (int c = 0; c < 3; ++c) A.fmat[(ia - 1) + n * ((ja - 1) + n * c)] += f[c];
    }
    if (A.want_fd && ia != ja) {
        const double r[3] = {A.xxx[ia - 1] - A.xxx[ja - 1], A.yyy[ia - 1] - A.yyy[ja - 1], A.zzz[ia - 1] - A.zzz[ja - 1]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) A.afd[(ia - 1) + (size_t)A.mxatms * (a + 3 * b)] += r[a] * f[b];
    }
}

// common tail of do_dimer / do_trimer for n molecules (2 or 3): mbpol.f90:519-654, 896-1067
void accumulate_cluster(Acc& A, int nmol, const int* o, const double* const* w, double* const* g, const double* fm,
                        double e, double& engacc, double& viracc)
{
    const int na = 3 * nmol;
    const double share = (nmol == 2) ? 1.0 / 6.0 : 1.0 / 9.0;     // sixth / nineth
    engacc = engacc + e * engunit;
    int idx[9];
    for (int m = 0; m < nmol; ++m)
        for (int k = 0; k < 3; ++k) idx[3 * m + k] = o[m] + k;
    for (int a = 0; a < na; ++a) A.aeng[idx[a] - 1] += share * e * engunit;   // update_energy_mbpol
    for (int a1 = 0; a1 < na; ++a1) {                                         // update_forces for all (a1,a2)
        const double* gg = g[a1 / 3] + 3 * (a1 % 3);
        double ft[3] = {share * gg[0], share * gg[1], share * gg[2]};
        for (int c = 0; c < 3; ++c) ft[c] = -(ft[c] * engunit);
        for (int a2 = 0; a2 < na; ++a2) update_forces(A, idx[a1], idx[a2], ft);
    }
    if (A.want_stress) {                                                       // HEAT_STRESS block
        const int ld = 3 * na;
        for (int a1 = 0; a1 < na; ++a1) {
            double st[9] = {0};
            for (int a2 = 0; a2 < na; ++a2) {
                if (a1 == a2) continue;
                double rij[3], ft[3];
                for (int c = 0; c < 3; ++c) {
                    rij[c] = w[a1 / 3][3 * (a1 % 3) + c] - w[a2 / 3][3 * (a2 % 3) + c];
                    ft[c] = -fm[ld * a1 + 3 * a2 + c];
                }
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) st[3 * a + b] = st[3 * a + b] + rij[a] * ft[b];
            }
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)   // update_stress_mbpol(iatom,a,b, -st*0.5*engunit); atomic_stress(iatom,a,b) col-major
                    A.astress[(idx[a1] - 1) + (size_t)A.mxatms * (a + 3 * b)] += -st[3 * a + b] * 0.5 * engunit;
        }
    }
    for (int k = 0; k < 9; ++k)
        for (int m = 0; m < nmol; ++m) {
            g[m][k] = g[m][k] * engunit;
            viracc = viracc + w[m][k] * g[m][k];
        }
    for (int m = 0; m < nmol; ++m)
        for (int k = 0; k < 3; ++k) {
            A.fx[o[m] - 1 + k] -= g[m][3 * k + 0];
            A.fy[o[m] - 1 + k] -= g[m][3 * k + 1];
            A.fz[o[m] - 1 + k] -= g[m][3 * k + 2];
        }
    double strs[9];
    auto dot = [&](int a, int b) {
        double sacc = 0.0;
        bool first = true;
        for (int m = 0; m < nmol; ++m)
            for (int k = 0; k < 3; ++k) {
                const double t = w[m][3 * k + a] * g[m][3 * k + b];
                sacc = first ? t : sacc + t;
                first = false;
            }
        return sacc;
    };
    strs[0] = dot(0, 0); strs[1] = dot(0, 1); strs[2] = dot(0, 2);
    strs[3] = strs[1];   strs[4] = dot(1, 1); strs[5] = dot(1, 2);
    strs[6] = strs[2];   strs[7] = strs[5];   strs[8] = dot(2, 2);
    for (int q = 0; q < 9; ++q) A.stress[q] = A.stress[q] + (-strs[q]);
}

struct Ctx {
    int imcon;
    const double* cell;
    bool use_ref;
    double r2f, r3f;
};

// do_dimer.  w1 is re-imaged IN PLACE when mutate is set (faithful mode), else a private copy is used.
bool do_dimer(const Ctx& C, Acc& A, double* w1_in, double* w2, int o1, int o2, bool mutate)
{
    double r12[3] = {w1_in[0] - w2[0], w1_in[1] - w2[1], w1_in[2] - w2[2]};
    image_vectors(C.imcon, C.cell, 1, r12);
    if ((r12[0] * r12[0] + r12[1] * r12[1] + r12[2] * r12[2]) > C.r2f * C.r2f) return false;
    double w1c[9];
    double* w1 = w1_in;
    if (!mutate) {
        std::copy(w1_in, w1_in + 9, w1c);
        w1 = w1c;
    }
    for (int k = 0; k < 3; ++k) {
        w1[3 + k] = (w1[3 + k] - w1[k]) + (w2[k] + r12[k]);
        w1[6 + k] = (w1[6 + k] - w1[k]) + (w2[k] + r12[k]);
        w1[k] = w2[k] + r12[k];
    }
    double e2 = 0, g1[9], g2[9], fm[108];
    if (C.use_ref) {
        std::fill(g1, g1 + 9, 0.0); std::fill(g2, g2 + 9, 0.0); std::fill(fm, fm + 108, 0.0);
        ref_2b(w1, w2, &e2, g1, g2, fm);
    } else {
        e2 = eval_2b(w1, w2, g1, g2, fm);
    }
    const int o[2] = {o1, o2};
    const double* w[2] = {w1, w2};
    double* g[2] = {g1, g2};
    accumulate_cluster(A, 2, o, w, g, fm, e2, A.eng2, A.vir2);
    return true;
}

bool trimer_accept(const Ctx& C, const double* w1, const double* w2, const double* w3, double* rab)
{
    for (int k = 0; k < 3; ++k) {
        rab[k + 0] = w1[k] - w2[k];
        rab[k + 3] = w2[k] - w3[k];
        rab[k + 6] = w1[k] - w3[k];
    }
    image_vectors(C.imcon, C.cell, 3, rab);
    double r12 = 0, r23 = 0, r13 = 0;
    for (int k = 0; k < 3; ++k) {
        r12 = r12 + rab[k + 0] * rab[k + 0];
        r23 = r23 + rab[k + 3] * rab[k + 3];
        r13 = r13 + rab[k + 6] * rab[k + 6];
    }
    r12 = std::sqrt(r12); r23 = std::sqrt(r23); r13 = std::sqrt(r13);
    const bool big1 = r12 > C.r3f || r13 > C.r3f;
    const bool big2 = r12 > C.r3f || r23 > C.r3f;
    const bool big3 = r13 > C.r3f || r23 > C.r3f;
    return !(big1 && big2 && big3);
}

bool do_trimer(const Ctx& C, Acc& A, double* w1_in, double* w2, double* w3_in, int o1, int o2, int o3, bool mutate)
{
    double rab[9];
    if (!trimer_accept(C, w1_in, w2, w3_in, rab)) return false;
    double w1c[9], w3c[9];
    double *w1 = w1_in, *w3 = w3_in;
    if (!mutate) {
        std::copy(w1_in, w1_in + 9, w1c);
        std::copy(w3_in, w3_in + 9, w3c);
        w1 = w1c;
        w3 = w3c;
    }
    for (int k = 0; k < 3; ++k) {
        w1[3 + k] = (w1[3 + k] - w1[k]) + (w2[k] + rab[k]);
        w1[6 + k] = (w1[6 + k] - w1[k]) + (w2[k] + rab[k]);
        w1[k] = w2[k] + rab[k];
        w3[3 + k] = (w3[3 + k] - w3[k]) + (w2[k] - rab[k + 3]);
        w3[6 + k] = (w3[6 + k] - w3[k]) + (w2[k] - rab[k + 3]);
        w3[k] = w2[k] - rab[k + 3];
    }
    double e3 = 0, g1[9] = {0}, g2[9] = {0}, g3[9] = {0}, fm[243];
    if (C.use_ref) {
        std::fill(fm, fm + 243, 0.0);
        ref_3b(w1, w2, w3, &e3, g1, g2, g3, fm);
    } else {
        e3 = eval_3b(w1, w2, w3, g1, g2, g3, fm);
    }
    const int o[3] = {o1, o2, o3};
    const double* w[3] = {w1, w2, w3};
    double* g[3] = {g1, g2, g3};
    accumulate_cluster(A, 3, o, w, g, fm, e3, A.eng3, A.vir3);
    return true;
}

}  // namespace

// ================================================================================================
// C interface (ctypes)
// ================================================================================================
extern "C" {

// load the reference evaluators compiled by oracle/Makefile (returns 0 on success)
int oracle_load_reference(const char* path)
{
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return 1;
    ref_2b = (ref2b_t)dlsym(h, "mbpol_2b_poly_");
    ref_3b = (ref3b_t)dlsym(h, "mbpol_3b_poly_");
    return (ref_2b && ref_3b) ? 0 : 2;
}
int oracle_have_reference() { return ref_2b && ref_3b; }

void oracle_cutoffs(double* r2f, double* r3f)   // mbpol_2b_cutoff_ / mbpol_3b_cutoff_
{
    *r2f = ORC_P2b_r2f;
    *r3f = ORC_P3b_r3f;
}

double oracle_poly2b(const double* a, const double* x, double* g) { return poly_eval<31>(ORC_MONO2B, ORC_NMONO2B, a, x, g); }
double oracle_poly3b(const double* a, const double* x, double* g) { return poly_eval<36>(ORC_MONO3B, ORC_NMONO3B, a, x, g); }
const double* oracle_coef2b() { return ORC_COEF2B; }
const double* oracle_coef3b() { return ORC_COEF3B; }

// same signatures as the reference's C ABI (mbpol-2b-poly.cpp:1888-1893, mbpol-3b-poly.cpp:1526-1531)
void oracle_mbpol_2b(const double* w1, const double* w2, double* E, double* g1, double* g2, double* fm)
{
    *E = eval_2b(w1, w2, g1, g2, fm);
}
void oracle_mbpol_3b(const double* w1, const double* w2, const double* w3, double* E, double* g1, double* g2, double* g3,
                     double* fm)
{
    *E = eval_3b(w1, w2, w3, g1, g2, g3, fm);
}
void oracle_ref_2b(const double* w1, const double* w2, double* E, double* g1, double* g2, double* fm)
{
    ref_2b(w1, w2, E, g1, g2, fm);
}
void oracle_ref_3b(const double* w1, const double* w2, const double* w3, double* E, double* g1, double* g2, double* g3,
                   double* fm)
{
    ref_3b(w1, w2, w3, E, g1, g2, g3, fm);
}

int oracle_image_vectors(int imcon, const double* cell, int n, double* dx) { return image_vectors(imcon, cell, n, dx); }
void oracle_image_monomers(int imcon, const double* cell, int nw, double* xyz) { image_monomers(imcon, cell, nw, xyz); }

// ------------------------------------------------------------------------------------------------
// The driver: restatement of mbpol::mbpol_forces (mbpol.f90:98-449).
//   mode 0 faithful, 1 listed.   All outputs are ACCUMULATED into (+=) like the reference.
//   atomic_stress / atomic_fdecomp / force_matrix / lists may be NULL.
//   max_clusters > 0 bounds the number of accepted dimers and of accepted trimers that are EVALUATED in
//   listed mode (the lists are still complete) -- used for bounded CPU-baseline timing samples only.
//   timing[0..2] (optional): seconds spent in list build, 2B, 3B.
// ------------------------------------------------------------------------------------------------
int oracle_mbpol_forces(int mode, int use_ref, int nthreads, int bead_rank, int bead_size, int imcon, const double* cell,
                        int nttm2, const int* listttm2, int mxatms, const double* xxx, const double* yyy,
                        const double* zzz, double* fxx, double* fyy, double* fzz, double* eng2, double* vir2,
                        double* eng3, double* vir3, double* stress, double* atomic_energy, double* atomic_stress,
                        double* atomic_fdecomp, double* force_matrix, long long* ndimer, long long* ntrimer,
                        int* dimer_list, long long dimer_cap, int* trimer_list, long long trimer_cap,
                        long long max_clusters, double* timing)
{
    if (use_ref && !(ref_2b && ref_3b)) return 10;
    if (imcon < 1 || imcon > 3) return 11;
    const int nw = nttm2 / 3;
    Ctx C{imcon, cell, use_ref != 0, ORC_P2b_r2f, ORC_P3b_r3f};
    std::vector<double> monomers(9 * (size_t)nw);
    for (int t = 1; t <= nttm2; ++t) {                                  // mbpol.f90:184-189
        const int i = listttm2[t - 1];
        monomers[3 * t - 3] = xxx[i - 1];
        monomers[3 * t - 2] = yyy[i - 1];
        monomers[3 * t - 1] = zzz[i - 1];
    }
    image_monomers(imcon, cell, nw, monomers.data());                     // mbpol.f90:196
    auto O = [&](int m) { return listttm2[3 * m]; };                      // 1-based O index of water m (0-based m)
    long long nd = 0, nt = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tl = 0, t2 = 0, t3 = 0;

    if (mode == 0) {
        Acc A;
        A.init(mxatms);
        A.fmat = force_matrix; A.xxx = xxx; A.yyy = yyy; A.zzz = zzz;
        A.want_stress = atomic_stress != nullptr; A.want_fd = atomic_fdecomp != nullptr;
        double t0 = now();
        {   // pair loop, mbpol.f90:214-272
            const long long nw2 = ((long long)nw * (nw - 1)) / 2;
            long long i = 1, j = 2 + bead_rank;
            if (j > nw) {   // normalise a start beyond the first row (the Fortran relies on nw > bead_size)
                while (j > nw) { i = i + 1; j = i + j - nw; }
            }
            for (long long t = bead_rank; t <= nw2 - 1; t += bead_size) {
                if (do_dimer(C, A, &monomers[9 * (i - 1)], &monomers[9 * (j - 1)], O(i - 1), O(j - 1), true)) {
                    if (dimer_list && nd < dimer_cap) { dimer_list[2 * nd] = (int)i - 1; dimer_list[2 * nd + 1] = (int)j - 1; }
                    ++nd;
                }
                j = j + bead_size;
                if (j > nw)
                    while (j > nw) { i = i + 1; j = i + j - nw; }
            }
        }
        double t1 = now();
        t2 = t1 - t0;
        {   // triple loop, mbpol.f90:306-374
            const long long nw3 = ((long long)nw * (nw - 1) * (nw - 2)) / 6;
            long long i = 1, j = 2, k = 3 + bead_rank;
            auto carry = [&]() {
                while (k > nw) {
                    j = j + 1;
                    if ((j + 1) > nw) { i = i + 1; j = i + 1; }
                    k = j + (k - nw);
                }
            };
            if (k > nw) carry();
            for (long long t = bead_rank; t <= nw3 - 1; t += bead_size) {
                if (do_trimer(C, A, &monomers[9 * (i - 1)], &monomers[9 * (j - 1)], &monomers[9 * (k - 1)], O(i - 1),
                              O(j - 1), O(k - 1), true)) {
                    if (trimer_list && nt < trimer_cap) {
                        trimer_list[3 * nt] = (int)i - 1; trimer_list[3 * nt + 1] = (int)j - 1; trimer_list[3 * nt + 2] = (int)k - 1;
                    }
                    ++nt;
                }
                k = k + bead_size;
                if (k > nw) carry();
            }
        }
        t3 = now() - t1;
        for (int a = 0; a < mxatms; ++a) { fxx[a] += A.fx[a]; fyy[a] += A.fy[a]; fzz[a] += A.fz[a]; }
        if (atomic_energy) for (int a = 0; a < mxatms; ++a) atomic_energy[a] += A.aeng[a];
        if (atomic_stress) for (size_t a = 0; a < 9 * (size_t)mxatms; ++a) atomic_stress[a] += A.astress[a];
        if (atomic_fdecomp) for (size_t a = 0; a < 9 * (size_t)mxatms; ++a) atomic_fdecomp[a] += A.afd[a];
        *eng2 += A.eng2; *vir2 += A.vir2; *eng3 += A.eng3; *vir3 += A.vir3;
        for (int q = 0; q < 9; ++q) stress[q] += A.stress[q];
    } else {
        // ---- listed mode: neighbour lists with the reference predicates on un-mutated coordinates ----
        if (nthreads < 1) nthreads = 1;
        double t0 = now();
        std::vector<std::vector<int>> nb2(nw), nb3(nw);       // j > i within r2f ; all j != i within r3f
        {
            std::atomic<int> next(0);
            auto work = [&]() {
                for (;;) {
                    int i = next.fetch_add(1);
                    if (i >= nw) break;
                    for (int j = 0; j < nw; ++j) {
                        if (j == i) continue;
                        double d[3];
                        const double *wi = &monomers[9 * (size_t)std::min(i, j)], *wj = &monomers[9 * (size_t)std::max(i, j)];
                        for (int k = 0; k < 3; ++k) d[k] = wi[k] - wj[k];
                        image_vectors(imcon, cell, 1, d);
                        const double rsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                        if (j > i && !(rsq > C.r2f * C.r2f)) nb2[i].push_back(j);
                        if (!(std::sqrt(rsq) > C.r3f)) nb3[i].push_back(j);
                    }
                }
            };
            std::vector<std::thread> th;
            for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
            for (auto& t : th) t.join();
        }
        std::vector<int> dl, tlst;
        for (int i = 0; i < nw; ++i)
            for (int j : nb2[i]) { dl.push_back(i); dl.push_back(j); }
        // trimers: centre c with two neighbours a<b; if a-b are neighbours too the triple has three valid
        // centres -> emit it only from the smallest index.  Equivalent to "at least two of three O-O <= r3f".
        // NOTE the reference evaluates the three distances from (i,j), (j,k), (i,k) differences in index order;
        // with un-mutated coordinates |d| does not depend on the order, so pairwise tests are identical.
        for (int c = 0; c < nw; ++c) {
            const std::vector<int>& n = nb3[c];
            for (size_t p = 0; p < n.size(); ++p)
                for (size_t q = p + 1; q < n.size(); ++q) {
                    const int a = n[p], b = n[q];
                    const bool closed = std::binary_search(nb3[a].begin(), nb3[a].end(), b);
                    if (closed && !(c < a && c < b)) continue;
                    int t[3] = {c, a, b};
                    std::sort(t, t + 3);
                    tlst.push_back(t[0]); tlst.push_back(t[1]); tlst.push_back(t[2]);
                }
        }
        {   // canonical order: lexicographic (i,j,k)
            std::vector<size_t> idx(tlst.size() / 3);
            for (size_t q = 0; q < idx.size(); ++q) idx[q] = q;
            std::sort(idx.begin(), idx.end(), [&](size_t x, size_t y) {
                return std::lexicographical_compare(&tlst[3 * x], &tlst[3 * x] + 3, &tlst[3 * y], &tlst[3 * y] + 3);
            });
            std::vector<int> s(tlst.size());
            for (size_t q = 0; q < idx.size(); ++q) std::copy(&tlst[3 * idx[q]], &tlst[3 * idx[q]] + 3, &s[3 * q]);
            tlst.swap(s);
        }
        nd = (long long)dl.size() / 2;
        nt = (long long)tlst.size() / 3;
        if (dimer_list) std::copy(dl.begin(), dl.begin() + 2 * std::min(nd, dimer_cap), dimer_list);
        if (trimer_list) std::copy(tlst.begin(), tlst.begin() + 3 * std::min(nt, trimer_cap), trimer_list);
        double t1 = now();
        tl = t1 - t0;
        const long long nd_eval = (max_clusters > 0) ? std::min(nd, max_clusters) : nd;
        const long long nt_eval = (max_clusters > 0) ? std::min(nt, max_clusters) : nt;
        // workers: cluster q goes to worker (q mod nworkers), like t = bead_rank, ..., bead_size (mbpol.f90:226, 314);
        // only the slice of this (bead_rank, bead_size) is evaluated; per-worker accumulators are summed at the end
        std::vector<Acc> accs(nthreads);
        for (auto& A : accs) {
            A.init(mxatms);
            A.fmat = nullptr; A.xxx = xxx; A.yyy = yyy; A.zzz = zzz;
            A.want_stress = atomic_stress != nullptr; A.want_fd = atomic_fdecomp != nullptr;
        }
        auto run = [&](bool three) {
            std::vector<std::thread> th;
            for (int w = 0; w < nthreads; ++w)
                th.emplace_back([&, w]() {
                    Acc& A = accs[w];
                    const long long n = three ? nt_eval : nd_eval;
                    for (long long q = (long long)bead_rank + (long long)bead_size * w; q < n; q += (long long)bead_size * nthreads) {
                        if (three) {
                            const int i = tlst[3 * q], j = tlst[3 * q + 1], k = tlst[3 * q + 2];
                            do_trimer(C, A, &monomers[9 * (size_t)i], &monomers[9 * (size_t)j], &monomers[9 * (size_t)k], O(i), O(j), O(k), false);
                        } else {
                            const int i = dl[2 * q], j = dl[2 * q + 1];
                            do_dimer(C, A, &monomers[9 * (size_t)i], &monomers[9 * (size_t)j], O(i), O(j), false);
                        }
                    }
                });
            for (auto& t : th) t.join();
        };
        run(false);
        double t2e = now();
        t2 = t2e - t1;
        run(true);
        t3 = now() - t2e;
        for (auto& A : accs) {
            for (int a = 0; a < mxatms; ++a) { fxx[a] += A.fx[a]; fyy[a] += A.fy[a]; fzz[a] += A.fz[a]; }
            if (atomic_energy) for (int a = 0; a < mxatms; ++a) atomic_energy[a] += A.aeng[a];
            if (atomic_stress) for (size_t a = 0; a < 9 * (size_t)mxatms; ++a) atomic_stress[a] += A.astress[a];
            if (atomic_fdecomp) for (size_t a = 0; a < 9 * (size_t)mxatms; ++a) atomic_fdecomp[a] += A.afd[a];
            *eng2 += A.eng2; *vir2 += A.vir2; *eng3 += A.eng3; *vir3 += A.vir3;
            for (int q = 0; q < 9; ++q) stress[q] += A.stress[q];
        }
    }
    if (ndimer) *ndimer = nd;
    if (ntrimer) *ntrimer = nt;
    if (timing) { timing[0] = tl; timing[1] = t2; timing[2] = t3; }
    return 0;
}

// heatcurrent.f90:464-483 compute_heat_flux, MB-pol slice: j1 = sum e_i v_i ; j2_a = sum_i sum_b S_i(a,b)_uppersym v_i,b
void oracle_heat_flux(int mxatms, const double* aeng, const double* astress, const double* vx, const double* vy,
                      const double* vz, double* j1, double* j2)
{
    for (int a = 0; a < 3; ++a) j1[a] = j2[a] = 0.0;
    const size_t n = mxatms;
    for (int i = 0; i < mxatms; ++i) {
        const double v[3] = {vx[i], vy[i], vz[i]};
        for (int a = 0; a < 3; ++a) j1[a] = j1[a] + aeng[i] * v[a];
        if (!astress) continue;
        auto S = [&](int a, int b) { return astress[i + n * (a + 3 * b)]; };
        j2[0] = j2[0] + S(0, 0) * v[0] + S(0, 1) * v[1] + S(0, 2) * v[2];
        j2[1] = j2[1] + S(0, 1) * v[0] + S(1, 1) * v[1] + S(1, 2) * v[2];
        j2[2] = j2[2] + S(0, 2) * v[0] + S(1, 2) * v[1] + S(2, 2) * v[2];
    }
}

// heatcurrent.f90:703-728 compute_heat_flux2, MB-pol slice of jpot, from the FULL force matrix (tiny systems only)
void oracle_heat_flux2_matrix(int mxatms, const double* fmat, const double* x, const double* y, const double* z,
                              const double* vx, const double* vy, const double* vz, double* jpot)
{
    const size_t n = mxatms;
    const int nreal = 3 * mxatms / 4;
    for (int a = 0; a < 3; ++a) jpot[a] = 0.0;
    for (int i = 0; i < nreal; ++i) {
        const double v[3] = {vx[i], vy[i], vz[i]};
        for (int j = 0; j < nreal; ++j) {
            if (i == j) continue;
            const double f[3] = {fmat[i + n * (j + n * 0)], fmat[i + n * (j + n * 1)], fmat[i + n * (j + n * 2)]};
            const double r[3] = {x[i] - x[j], y[i] - y[j], z[i] - z[j]};
            const double fv = f[0] * v[0] + f[1] * v[1] + f[2] * v[2];
            for (int a = 0; a < 3; ++a) jpot[a] = jpot[a] + r[a] * fv;
        }
    }
}

// the same contraction from the compressed per-atom tensor W (what the GPU path produces)
void oracle_heat_flux2_fdecomp(int mxatms, const double* afd, const double* vx, const double* vy, const double* vz,
                               double* jpot)
{
    const size_t n = mxatms;
    const int nreal = 3 * mxatms / 4;
    for (int a = 0; a < 3; ++a) jpot[a] = 0.0;
    for (int i = 0; i < nreal; ++i) {
        const double v[3] = {vx[i], vy[i], vz[i]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) jpot[a] += afd[i + n * (a + 3 * b)] * v[b];
    }
}

}  // extern "C"
