// oracle/mbnrg_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile): CPU restatement of the reference's MB-nrg
// ion-water / ion-ion 2B path.  Never linked into the product library; only tests/,
// __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may load it.
//
//   evaluators : mbnrg-2b-h2o-<ion>-poly.cpp:220-329 (8 variables, switch on the ion-oxygen distance),
//                mbnrg-2b-<ion>-<ion>-poly.cpp (1 variable), polynomials from the monomial tables of data/mbnrg_model.npz
//   driver     : mbnrg.f90:170-485 (mbnrg_forces), :489-775 (do_dimer); the 3B part is compiled out in the reference
//                (mbnrg.f90:4) and so is it here.
// Pinned against the reference's own objects (oracle/_ref/libmbnrg_ref.so, built by `make ref-nrg`) in tests/test_mbnrg.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <vector>

#include "oracle_mbnrg.inc"

extern "C" int oracle_image_vectors(int imcon, const double* cell, int n, double* dx);   // mbpol_oracle.cpp (same routine
extern "C" void oracle_image_monomers(int imcon, const double* cell, int nw, double* xyz);  // in mbnrg.f90:1016-1119)

namespace {

const double engunit = 418.4;   // mbnrg.f90:33

typedef void (*refpoly_t)(const double*, const double*, double*, double*, double*);
void* ref_handle = nullptr;
refpoly_t ref_hi[9];
refpoly_t ref_ii[9][9];
const char* ION[9] = {"f", "cl", "br", "i", "li", "na", "k", "rb", "cs"};

// E = sum_m a_k prod x_v, g = dE/dx from the raw monomial table (degree <= 5)
double poly_hi(const double* a, const double* x, double* g)
{
    double E = 0.0;
    for (int v = 0; v < 8; ++v) g[v] = 0.0;
    for (int m = 0; m < ORC_NRG_NMONO; ++m) {
        const short* r = ORC_NRG_MONO[m];
        int n = 0;
        while (n < 5 && r[1 + n] >= 0) ++n;
        double p = a[r[0]];
        for (int i = 0; i < n; ++i) p *= x[r[1 + i]];
        E += p;
        for (int i = 0; i < n; ++i) {
            double q = a[r[0]];
            for (int j = 0; j < n; ++j)
                if (j != i) q *= x[r[1 + j]];
            g[r[1 + i]] += q;
        }
    }
    return E;
}

struct Var {   // mbnrg-2b-h2o-na-poly.cpp:12-80
    double g[3];
    double v_exp(double r0, double k, const double* c, int o, int x)
    {
        for (int i = 0; i < 3; ++i) g[i] = c[o + i] - c[x + i];
        const double r = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        const double e = std::exp(k * (r0 - r));
        const double gg = -k * e / r;
        for (int i = 0; i < 3; ++i) g[i] *= gg;
        return e;
    }
    double v_coul(double r0, double k, const double* c, int o, int x)
    {
        for (int i = 0; i < 3; ++i) g[i] = c[o + i] - c[x + i];
        const double rsq = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        const double r = std::sqrt(rsq);
        const double e = std::exp(k * (r0 - r));
        const double rinv = 1.0 / r;
        const double val = e * rinv;
        const double gg = -(k + rinv) * val * rinv;
        for (int i = 0; i < 3; ++i) g[i] *= gg;
        return val;
    }
    void grads(double gg, double* xg, int o, int x) const
    {
        for (int i = 0; i < 3; ++i) {
            const double d = gg * g[i];
            xg[o + i] += d;
            xg[x + i] -= d;
        }
    }
};

double f_switch(double r, double ri, double rf, double& g)   // mbnrg-2b-h2o-na-poly.cpp:202-216
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

// water w[9] + ion x[3] -> E, g1[9] (overwritten), g2[3] (ACCUMULATED, as the reference does); ion = code 1..9
double eval_h2o_ion(int ion, const double* w, const double* x, double* g1, double* g2)
{
    const double* P = ORC_NRG_PAR_HI + 14 * (ion - 1);
    const double kHH = P[0], kOH = P[1], kXH = P[2], kXO = P[3], kXL = P[4], dHH = P[5], dOH = P[6], dXH = P[7], dXO = P[8],
                 dXL = P[9], gin = P[10], gout = P[11], r2i = P[12], r2f = P[13];
    const double dXO3[3] = {w[0] - x[0], w[1] - x[1], w[2] - x[2]};
    const double rXO = std::sqrt(dXO3[0] * dXO3[0] + dXO3[1] * dXO3[1] + dXO3[2] * dXO3[2]);
    if (rXO > r2f) return 0.0;
    enum { O = 0, H1 = 3, H2 = 6, X = 9, L1 = 12, L2 = 15 };
    double c[18];
    std::copy(w, w + 9, c);
    std::copy(x, x + 3, c + 9);      // (the reference copies 9 doubles here, :249; slots 12..17 are overwritten below)
    double oh1[3], oh2[3];
    for (int i = 0; i < 3; ++i) {
        oh1[i] = c[i + 3] - c[i];
        oh2[i] = c[i + 6] - c[i];
    }
    const double vv[3] = {oh1[1] * oh2[2] - oh1[2] * oh2[1], oh1[2] * oh2[0] - oh1[0] * oh2[2], oh1[0] * oh2[1] - oh1[1] * oh2[0]};
    for (int i = 0; i < 3; ++i) {
        const double inp = c[i] + 0.5 * gin * (oh1[i] + oh2[i]);
        const double outp = gout * vv[i];
        c[L1 + i] = inp + outp;
        c[L2 + i] = inp - outp;
    }
    double v[8], g[8];
    Var ct[8];
    v[0] = ct[0].v_exp(dHH, kHH, c, H1, H2);
    v[1] = ct[1].v_exp(dOH, kOH, c, O, H1);
    v[2] = ct[2].v_exp(dOH, kOH, c, O, H2);
    v[3] = ct[3].v_coul(dXH, kXH, c, X, H1);
    v[4] = ct[4].v_coul(dXH, kXH, c, X, H2);
    v[5] = ct[5].v_coul(dXO, kXO, c, X, O);
    v[6] = ct[6].v_exp(dXL, kXL, c, X, L1);
    v[7] = ct[7].v_exp(dXL, kXL, c, X, L2);
    const double Ep = poly_hi(ORC_NRG_COEF_HI + 429 * (ion - 1), v, g);
    double xg[18];
    std::fill(xg, xg + 18, 0.0);
    ct[0].grads(g[0], xg, H1, H2);
    ct[1].grads(g[1], xg, O, H1);
    ct[2].grads(g[2], xg, O, H2);
    ct[3].grads(g[3], xg, X, H1);
    ct[4].grads(g[4], xg, X, H2);
    ct[5].grads(g[5], xg, X, O);
    ct[6].grads(g[6], xg, X, L1);
    ct[7].grads(g[7], xg, X, L2);
    {   // lone-pair gradients back onto O, H1, H2 (monomer::grads, :125-158)
        const double* a = xg + L1;
        const double* b = xg + L2;
        const double gm[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
        const double t1[3] = {oh2[1] * gm[2] - oh2[2] * gm[1], oh2[2] * gm[0] - oh2[0] * gm[2], oh2[0] * gm[1] - oh2[1] * gm[0]};
        const double t2[3] = {oh1[1] * gm[2] - oh1[2] * gm[1], oh1[2] * gm[0] - oh1[0] * gm[2], oh1[0] * gm[1] - oh1[1] * gm[0]};
        for (int i = 0; i < 3; ++i) {
            const double gsum = a[i] + b[i];
            const double inp = 0.5 * gin * gsum;
            const double gh1 = inp + gout * t1[i];
            const double gh2 = inp - gout * t2[i];
            xg[i + 0] += gsum - (gh1 + gh2);
            xg[i + 3] += gh1;
            xg[i + 6] += gh2;
        }
    }
    double gsw;
    const double sw = f_switch(rXO, r2i, r2f, gsw);
    for (int i = 0; i < 9; ++i) g1[i] = sw * xg[i];
    for (int i = 0; i < 3; ++i) g2[i] += sw * xg[i + 9];
    gsw *= Ep / rXO;
    for (int i = 0; i < 3; ++i) {
        const double d = gsw * dXO3[i];
        g1[i] += d;
        g2[i] -= d;
    }
    return sw * Ep;
}

// two ions (codes a, b) -> E, g1[3] (overwritten), g2[3] (accumulated): mbnrg-2b-na-cl-poly.cpp:129-186
double eval_ion_ion(int a, int b, const double* w, const double* x, double* g1, double* g2, bool& have)
{
    have = ORC_NRG_HAVE_II[9 * (a - 1) + (b - 1)] != 0;
    if (!have) return 0.0;
    const double* P = ORC_NRG_PAR_II + 4 * (9 * (a - 1) + (b - 1));
    const double* A = ORC_NRG_COEF_II + 13 * (9 * (a - 1) + (b - 1));
    const double d3[3] = {w[0] - x[0], w[1] - x[1], w[2] - x[2]};
    const double r = std::sqrt(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
    if (r > P[3]) return 0.0;
    double c[6] = {w[0], w[1], w[2], x[0], x[1], x[2]};
    Var ct;
    const double v = ct.v_exp(P[1], P[0], c, 0, 3);
    // E = sum_n a[n-1] v^n by Horner, dE/dv by the product rule along the same chain -- the association of
    // poly-2b-A1-B1-v1x.cpp:9-37 (the fitted coefficients cancel heavily, so the order of operations matters at 1e-10)
    double T = A[12] * v;                                   // t13
    double D = 2.0 * T + A[11];
    T = (T + A[11]) * v;                                    // t16
    for (int n = 10; n >= 0; --n) {
        D = D * v + A[n] + T;
        T = (A[n] + T) * v;
    }
    const double Ep = T, gp = D;
    double xg[6] = {0, 0, 0, 0, 0, 0};
    ct.grads(gp, xg, 0, 3);
    double gsw;
    const double sw = f_switch(r, P[2], P[3], gsw);
    for (int i = 0; i < 3; ++i) g1[i] = sw * xg[i];
    for (int i = 0; i < 3; ++i) g2[i] += sw * xg[i + 3];
    gsw *= Ep / r;
    for (int i = 0; i < 3; ++i) {
        const double d = gsw * d3[i];
        g1[i] += d;
        g2[i] -= d;
    }
    return sw * Ep;
}


// ---- 3B ion-water-water (mbnrg-3b-h2o-h2o-<ion>-poly.cpp) -----------------------------------------------------------
typedef void (*refpoly3_t)(const double*, const double*, const double*, double*, double*, double*, double*);
refpoly3_t ref_3b[9];

// E = sum_m a_k prod x_v, g = dE/dx from the raw monomial table of poly-3b-h2o-ion-v1x.cpp (21 variables, degree <= 5)
double poly_3b(const double* a, const double* x, double* g)
{
    double E = 0.0;
    for (int v = 0; v < 21; ++v) g[v] = 0.0;
    for (int m = 0; m < ORC_NRG_NMONO3; ++m) {
        const short* r = ORC_NRG_MONO3[m];
        int n = 0;
        while (n < 5 && r[1 + n] >= 0) ++n;
        double p = a[r[0]];
        for (int i = 0; i < n; ++i) p *= x[r[1 + i]];
        E += p;
        for (int i = 0; i < n; ++i) {
            double q = a[r[0]];
            for (int j = 0; j < n; ++j)
                if (j != i) q *= x[r[1 + j]];
            g[r[1 + i]] += q;
        }
    }
    return E;
}

// x3b_h2o_ion_v2x_p::operator() (mbnrg-3b-h2o-h2o-na-poly.cpp:149-305): sites 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 X;
// 21 variables exp(-k (d - d0)) (:168-189); s = sab sac + sab sbc + sac sbc on |Oa-Ob|, |Oa-X|, |Ob-X| (:196-219);
// gradients are ACCUMULATED into g1, g2, g3 (:28-41, 61-78); parameters in the order of tools/extract_mbnrg.py PAR_3B
const int V3N_A[21] = {1, 4, 0, 0, 3, 3, 1, 1, 2, 2, 0, 0, 3, 3, 0, 6, 6, 6, 6, 6, 6};
const int V3N_B[21] = {2, 5, 1, 2, 4, 5, 4, 5, 4, 5, 4, 5, 1, 2, 3, 1, 2, 4, 5, 0, 3};
const int V3N_K[21] = {0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 5, 5, 5, 5, 6, 6};   // HH_intra OH_intra HH OH OO XH XO
double eval_h2o_h2o_ion(int ion, const double* w1, const double* w2, const double* x, double* g1, double* g2, double* g3)
{
    const double* P = ORC_NRG_PAR_3B + 16 * (ion - 1);
    const double r3i = P[0], r3f = P[1];
    const double* kk = P + 2;     // 7 k's
    const double* dd = P + 9;     // 7 d0's
    double c[21];
    for (int i = 0; i < 9; ++i) { c[i] = w1[i]; c[9 + i] = w2[i]; }
    for (int i = 0; i < 3; ++i) c[18 + i] = x[i];
    double xi[21], g[21], dx[21][3], dist[21];
    for (int n = 0; n < 21; ++n) {
        const int a = 3 * V3N_A[n], b = 3 * V3N_B[n];
        for (int i = 0; i < 3; ++i) dx[n][i] = c[a + i] - c[b + i];
        dist[n] = std::sqrt(dx[n][0] * dx[n][0] + dx[n][1] * dx[n][1] + dx[n][2] * dx[n][2]);
        xi[n] = std::exp(-kk[V3N_K[n]] * (dist[n] - dd[V3N_K[n]]));
    }
    double E = poly_3b(ORC_NRG_COEF_3B + 1831 * (ion - 1), xi, g);
    double rab[3], rac[3], rbc[3], drab = 0, drac = 0, drbc = 0;
    for (int i = 0; i < 3; ++i) {
        rab[i] = c[i] - c[9 + i]; drab += rab[i] * rab[i];
        rac[i] = c[i] - c[18 + i]; drac += rac[i] * rac[i];
        rbc[i] = c[9 + i] - c[18 + i]; drbc += rbc[i] * rbc[i];
    }
    drab = std::sqrt(drab); drac = std::sqrt(drac); drbc = std::sqrt(drbc);
    double gab, gac, gbc;
    const double sab = f_switch(drab, r3i, r3f, gab), sac = f_switch(drac, r3i, r3f, gac), sbc = f_switch(drbc, r3i, r3f, gbc);
    const double sw = sab * sac + sab * sbc + sac * sbc;
    double gr[21];
    for (int i = 0; i < 21; ++i) gr[i] = 0.0;
    for (int n = 0; n < 21; ++n) {
        const double gg = -kk[V3N_K[n]] * (g[n] * sw) * xi[n] / dist[n];
        for (int i = 0; i < 3; ++i) {
            gr[3 * V3N_A[n] + i] += gg * dx[n][i];
            gr[3 * V3N_B[n] + i] -= gg * dx[n][i];
        }
    }
    gab *= (sac + sbc) * E / drab;
    gac *= (sab + sbc) * E / drac;
    gbc *= (sab + sac) * E / drbc;
    for (int i = 0; i < 3; ++i) {
        gr[i] += gab * rab[i] + gac * rac[i];
        gr[9 + i] += gbc * rbc[i] - gab * rab[i];
        gr[18 + i] -= gac * rac[i] + gbc * rbc[i];
    }
    for (int i = 0; i < 9; ++i) { g1[i] += gr[i]; g2[i] += gr[9 + i]; }
    for (int i = 0; i < 3; ++i) g3[i] += gr[18 + i];
    return E * sw;
}

}  // namespace

extern "C" {

// dlopen the reference's own MB-nrg evaluator objects (oracle/_ref/libmbnrg_ref.so); 0 = ok
int oracle_mbnrg_load_reference(const char* path)
{
    ref_handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!ref_handle) return 1;
    char name[96];
    for (int i = 0; i < 9; ++i) {
        snprintf(name, sizeof name, "mbnrg_2b_h2o_%s_poly_", ION[i]);
        ref_hi[i] = (refpoly_t)dlsym(ref_handle, name);
        if (!ref_hi[i]) return 2;
        snprintf(name, sizeof name, "mbnrg_3b_h2o_h2o_%s_poly_", ION[i]);
        ref_3b[i] = (refpoly3_t)dlsym(ref_handle, name);      // absent in a 2B-only reference build: the 3B calls then use the restatement
        for (int j = 0; j < 9; ++j) {
            snprintf(name, sizeof name, "mbnrg_2b_%s_%s_poly_", ION[i], ION[j]);
            ref_ii[i][j] = (refpoly_t)dlsym(ref_handle, name);
            if (!ref_ii[i][j]) {
                snprintf(name, sizeof name, "mbnrg_2b_%s_%s_poly_", ION[j], ION[i]);
                ref_ii[i][j] = (refpoly_t)dlsym(ref_handle, name);
            }
        }
    }
    return 0;
}
int oracle_mbnrg_have_reference() { return ref_handle != nullptr; }

double oracle_mbnrg_polyhi(const double* a, const double* x, double* g) { return poly_hi(a, x, g); }
const double* oracle_mbnrg_coef_hi(int ion) { return ORC_NRG_COEF_HI + 429 * (ion - 1); }

// single clusters; use_ref != 0 calls the reference object instead of the restatement.  g2 is zeroed first (do_dimer :514)
void oracle_mbnrg_h2o_ion(int use_ref, int ion, const double* w, const double* x, double* E, double* g1, double* g2)
{
    std::fill(g1, g1 + 9, 0.0);
    std::fill(g2, g2 + 3, 0.0);
    if (use_ref) {
        double xx[9] = {x[0], x[1], x[2], 0, 0, 0, 0, 0, 0};   // the reference reads 9 doubles from the ion pointer (:249)
        ref_hi[ion - 1](w, xx, E, g1, g2);
    } else {
        *E = eval_h2o_ion(ion, w, x, g1, g2);
    }
}
int oracle_mbnrg_ion_ion(int use_ref, int a, int b, const double* w, const double* x, double* E, double* g1, double* g2)
{
    std::fill(g1, g1 + 3, 0.0);
    std::fill(g2, g2 + 3, 0.0);
    *E = 0.0;
    if (use_ref) {
        if (!ref_ii[a - 1][b - 1]) return 0;
        ref_ii[a - 1][b - 1](w, x, E, g1, g2);
        return 1;
    }
    bool have;
    *E = eval_ion_ion(a, b, w, x, g1, g2, have);
    return have ? 1 : 0;
}

// mbnrg_forces (mbnrg.f90:170-485) for one rank of `nranks`:
//   * ions are the FIRST nions atoms (positions read from xxx(1..nions), :226-233), listions(j) = their force slots;
//   * water-ion pairs: every ion j with the waters of this rank's contiguous slice, skipped when the imaged ion-oxygen
//     distance exceeds r2f = 10 (:130, :525); the water is moved next to the ion before the evaluator call (:530-541);
//   * ion-ion pairs j < k: evaluated by rank 0 only (the reference lets EVERY bead rank add them, :319-337, which
//     over-counts after its gdsum when bead_size > 1; with one rank the two agree);
//   * E, virial (sum r.g over the imaged positions), stress and forces accumulate into the caller's values like do_dimer.
// counts[0] = water-ion pairs inside r2f, counts[1] = pairs inside the evaluator's own cutoff, counts[2] = ion-ion pairs.
int oracle_mbnrg_forces(int use_ref, int rank, int nranks, int imcon, const double* cell, int nttm2, const int* listttm2, int nions,
                        const int* listions, const int* ion_code, const double* xxx, const double* yyy, const double* zzz, double* fxx,
                        double* fyy, double* fzz, double* eng2, double* vir2, double* stress, long long* counts)
{
    const double r2f = 10.0;
    const int nw = nttm2 / 3;
    std::vector<double> wat(9 * (size_t)nw), ion(3 * (size_t)nions);
    for (int t = 0; t < nttm2; ++t) {
        const int i = listttm2[t] - 1;
        wat[3 * t + 0] = xxx[i];
        wat[3 * t + 1] = yyy[i];
        wat[3 * t + 2] = zzz[i];
    }
    for (int t = 0; t < nions; ++t) {
        ion[3 * t + 0] = xxx[t];
        ion[3 * t + 1] = yyy[t];
        ion[3 * t + 2] = zzz[t];
    }
    oracle_image_monomers(imcon, cell, nw, wat.data());
    const long long lo = (long long)nw * rank / nranks, hi = (long long)nw * (rank + 1) / nranks;
    double e2 = 0.0, v2 = 0.0;
    long long c0 = 0, c1 = 0, c2 = 0;
    auto finish = [&](int n1, const double* w1, const double* w2, double* g1, double* g2, double E, int o1, int o2) {
        e2 += E * engunit;
        for (int k = 0; k < n1; ++k) {
            g1[k] *= engunit;
            v2 += w1[k] * g1[k];
        }
        for (int k = 0; k < 3; ++k) {
            g2[k] *= engunit;
            v2 += w2[k] * g2[k];
        }
        for (int k = 0; k < n1 / 3; ++k) {
            fxx[o1 - 1 + k] -= g1[3 * k + 0];
            fyy[o1 - 1 + k] -= g1[3 * k + 1];
            fzz[o1 - 1 + k] -= g1[3 * k + 2];
        }
        fxx[o2 - 1] -= g2[0];
        fyy[o2 - 1] -= g2[1];
        fzz[o2 - 1] -= g2[2];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double s = w2[a] * g2[b];
                for (int k = 0; k < n1 / 3; ++k) s += w1[3 * k + a] * g1[3 * k + b];
                if (b >= a) stress[3 * a + b] -= s;          // strs(1,2,3,5,6,9) computed, (4,7,8) mirrored (:745-771)
                if (b > a) stress[3 * b + a] -= s;
            }
    };
    for (int j = 0; j < nions; ++j) {
        double* x = &ion[3 * j];
        for (long long i = lo; i < hi; ++i) {
            double* w = &wat[9 * i];
            double r12[3] = {w[0] - x[0], w[1] - x[1], w[2] - x[2]};
            if (oracle_image_vectors(imcon, cell, 1, r12)) return 1;
            if (r12[0] * r12[0] + r12[1] * r12[1] + r12[2] * r12[2] > r2f * r2f) continue;
            for (int k = 0; k < 3; ++k) {
                w[3 + k] = (w[3 + k] - w[k]) + (x[k] + r12[k]);
                w[6 + k] = (w[6 + k] - w[k]) + (x[k] + r12[k]);
                w[k] = x[k] + r12[k];
            }
            ++c0;
            double E, g1[9], g2[3];
            oracle_mbnrg_h2o_ion(use_ref, ion_code[j], w, x, &E, g1, g2);
            if (E != 0.0) ++c1;
            finish(9, w, x, g1, g2, E, listttm2[3 * i], listions[j]);
        }
        if (rank != 0) continue;
        for (int k = j + 1; k < nions; ++k) {
            double* y = &ion[3 * k];
            double r12[3] = {x[0] - y[0], x[1] - y[1], x[2] - y[2]};
            if (oracle_image_vectors(imcon, cell, 1, r12)) return 1;
            if (r12[0] * r12[0] + r12[1] * r12[1] + r12[2] * r12[2] > r2f * r2f) continue;
            for (int q = 0; q < 3; ++q) x[q] = y[q] + r12[q];       // the first ion moves next to the second (:541)
            double E, g1[3], g2[3];
            if (!oracle_mbnrg_ion_ion(use_ref, ion_code[j], ion_code[k], x, y, &E, g1, g2)) continue;
            ++c2;
            finish(3, x, y, g1, g2, E, listions[j], listions[k]);
        }
    }
    *eng2 += e2;
    *vir2 += v2;
    if (counts) {
        counts[0] = c0;
        counts[1] = c1;
        counts[2] = c2;
    }
    return 0;
}


double oracle_mbnrg_poly3b(const double* a, const double* x, double* g) { return poly_3b(a, x, g); }
const double* oracle_mbnrg_coef_3b(int ion) { return ORC_NRG_COEF_3B + 1831 * (ion - 1); }
int oracle_mbnrg_have_reference_3b() { return ref_handle != nullptr && ref_3b[0] != nullptr; }

// one ion-water-water cluster; the gradients are zeroed first, as do_trimer does (mbnrg.f90:864-866)
void oracle_mbnrg_h2o_h2o_ion(int use_ref, int ion, const double* w1, const double* w2, const double* x, double* E, double* g1, double* g2,
                              double* g3)
{
    std::fill(g1, g1 + 9, 0.0);
    std::fill(g2, g2 + 9, 0.0);
    std::fill(g3, g3 + 3, 0.0);
    if (use_ref && ref_3b[ion - 1]) ref_3b[ion - 1](w1, w2, x, E, g1, g2, g3);
    else *E = eval_h2o_h2o_ion(ion, w1, w2, x, g1, g2, g3);
}

// The 3B part of mbnrg_forces (mbnrg.f90:354-427, do_trimer :781-993; disabled in the reference build by `!#define ENABLE_X3B`, :4):
// for every ion k and every water pair i < j, with the minimum-image vectors O_i - O_j, O_j - X, O_i - X: the trimer is evaluated iff
// |O_i - X| < r3f AND |O_j - X| < r3f (r3f = 7, :135; strict <, :840-842); water i and the ion are placed next to water j (:849-858);
// energy, virial, stress (computed 1,2,3,5,6,9; mirrored 4,7,8) and forces accumulate as in do_trimer.  Canonical semantics:
// the stored monomers are not mutated.  Rank r of nranks takes the contiguous range [T r/N, T (r+1)/N) of the accepted trimers in
// (k, i, j) order.  counts[0] = accepted trimers (all ranks), counts[1] = evaluated by this rank.
int oracle_mbnrg_forces3(int use_ref, int rank, int nranks, int imcon, const double* cell, int nttm2, const int* listttm2, int nions,
                         const int* listions, const int* ion_code, const double* xxx, const double* yyy, const double* zzz, double* fxx,
                         double* fyy, double* fzz, double* eng3, double* vir3, double* stress, long long* counts)
{
    const double r3f = 7.0;
    const int nw = nttm2 / 3;
    std::vector<double> wat(9 * (size_t)nw), ion(3 * (size_t)nions);
    for (int t = 0; t < nttm2; ++t) {
        const int i = listttm2[t] - 1;
        wat[3 * t + 0] = xxx[i]; wat[3 * t + 1] = yyy[i]; wat[3 * t + 2] = zzz[i];
    }
    for (int t = 0; t < nions; ++t) { ion[3 * t + 0] = xxx[t]; ion[3 * t + 1] = yyy[t]; ion[3 * t + 2] = zzz[t]; }
    oracle_image_monomers(imcon, cell, nw, wat.data());
    struct Tri { int k, i, j; };
    std::vector<Tri> tris;
    for (int k = 0; k < nions; ++k) {
        std::vector<int> near;
        for (int i = 0; i < nw; ++i) {
            double r[3] = {wat[9 * i] - ion[3 * k], wat[9 * i + 1] - ion[3 * k + 1], wat[9 * i + 2] - ion[3 * k + 2]};
            if (oracle_image_vectors(imcon, cell, 1, r)) return 1;
            if (std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) < r3f) near.push_back(i);
        }
        for (size_t p = 0; p < near.size(); ++p)
            for (size_t q = p + 1; q < near.size(); ++q) tris.push_back({k, near[p], near[q]});
    }
    const long long T = (long long)tris.size(), lo = T * rank / nranks, hi = T * (rank + 1) / nranks;
    double e3 = 0.0, v3 = 0.0;
    for (long long t = lo; t < hi; ++t) {
        const Tri& tr = tris[t];
        const double* wi = &wat[9 * (size_t)tr.i];
        const double* w2 = &wat[9 * (size_t)tr.j];
        const double* xk = &ion[3 * (size_t)tr.k];
        double rab[6] = {wi[0] - w2[0], wi[1] - w2[1], wi[2] - w2[2], w2[0] - xk[0], w2[1] - xk[1], w2[2] - xk[2]};
        if (oracle_image_vectors(imcon, cell, 2, rab)) return 1;
        double w1[9], x[3];
        for (int q = 0; q < 3; ++q) {
            w1[3 + q] = (wi[3 + q] - wi[q]) + (w2[q] + rab[q]);
            w1[6 + q] = (wi[6 + q] - wi[q]) + (w2[q] + rab[q]);
            w1[q] = w2[q] + rab[q];
            x[q] = w2[q] - rab[3 + q];
        }
        double E, g1[9], g2[9], g3[3];
        oracle_mbnrg_h2o_h2o_ion(use_ref, ion_code[tr.k], w1, w2, x, &E, g1, g2, g3);
        e3 += E * engunit;
        for (int q = 0; q < 9; ++q) { g1[q] *= engunit; g2[q] *= engunit; v3 += w1[q] * g1[q] + w2[q] * g2[q]; }
        for (int q = 0; q < 3; ++q) { g3[q] *= engunit; v3 += x[q] * g3[q]; }
        const int o1 = listttm2[3 * tr.i], o2 = listttm2[3 * tr.j], o3 = listions[tr.k];
        for (int q = 0; q < 3; ++q) {
            fxx[o1 - 1 + q] -= g1[3 * q]; fyy[o1 - 1 + q] -= g1[3 * q + 1]; fzz[o1 - 1 + q] -= g1[3 * q + 2];
            fxx[o2 - 1 + q] -= g2[3 * q]; fyy[o2 - 1 + q] -= g2[3 * q + 1]; fzz[o2 - 1 + q] -= g2[3 * q + 2];
        }
        fxx[o3 - 1] -= g3[0]; fyy[o3 - 1] -= g3[1]; fzz[o3 - 1] -= g3[2];
        for (int a = 0; a < 3; ++a)
            for (int b = a; b < 3; ++b) {
                double sacc = x[a] * g3[b];
                for (int q = 0; q < 3; ++q) sacc += w1[3 * q + a] * g1[3 * q + b] + w2[3 * q + a] * g2[3 * q + b];
                stress[3 * a + b] -= sacc;
                if (b > a) stress[3 * b + a] -= sacc;
            }
    }
    *eng3 += e3;
    *vir3 += v3;
    if (counts) { counts[0] = T; counts[1] = hi - lo; }
    return 0;
}

}  // extern "C"
