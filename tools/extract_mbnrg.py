#!/usr/bin/env python3
"""Derive the MB-nrg ion-water / ion-ion 2B model tables from the reference tree (run ONCE, in the build container).

Same approach as tools/extract_model.py: the Maple-generated bodies (poly-2b-h2o-ion-v1x.cpp:5-1403, 8 variables,
429 linear coefficients, degree 5; poly-2b-A1-A1-v1x.cpp / poly-2b-A1-B1-v1x.cpp, 1 variable, degree 13) are
*symbolically executed* into monomial tables, and the fitted literals of every evaluator file
(mbnrg-2b-h2o-<ion>-poly.cpp:335-793, mbnrg-2b-<ion>-<ion>-poly.cpp) are parsed.  No source is copied.

Output: dlpoly2-mbpol-heat_b200/data/mbnrg_model.npz
    mono_hi   int16 [n, 6]  (k, v0..v4), -1 padded      w_hi float64 [n]
    coef_hi   float64 [9, 429]       one row per ion, ion codes 1..9 = f cl br i li na k rb cs (mbnrg.f90:541-565)
    par_hi    float64 [9, 14]        k_HH_intra k_OH_intra k_XH_coul k_XO_coul k_XLp_main d_HH_intra d_OH_intra d_XH_coul
                                     d_XO_coul d_XLp_main in_plane_gamma out_of_plane_gamma r2i r2f
    mono_ii   int16 [13, 14] / w_ii  the 1-variable ion-ion polynomial  sum_n a[n-1] x^n
    coef_ii   float64 [9, 9, 13], par_ii float64 [9, 9, 4] (k_AB d_AB r2i r2f), have_ii bool [9, 9] (symmetric)
    mono_3b   int16 [n, 6] / w_3b    the 21-variable, degree-5 ion-water-water polynomial (poly-3b-h2o-ion-v1x.cpp)
    coef_3b   float64 [9, 1831]      mbnrg-3b-h2o-h2o-<ion>-poly.cpp
    par_3b    float64 [9, 16]        r3i r3f kHH_intra kOH_intra kHH kOH kOO kXH kXO dHH_intra dOH_intra dHH dOH dOO dXH dXO
"""
import os
import re
import sys

import numpy as np

import extract_model as X

IONS = ["f", "cl", "br", "i", "li", "na", "k", "rb", "cs"]
PAR_HI = ["m_k_HH_intra", "m_k_OH_intra", "m_k_XH_coul", "m_k_XO_coul", "m_k_XLp_main", "m_d_HH_intra", "m_d_OH_intra",
          "m_d_XH_coul", "m_d_XO_coul", "m_d_XLp_main", "m_in_plane_gamma", "m_out_of_plane_gamma", "m_r2i", "m_r2f"]


PAR_3B = ["m_r3i", "m_r3f", "m_kHH_intra", "m_kOH_intra", "m_kHH", "m_kOH", "m_kOO", "m_kXH", "m_kXO", "m_dHH_intra", "m_dOH_intra", "m_dHH",
          "m_dOH", "m_dOO", "m_dXH", "m_dXO"]


def table(p, ncoef, maxdeg):
    rows, w = [], []
    for (k, m), wt in p.t.items():
        if wt == 0.0:
            continue
        assert 0 <= k < ncoef and 1 <= len(m) <= maxdeg, (k, m)
        rows.append((k,) + m + (-1,) * (maxdeg - len(m)))
        w.append(wt)
    order = sorted(range(len(rows)), key=lambda i: (sum(v >= 0 for v in rows[i][1:]), rows[i][1:], rows[i][0]))
    return np.array([rows[i] for i in order], dtype=np.int16), np.array([w[i] for i in order], dtype=np.float64)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/SOURCE"
    out = {}
    p, nused, ntot = X.symbolic_energy(os.path.join(ref, "poly-2b-h2o-ion-v1x.cpp"))
    out["mono_hi"], out["w_hi"] = table(p, 429, 5)
    deg = (out["mono_hi"][:, 1:] >= 0).sum(1)
    print("h2o-ion: %d/%d statements feed E; %d monomials (deg hist %s); weights %s" % (
        nused, ntot, len(out["mono_hi"]), np.bincount(deg).tolist(), np.unique(out["w_hi"]).tolist()))
    tabs = []
    for fn in ("poly-2b-A1-A1-v1x.cpp", "poly-2b-A1-B1-v1x.cpp"):
        p, _, _ = X.symbolic_energy(os.path.join(ref, fn))
        tabs.append(table(p, 13, 13))
    assert np.array_equal(tabs[0][0], tabs[1][0]) and np.array_equal(tabs[0][1], tabs[1][1])
    out["mono_ii"], out["w_ii"] = tabs[0]
    assert len(out["mono_ii"]) == 13 and np.all(out["w_ii"] == 1.0)
    for n, r in enumerate(out["mono_ii"]):            # a[n] multiplies x^(n+1)
        assert r[0] == n and (r[1:] >= 0).sum() == n + 1, r
    coef, par = np.zeros((9, 429)), np.zeros((9, 14))
    for i, ion in enumerate(IONS):
        src = open(os.path.join(ref, "mbnrg-2b-h2o-%s-poly.cpp" % ion)).read()
        coef[i] = X.literal_block(src, "the_poly[] = {")
        pp = X.named_params(src, PAR_HI)
        par[i] = [pp[k] for k in PAR_HI]
    out["coef_hi"], out["par_hi"] = coef, par
    cii, pii, have = np.zeros((9, 9, 13)), np.zeros((9, 9, 4)), np.zeros((9, 9), dtype=bool)
    for i, a in enumerate(IONS):
        for j, b in enumerate(IONS):
            fn = os.path.join(ref, "mbnrg-2b-%s-%s-poly.cpp" % (a, b))
            if not os.path.exists(fn):
                continue
            src = open(fn).read()
            c = X.literal_block(src, "the_poly[] = {")
            assert c.shape == (13,), (fn, c.shape)
            names = ["k_AA", "d_AA"] if "k_AA" in src else ["k_AB", "d_AB"]
            pp = X.named_params(src, names + ["m_r2i", "m_r2f"])
            for (u, v) in ((i, j), (j, i)):
                cii[u, v], pii[u, v], have[u, v] = c, [pp[n] for n in names + ["m_r2i", "m_r2f"]], True
    out["coef_ii"], out["par_ii"], out["have_ii"] = cii, pii, have
    # ---- 3B ion-water-water (mbnrg-3b-h2o-h2o-<ion>-poly.cpp, poly-3b-h2o-ion-v1x.cpp: 21 variables, 1831 coefficients, degree 5)
    p, nused, ntot = X.symbolic_energy(os.path.join(ref, "poly-3b-h2o-ion-v1x.cpp"))
    out["mono_3b"], out["w_3b"] = table(p, 1831, 5)
    assert np.all(out["w_3b"] == 1.0)
    deg = (out["mono_3b"][:, 1:] >= 0).sum(1)
    print("h2o-h2o-ion: %d/%d statements feed E; %d monomials (deg hist %s)" % (nused, ntot, len(out["mono_3b"]), np.bincount(deg).tolist()))
    c3, p3 = np.zeros((9, 1831)), np.zeros((9, 16))
    for i, ion in enumerate(IONS):
        src = open(os.path.join(ref, "mbnrg-3b-h2o-h2o-%s-poly.cpp" % ion)).read()
        c3[i] = X.literal_block(src, "the_poly[] = {")
        pp = X.named_params(src, PAR_3B)
        p3[i] = [pp[k] for k in PAR_3B]
    out["coef_3b"], out["par_3b"] = c3, p3
    print("ion-ion pairs with a fitted evaluator: %d of 45" % (np.triu(have).sum()))
    dst = os.path.join(X.DATA, "mbnrg_model.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.setrecursionlimit(100000)
    main()
