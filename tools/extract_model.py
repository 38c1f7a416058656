#!/usr/bin/env python3
"""Derive the MB-pol model tables from the reference tree (run ONCE, in the build container).

The reference ships its 2B/3B permutationally invariant polynomials only as
Maple-generated straight-line C++ (poly-2b-v6x.cpp:5-13843, poly-3b-v2x.cpp:5-28380)
and the fitted linear coefficients / non-linear parameters as literals in
mbpol-2b-poly.cpp:688-1866 and mbpol-3b-poly.cpp:318-1503.  Nothing of that source
is copied.  Instead this script *symbolically executes* the energy part of each
generated body with a sparse polynomial class, which recovers the mathematical
definition of the model:

    E(x) = sum_k a_k * sum_{m in orbit_k} w_m * prod_{v in m} x_v

i.e. for every linear coefficient k the list of monomials (multisets of variable
indices, degree <= 4) it multiplies.  From that table our own code generator
(tools/gen_poly.py) builds a completely different evaluation schedule.

Outputs (committed; the model definition, not code):
    dlpoly2-mbpol-heat_b200/data/mbpol_model.npz
        mono2b  int16 [n_mono, 5]   (k, v0, v1, v2, v3), -1 padded, sorted
        w2b     float64 [n_mono]    monomial weights (all exactly 1.0 for MB-pol)
        coef2b  float64 [1153]
        mono3b / w3b / coef3b       same for the 3B polynomial
        par_*   scalar non-linear parameters (k's, d0's, gammas, switch radii)

Usage:  python tools/extract_model.py [--ref /root/reference/SOURCE] [--check]
"""
import argparse
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DATA = os.path.join(ROOT, "dlpoly2-mbpol-heat_b200", "data")

TOKEN = re.compile(r"\s*(?:(t\d+)|a\[(\d+)\]|x\[(\d+)\]|df\[(\d+)\]|(\d+\.\d*)|([()+*]))")


class SparsePoly:
    """Polynomial that is at most linear in the a_k: {(k, monomial-tuple): weight}; k=-1 = no a factor."""
    __slots__ = ("t",)

    def __init__(self, t=None):
        self.t = t if t is not None else {}

    @staticmethod
    def coef(k):
        return SparsePoly({(k, ()): 1.0})

    @staticmethod
    def var(v):
        return SparsePoly({(-1, (v,)): 1.0})

    @staticmethod
    def const(c):
        return SparsePoly({(-1, ()): float(c)})

    def __add__(self, o):
        if len(self.t) < len(o.t):
            self, o = o, self
        r = dict(self.t)
        for key, w in o.t.items():
            r[key] = r.get(key, 0.0) + w
        return SparsePoly(r)

    def __mul__(self, o):
        r = {}
        for (k1, m1), w1 in self.t.items():
            for (k2, m2), w2 in o.t.items():
                if k1 >= 0 and k2 >= 0:
                    raise ValueError("polynomial is not linear in the coefficients")
                k = k1 if k1 >= 0 else k2
                m = tuple(sorted(m1 + m2))
                key = (k, m)
                r[key] = r.get(key, 0.0) + w1 * w2
        return SparsePoly(r)


def split_statements(path):
    """Return ({name: expr-string}, return-expr) of the generated eval body."""
    src = open(path).read()
    body = src[src.index("{", src.index("::eval")) + 1: src.rindex("return")]
    ret = src[src.rindex("return") + len("return"): src.index(";", src.rindex("return"))]
    stm = {}
    for s in body.split(";"):
        s = " ".join(s.split())
        if "=" not in s:
            continue
        lhs, rhs = s.split("=", 1)
        lhs = lhs.replace("const double", "").strip()
        stm[lhs] = rhs.strip()
    return stm, " ".join(ret.split())


def tokenize(expr):
    out, pos = [], 0
    while pos < len(expr):
        m = TOKEN.match(expr, pos)
        if not m:
            if expr[pos:].strip() == "":
                break
            raise ValueError("cannot tokenize: %r" % expr[pos:pos + 40])
        out.append(m)
        pos = m.end()
    return out


def symbolic_energy(path):
    stm, ret = split_statements(path)
    cache = {}

    def value_of(name):
        if name in cache:
            return cache[name]
        # iterative dependency resolution (the chains are deep: avoid recursion limits)
        stack = [name]
        while stack:
            cur = stack[-1]
            if cur in cache:
                stack.pop()
                continue
            deps = [d for d in set(re.findall(r"t\d+|df\[\d+\]", stm[cur])) if d not in cache]
            if deps:
                stack.extend(deps)
                continue
            cache[cur] = evaluate(stm[cur])
            stack.pop()
        return cache[name]

    def evaluate(expr):
        toks = tokenize(expr)
        pos = 0

        def atom():
            nonlocal pos
            m = toks[pos]
            pos += 1
            if m.group(1):
                return cache[m.group(1)]
            if m.group(2):
                return SparsePoly.coef(int(m.group(2)))
            if m.group(3):
                return SparsePoly.var(int(m.group(3)))
            if m.group(4):
                return cache["df[%s]" % m.group(4)]
            if m.group(5):
                return SparsePoly.const(m.group(5))
            if m.group(6) == "(":
                v = summ()
                assert toks[pos].group(6) == ")"
                pos += 1
                return v
            raise ValueError("unexpected token %r" % m.group(0))

        def prod():
            nonlocal pos
            v = atom()
            while pos < len(toks) and toks[pos].group(6) == "*":
                pos += 1
                v = v * atom()
            return v

        def summ():
            nonlocal pos
            v = prod()
            while pos < len(toks) and toks[pos].group(6) == "+":
                pos += 1
                v = v + prod()
            return v

        v = summ()
        assert pos == len(toks), expr
        return v

    for d in set(re.findall(r"t\d+|df\[\d+\]", ret)):
        value_of(d)
    return evaluate(ret), len(cache), len(stm)


def poly_to_table(p, ncoef):
    rows, w = [], []
    for (k, m), wt in p.t.items():
        if wt == 0.0:
            continue
        assert 0 <= k < ncoef and 1 <= len(m) <= 4, (k, m)
        rows.append((k,) + m + (-1,) * (4 - len(m)))
        w.append(wt)
    order = sorted(range(len(rows)), key=lambda i: (len([v for v in rows[i][1:] if v >= 0]), rows[i][1:], rows[i][0]))
    return np.array([rows[i] for i in order], dtype=np.int16), np.array([w[i] for i in order], dtype=np.float64)


def literal_block(src, start_marker, end_marker="};"):
    i = src.index(start_marker)
    j = src.index(end_marker, i)
    nums = re.findall(r"^\s*([-+]?\d\.\d+e[-+]\d+)\s*,?\s*(?://.*)?$", src[i:j], flags=re.M)
    return np.array([float(s) for s in nums], dtype=np.float64)


def named_params(src, names):
    out = {}
    for n in names:
        m = re.search(r"\b%s\s*=\s*([-+]?\d\.\d+e[-+]\d+)\s*;" % re.escape(n), src)
        if not m:
            raise KeyError(n)
        out[n] = float(m.group(1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/SOURCE")
    ap.add_argument("--out", default=os.path.join(DATA, "mbpol_model.npz"))
    args = ap.parse_args()

    out = {}
    for tag, fn, ncoef, nvar in (("2b", "poly-2b-v6x.cpp", 1153, 31), ("3b", "poly-3b-v2x.cpp", 1163, 36)):
        t0 = time.time()
        p, nused, ntot = symbolic_energy(os.path.join(args.ref, fn))
        mono, w = poly_to_table(p, ncoef)
        deg = (mono[:, 1:] >= 0).sum(1)
        print("%s: %d/%d statements feed E; %d monomials (deg hist %s); weights %s; vars<=%d; %.1fs" % (
            tag, nused, ntot, len(mono), np.bincount(deg).tolist(), np.unique(w).tolist(), mono[:, 1:].max(),
            time.time() - t0))
        assert mono[:, 1:].max() == nvar - 1
        out["mono" + tag], out["w" + tag] = mono, w

    s2 = open(os.path.join(args.ref, "mbpol-2b-poly.cpp")).read()
    s3 = open(os.path.join(args.ref, "mbpol-3b-poly.cpp")).read()
    out["coef2b"] = literal_block(s2, "the_poly[] = {")
    out["coef3b"] = literal_block(s3, "the_poly[] = {")
    assert out["coef2b"].shape == (1153,) and out["coef3b"].shape == (1163,), (out["coef2b"].shape, out["coef3b"].shape)
    p2 = named_params(s2, ["m_k_HH_intra", "m_k_OH_intra", "m_k_HH_coul", "m_k_OH_coul", "m_k_OO_coul",
                           "m_k_XH_main", "m_k_XO_main", "m_k_XX_main", "m_in_plane_gamma",
                           "m_out_of_plane_gamma", "m_r2i", "m_r2f"])
    p3 = named_params(s3, ["m_r3i", "m_r3f", "m_kHH_intra", "m_kOH_intra", "m_kHH", "m_kOH", "m_kOO",
                           "m_dHH_intra", "m_dOH_intra", "m_dHH", "m_dOH", "m_dOO"])
    for k, v in p2.items():
        out["par2b_" + k[2:]] = np.float64(v)
    for k, v in p3.items():
        out["par3b_" + k[2:]] = np.float64(v)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes")


if __name__ == "__main__":
    sys.setrecursionlimit(100000)
    main()
