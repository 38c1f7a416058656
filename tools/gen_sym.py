#!/usr/bin/env python3
"""Generate the SYMMETRY-BATCHED fat-row engine of the 3B polynomial (engine 4, csrc/mbpol_sym.cuh).

Why.  ncu of the fat-row engine (tools/gen_fat.py): the FP64 pipe is 38 % busy and half of all DFMA stall samples are
`short_scoreboard`, i.e. waiting for the broadcast LDS that delivers the next coefficient; one LDS.128 feeds only four
DFMAs and ptxas keeps few of them in flight at 255 registers.  The polynomial is invariant under the permutations of the
three water molecules (and of the two H of each water): a_{sigma(m)} = a_m.  So a subgroup K of that group is used to
evaluate |K| monomials with ONE coefficient load:

  * monomials are grouped into K-orbits; the trie roles (v, w = dynamic row pair, u = child, c = leaf) are chosen for ONE
    representative per orbit and PROPAGATED to its images, so the fat rows come in K-orbits ("super rows") whose members
    have identical coefficient tapes and differ only by a static renaming of x[] / g[] registers;
  * a super row loads |K| dynamic (w,v) pairs and runs one body in which every coefficient feeds 2|K| DFMAs on |K|
    independent chains:     for j < |K|:  S_kj += a x[sigma_j(c)];  g[sigma_j(c)] += a P_kj;
  * a monomial fixed by an element of K is produced by several images: its tape value is a |orbit| / |K|.

Only the molecule permutations are used (S3 or a subgroup): they cost 0.1-1.2 % extra monomial evaluations, an H swap 18 %.
The variable order of the trie lists K-orbits of variables contiguously, which keeps the propagated trie as regular as
the plain one (few distinct bodies, several super rows per body -- the bodies must be re-executed while they are still in
the instruction cache, experiments/README.md section 2).

Output: csrc/gen/sym_3b.h.  Same staging as engines 1 and 3 (cp.async double-buffered chunks, broadcast LDS.128).
"""
import argparse
import collections
import os

import numpy as np

import gen_poly as G

CHUNK = 2048
DISPATCH = os.environ.get("SYM_DISPATCH", "if")
ASMFMA = int(os.environ.get("SYM_ASMFMA", "0"))
POST = os.environ.get("SYM_POST", "inbody")       # row epilogue: "serial" (round-2 first form) | "batch" (all gradient-column loads first) | "inbody" (batch, inside the shape bodies)
WORDS_AHEAD = int(os.environ.get("SYM_WORDS_AHEAD", "1"))   # row words are read from constant memory this many rows before they are unpacked
# trie order and representative rule found by tools/search_sym_order.py for K = z (objective: FP64 instructions + 0.6 * leaf pairs of
# code, pair sums included): 71 589 instructions per trimer, 120 super rows of 47 shapes, 11 079 leaf pairs of code -- measured
# 14.75 ms; "1,2,0,11,8,7,10,9,4,5,6,3" / min (72 414 / 138 / 52 / 9 204): 14.97 ms; the reversed orbit order of the first version
# (79 791 / 207 / 61 / 9 408): 17.4 ms.  Orders must be MEASURED: "0,11,10,9,8,7,1,2,6,5,3,4" / min (77 556 instructions, only 6 744
# leaf pairs of code) runs at 16.2 ms in lock step and at 32.8 ms with SYM_SPLIT_ORDER (experiments/README.md section 9).
DEFAULT_ORDER = "1,2,0,10,9,8,7,5,11,6,4,3"
DEFAULT_REPSEL = "max"
NV = 36
UNIT = NV
# atoms: 0 Oa 1 Ha1 2 Ha2 3 Ob 4 Hb1 5 Hb2 6 Oc 7 Hc1 8 Hc2 ; variable n = distance (V3A[n], V3B[n]) (csrc/mbpol_device.cuh)
V3A = [1, 4, 7, 0, 0, 3, 3, 6, 6, 1, 1, 1, 1, 2, 2, 2, 2, 4, 4, 5, 5, 0, 0, 0, 0, 3, 3, 3, 3, 6, 6, 6, 6, 0, 0, 3]
V3B = [2, 5, 8, 1, 2, 4, 5, 7, 8, 4, 5, 7, 8, 4, 5, 7, 8, 7, 8, 7, 8, 4, 5, 7, 8, 1, 2, 7, 8, 1, 2, 4, 5, 3, 6, 6]


def molperm(p):
    """atom permutation that sends molecule m to molecule p[m]"""
    return tuple(3 * p[a // 3] + a % 3 for a in range(9))


def varperm(ap):
    idx = {frozenset((a, b)): n for n, (a, b) in enumerate(zip(V3A, V3B))}
    return tuple(idx[frozenset((ap[a], ap[b]))] for a, b in zip(V3A, V3B)) + (UNIT,)


GROUPS = {   # static images sigma_j (as molecule permutations); image 0 is the identity
    "1": [(0, 1, 2)],
    "t": [(0, 1, 2), (1, 0, 2)],
    "z": [(0, 1, 2), (1, 2, 0), (2, 0, 1)],
    "s3": [(0, 1, 2), (1, 2, 0), (2, 0, 1), (1, 0, 2), (0, 2, 1), (2, 1, 0)],
}


def orbit_order(vps):
    """variable order in which the K-orbits of variables are contiguous (orbits sorted by their smallest member)"""
    seen, order = set(), []
    for v in range(NV):
        if v in seen:
            continue
        orb = []
        for vp in vps:
            if vp[v] not in orb:
                orb.append(vp[v])
        seen |= set(orb)
        order.extend(orb)
    return order


def build(mono, group="z", reverse=True, repsel="max", order=None):
    """-> (super rows, shapes, vps).  super row = (shape id, [(w_j, v_j)], children); child = (us, own, leaves) with
    us = (u_j), own = (coef index, num) or None, leaves = [((c_j), coef index, num)]; the tape value is a * num / |K|."""
    vps = [varperm(molperm(p)) for p in GROUPS[group]]
    K = len(vps)
    if order is None:
        order = orbit_order(vps)
        if reverse:
            order = order[::-1]
    else:      # explicit trie order: a permutation of the K-orbits of variables (tools/search_sym_order.py)
        orbs = orbit_order(vps)
        n = len(vps)
        blocks = [orbs[i:i + n] for i in range(0, len(orbs), n)]
        assert all(len(set(vp[v] for vp in vps)) == n for v in range(NV)), "explicit orders need a group without fixed variables"
        order = [v for b in order for v in blocks[b]]
        assert sorted(order) == list(range(NV))
    rank = {v: i for i, v in enumerate(order)}
    monos = {}
    for r in mono:
        monos[tuple(sorted(int(v) for v in r[1:] if v >= 0))] = int(r[0])
    pick = max if repsel == "max" else min
    seen = set()
    rows = collections.defaultdict(lambda: collections.defaultdict(lambda: [None, []]))
    pend3 = []
    for m, k in monos.items():
        if m in seen:
            continue
        imgs = [tuple(sorted(vp[v] for v in m)) for vp in vps]
        orbit = set(imgs)
        assert all(monos[i] == k for i in orbit), "the monomial table is not invariant under the group"
        seen |= orbit
        rep = pick(orbit, key=lambda im: tuple(sorted((rank[v] for v in im), reverse=True)))
        roles = sorted(rep, key=lambda v: rank[v], reverse=True)          # v, w, (u, (c))
        num = len(orbit)
        if len(roles) == 4:
            v, w, u, c = roles
            rows[tuple((vp[w], vp[v]) for vp in vps)][tuple(vp[u] for vp in vps)][1].append((tuple(vp[c] for vp in vps), k, num))
        elif len(roles) == 3:
            pend3.append((roles, k, num))
        else:
            v, w = roles
            key = tuple((vp[w], vp[v]) for vp in vps)
            assert rows[key][(UNIT,) * K][0] is None
            rows[key][(UNIT,) * K][0] = (k, num)
    for (v, w, u), k, num in pend3:     # degree 3: the own coefficient of child u if that child exists, else a leaf of the unit child
        key = tuple((vp[w], vp[v]) for vp in vps)
        us = tuple(vp[u] for vp in vps)
        if us in rows[key]:
            assert rows[key][us][0] is None
            rows[key][us][0] = (k, num)
        else:
            rows[key][(UNIT,) * K][1].append((us, k, num))
    shapes, out = {}, []
    for key, ch in rows.items():
        children = []
        for us, (own, leaves) in sorted(ch.items()):
            leaves.sort()
            children.append((us, own, leaves))
        shape = tuple((us, tuple(cs for cs, _, _ in leaves)) for us, own, leaves in children)
        s = shapes.setdefault(shape, len(shapes))
        out.append((s, list(key), children))
    out.sort(key=lambda r: (r[0], r[1]))
    return out, {s: sh for sh, s in shapes.items()}, vps


# Pair sums.  The polynomial is also invariant under the exchange of the two H of one water; a prefix (w, v, u) that does not
# tell them apart has sibling leaves c, c' = h(c) with the SAME coefficient:  a x_c + a x_c' = a (x_c + x_c').  For the pairs listed
# here (one K-orbit: the two intramolecular O-H variables of each water -- 52 % of all such siblings) the sum y = x_c + x_c' is
# a virtual variable NV + 1 + p with its own gradient accumulator: one leaf (2 |K| DFMAs) instead of two, and at the end
# g_c += g_y, g_c' += g_y.  Costs 4 registers per pair.
PAIRS = [(3, 4), (5, 6), (7, 8)] if os.environ.get("SYM_PAIRS", "1") == "1" else []


def merge_pairs(rows):
    """-> (rows, shapes) with sibling leaves of equal coefficient over a listed pair merged into one leaf of the virtual variable.
    Only merges that EVERY super row of a shape can do are applied, so that rows keep sharing their bodies (merging row by row
    gives 5.3 % fewer FP64 instructions but 85 instead of 52 shapes and 1.8 x the code, which costs more than it saves)."""
    virt = {frozenset(p): NV + 1 + i for i, p in enumerate(PAIRS)}

    def candidates(children):
        """set of (child index, leaf a, leaf b) that could merge in this row"""
        c = set()
        for ci, (us, own, leaves) in enumerate(children):
            by = collections.defaultdict(list)
            for cs, k, num in leaves:
                by[(k, num)].append(cs)
            for lst in by.values():
                lst = sorted(lst)
                used = set()
                for i, a in enumerate(lst):
                    if i in used:
                        continue
                    for j in range(i + 1, len(lst)):
                        if j not in used and all(frozenset((a[q], lst[j][q])) in virt for q in range(len(a))):
                            c.add((ci, a, lst[j]))
                            used |= {i, j}
                            break
        return c
    common = {}
    for (s, key, children) in rows:
        c = candidates(children)
        common[s] = c if s not in common else (common[s] & c)
    shapes, out = {}, []
    for (s, key, children) in rows:
        nch = []
        for ci, (us, own, leaves) in enumerate(children):
            coef = {cs: (k, num) for cs, k, num in leaves}
            drop, add = set(), []
            for (cj, a, b) in common[s]:
                if cj == ci:
                    assert coef[a] == coef[b]
                    drop |= {a, b}
                    add.append((tuple(virt[frozenset((a[q], b[q]))] for q in range(len(a))),) + coef[a])
            merged = sorted([l for l in leaves if l[0] not in drop] + add)
            nch.append((us, own, merged))
        shape = tuple((us, tuple(cs for cs, _, _ in leaves)) for us, own, leaves in nch)
        sid = shapes.setdefault(shape, len(shapes))
        out.append((sid, key, nch))
    out.sort(key=lambda r: (r[0], r[1]))
    return out, {s: sh for sh, s in shapes.items()}


def xn(c):
    return "x[%d]" % c if c <= NV else "xq[%d]" % (c - NV - 1)


def gn(c):
    return "g[%d]" % c if c <= NV else "gq[%d]" % (c - NV - 1)


def stats(rows, shapes, K):
    """(FP64 instructions per cluster, leaf pairs of code, super rows, shapes) of a schedule"""
    nfp = sum(K * (4 + sum(2 * len(ls) + (4 if us[0] != UNIT else 1) for us, own, ls in ch)) for s, key, ch in rows)
    code = sum(sum(len(ls) for us, ls in sh) for sh in shapes.values()) * K
    return nfp, code, len(rows), len(shapes)


def slots_of(shape):
    n = sum(1 + len(ls) for us, ls in shape)
    return n + (n % 2)


def emit(rows, shapes, vps, ahead, nsum):
    K = len(vps)
    tape, rowvars, rowmeta = [], [], []
    for (s, key, children) in rows:
        start = len(tape)
        for us, own, leaves in children:
            tape.append(own if own else (-1, 0))
            tape.extend((k, num) for cs, k, num in leaves)
        if (len(tape) - start) % 2:
            tape.append((-1, 0))
        assert len(tape) - start == slots_of(shapes[s])
        cols = [c for wv in key for c in wv]
        alias = 1 if len(set(cols)) < len(cols) else 0      # some dynamic column is updated twice by this super row: the epilogue must not batch its loads
        rowmeta.append((s, alias, key))
    segs, chunks, rowoff = [], [], []
    used, pos, cur = CHUNK + 1, 0, None
    for i, (s, key, children) in enumerate(rows):
        n = slots_of(shapes[s])
        assert n <= CHUNK
        if used + n > CHUNK:
            chunks.append((len(segs), i, pos))
            used, cur = 0, None
        rowoff.append(8 * used)
        if cur != s:
            segs.append(s)
            cur = s
        segs[-1] += 1 << 8
        used += n
        pos += n
    chunks.append((len(segs), len(rows), len(tape)))
    # row words: w | v << 6 | shape << 12 | alias << 19 | (byte offset of the row's tape slots inside its chunk / 16) << 20
    assert len(shapes) <= 128 and NV + 1 + len(PAIRS) <= 64 and CHUNK * 8 // 16 <= 1024
    for (sid, alias, key), off in zip(rowmeta, rowoff):
        assert off % 16 == 0
        rowvars.extend(w | (v << 6) | (sid << 12) | (alias << 19) | ((off // 16) << 20) for (w, v) in key)
    nfp = 0
    body_of = {}
    jj = range(K)
    shape_alias = collections.defaultdict(set)
    for (s, key, children) in rows:
        cols = [c for wv in key for c in wv]
        shape_alias[s].add(len(set(cols)) < len(cols))
    # row epilogue pieces (see the POST comment further down)
    ldw = " ".join("const double gw_%d = tape_lds(aw_%d + gs_delta);" % (j, j) for j in jj)
    ldv = " ".join("const double gv_%d = tape_lds(av_%d + gs_delta);" % (j, j) for j in jj)
    stw = " ".join("tape_sts(aw_%d + gs_delta, fma(Tw_%d, xv_%d, gw_%d));" % (j, j, j, j) for j in jj)
    stv = " ".join("tape_sts(av_%d + gs_delta, fma(Tw_%d, xw_%d, gv_%d));" % (j, j, j, j) for j in jj)
    ee = " ".join("e = fma(Tw_%d, p2_%d, e);" % (j, j) for j in jj)
    serial = " ".join("tape_sts(aw_%d + gs_delta, fma(Tw_%d, xv_%d, tape_lds(aw_%d + gs_delta))); tape_sts(av_%d + gs_delta, fma(Tw_%d, xw_%d, tape_lds(av_%d + gs_delta)));"
                      % (j, j, j, j, j, j, j, j) for j in jj)
    for s, shape in shapes.items():
        body, off = [], 0
        npairs = slots_of(shape) // 2
        loaded = [0]

        def load_upto(n):
            while loaded[0] < min(n, npairs):
                k = loaded[0]
                body.append("double q%dx, q%dy; TAPE_LDS_V2(q%dx, q%dy, cp, %d);" % (k, k, k, k, 16 * k))
                loaded[0] += 1

        def slot(i):
            pair, lane = i // 2, i % 2
            load_upto(pair + 1 + ahead)
            return "q%d%s" % (pair, "xy"[lane])
        load_upto(ahead)
        hoist = POST == "inbody" and shape_alias[s] == {False}
        for k, (us, ls) in enumerate(shape):
            if hoist and k == len(shape) - 1:      # gradient columns of the dynamic variables: loaded a whole child ahead of their use
                body.append(ldw + " " + ldv)
            own = slot(off)
            for j in range(K):
                if us[j] != UNIT:
                    body.append("const double P%d_%d = p2_%d * x[%d];" % (j, k, j, us[j]))
            acc = [["S%d_%d_%d" % (j, k, n) for n in range(nsum)] for j in range(K)]
            for j in range(K):
                body.append("double %s = %s%s;" % (acc[j][0], own, "".join(", %s = 0.0" % a for a in acc[j][1:])))
            for i, cs in enumerate(ls):
                a = slot(off + 1 + i)
                for j in range(K):
                    pu = "p2_%d" % j if us[j] == UNIT else "P%d_%d" % (j, k)
                    if ASMFMA:      # inline PTX keeps the coefficient in the FIRST multiplicand slot of all 2|K| FMAs (operand reuse cache)
                        body.append("SYM_FMA(%s, %s, %s); SYM_FMA(%s, %s, %s);" % (acc[j][i % nsum], a, xn(cs[j]), gn(cs[j]), a, pu))
                    else:
                        body.append("%s = fma(%s, %s, %s); %s = fma(%s, %s, %s);" % (acc[j][i % nsum], a, xn(cs[j]), acc[j][i % nsum], gn(cs[j]), a, pu, gn(cs[j])))
            nacc = max(1, min(nsum, len(ls)))
            for j in range(K):
                ssum = " + ".join(acc[j][:nacc])
                if us[j] == UNIT:
                    body.append("Tw_%d += %s;" % (j, ssum))
                else:
                    body.append("{ const double S_ = %s; g[%d] = fma(S_, p2_%d, g[%d]); Tw_%d = fma(S_, x[%d], Tw_%d); }" % (ssum, us[j], j, us[j], j, us[j], j))
            off += 1 + len(ls)
        if POST == "inbody":
            if hoist:
                body.append("%s %s %s" % (ee, stw, stv))
            elif shape_alias[s] == {True}:
                body.append("%s %s" % (ee, serial))
            else:
                body.append("%s if (alias) { %s } else { %s %s %s %s }" % (ee, serial, ldw, ldv, stw, stv))
        body_of[s] = " ".join(body)
    for (s, key, children) in rows:
        nfp += K * (4 + sum(2 * len(ls) + (4 if us[0] != UNIT else 1) for us, own, ls in children))
    Tg = "3B"
    L = ["// generated by tools/gen_sym.py (|K| = %d static images) -- do not edit" % K, "#pragma once",
         "#define SYM%s_NV %d" % (Tg, NV), "#define SYM%s_K %d" % (Tg, K), "#define SYM%s_NROWS %d" % (Tg, len(rows)),
         "#define SYM%s_NTAPE %d" % (Tg, len(tape)), "#define SYM%s_NSEGS %d" % (Tg, len(segs)), "#define SYM%s_NCHUNKS %d" % (Tg, len(chunks) - 1),
         "#define SYM%s_NSHAPES %d" % (Tg, len(shapes)), "#define SYM%s_FP64_INSTR %d" % (Tg, nfp), "#define SYM_CHUNK_DOUBLES %d" % CHUNK]
    rv = rowvars + rowvars[-K:] * (1 + WORDS_AHEAD)      # spare super rows: the header of the next row(s) is prefetched
    L.append("static const unsigned SYM%s_ROWVARS_HOST[%d] = {%s};" % (Tg, len(rv), ",".join(str(w) for w in rv)))
    L.append("static const unsigned SYM%s_SEGS_HOST[%d] = {%s};" % (Tg, len(segs), ",".join(str(w) for w in segs)))
    L.append("// tape slot = coefficient index (< 0: zero) and numerator: value = a * num / |K|")
    L.append("static const short SYM%s_COEFIDX_HOST[%d] = {%s};" % (Tg, len(tape), ",".join(str(c) for c, n in tape)))
    L.append("static const unsigned char SYM%s_COEFNUM_HOST[%d] = {%s};" % (Tg, len(tape), ",".join(str(n) for c, n in tape)))
    L.append("static const int SYM%s_CHUNKS_HOST[%d][3] = {%s};" % (Tg, len(chunks), ",".join("{%d,%d,%d}" % c for c in chunks)))
    L.append("#define SYM_ROWVARS_SPARE %d" % (K * (1 + WORDS_AHEAD)))

    if PAIRS:
        L.append("#define SYM_PAIR_INIT double xq[%d] = {%s}, gq[%d] = {%s};" % (len(PAIRS), ", ".join("x[%d] + x[%d]" % p for p in PAIRS), len(PAIRS), ", ".join("0.0" for p in PAIRS)))
        L.append("#define SYM_PAIR_FINI %s" % " ".join("g[%d] += gq[%d]; g[%d] += gq[%d];" % (p[0], i, p[1], i) for i, p in enumerate(PAIRS)))
    else:
        L.append("#define SYM_PAIR_INIT")
        L.append("#define SYM_PAIR_FINI")
    # row header macros for K images.  The header of the NEXT super row (row words -> column addresses -> the x values of its dynamic
    # variables) is fetched while this row's body runs; with WORDS_AHEAD the row words themselves are read one row earlier
    # still, so that the constant-memory latency is off the address computation.
    jj = range(K)
    L.append("#define SYM_ROW_DECL unsigned nshape, nalias, noff, %s; double %s;%s" % (
        ", ".join("nw_%d, nv_%d" % (j, j) for j in jj), ", ".join("nxw_%d, nxv_%d" % (j, j) for j in jj),
        (" unsigned %s;" % ", ".join("rw_%d" % j for j in jj)) if WORDS_AHEAD else ""))
    unpack = " ".join(
        "{ %snw_%d = xs_t + (vw_%d & 63u) * TAPE_COL; nv_%d = xs_t + ((vw_%d >> 6) & 63u) * TAPE_COL; nxw_%d = tape_lds(nw_%d); nxv_%d = tape_lds(nv_%d); }"
        % ("nshape = (vw_%d >> 12) & 127u; nalias = (vw_%d >> 19) & 1u; noff = (vw_%d >> 20) << 4; " % (j, j, j) if j == 0 else "", j, j, j, j, j, j, j, j) for j in jj)
    if WORDS_AHEAD:
        L.append("#define SYM_ROW_WORDS(ROWIDX) { %s }" % " ".join("rw_%d = rowvars[(ROWIDX) * %d + %d];" % (j, K, j) for j in jj))
        L.append("#define SYM_ROW_FETCH(ROWIDX, ROWAFTER) { %s %s SYM_ROW_WORDS(ROWAFTER) }" % (" ".join("const unsigned vw_%d = rw_%d;" % (j, j) for j in jj), unpack))
    else:
        L.append("#define SYM_ROW_WORDS(ROWIDX)")
        L.append("#define SYM_ROW_FETCH(ROWIDX, ROWAFTER) { %s %s }" % (" ".join("const unsigned vw_%d = rowvars[(ROWIDX) * %d + %d];" % (j, K, j) for j in jj), unpack))
    if POST == "inbody":
        # the epilogue lives in the shape bodies; the header of the next row is fetched AFTER the body (its x loads fly during the
        # dispatch of the next row), straight into the registers the next row reads: no double-buffered header, no register moves
        L.append("#define SYM_ROW_PRE const unsigned shape = nshape; const bool alias = nalias != 0u; cp = cbase + noff; %s" % " ".join(
            "const unsigned aw_%d = nw_%d, av_%d = nv_%d; const double xw_%d = nxw_%d, xv_%d = nxv_%d; const double p2_%d = xv_%d * xw_%d; double Tw_%d = 0.0;"
            % (j, j, j, j, j, j, j, j, j, j, j, j) for j in jj))
        L.append("#define SYM_ROW_POST SYM_ROW_FETCH(row_next, row_next2)")
    else:
      L.append("#define SYM_ROW_PRE const unsigned shape = nshape; const bool alias = nalias != 0u; cp = cbase + noff; %s SYM_ROW_FETCH(row_next, row_next2)" % " ".join(
        "const unsigned aw_%d = nw_%d, av_%d = nv_%d; const double xw_%d = nxw_%d, xv_%d = nxv_%d; const double p2_%d = xv_%d * xw_%d; double Tw_%d = 0.0;"
        % (j, j, j, j, j, j, j, j, j, j, j, j) for j in jj))
    if POST == "inbody":
        pass
    elif POST == "serial":      # first form: six dependent load -> FMA -> store round trips (ptxas cannot reorder the possibly aliasing columns)
        L.append("#define SYM_ROW_POST %s" % " ".join(
            "e = fma(Tw_%d, p2_%d, e); tape_sts(aw_%d + gs_delta, fma(Tw_%d, xv_%d, tape_lds(aw_%d + gs_delta))); tape_sts(av_%d + gs_delta, fma(Tw_%d, xw_%d, tape_lds(av_%d + gs_delta)));"
            % (j, j, j, j, j, j, j, j, j, j) for j in jj))
    else:
        # batched: unless two of the 2|K| dynamic columns of the super row coincide (squares x_w^2, w_i = v_j, a variable fixed by
        # an element of K; 18 of the 207 rows for K = z, flagged `alias` in the row word) all 2|K| loads go first, then the FMAs,
        # then the stores -- one shared-memory round trip per row instead of 2|K| dependent ones.  Flagged rows stay serial.
        ldw = " ".join("const double gw_%d = tape_lds(aw_%d + gs_delta);" % (j, j) for j in jj)
        ldv = " ".join("const double gv_%d = tape_lds(av_%d + gs_delta);" % (j, j) for j in jj)
        stw = " ".join("tape_sts(aw_%d + gs_delta, fma(Tw_%d, xv_%d, gw_%d));" % (j, j, j, j) for j in jj)
        stv = " ".join("tape_sts(av_%d + gs_delta, fma(Tw_%d, xw_%d, gv_%d));" % (j, j, j, j) for j in jj)
        ee = " ".join("e = fma(Tw_%d, p2_%d, e);" % (j, j) for j in jj)
        serial = " ".join("tape_sts(aw_%d + gs_delta, fma(Tw_%d, xv_%d, tape_lds(aw_%d + gs_delta))); tape_sts(av_%d + gs_delta, fma(Tw_%d, xw_%d, tape_lds(av_%d + gs_delta)));"
                          % (j, j, j, j, j, j, j, j) for j in jj)
        L.append("#define SYM_ROW_POST if (!alias) { %s %s %s %s %s } else { %s %s }" % (ldw, ldv, ee, stw, stv, ee, serial))
    if DISPATCH == "switch":      # one loop over the super rows, dispatch inside (see tools/gen_fat.py); jump table instead of a compare tree
        L.append("#define SYM%s_ROW_SWITCH(shape) switch (shape) { \\" % Tg)
        for s in sorted(shapes):
            nb = 8 * slots_of(shapes[s])
            L.append("case %d: { %s cp += %d; } break; \\" % (s, body_of[s], nb))
        L.append("default: break; }")
    else:
        L.append("#define SYM%s_ROW_SWITCH(shape) { \\" % Tg)
        for s in sorted(shapes):
            nb = 8 * slots_of(shapes[s])
            L.append("%sif ((shape) == %d) { %s cp += %d; } \\" % ("else " if s else "", s, body_of[s], nb))
        L.append("}")
    return "\n".join(L) + "\n", nfp, len(segs), len(chunks) - 1, len(tape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=os.environ.get("SYM_GROUP", "z"), choices=sorted(GROUPS))
    ap.add_argument("--forward", action="store_true", help="trie order = orbit order (default: reversed)")
    ap.add_argument("--repsel", default=os.environ.get("SYM_REPSEL", DEFAULT_REPSEL), choices=["max", "min"])
    ap.add_argument("--order", default=os.environ.get("SYM_ORDER", DEFAULT_ORDER), help="trie order: comma-separated permutation of the K-orbits of variables ('' = --forward / reversed orbit order)")
    ap.add_argument("--ahead", type=int, default=int(os.environ.get("SYM_AHEAD", "6")))
    ap.add_argument("--nsum", type=int, default=int(os.environ.get("SYM_NSUM", "1")))
    ap.add_argument("--dispatch", default=DISPATCH, choices=["if", "switch"])
    ap.add_argument("--asmfma", type=int, default=ASMFMA)
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("SYM_CHUNK", str(CHUNK))), help="doubles per staged tape chunk")
    ap.add_argument("--out", default=os.path.join(G.OUTD, "sym_3b.h"))
    args = ap.parse_args()
    d = np.load(G.DATA)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    ap_dispatch = args.dispatch
    globals()["CHUNK"] = args.chunk
    globals()["DISPATCH"] = ap_dispatch
    globals()["ASMFMA"] = args.asmfma
    order = [int(v) for v in args.order.split(",")] if args.order and args.group == "z" else None
    rows, shapes, vps = build(d["mono3b"], args.group, not args.forward, args.repsel, order=order)
    if PAIRS and args.group == "z":
        rows, shapes = merge_pairs(rows)
    txt, nfp, nseg, nch, ntape = emit(rows, shapes, vps, args.ahead, args.nsum)
    open(args.out, "w").write(txt)
    code = sum(sum(len(ls) for us, ls in sh) for sh in shapes.values()) * len(vps)
    print("3b: |K| = %d: %d super rows of %d shapes in %d segments / %d chunks, %d tape slots, %d FP64 instr per cluster, %d leaf pairs of code (largest body %d)" % (
        len(vps), len(rows), len(shapes), nseg, nch, ntape, nfp, code, max(sum(len(ls) for us, ls in sh) for sh in shapes.values()) * len(vps)))


if __name__ == "__main__":
    main()
