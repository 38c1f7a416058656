#!/usr/bin/env python3
"""Generate the "coefficient tape" polynomial engine (energy + full gradient) from the model tables.

Input : dlpoly2-mbpol-heat_b200/data/mbpol_model.npz
Output: dlpoly2-mbpol-heat_b200/csrc/gen/tape_{2b,3b}.h

Why: measured on B200 (experiments/README.md) a straight-line FMA stream with one constant-bank coefficient per
FMA pair runs at ~15 % of the FP64 pipe (the uniform-constant path starves it), and a loop body stays fast only
while it is a few tens of KB.  So the *program* is moved out of the instruction stream into a data stream:

  * the trie of tools/gen_poly.py (same polynomial, same sharing of the degree-3 prefixes) is flattened into ROWS;
  * every coefficient use becomes one slot of a TAPE of doubles that the kernel streams global -> shared with
    cp.async (double buffered) and reads with broadcast LDS.128;
  * the only code is an interpreter loop plus one small straight-line body per ROW TYPE.  A row = the leaf children
    of one trie node: S += a_c * x_c ; g_c += a_c * P for the variables c of the row.  Permutational symmetry makes
    the distinct variable sets few (49 for the 3B polynomial with the descending-index trie, 25 for 2B ascending),
    so all bodies together are ~1.3k FMAs of code and x[], g[] keep static register names.

Rows.  Every trie node with leaf children is one independent ROW
      (type t, u, w, v, own, a_1..a_n):   P = x_u x_w x_v,  S = own + sum_c a_c x_c,
      E += S*P;  g_c += a_c*P (c in the row's static variable set);  g_u += S x_w x_v;  g_w += S x_u x_v;  g_v += S x_u x_w
(missing levels use a pseudo-variable UNIT = NV that is identically 1 and whose gradient column is a dump slot).
Rows carry no state between each other, so they are SORTED BY (type, u, w, v): the kernel runs one small counted loop per
(chunk, type, u) segment -- no per-row dispatch, x_u and the segment's g_u sum in registers, loop bodies of a few hundred
bytes that live in the L0 instruction cache, coefficients from the staged tape with static offsets, the row word (w,v)
from constant memory.  (tools/gen_fat.py goes one step further for the 3B polynomial: rows with the same (w,v) merged.)
Tape slots per row: [own, a_1..a_n, pad to even] so that every row starts 16-byte aligned.
"""
import argparse
import collections
import os

import numpy as np

import gen_poly as G

OUTD = G.OUTD
CHUNK = 2048          # doubles per staged tape chunk (16 KB), shared-memory tape engine (1024: +2 %, 4096: +7 % time)
CWINDOW = G.CWINDOW    # doubles of the constant-bank window (41 KB), constant tape engine
NPF = 4               # tape pairs of the NEXT row prefetched by every row (software pipelining across rows)
PAIR = False          # see emit_header
NSUM = 2              # independent partial sums of the Horner accumulator S (shortens the dependent DFMA chain)


def build(tag, mono, nv, mode):
    """Flatten the trie into independent ROWS.  A row belongs to one trie node that has leaf children:
         (type, u, w, v, own, [coefficient index per leaf])     with type = index of the leaf-variable set,
       P = x_u x_w x_v (missing levels use the pseudo-variable UNIT = nv, identically 1), S = own + sum_c a_c x_c,
       contribution E += S*P.  Internal nodes without leaf children but with an own coefficient become rows of the
       empty type.  Every monomial of the polynomial appears in exactly one row."""
    monos = [tuple(int(v) for v in r[1:] if v >= 0) for r in mono]
    coefs = [int(r[0]) for r in mono]
    seqs = G.choose_orders(monos, nv, mode)
    key = lambda m: tuple(sorted(m))
    lut = {key(m): k for m, k in zip(monos, coefs)}
    root = G.build_trie(seqs, [lut[key(s)] for s in seqs])
    UNIT = nv
    types = {}
    rows = []

    def visit(node, path):
        leaves = sorted((c for c in node.children.values() if not c.children), key=lambda c: c.var)
        if path and (leaves or node.coef >= 0):
            vs = tuple(c.var for c in leaves)
            t = types.setdefault(vs, len(types))
            p = list(path) + [UNIT] * (3 - len(path))          # (v, w, u)
            rows.append((t, p[2], p[1], p[0], node.coef, [c.coef for c in leaves]))
        for c in node.children.values():
            if c.children:
                assert len(path) < 3, "degree > 4 not supported"
                visit(c, path + [c.var])

    # leaves hanging directly off the root (bare linear terms a*x_v) are rows of the empty type with P = x_v
    for c in root.children.values():
        if not c.children:
            t = types.setdefault((), len(types))
            rows.append((t, UNIT, UNIT, c.var, c.coef, []))
    visit(root, [])
    # (type, u) first: u, the variable next to the leaves, is almost a function of the type (72 (type,u) runs for the
    # 3B polynomial, 51 for 2B), so it becomes a per-SEGMENT constant; inside a segment rows with the same w are adjacent
    rows.sort(key=lambda r: (r[0], r[1], r[2], r[3]))
    nmono = sum(len(r[5]) + (1 if r[4] >= 0 else 0) for r in rows)
    assert nmono == len(monos), (nmono, len(monos))
    return rows, types


def row_slots(n):
    n += 1
    return n + (n % 2)


def emit_header(tag, nv, rows, types, mode):
    T = tag.upper()
    by_id = {t: vs for vs, t in types.items()}
    nt = len(types)
    tape = []
    nfma = 0
    for (t, u, w, v, own, cs) in rows:
        tape.append(own)
        tape.extend(cs)
        if len(tape) % 2:
            tape.append(-1)
        nfma += 2 * len(cs) + 7

    def partition(limit):
        """Cut the sorted row list into pieces of <= limit tape slots; inside a piece, runs of equal (type, u) are
        segments (type | u << 8 | nrows << 16).  Returns (segs, pieces) with pieces = (first seg, first row, first slot)."""
        segs, pieces = [], []
        used, pos, cur = limit + 1, 0, None
        for i, (t, u, w, v, own, cs) in enumerate(rows):
            n = row_slots(len(cs))
            if used + n > limit:
                pieces.append((len(segs), i, pos))
                used, cur = 0, None
            if cur != (t, u):
                segs.append(t | (u << 8))
                cur = (t, u)
            segs[-1] += 1 << 16
            used += n
            pos += n
        assert pos == len(tape)
        pieces.append((len(segs), len(rows), len(tape)))
        return segs, pieces

    segs, chunks = partition(CHUNK)
    csegs, cpass = partition(CWINDOW - 2)      # two slots of every window hold the pass header
    rowvars = [w | (v << 8) for (t, u, w, v, own, cs) in rows]      # row word: the two per-row dynamic variables
    nfma += 2 * len(segs)
    L = ["// generated by tools/gen_tape.py (%s, trie order=%s) -- do not edit" % (tag, mode), "#pragma once",
         "#define TAPE%s_NV %d" % (T, nv), "#define TAPE%s_NROWS %d" % (T, len(rows)), "#define TAPE%s_NTAPE %d" % (T, len(tape)),
         "#define TAPE%s_NSEGS %d" % (T, len(segs)), "#define TAPE%s_NCHUNKS %d" % (T, len(chunks) - 1),
         "#define TAPE%s_NTYPES %d" % (T, nt), "#define TAPE%s_FP64_INSTR %d" % (T, nfma), "#define TAPE_CHUNK_DOUBLES %d" % CHUNK, "#define TAPE_NSUM %d" % NSUM, "#define TAPE_NPF %d" % NPF]
    L.append("static const unsigned TAPE%s_ROWVARS_HOST[%d] = {%s};" % (T, len(rowvars), ",".join(str(w) for w in rowvars)))
    L.append("static const unsigned TAPE%s_SEGS_HOST[%d] = {%s};" % (T, len(segs), ",".join(str(w) for w in segs)))
    L.append("static const short TAPE%s_COEFIDX_HOST[%d] = {%s};" % (T, len(tape), ",".join(str(c) for c in tape)))
    L.append("static const int TAPE%s_CHUNKS_HOST[%d][3] = {%s};" % (T, len(chunks), ",".join("{%d,%d,%d}" % c for c in chunks)))
    L.append("// constant-bank engine: passes of <= TAPE_CWINDOW_DOUBLES tape slots")
    L.append("#define TAPE%s_NCSEGS %d" % (T, len(csegs)))
    L.append("#define TAPE%s_NCPASS %d" % (T, len(cpass) - 1))
    L.append("#define TAPE_CWINDOW_DOUBLES %d" % CWINDOW)
    L.append("static const unsigned TAPE%s_CSEGS_HOST[%d] = {%s};" % (T, len(csegs), ",".join(str(w) for w in csegs)))
    L.append("static const int TAPE%s_CPASS_HOST[%d][3] = {%s};" % (T, len(cpass), ",".join("{%d,%d,%d}" % c for c in cpass)))
    L.append("// One counted loop per row type: TAPE_ROW_PRE takes (x_w, x_v) of the row, forms P and the first tape pair (q0x = own,")
    L.append("// q0y = first coefficient); the body accumulates S0/S1 (two partial Horner sums) and g[c] += a_c*P with static")
    L.append("// register names; TAPE_ROW_POST(bytes) adds S*x_w*x_v to the segment sum, updates g_v / the w-run sum and advances the tape.")
    # the NPF prefetched pairs: declaration list, loads (shared-memory engine), hand-over to the row, direct loads (constant engine)
    pf = ["nq%d%s" % (k, c) for k in range(NPF) for c in "xy"]
    L.append("#define TAPE_PF_DECL double %s;" % ", ".join(pf))
    L.append("#define TAPE_PF_LOADS(CP) %s" % " ".join("TAPE_LDS_V2(nq%dx, nq%dy, CP, %d);" % (k, k, 16 * k) for k in range(NPF)))
    L.append("#define TAPE_PF_TAKE(S) const double %s;" % ", ".join("q%d%s##S = nq%d%s" % (k, c, k, c) for k in range(NPF) for c in "xy"))
    L.append("#define TAPE_PF_CLOADS(S, OFF) double %s; %s" % (", ".join("q%d%s##S" % (k, c) for k in range(NPF) for c in "xy"),
             " ".join("TAPE_LD2(q%dx##S, q%dy##S, cp, (OFF) + %d);" % (k, k, 16 * k) for k in range(NPF))))
    # an if-chain, not a switch: an indirect branch (BRX) makes ptxas give up on warp-uniformity of the tape pointer.
    # --pair (experiment, slower: 35.3 vs 31.5 ms): rows evaluated in explicit PAIRS (a, b) with all loads of both rows
    # ahead of the gradient stores of either; the default leaves the interleaving to `#pragma unroll 2`.
    def row_body(vs, sfx, k, nbytes):
        body, have = [], set(range(NPF))          # pairs 0..NPF-1 were prefetched while the previous row was computed
        for i, v in enumerate(vs):
            slot = 1 + i
            pair, lane = slot // 2, slot % 2
            if pair not in have:
                have.add(pair)
                body.append("double q%dx%s, q%dy%s; TAPE_LD2(q%dx%s, q%dy%s, cp, %d);" % (pair, sfx, pair, sfx, pair, sfx, pair, sfx, k * nbytes + 16 * pair))
            a = "q%d%s%s" % (pair, "xy"[lane], sfx)
            acc = "S%d%s" % (i % NSUM, sfx)
            body.append("%s = fma(%s, x[%d], %s); g[%d] = fma(%s, P%s, g[%d]);" % (acc, a, v, acc, v, a, sfx, v))
        return " ".join(body)

    L.append("#define TAPE%s_SEGMENT_SWITCH(type, nrows) { \\" % T)
    for t in range(min(nt, int(os.environ.get("TAPE_MAXTYPES", "1000")))):      # TAPE_MAXTYPES: compile-time experiments only
        vs = by_id[t]
        nb = 8 * row_slots(len(vs))
        if PAIR:
            L.append("%sif ((type) == %d) { unsigned r_ = 0; \\" % ("else " if t else "", t))
            L.append("  TAPE_PAIR_LOOP(nrows) { TAPE_ROW_PRE(%d, a, 0) %s TAPE_ROW_PRE(%d, b, 1) %s TAPE_ROW_POST(a) TAPE_ROW_POST(b) TAPE_ROWS_ADVANCE(%d, 2) } \\" % (
                nb, row_body(vs, "a", 0, nb), nb, row_body(vs, "b", 1, nb), 2 * nb))
            L.append("  if (r_ < (nrows)) { TAPE_ROW_PRE(%d, a, 0) %s TAPE_ROW_POST(a) TAPE_ROWS_ADVANCE(%d, 1) } } \\" % (nb, row_body(vs, "a", 0, nb), nb))
        else:
            L.append("%sif ((type) == %d) { TAPE_ROW_LOOP(nrows) { TAPE_ROW_PRE(%d, a, 0) %s TAPE_ROW_POST(a) TAPE_ROWS_ADVANCE(%d, 1) } } \\" % (
                "else " if t else "", t, nb, row_body(vs, "a", 0, nb), nb))
    L.append("}")
    return "\n".join(L) + "\n", nfma, len(chunks) - 1, len(segs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", default=OUTD)
    ap.add_argument("--order2b", default="identity")
    ap.add_argument("--order3b", default="reverse")
    ap.add_argument("--nsum", type=int, default=NSUM)
    ap.add_argument("--npf", type=int, default=NPF)
    ap.add_argument("--pair", action="store_true")
    args = ap.parse_args()
    globals()["NSUM"] = args.nsum
    globals()["NPF"] = args.npf
    globals()["PAIR"] = args.pair
    d = np.load(G.DATA)
    os.makedirs(args.outdir, exist_ok=True)
    for tag, nv, mode in (("2b", 31, args.order2b), ("3b", 36, args.order3b)):
        rows, types = build(tag, d["mono" + tag], nv, mode)
        txt, nfma, nch, nseg = emit_header(tag, nv, rows, types, mode)
        with open(os.path.join(args.outdir, "tape_%s.h" % tag), "w") as f:
            f.write(txt)
        print("%s: %d rows of %d types in %d segments / %d chunks, %d FP64 instr per cluster" % (
            tag, len(rows), len(types), nseg, nch, nfma))


if __name__ == "__main__":
    main()
