#!/usr/bin/env python3
"""Generate the "coefficient tape" polynomial engine (energy + full gradient) from the model tables.

Input : dlpoly2-mbpol-heat_b200/data/mbpol_model.npz
Output: dlpoly2-mbpol-heat_b200/csrc/gen/tape_{2b,3b}.h

Why: measured on B200 (experiments/README.md) a straight-line FMA stream with one constant-bank coefficient per
FMA pair runs at ~15 % of the FP64 pipe (the uniform-constant path starves it), and a loop body stays fast only
while it is a few tens of KB.  So the *program* is moved out of the instruction stream into a data stream:

  * the trie walk of tools/gen_poly.py (same math: p top-down, s = multivariate Horner bottom-up,
    dE/dx_c += p(parent) * s(child)) is serialised into a short list of OPS (one 32-bit word each, constant memory);
  * every coefficient use becomes one slot of a TAPE of doubles that the kernel streams global -> shared with
    cp.async (double buffered) and reads with broadcast LDS.128;
  * the only code is an interpreter loop plus one small straight-line body per ROW TYPE.  A row = the leaf children
    of one trie node: S += a_c * x_c ; g_c += a_c * P for the variables c of the row.  Permutational symmetry makes
    the distinct variable sets few (49 for the 3B polynomial with the descending-index trie, 25 for 2B ascending),
    so all bodies together are ~1.3k FMAs of code and x[], g[] keep static register names.

Op word: bits 0-7 handler (row type t for NODE3, else NTYPES + {0 ENTER1, 1 ENTER2, 2 EXIT2, 3 EXIT1, 4 LIN}),
bits 8-19 tape bytes consumed, bits 20-25 variable.  Tape slots in brackets; every op consumes an even number of
slots so that each row starts 16-byte aligned.
   NODE3  u,t    xu = x[u]; P = p2*xu; S = own; row body t (S += a_c x_c, g_c += a_c P);   [own, row t]
                 s2 += S*xu; g[u] += S*p2
   ENTER1 v      xv = x[v]; s1 = own                                                       [own, pad]
   ENTER2 w      xw = x[w]; p2 = xv*xw; s2 = own                                           [own, pad]
   EXIT2  w      s1 += s2*xw; g[w] += s2*xv
   EXIT1  v      E += s1*xv; g[v] += s1
   LIN    v      E += a*x[v]; g[v] += a                                                    [a, pad]
Leaves that hang directly under a depth-1 or depth-2 node are expressed with the same NODE3 handler through a
pseudo-variable UNIT = NV that is identically 1 (its gradient column is a dump slot).
Dynamic-index reads of x and updates of g (once per op) go through per-thread shared-memory columns.
"""
import argparse
import collections
import os

import numpy as np

import gen_poly as G

OUTD = G.OUTD
CHUNK = 1024          # doubles per staged tape chunk (8 KB)
H_ENTER1, H_ENTER2, H_EXIT2, H_EXIT1, H_LIN = range(5)   # handler ids after the NTYPES row types


def build(tag, mono, nv, mode):
    """Serialise the trie walk.  Returns (ops, tape, types): ops are (kind, var, type) tuples with kind in
    {"node3","enter1","enter2","exit2","exit1","lin"}; tape entries are coefficient indices or -1 (0.0)."""
    monos = [tuple(int(v) for v in r[1:] if v >= 0) for r in mono]
    coefs = [int(r[0]) for r in mono]
    seqs = G.choose_orders(monos, nv, mode)
    key = lambda m: tuple(sorted(m))
    lut = {key(m): k for m, k in zip(monos, coefs)}
    root = G.build_trie(seqs, [lut[key(s)] for s in seqs])
    UNIT = nv             # pseudo-variable that is identically 1 (lets ROW1/ROW2 reuse the NODE3 handler)

    types = {}
    ops, tape = [], []

    def leaf_row(node):
        leaves = sorted((c for c in node.children.values() if not c.children), key=lambda c: c.var)
        if not leaves:
            return None, []
        vs = tuple(c.var for c in leaves)
        t = types.setdefault(vs, len(types))
        return t, [c.coef for c in leaves]

    def pad():
        if len(tape) % 2:
            tape.append(-1)

    def node3(var, t, own, cs):
        ops.append(("node3", var, t))
        tape.append(own)
        tape.extend(cs)
        pad()

    for v in sorted(root.children):
        n1 = root.children[v]
        if not n1.children:
            ops.append(("lin", v, 0))
            tape.extend([n1.coef, -1])
            continue
        ops.append(("enter1", v, 0))
        tape.extend([n1.coef, -1])
        t, cs = leaf_row(n1)
        if t is not None:          # leaves directly under x_v: a depth-2 "unit" node holding one row
            ops.append(("enter2", UNIT, 0))
            tape.extend([-1, -1])
            node3(UNIT, t, -1, cs)
            ops.append(("exit2", UNIT, 0))
        for w in sorted(c.var for c in n1.children.values() if c.children):
            n2 = n1.children[w]
            ops.append(("enter2", w, 0))
            tape.extend([n2.coef, -1])
            t, cs = leaf_row(n2)
            if t is not None:      # leaves directly under x_v x_w: a row multiplied by the unit variable
                node3(UNIT, t, -1, cs)
            for u in sorted(c.var for c in n2.children.values() if c.children):
                n3 = n2.children[u]
                assert all(not c.children for c in n3.children.values()), "degree > 4 not supported"
                t, cs = leaf_row(n3)
                node3(u, t, n3.coef, cs)
            ops.append(("exit2", w, 0))
        ops.append(("exit1", v, 0))
    return ops, tape, types


def op_slots(op, by_id):
    kind, _, typ = op
    if kind in ("enter1", "enter2", "lin"):
        return 2
    if kind == "node3":
        n = 1 + len(by_id[typ])
        return n + (n % 2)
    return 0


def emit_header(tag, nv, ops, tape, types, mode):
    T = tag.upper()
    by_id = {t: vs for vs, t in types.items()}
    nt = len(types)
    hid = {"enter1": nt + H_ENTER1, "enter2": nt + H_ENTER2, "exit2": nt + H_EXIT2, "exit1": nt + H_EXIT1, "lin": nt + H_LIN}
    chunks = [(0, 0)]
    used = pos = 0
    words = []
    nfma = 0
    for i, op in enumerate(ops):
        n = op_slots(op, by_id)
        if used + n > CHUNK:
            chunks.append((i, pos))
            used = 0
        used += n
        pos += n
        kind, var, typ = op
        h = typ if kind == "node3" else hid[kind]
        assert h < 256 and 8 * n < 4096 and var < 64
        words.append(h | ((8 * n) << 8) | (var << 20))       # handler | tape bytes consumed | variable
        nfma += {"node3": 2 * len(by_id[typ]) + 4 if kind == "node3" else 0, "enter2": 1, "exit2": 2, "exit1": 2, "lin": 2,
                 "enter1": 0}[kind]
    assert pos == len(tape), (pos, len(tape))
    chunks.append((len(ops), len(tape)))
    L = ["// generated by tools/gen_tape.py (%s, trie order=%s) -- do not edit" % (tag, mode), "#pragma once",
         "#define TAPE%s_NV %d" % (T, nv), "#define TAPE%s_NOPS %d" % (T, len(ops)), "#define TAPE%s_NTAPE %d" % (T, len(tape)),
         "#define TAPE%s_NCHUNKS %d" % (T, len(chunks) - 1), "#define TAPE%s_NTYPES %d" % (T, nt),
         "#define TAPE%s_FP64_INSTR %d" % (T, nfma), "#define TAPE_CHUNK_DOUBLES %d" % CHUNK]
    L.append("static const unsigned TAPE%s_OPS_HOST[%d] = {%s};" % (T, len(words), ",".join(str(w) for w in words)))
    L.append("static const short TAPE%s_COEFIDX_HOST[%d] = {%s};" % (T, len(tape), ",".join(str(c) for c in tape)))
    L.append("static const int TAPE%s_CHUNKS_HOST[%d][2] = {%s};" % (T, len(chunks), ",".join("{%d,%d}" % c for c in chunks)))
    L.append("// row bodies.  cp = 32-bit shared address of the row's [own, c1, c2, ...] slots (16-byte aligned); the caller has")
    L.append("// already loaded the first pair into (q0x = own, q0y = c1).  S0/S1: two partial Horner sums (halves the DFMA chain).")
    L.append("#define TAPE%s_ROW_SWITCH(type, cp, S0, S1, P, q0x, q0y) switch (type) { \\" % T)
    for t in range(nt):
        vs = by_id[t]
        body, have = [], {0}
        for i, v in enumerate(vs):
            slot = 1 + i
            pair, lane = slot // 2, slot % 2
            if pair not in have:
                have.add(pair)
                body.append("double q%dx, q%dy; TAPE_LDS_V2(q%dx, q%dy, cp, %d);" % (pair, pair, pair, pair, 16 * pair))
            a = "q%d%s" % (pair, "xy"[lane])
            acc = "S0" if i % 2 == 0 else "S1"
            body.append("%s = fma(%s, x[%d], %s); g[%d] = fma(%s, P, g[%d]);" % (acc, a, v, acc, v, a, v))
        L.append("case %d: { %s } break; \\" % (t, " ".join(body)))
    L.append("default: break; }")
    return "\n".join(L) + "\n", nfma, len(chunks) - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", default=OUTD)
    ap.add_argument("--order2b", default="identity")
    ap.add_argument("--order3b", default="reverse")
    args = ap.parse_args()
    d = np.load(G.DATA)
    os.makedirs(args.outdir, exist_ok=True)
    for tag, nv, mode in (("2b", 31, args.order2b), ("3b", 36, args.order3b)):
        ops, tape, types = build(tag, d["mono" + tag], nv, mode)
        txt, nfma, nch = emit_header(tag, nv, ops, tape, types, mode)
        with open(os.path.join(args.outdir, "tape_%s.h" % tag), "w") as f:
            f.write(txt)
        hist = collections.Counter(o[0] for o in ops)
        print("%s: %d ops %s, %d tape slots (%.0f KB), %d row types, %d chunks, %d FP64 instr/cluster" % (
            tag, len(ops), dict(sorted(hist.items())), len(tape), len(tape) * 8 / 1024, len(types), nch, nfma))


if __name__ == "__main__":
    main()
