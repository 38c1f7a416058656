#!/usr/bin/env python3
"""Generate the straight-line polynomial evaluators (energy + full gradient) from the model tables.

Input : dlpoly2-mbpol-heat_b200/data/mbpol_model.npz   (monomial tables, see tools/extract_model.py)
Output: dlpoly2-mbpol-heat_b200/csrc/gen/
            poly{2b,3b}_part<k>.ptx   device functions for sm_100a (assembled by ptxas, linked with -rdc)
            poly{2b,3b}_body.inc      the same schedule as C macros (CPU schedule test only)
            poly_gen.h                prototypes + instruction statistics

Evaluation schedule (ours; unrelated to the reference's Maple temporaries in
poly-2b-v6x.cpp:5-13843 / poly-3b-v2x.cpp:5-28380 -- only the polynomial is the same):

Every monomial is a path in a prefix tree ("trie") over variable indices.  With
    p(node)  = product of the variables on the path root -> node      (top-down)
    s(node)  = a(node) + sum_{child c via x_c} x_c * s(c)             (bottom-up, multivariate Horner)
we have  E = s(root)  and, summing over all trie edges (parent P --x_c--> child C),
    dE/dx_c += p(P) * s(C).
For a leaf, s(C) is the literal coefficient, so a leaf costs exactly two FMAs whose
coefficient operand is a warp-uniform constant-bank value:
    s(P)  = fma(a, x_c, s(P));     g[c] = fma(a, p(P), g[c]);
An internal node costs one multiply (p) and the same two FMAs with s(C) in a register.
A depth-first walk keeps only (p, s) of the <= 3 open levels live, so the working set is
x[nv] + g[nv] + ~8 doubles: it fits the 255-register budget of one thread, and there is
no adjoint tape (the reference needs df[4138] / ~22k temporaries).

Which variable of a top-degree monomial becomes the leaf is chosen by a greedy set cover
that minimises the number of distinct internal (degree-3) nodes.

FP64-pipe instructions per cluster: 2*leaves + 3*internal = 26 415 (2B), 69 936 (3B), versus
30 601 / 111 818 counted flops in the reference bodies.

Why PTX and not CUDA C++: nvcc's front end (cicc) is super-linear in the size of this
function (> 1 h for the 3B body); ptxas alone assembles it in minutes.  The body is cut
into NPART device functions at top-level subtree boundaries so the parts assemble in
parallel; they communicate through a per-thread scratch record in local memory
    buf[0..nv-1] = x,  buf[nv] = E,  buf[nv+1 .. 2nv] = g,  then s1..s3, p2..p3   (doubles)
Each part loads only what it reads before writing and stores only what it wrote and a later
part (or the caller) still needs (plain liveness analysis over the op list), so a cut is
legal anywhere.  Every part loops over `nbatch` consecutive records: one fetch of the part's
code is reused nbatch times -- on B200 straight-line code bigger than the instruction cache
is limited by instruction fetch (measured: 15 % FP64 pipe, 61 % "no instruction" stalls at
nbatch = 1), not by the FP64 pipe.
"""
import argparse
import collections
import heapq
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CWINDOW = 5248        # doubles of the shared constant-bank window (see tools/gen_tape.py)
DATA = os.path.join(ROOT, "dlpoly2-mbpol-heat_b200", "data", "mbpol_model.npz")
OUTD = os.path.join(ROOT, "dlpoly2-mbpol-heat_b200", "csrc", "gen")


# ----------------------------------------------------------------------------------------------
# trie construction
# ----------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("var", "coef", "children")

    def __init__(self, var):
        self.var = var
        self.coef = -1
        self.children = {}


def choose_orders(monos, nv, mode):
    """Return the variable sequence (root->leaf) to use for every monomial."""
    if mode == "identity":
        return [tuple(sorted(m)) for m in monos]
    if mode == "reverse":
        return [tuple(sorted(m, reverse=True)) for m in monos]
    if mode == "greedy":
        return greedy_orders(monos)
    raise ValueError(mode)


def greedy_orders(monos):
    """Pick, per top-degree monomial, which variable is the leaf so that few distinct parents are needed.

    Greedy set cover: repeatedly open the parent multiset covering the most still-uncovered
    monomials; a parent that is itself a monomial of the polynomial gets a small bonus.
    """
    canon = lambda t: tuple(sorted(t, reverse=True))
    maxdeg = max(len(m) for m in monos)
    top = [canon(m) for m in monos if len(m) == maxdeg]
    rest = [canon(m) for m in monos if len(m) < maxdeg]
    cover = collections.defaultdict(set)
    for i, m in enumerate(top):
        for j in range(len(m)):
            cover[m[:j] + m[j + 1:]].add(i)
    bonus = set(m for m in rest if len(m) == maxdeg - 1)
    uncovered = set(range(len(top)))
    assign = {}
    heap = [(-(len(s) + (0.5 if p in bonus else 0.0)), p) for p, s in cover.items()]
    heapq.heapify(heap)
    while uncovered:
        negn, p = heapq.heappop(heap)
        live = cover[p] & uncovered
        if not live:
            continue
        score = len(live) + (0.5 if p in bonus else 0.0)
        if score < -negn:
            heapq.heappush(heap, (-score, p))
            continue
        for i in live:
            assign[i] = p
        uncovered -= live
    out = []
    for i, m in enumerate(top):
        rem = list(m)
        for v in assign[i]:
            rem.remove(v)
        out.append(assign[i] + (rem[0],))
    out.extend(rest)
    return out


def build_trie(seqs, coefs):
    root = Node(-1)
    for seq, k in zip(seqs, coefs):
        n = root
        for v in seq:
            if v not in n.children:
                n.children[v] = Node(v)
            n = n.children[v]
        assert n.coef < 0, "duplicate monomial"
        n.coef = k
    return root


# ----------------------------------------------------------------------------------------------
# schedule -> IR.  Operands: ("x",i) ("g",i) ("s",d) ("p",d) ("E",) ("c",k)
# Ops: ("fma",dst,a,b,c) ("mul",dst,a,b) ("add",dst,a,b) ("mov",dst,a) ("cut",)  [cut = part boundary allowed here]
# ----------------------------------------------------------------------------------------------
class Schedule:
    def __init__(self, nv):
        self.nv = nv
        self.ops = []
        self.g_init = [False] * nv
        self.E_init = False

    def count(self):
        c = collections.Counter(op[0] for op in self.ops)
        return c["fma"], c["mul"], c["add"]

    def acc(self, dst, a, b, started):
        if started:
            self.ops.append(("fma", dst, a, b, dst))
        else:
            self.ops.append(("mul", dst, a, b))

    def acc_g(self, v, a, b):
        self.acc(("g", v), a, b, self.g_init[v])
        self.g_init[v] = True

    def node(self, node, depth, p):
        """Emit the subtree below internal `node`; leaves s(node) in ("s",depth)."""
        s = ("s", depth)
        started = False
        own = ("c", node.coef) if node.coef >= 0 else None
        kids = sorted(node.children.values(), key=lambda c: (len(c.children) > 0, c.var))
        for c in kids:
            xv = ("x", c.var)
            if c.children:
                cp = ("p", depth + 1)
                self.ops.append(("mul", cp, p, xv))
                self.node(c, depth + 1, cp)
                sc = ("s", depth + 1)
            else:
                sc = ("c", c.coef)
            if not started and own is not None:
                self.ops.append(("fma", s, sc, xv, own))
            else:
                self.acc(s, sc, xv, started)
            started = True
            self.acc_g(c.var, sc, p)
            if depth == 1:
                self.ops.append(("cut",))

    def top(self, root):
        for v in sorted(root.children):
            c = root.children[v]
            xv = ("x", v)
            if not c.children:
                self.acc(("E",), ("c", c.coef), xv, self.E_init)
                self.E_init = True
                src = ("c", c.coef)
            else:
                self.node(c, 1, xv)
                self.acc(("E",), ("s", 1), xv, self.E_init)
                self.E_init = True
                src = ("s", 1)
            if self.g_init[v]:
                self.ops.append(("add", ("g", v), ("g", v), src))
            else:
                self.ops.append(("mov", ("g", v), src))
                self.g_init[v] = True
            self.ops.append(("cut",))
        assert all(self.g_init) and self.E_init


def make_schedule(mono, nv, mode):
    monos = [tuple(int(v) for v in r[1:] if v >= 0) for r in mono]
    coefs = [int(r[0]) for r in mono]
    seqs = choose_orders(monos, nv, mode)
    key = lambda m: tuple(sorted(m))
    lut = {key(m): k for m, k in zip(monos, coefs)}
    assert len(lut) == len(monos)
    root = build_trie(seqs, [lut[key(s)] for s in seqs])
    sch = Schedule(nv)
    sch.top(root)
    return sch


# ----------------------------------------------------------------------------------------------
# backends
# ----------------------------------------------------------------------------------------------
def c_operand(o):
    if o[0] in ("x", "g"):
        return "%s[%d]" % o
    if o[0] in ("s", "p"):
        return "%s%d" % o
    if o[0] == "E":
        return "E"
    return "PCOEF(%d)" % o[1]


def emit_c(sch, tag, mode):
    out = ["// generated by tools/gen_poly.py (%s, order=%s) -- CPU schedule test only" % (tag, mode),
           "double s1, s2, s3, p2, p3;"]
    for op in sch.ops:
        if op[0] == "fma":
            out.append("%s = PFMA(%s, %s, %s);" % tuple(c_operand(o) for o in op[1:]))
        elif op[0] == "mul":
            out.append("%s = PMUL(%s, %s);" % tuple(c_operand(o) for o in op[1:]))
        elif op[0] == "add":
            out.append("%s = PADD(%s, %s);" % tuple(c_operand(o) for o in op[1:]))
        elif op[0] == "mov":
            out.append("%s = %s;" % tuple(c_operand(o) for o in op[1:]))
    return "\n".join(out) + "\n"


def reg_slot(o, nv):
    """Record slot (in doubles) of a register operand."""
    if o[0] == "x":
        return o[1]
    if o[0] == "E":
        return nv
    if o[0] == "g":
        return nv + 1 + o[1]
    if o[0] == "s":
        return 2 * nv + o[1]          # s1..s3 -> 2nv+1..2nv+3
    if o[0] == "p":
        return 2 * nv + 2 + o[1]      # p2..p3 -> 2nv+4..2nv+5
    raise ValueError(o)


def record_doubles(nv):
    return 2 * nv + 6


def split_parts(ops, nparts, nv):
    """Cut the op list into nparts pieces of equal FP64 instruction count (any boundary is legal:
    whatever is live across a cut travels through the per-thread record).

    Returns [(ops, loads, stores)] with loads = upward-exposed non-x registers, stores = written & live-out.
    """
    ops = [o for o in ops if o[0] != "cut"]
    n = len(ops)
    bounds = [round(k * n / nparts) for k in range(nparts + 1)]
    chunks = [ops[bounds[k]:bounds[k + 1]] for k in range(nparts)]
    # backward liveness at chunk granularity
    live = set([("E",)] + [("g", i) for i in range(nv)])
    info = [None] * nparts
    for k in range(nparts - 1, -1, -1):
        written, exposed = set(), set()
        for o in chunks[k]:
            for a in o[2:]:
                if a[0] in ("g", "s", "p", "E") and a not in written:
                    exposed.add(a)
            written.add(o[1])
        stores = sorted(written & live)
        info[k] = (chunks[k], sorted(exposed), stores)
        live = (live - written) | exposed
    assert not live, "registers read before ever written: %s" % sorted(live)
    return info


def emit_ptx_part(ops, loads, stores, tag, part, nv, ncoef, barrier_period):
    """One device function  poly<tag>_part<k>(u64 buf, u32 nbatch).

    buf -> nbatch consecutive per-thread records (generic address of local memory).  The body is
    executed once per record inside a loop, so the code of one part is fetched once and reused
    nbatch times (the instruction stream, not the FP64 pipe, is what limits straight-line code
    that is larger than the instruction cache).
    """
    name = "poly%s_part%d" % (tag, part)
    # both coefficient tables live in the library's single constant-bank window: 2B at [0, 1153), 3B at [1153, 2316)
    cname = "mbpol_cwin"
    cbase = 0 if tag == "2b" else 1153
    stride = 8 * record_doubles(nv)
    L = []
    L.append("// generated by tools/gen_poly.py -- do not edit")
    L.append(".version 8.8")
    L.append(".target sm_100a")
    L.append(".address_size 64")
    L.append(".extern .const .align 16 .f64 %s[%d];" % (cname, CWINDOW))
    L.append(".visible .func %s(.param .b64 %s_buf, .param .b32 %s_nb)" % (name, name, name))
    L.append("{")
    ncref = sum(1 for o in ops for a in o[1:] if a[0] == "c")
    L.append(".reg .b64 %rd<3>;")
    L.append(".reg .b32 %nb, %ib;")
    L.append(".reg .pred %more;")
    L.append(".reg .f64 %%x<%d>, %%g<%d>, %%s<4>, %%p<4>, %%E, %%k<%d>;" % (nv, nv, ncref + 1))
    L.append("ld.param.u64 %%rd2, [%s_buf];" % name)
    L.append("ld.param.u32 %%nb, [%s_nb];" % name)
    L.append("cvta.to.local.u64 %rd1, %rd2;")
    L.append("mov.u32 %ib, 0;")
    L.append("BATCH_%s:" % name)

    def rname(o):
        return "%E" if o[0] == "E" else "%%%s%d" % o

    used_x = sorted(set(a for o in ops for a in o[1:] if a[0] == "x"))
    for o in used_x + list(loads):
        L.append("ld.local.f64 %s, [%%rd1+%d];" % (rname(o), 8 * reg_slot(o, nv)))
    kc = [0]
    since_bar = [0]

    def reg(o, pre):
        if o[0] != "c":
            return rname(o)
        kc[0] += 1
        pre.append("ld.const.f64 %%k%d, [%s+%d];" % (kc[0], cname, 8 * (cbase + o[1])))
        return "%%k%d" % kc[0]

    mn = {"fma": "fma.rn.f64", "mul": "mul.rn.f64", "add": "add.rn.f64", "mov": "mov.f64"}
    for o in ops:
        pre = []
        srcs = [reg(a, pre) for a in o[2:]]
        L.extend(pre)
        L.append("%s %s, %s;" % (mn[o[0]], rname(o[1]), ", ".join(srcs)))
        since_bar[0] += 1
        if barrier_period and since_bar[0] >= barrier_period:
            L.append("bar.sync 0;")
            since_bar[0] = 0
    for o in stores:
        L.append("st.local.f64 [%%rd1+%d], %s;" % (8 * reg_slot(o, nv), rname(o)))
    L.append("add.u64 %%rd1, %%rd1, %d;" % stride)
    L.append("add.u32 %ib, %ib, 1;")
    L.append("setp.lt.u32 %more, %ib, %nb;")
    L.append("@%%more bra BATCH_%s;" % name)
    L.append("ret;")
    L.append("}")
    return "\n".join(L) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", default="greedy", choices=["identity", "reverse", "greedy"])
    ap.add_argument("--outdir", default=OUTD)
    ap.add_argument("--parts2b", type=int, default=2)
    ap.add_argument("--parts3b", type=int, default=8)
    ap.add_argument("--barrier-period", type=int, default=0,
                    help="insert bar.sync 0 every N FP64 instructions (keeps the warps of a CTA in one I-cache window)")
    args = ap.parse_args()
    d = np.load(DATA)
    os.makedirs(args.outdir, exist_ok=True)
    hdr = ["// generated by tools/gen_poly.py -- do not edit", "#pragma once"]
    for tag, nv, nparts in (("2b", 31, args.parts2b), ("3b", 36, args.parts3b)):
        sch = make_schedule(d["mono" + tag], nv, args.order)
        nf, nm, na = sch.count()
        with open(os.path.join(args.outdir, "poly%s_body.inc" % tag), "w") as f:
            f.write(emit_c(sch, tag, args.order))
        parts = split_parts(sch.ops, nparts, nv)
        for k, (ops, loads, stores) in enumerate(parts):
            txt = emit_ptx_part(ops, loads, stores, tag, k, nv, len(d["coef" + tag]), args.barrier_period)
            with open(os.path.join(args.outdir, "poly%s_part%d.ptx" % (tag, k)), "w") as f:
                f.write(txt)
        sizes = [len(p[0]) for p in parts]
        traffic = sum(len(p[1]) + len(p[2]) + len(set(a for o in p[0] for a in o[1:] if a[0] == "x")) for p in parts)
        print("%s: fma=%d mul=%d add=%d total=%d ; %d parts of ~%d ops; record ld/st per cluster %d" % (
            tag, nf, nm, na, nf + nm + na, len(parts), sizes[0], traffic))
        T = tag.upper()
        hdr.append("#define POLY%s_FP64_INSTR %d" % (T, nf + nm + na))
        hdr.append("#define POLY%s_NVARS %d" % (T, nv))
        hdr.append("#define POLY%s_NPARTS %d" % (T, len(parts)))
        hdr.append("#define POLY%s_RECORD_DOUBLES %d" % (T, record_doubles(nv)))
        for k in range(len(parts)):
            hdr.append('extern "C" __device__ void poly%s_part%d(double* buf, unsigned nbatch);' % (tag, k))
        hdr.append("#define POLY%s_CALL_ALL(buf, nb) do { %s } while (0)" % (
            T, " ".join("poly%s_part%d(buf, nb);" % (tag, k) for k in range(len(parts)))))
    with open(os.path.join(args.outdir, "poly_gen.h"), "w") as f:
        f.write("\n".join(hdr) + "\n")


if __name__ == "__main__":
    main()
