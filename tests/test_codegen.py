"""CPU test of OUR polynomial evaluation schedule (tools/gen_poly.py): the generated straight-line body, compiled
as plain C++, must reproduce the table-driven oracle polynomial and its gradient."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#define PFMA(a,b,c) ((a)*(b)+(c))
#define PMUL(a,b) ((a)*(b))
#define PADD(a,b) ((a)+(b))
#define PCOEF(k) a[k]
extern "C" double sched2b(const double* a, const double* x, double* g) { double E;
#include "poly2b_body.inc"
return E; }
extern "C" double sched3b(const double* a, const double* x, double* g) { double E;
#include "poly3b_body.inc"
return E; }
"""


def test_generated_schedule_matches_tables(tmp_path, oracle):
    gen = tmp_path / "gen"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_poly.py"), "--outdir", str(gen),
                           "--parts2b", "3", "--parts3b", "5"])
    # structural checks on the PTX parts: every part is a .func taking (buf, nbatch) and all FP64 ops are accounted for
    hdr = open(gen / "poly_gen.h").read()
    n3 = int(hdr.split("POLY3B_FP64_INSTR")[1].split()[0])
    ptx_ops = 0
    for k in range(5):
        txt = open(gen / ("poly3b_part%d.ptx" % k)).read()
        assert ".visible .func poly3b_part%d" % k in txt and ".target sm_100a" in txt
        ptx_ops += sum(txt.count(m) for m in ("fma.rn.f64", "mul.rn.f64", "add.rn.f64"))
    assert ptx_ops == n3 and n3 < 111818 // 1.5        # fewer FP64 instructions than the reference has flops
    src = tmp_path / "h.cpp"
    src.write_text(HARNESS)
    lib = tmp_path / "libsched.so"
    subprocess.check_call(["g++", "-O0", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(gen), str(src), "-o", str(lib)])
    L = C.CDLL(str(lib))
    rng = np.random.default_rng(3)
    dp = C.POINTER(C.c_double)
    for tag, nv, f in (("2b", 31, L.sched2b), ("3b", 36, L.sched3b)):
        f.restype = C.c_double
        for trial in range(3):
            a = oracle.coef(tag) if trial == 0 else rng.normal(size=len(oracle.coef(tag)))
            x = rng.uniform(0.05, 1.3, size=nv)
            g = np.zeros(nv)
            E = f(a.ctypes.data_as(dp), x.ctypes.data_as(dp), g.ctypes.data_as(dp))
            Er, gr = oracle.poly(tag, a, x)
            assert abs(E - Er) <= 1e-11 * max(abs(Er), np.abs(gr).max())
            assert np.abs(g - gr).max() <= 1e-11 * np.abs(gr).max()


def test_tape_rows_reproduce_the_polynomials(oracle):
    """tools/gen_tape.py: the flattened rows (E += S*P, S = own + sum a_c x_c, P = x_u x_w x_v) of the tape engines are
    the same polynomial and gradient as the monomial tables -- checked in numpy from the generator's own row list."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_poly
    import gen_tape
    d = np.load(gen_poly.DATA)
    rng = np.random.default_rng(11)
    for tag, nv, mode in (("2b", 31, "identity"), ("3b", 36, "reverse")):
        rows, types = gen_tape.build(tag, d["mono" + tag], nv, mode)
        by_id = {t: vs for vs, t in types.items()}
        a = oracle.coef(tag)
        x = rng.uniform(0.1, 1.2, size=nv)
        xe = np.concatenate([x, [1.0]])                 # pseudo-variable UNIT = nv
        E, g = 0.0, np.zeros(nv + 1)
        for (t, u, w, v, own, cs) in rows:
            vs = by_id[t]
            assert len(vs) == len(cs)
            P = xe[u] * xe[w] * xe[v]
            S = (a[own] if own >= 0 else 0.0) + sum(a[k] * x[c] for k, c in zip(cs, vs))
            E += S * P
            for k, c in zip(cs, vs):
                g[c] += a[k] * P
            g[u] += S * xe[w] * xe[v]
            g[w] += S * xe[u] * xe[v]
            g[v] += S * xe[u] * xe[w]
        Er, gr = oracle.poly(tag, a, x)
        assert abs(E - Er) <= 1e-11 * max(abs(Er), np.abs(gr).max())
        assert np.abs(g[:nv] - gr).max() <= 1e-11 * np.abs(gr).max()
        assert len(types) <= 64 and max(len(v) for v in by_id.values()) <= nv


def test_fat_rows_reproduce_the_3b_polynomial(oracle):
    """tools/gen_fat.py: the fat rows of engine 3 (all rows sharing the (w,v) prefix merged; children (u, leaf set) static)
    are the same polynomial and gradient as the monomial table -- checked in numpy from the generator's own structures,
    following exactly the arithmetic of the generated code (T = sum_k S_k x_u, E += T p2, g_u += S_k p2, ...)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_fat
    import gen_poly
    d = np.load(gen_poly.DATA)
    nv = 36
    fat, shapes = gen_fat.build("3b", d["mono3b"], nv, "reverse")
    assert len(fat) == 361 and len(shapes) <= 64 and len(set((w, v) for (_, w, v, _) in fat)) == len(fat)
    a = oracle.coef("3b")
    rng = np.random.default_rng(12)
    x = rng.uniform(0.1, 1.2, size=nv)
    xe = np.concatenate([x, [1.0]])
    E, g = 0.0, np.zeros(nv + 1)
    nmono = 0
    for (s, w, v, ch) in fat:
        assert tuple((c[0], c[1]) for c in ch) == shapes[s]
        p2 = xe[w] * xe[v]
        Tw = 0.0
        for (u, ls, own, cs) in ch:
            P = p2 * xe[u]
            S = (a[own] if own >= 0 else 0.0) + sum(a[k] * x[c] for k, c in zip(cs, ls))
            for k, c in zip(cs, ls):
                g[c] += a[k] * P
            g[u] += S * p2
            Tw += S * xe[u]
            nmono += len(cs) + (own >= 0)
        E += Tw * p2
        g[w] += Tw * xe[v]
        g[v] += Tw * xe[w]
    assert nmono == len(d["mono3b"])
    Er, gr = oracle.poly("3b", a, x)
    assert abs(E - Er) <= 1e-11 * max(abs(Er), np.abs(gr).max())
    assert np.abs(g[:nv] - gr).max() <= 1e-11 * np.abs(gr).max()


@pytest.mark.parametrize("group", ["t", "z", "s3"])
def test_symmetry_batched_rows_reproduce_the_3b_polynomial(oracle, group):
    """tools/gen_sym.py: the super rows of engine 4 (fat rows batched over a subgroup K of the molecule permutations; one
    coefficient a*|orbit|/|K| feeds |K| statically renamed images) are the same polynomial and gradient as the monomial
    table -- checked in numpy from the generator's own structures, following the arithmetic of the generated code."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_poly
    import gen_sym
    d = np.load(gen_poly.DATA)
    nv = 36
    rows, shapes, vps = gen_sym.build(d["mono3b"], group)
    K = len(vps)
    assert K == {"t": 2, "z": 3, "s3": 6}[group] and len(shapes) <= 255
    a = oracle.coef("3b")
    rng = np.random.default_rng(13)
    x = rng.uniform(0.1, 1.2, size=nv)
    xe = np.concatenate([x, [1.0]])
    E, g = 0.0, np.zeros(nv + 1)
    nmono = 0
    for (s, key, children) in rows:
        assert tuple((us, tuple(cs for cs, _, _ in leaves)) for us, own, leaves in children) == shapes[s]
        for j in range(K):
            w, v = key[j]
            p2 = xe[w] * xe[v]
            Tw = 0.0
            for (us, own, leaves) in children:
                u = us[j]
                P = p2 * xe[u]
                S = a[own[0]] * own[1] / K if own else 0.0
                for (cs, k, num) in leaves:
                    c = cs[j]
                    S += a[k] * num / K * xe[c]
                    g[c] += a[k] * num / K * P
                g[u] += S * p2
                Tw += S * xe[u]
            E += Tw * p2
            g[w] += Tw * xe[v]
            g[v] += Tw * xe[w]
        for (us, own, leaves) in children:
            nmono += sum(num for _, _, num in leaves) + (own[1] if own else 0)
    assert nmono == len(d["mono3b"])          # every monomial exactly once (orbit sizes add up)
    Er, gr = oracle.poly("3b", a, x)
    assert abs(E - Er) <= 1e-11 * max(abs(Er), np.abs(gr).max())
    assert np.abs(g[:nv] - gr).max() <= 1e-11 * np.abs(gr).max()


SYM_HOST_HARNESS = r"""
// host emulation of csrc/mbpol_sym.cuh's row loop around the GENERATED header: shared memory is a byte array, one "thread"
#include <cmath>
#include <cstring>
#include <vector>
static unsigned char SM[1 << 21];
static inline double tape_lds(unsigned a) { double v; std::memcpy(&v, SM + a, 8); return v; }
static inline void tape_sts(unsigned a, double v) { std::memcpy(SM + a, &v, 8); }
#define TAPE_LDS_V2(vx, vy, addr, imm) { vx = tape_lds((addr) + (imm)); vy = tape_lds((addr) + (imm) + 8); }
#define SYM_FMA(acc, a, b) acc = std::fma(a, b, acc)
#define TAPE_COL 8
using std::fma;
#include "sym_3b.h"
extern "C" double sym_host_eval(const double* a, const double* xin, double* gout, int split)
{
    const unsigned* rowvars = SYM3B_ROWVARS_HOST;
    const unsigned xs_t = 64, gs_delta = 37 * 8, buf0 = 4096;
    double x[36], g[36];
    for (int c = 0; c < 36; ++c) { x[c] = xin[c]; g[c] = 0.0; tape_sts(xs_t + c * TAPE_COL, x[c]); tape_sts(xs_t + gs_delta + c * TAPE_COL, 0.0); }
    tape_sts(xs_t + 36 * TAPE_COL, 1.0); tape_sts(xs_t + gs_delta + 36 * TAPE_COL, 0.0);
    SYM_PAIR_INIT
    for (int i = 0; i < SYM3B_NTAPE; ++i)
        tape_sts(buf0 + 8 * i, SYM3B_COEFIDX_HOST[i] < 0 ? 0.0 : a[SYM3B_COEFIDX_HOST[i]] * (double)SYM3B_COEFNUM_HOST[i] / (double)SYM3B_K);
    double e = 0.0;
    for (int k = 0; k < SYM3B_NCHUNKS; ++k) {      // the row loop of csrc/mbpol_sym.cuh; `split`: the cyclic order of warps 4..7
        const unsigned cbase = buf0 + 8 * SYM3B_CHUNKS_HOST[k][2];
        const unsigned row0 = SYM3B_CHUNKS_HOST[k][1], row1 = SYM3B_CHUNKS_HOST[k + 1][1];
        unsigned row = row0 + (split ? (row1 - row0) / 2 : 0);
        unsigned row_next = row + 1 == row1 ? row0 : row + 1;
        unsigned cp;
        SYM_ROW_DECL
        SYM_ROW_WORDS(row)
        SYM_ROW_FETCH(row, row_next)
        for (unsigned it = row0; it < row1; ++it) {
            const unsigned row_next2 = row_next + 1 == row1 ? row0 : row_next + 1;
            SYM_ROW_PRE
            SYM3B_ROW_SWITCH(shape)
            SYM_ROW_POST
            row = row_next;
            row_next = row_next2;
        }
    }
    SYM_PAIR_FINI
    for (int c = 0; c < 36; ++c) gout[c] = g[c] + tape_lds(xs_t + gs_delta + c * TAPE_COL);
    return e;
}
"""


@pytest.mark.parametrize("env", [{}, {"SYM_POST": "serial", "SYM_WORDS_AHEAD": "0"}, {"SYM_POST": "batch", "SYM_PAIRS": "0"}, {"SYM_GROUP": "s3"}])
def test_generated_sym_header_on_the_host(tmp_path, oracle, env):
    """The GENERATED header of engine 4 (macros included: row header prefetch, batched / serial row epilogue, alias flag, shape
    dispatch) compiled for the host around an emulation of the shared-memory columns reproduces the oracle polynomial."""
    hdr = tmp_path / "sym_3b.h"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_sym.py"), "--out", str(hdr)], env=dict(os.environ, **env),
                          stdout=subprocess.DEVNULL)
    (tmp_path / "h.cpp").write_text(SYM_HOST_HARNESS)
    so = tmp_path / "libsymhost.so"
    subprocess.check_call(["g++", "-O0", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(tmp_path), "-o", str(so), str(tmp_path / "h.cpp")])
    lib = C.CDLL(str(so))
    lib.sym_host_eval.restype = C.c_double
    lib.sym_host_eval.argtypes = [C.c_void_p] * 3 + [C.c_int]
    a = np.ascontiguousarray(oracle.coef("3b"))
    rng = np.random.default_rng(17)
    for _ in range(3):
        x = rng.uniform(0.1, 1.2, size=36)
        g = np.zeros(36)
        Er, gr = oracle.poly("3b", a, x)
        for split in (0, 1):
            E = lib.sym_host_eval(a.ctypes.data, x.ctypes.data, g.ctypes.data, split)
            assert abs(E - Er) <= 1e-11 * max(abs(Er), np.abs(gr).max())
            assert np.abs(g - gr).max() <= 1e-11 * np.abs(gr).max()


def test_default_sym_schedule_is_the_documented_one():
    """tools/gen_sym.py: the default trie order gives the schedule DESIGN.md / experiments/README.md quote, and the pair-sum merge
    neither loses nor duplicates a monomial (a merged leaf stands for two monomials of equal coefficient)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_poly
    import gen_sym
    mono = np.load(gen_poly.DATA)["mono3b"]
    order = [int(v) for v in gen_sym.DEFAULT_ORDER.split(",")]
    rows, shapes, vps = gen_sym.build(mono, "z", True, gen_sym.DEFAULT_REPSEL, order=order)
    K = len(vps)
    rows, shapes = gen_sym.merge_pairs(rows)
    nfp, code, nrows, nshapes = gen_sym.stats(rows, shapes, K)
    assert (nfp, code, nrows, nshapes) == (71589, 11079, 120, 47)
    nmono = 0
    for s, key, children in rows:
        for us, own, leaves in children:
            nmono += own[1] if own else 0
            for cs, k, num in leaves:
                assert all((c > gen_sym.NV) == (cs[0] > gen_sym.NV) for c in cs)      # an image of a pair sum is a pair sum
                nmono += num * (2 if cs[0] > gen_sym.NV else 1)
    assert nmono == len(mono)
    # every row word field fits its bits (w, v: 6 bits; shape: 7 bits; tape offset / 16: 10 bits)
    assert nshapes <= 128 and gen_sym.NV + 1 + len(gen_sym.PAIRS) <= 64
