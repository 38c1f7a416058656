"""CPU tests of the MB-nrg ion-water / ion-ion 2B row: the oracle restatement
(oracle/mbnrg_oracle.cpp) against the golden vectors taken from the reference's own evaluator objects and, where
oracle/_ref/libmbnrg_ref.so is present, against those objects live; our evaluation schedule of the 429-coefficient
polynomial against the raw monomial table; the C-ABI surface of include/mbnrg_b200.h."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import mbpol_b200 as mb
from mbpol_b200 import mbnrg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The fitted coefficients cancel heavily (|a_k| up to 1e3, E ~ 1e1): two summation orders of the same polynomial
# differ by a few 1e-11 of max(|E|, |g|).  Evaluator-level tolerance; totals over a box are held to 1e-10 below.
TOL_EVAL = 5e-10


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "mbnrg.npz"))


def test_model_tables():
    d = np.load(os.path.join(ROOT, "dlpoly2-mbpol-heat_b200", "data", "mbnrg_model.npz"))
    assert d["mono_hi"].shape == (1071, 6) and d["coef_hi"].shape == (9, 429) and d["par_hi"].shape == (9, 14)
    assert set(np.unique(d["w_hi"])) == {1.0} and d["mono_hi"][:, 1:].max() == 7
    assert sorted(set(d["mono_hi"][:, 0])) == list(range(429))          # every coefficient is used
    assert d["have_ii"].sum() == 2 * 29 - 9 and np.array_equal(d["have_ii"], d["have_ii"].T)   # 29 fitted pairs, 9 of them A-A
    assert np.all(d["par_hi"][:, 13] - d["par_hi"][:, 12] == 1.0) and d["par_hi"][:, 13].tolist() == [6, 6.5, 6.5, 7, 6, 6.5, 7, 7, 7]
    assert np.all(d["par_ii"][d["have_ii"]][:, 2:] == [7.0, 8.0])                              # ion-ion switch 7 -> 8 A


def test_oracle_h2o_ion_matches_reference_vectors(oracle, gold):
    for ion, w, x, E, g1, g2 in zip(gold["h2o_ion_ion"], gold["h2o_ion_w"], gold["h2o_ion_x"], gold["h2o_ion_E"], gold["h2o_ion_g1"],
                                    gold["h2o_ion_g2"]):
        e, a, b = oracle.nrg_h2o_ion(int(ion), w, x)
        sc = max(abs(E), np.abs(g1).max())
        assert abs(e - E) <= TOL_EVAL * sc and np.abs(a - g1).max() <= TOL_EVAL * sc and np.abs(b - g2).max() <= TOL_EVAL * sc
    assert np.count_nonzero(gold["h2o_ion_E"]) >= len(gold["h2o_ion_E"]) - 2      # f, li: the 6.4 A sample is beyond their 6.0 A cutoff


def test_oracle_ion_ion_is_bit_exact(oracle, gold):
    assert len(gold["ion_ion_E"]) == 29 * 4
    for a, b, w, x, E, g1, g2 in zip(gold["ion_ion_a"], gold["ion_ion_b"], gold["ion_ion_w"], gold["ion_ion_x"], gold["ion_ion_E"],
                                     gold["ion_ion_g1"], gold["ion_ion_g2"]):
        have, e, p, q = oracle.nrg_ion_ion(int(a), int(b), w, x)
        assert have and e == E and np.array_equal(p, g1) and np.array_equal(q, g2)
        have, e2, p2, q2 = oracle.nrg_ion_ion(int(b), int(a), x, w)               # the pair function is symmetric
        assert have and abs(e2 - E) <= 1e-13 * max(abs(E), 1e-30) and np.abs(p2 - g2).max() <= 1e-12 * max(np.abs(g2).max(), 1e-30)


def test_oracle_box_matches_reference_driver(oracle, gold, synth):
    box = synth.make_ion_box(216, 8)
    r = oracle.nrg_forces(box)
    assert [r["npairs"], r["nactive"], r["nionion"]] == gold["box_counts"].tolist()
    assert abs(r["eng2"] - gold["box_scalars"][0]) <= 1e-10 * abs(gold["box_scalars"][0])
    assert abs(r["vir2"] - gold["box_scalars"][1]) <= 1e-10 * abs(gold["box_scalars"][1])
    assert np.abs(r["stress"] - gold["box_stress"]).max() <= 1e-10 * np.abs(gold["box_stress"]).max()
    for k in ("fxx", "fyy", "fzz"):
        assert np.sqrt(np.mean((r[k] - gold["box_" + k]) ** 2) / np.mean(gold["box_" + k] ** 2)) <= 1e-9
    # translation invariance: net force zero, stress symmetric, virial = -trace(stress)
    assert abs(sum(r[k].sum() for k in ("fxx", "fyy", "fzz"))) <= 1e-8 * np.abs(r["fxx"]).max()
    assert abs(r["vir2"] + r["stress"][[0, 4, 8]].sum()) <= 1e-10 * abs(r["vir2"])
    # shards partition the waters; ion-ion pairs live on rank 0
    parts = [oracle.nrg_forces(box, rank=k, nranks=3) for k in range(3)]
    assert sum(p["npairs"] for p in parts) == r["npairs"] and [p["nionion"] for p in parts] == [r["nionion"], 0, 0]
    assert abs(sum(p["eng2"] for p in parts) - r["eng2"]) <= 1e-12 * abs(r["eng2"])


def test_oracle_against_live_reference_objects(oracle):
    if not oracle.have_nrg_reference:
        pytest.skip("oracle/_ref/libmbnrg_ref.so not built here")
    rng = np.random.default_rng(9)
    w0 = np.array([0.00, 0.00, 0.00, 0.9572, 0.00, 0.00, -0.2400, 0.9266, 0.00])
    for ion in range(1, 10):
        for _ in range(20):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            w = w0 + rng.normal(scale=0.04, size=9)
            x = d * rng.uniform(2.0, 6.8)
            e, a, b = oracle.nrg_h2o_ion(ion, w, x)
            E, g1, g2 = oracle.nrg_h2o_ion(ion, w, x, ref=True)
            sc = max(abs(E), np.abs(g1).max(), 1e-300)
            assert abs(e - E) <= TOL_EVAL * sc and np.abs(a - g1).max() <= TOL_EVAL * sc and np.abs(b - g2).max() <= TOL_EVAL * sc
            assert (E == 0.0) == (np.linalg.norm(w[:3] - x) > [6, 6.5, 6.5, 7, 6, 6.5, 7, 7, 7][ion - 1])


HARNESS = r"""
#define PFMA(a,b,c) ((a)*(b)+(c))
#define PMUL(a,b) ((a)*(b))
#define PADD(a,b) ((a)+(b))
#define PCOEF(k) a[k]
extern "C" double schedhi(const double* a, const double* x, double* g) { double E;
#include "polyhi_body.inc"
return E; }
"""


def test_generated_schedule_matches_table(tmp_path, oracle):
    """tools/gen_mbnrg.py: OUR shared-prefix Horner schedule of the ion-water polynomial, compiled as plain C++, against
    the monomial-table evaluation of the oracle (fitted and random coefficients)."""
    gen = os.path.join(ROOT, "dlpoly2-mbpol-heat_b200", "csrc", "gen")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_mbnrg.py")], cwd=os.path.join(ROOT, "tools"))
    hdr = open(os.path.join(gen, "mbnrg_model_data.h")).read()
    n = int(re.search(r"MBNRG_POLYHI_FP64_INSTR (\d+)", hdr).group(1))
    assert n < 3000                      # 1071 monomials, energy + 8 gradient components
    src = tmp_path / "h.cpp"
    src.write_text(HARNESS)
    lib = tmp_path / "libschedhi.so"
    subprocess.check_call(["g++", "-O0", "-ffp-contract=off", "-shared", "-fPIC", "-I", gen, str(src), "-o", str(lib)])
    f = C.CDLL(str(lib)).schedhi
    f.restype = C.c_double
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(4)
    for trial in range(4):
        a = oracle.nrg_coef_hi(1 + trial) if trial < 2 else rng.normal(size=429)
        x = rng.uniform(0.05, 1.2, size=8)
        g = np.zeros(8)
        E = f(a.ctypes.data_as(dp), x.ctypes.data_as(dp), g.ctypes.data_as(dp))
        Er, gr = oracle.nrg_polyhi(a, x)
        sc = max(abs(Er), np.abs(gr).max())
        assert abs(E - Er) <= 1e-11 * sc * (100 if trial < 2 else 1) and np.abs(g - gr).max() <= 1e-11 * sc * (100 if trial < 2 else 1)


def test_abi_symbols_and_loud_failure():
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mbnrg_b200.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(mbnrg_b200_\w+)\s*\(", txt)))
    assert len(names) == 12 and "mbnrg_b200_forces" in names and "mbnrg_b200_forces_3b" in names
    lib = mb.load_library()
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/mbnrg_b200.h but not exported" % n
    ptr = C.c_void_p()
    ic = (C.c_int * 2)(6, 2)
    assert lib.mbnrg_b200_init(7, 2, ic, -1, C.byref(ptr)) == 1 and not ptr           # nttm2 not a multiple of 3
    bad = (C.c_int * 2)(6, 10)
    assert lib.mbnrg_b200_init(9, 2, bad, -1, C.byref(ptr)) == 1 and not ptr          # unknown ion code
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(mb.MBPolError):
            mbnrg.mbnrg_init(9, [6, 2])


def test_ion_box_generator(synth):
    box = synth.make_ion_box(216, 8)
    assert box["nions"] == 8 and box["nw"] == 208 and box["mxatms"] == 8 + 4 * 208
    assert box["listions"].tolist() == list(range(1, 9)) and box["listttm2"][0] == 9          # ions are the first atoms
    assert box["ion_code"].tolist() == [6, 2] * 4
    b2 = synth.make_ion_box(216, 8)
    assert np.array_equal(box["xxx"], b2["xxx"])


# ---- 3B ion-water-water (mbnrg-3b-h2o-h2o-<ion>-poly.cpp; compiled out in the reference build, mbnrg.f90:4) -------------------
@pytest.fixture(scope="module")
def gold3():
    return np.load(os.path.join(ROOT, "tests", "golden", "mbnrg3.npz"))


def test_oracle_3b_matches_reference_vectors(oracle, gold3):
    """The restated x3b_h2o_ion_v2x_p::operator() (21 variables, table-driven 1831-coefficient polynomial) against the golden vectors the
    reference's own objects produced for all nine ion species.  The fitted coefficients cancel heavily (|a| up to 1e3 for |E| ~ 1),
    so two summation orders agree to ~1e-11 of max(|E|, |g|) per cluster, not better."""
    for q in range(len(gold3["tri_ion"])):
        w = gold3["tri_w"][q]
        E, g1, g2, g3 = oracle.nrg_h2o_h2o_ion(int(gold3["tri_ion"][q]), w[:9], w[9:18], w[18:])
        sc = max(abs(gold3["tri_E"][q]), np.abs(gold3["tri_g"][q]).max())
        assert abs(E - gold3["tri_E"][q]) <= 2e-10 * sc, q
        assert np.abs(np.r_[g1, g2, g3] - gold3["tri_g"][q]).max() <= 2e-10 * sc, q
        assert np.abs(np.r_[g1, g2, g3].reshape(7, 3).sum(0)).max() <= 1e-9 * sc          # no net force on the cluster


def test_oracle_3b_box_matches_reference_driver(oracle, gold3, synth):
    """The restated 3B driver (both oxygen-ion distances < 7 A, water i and ion placed next to water j) on the 208-water + 8-ion box."""
    box = synth.make_ion_box(216, 8)
    r = oracle.nrg_forces3(box, use_ref=oracle.have_nrg_reference_3b)
    assert r["ntrimers"] == int(gold3["box_counts"][0]) and r["ntrimers"] > 1000
    assert abs(r["eng3"] - gold3["box_scalars"][0]) <= 1e-10 * abs(gold3["box_scalars"][0])
    assert abs(r["vir3"] - gold3["box_scalars"][1]) <= 1e-9 * abs(gold3["box_scalars"][1])
    for k in ("fxx", "fyy", "fzz"):
        assert np.sqrt(np.mean((r[k] - gold3["box_" + k]) ** 2) / np.mean(gold3["box_" + k] ** 2)) <= 1e-9
    assert np.abs(r["stress"] - gold3["box_stress"]).max() <= 1e-9 * np.abs(gold3["box_stress"]).max()
    # two ranks partition the trimer list and add up
    a, b = oracle.nrg_forces3(box, use_ref=oracle.have_nrg_reference_3b, rank=0, nranks=2), oracle.nrg_forces3(box, use_ref=oracle.have_nrg_reference_3b, rank=1, nranks=2)
    assert a["nshard"] + b["nshard"] == r["ntrimers"]
    assert abs(a["eng3"] + b["eng3"] - r["eng3"]) <= 1e-11 * abs(r["eng3"])


def test_generated_3b_schedule_matches_table(oracle):
    """tools/gen_mbnrg.py: the 24-part shared-prefix Horner schedule of the 3B polynomial (what the GPU runs), executed here in Python on
    the per-thread record, equals the raw monomial table."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_mbnrg
    import gen_poly
    d = np.load(gen_mbnrg.DATA)
    sch = gen_poly.make_schedule(d["mono_3b"], 21, "greedy")
    rng = np.random.default_rng(5)
    x = rng.uniform(0.2, 1.1, size=21)
    a = oracle.nrg_coef_3b(6)
    reg = {}
    val = lambda o: a[o[1]] if o[0] == "c" else (x[o[1]] if o[0] == "x" else reg[o])
    for op in sch.ops:
        if op[0] == "fma":
            reg[op[1]] = val(op[2]) * val(op[3]) + val(op[4])
        elif op[0] == "mul":
            reg[op[1]] = val(op[2]) * val(op[3])
        elif op[0] == "add":
            reg[op[1]] = val(op[2]) + val(op[3])
        elif op[0] == "mov":
            reg[op[1]] = val(op[2])
    E, g = oracle.nrg_poly3b(a, x)
    gs = np.array([reg[("g", i)] for i in range(21)])
    sc = max(abs(E), np.abs(g).max())
    assert abs(reg[("E",)] - E) <= 1e-9 * sc and np.abs(gs - g).max() <= 1e-9 * sc
    # every part of the split loads only what an earlier part stored
    parts = gen_poly.split_parts(sch.ops, gen_mbnrg.NPART3, 21)
    assert len(parts) == gen_mbnrg.NPART3 and not parts[0][1]
