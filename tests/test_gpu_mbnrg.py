"""GPU parity tests of the MB-nrg ion-water / ion-ion 2B row: the CUDA path, called through the
C ABI of include/mbnrg_b200.h, against the CPU oracle (oracle/mbnrg_oracle.cpp) and the golden vectors produced by the
reference's own evaluator objects.  Tolerances as for MB-pol: 1e-10 relative on total energies / virial / stress,
1e-9 relative RMS on forces, pair counts bit-exact."""
import os

import numpy as np
import pytest

import mbpol_b200 as mb
from mbpol_b200 import mbnrg
from conftest import rel_rms

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_EVAL = 5e-10       # evaluator level: see tests/test_mbnrg.py (cancellation inside the fitted polynomial)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "mbnrg.npz"))


@pytest.fixture(scope="module")
def handle():
    h = mbnrg.mbnrg_init(9, [6, 2])
    yield h
    mbnrg.mbnrg_fini(h)


def test_h2o_ion_golden_vectors(handle, gold):
    for ion in range(1, 10):
        sel = gold["h2o_ion_ion"] == ion
        E, g1, g2 = mbnrg.mbnrg_eval_h2o_ion(handle, ion, gold["h2o_ion_w"][sel], gold["h2o_ion_x"][sel])
        for k in range(sel.sum()):
            Er, g1r, g2r = gold["h2o_ion_E"][sel][k], gold["h2o_ion_g1"][sel][k], gold["h2o_ion_g2"][sel][k]
            sc = max(abs(Er), np.abs(g1r).max(), 1e-300)
            assert abs(E[k] - Er) <= TOL_EVAL * sc and np.abs(g1[k] - g1r).max() <= TOL_EVAL * sc and np.abs(g2[k] - g2r).max() <= TOL_EVAL * sc
            assert (E[k] == 0.0) == (Er == 0.0)


def test_ion_ion_golden_vectors(handle, gold):
    for a, b in sorted(set(zip(gold["ion_ion_a"].tolist(), gold["ion_ion_b"].tolist()))):
        sel = (gold["ion_ion_a"] == a) & (gold["ion_ion_b"] == b)
        E, g1, g2 = mbnrg.mbnrg_eval_ion_ion(handle, a, b, gold["ion_ion_w"][sel], gold["ion_ion_x"][sel])
        sc = np.maximum(np.abs(gold["ion_ion_E"][sel]), np.abs(gold["ion_ion_g1"][sel]).max(1))
        ok = sc > 0
        assert np.all(np.abs(E - gold["ion_ion_E"][sel])[ok] <= TOL_EVAL * sc[ok])
        assert np.all(np.abs(g1 - gold["ion_ion_g1"][sel]).max(1)[ok] <= TOL_EVAL * sc[ok])
        assert np.all(np.abs(g2 - gold["ion_ion_g2"][sel]).max(1)[ok] <= TOL_EVAL * sc[ok])
        assert np.array_equal(E == 0.0, gold["ion_ion_E"][sel] == 0.0)
    with pytest.raises(mb.MBPolError):
        mbnrg.mbnrg_eval_ion_ion(handle, 1, 2, np.zeros((1, 3)), np.ones((1, 3)))      # f-cl: no evaluator in the reference


def test_random_clusters_against_oracle(handle, oracle):
    rng = np.random.default_rng(21)
    w0 = np.array([0.00, 0.00, 0.00, 0.9572, 0.00, 0.00, -0.2400, 0.9266, 0.00])
    for ion in (2, 6, 4, 5):
        n = 200
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        w = w0 + rng.normal(scale=0.04, size=(n, 9))
        x = w[:, :3] + d * rng.uniform(2.0, 7.2, size=(n, 1))
        E, g1, g2 = mbnrg.mbnrg_eval_h2o_ion(handle, ion, w, x)
        for k in range(n):
            Er, g1r, g2r = oracle.nrg_h2o_ion(ion, w[k], x[k], ref=oracle.have_nrg_reference)
            sc = max(abs(Er), np.abs(g1r).max(), 1e-300)
            assert abs(E[k] - Er) <= TOL_EVAL * sc and np.abs(g1[k] - g1r).max() <= TOL_EVAL * sc and np.abs(g2[k] - g2r).max() <= TOL_EVAL * sc
            assert (E[k] == 0.0) == (Er == 0.0)


def check_box(box, oracle, out, ref, cnt):
    assert [cnt["npairs"], cnt["nactive"], cnt["nionion"]] == [ref["npairs"], ref["nactive"], ref["nionion"]]
    assert abs(out["eng2"] - ref["eng2"]) <= 1e-10 * abs(ref["eng2"])
    assert abs(out["vir2"] - ref["vir2"]) <= 1e-10 * abs(ref["vir2"])
    assert np.abs(out["stress"] - ref["stress"]).max() <= 1e-10 * np.abs(ref["stress"]).max()
    for k in ("fxx", "fyy", "fzz"):
        assert rel_rms(out[k], ref[k]) <= 1e-9
    n0 = box["nions"] + box["nttm2"]
    assert not np.any(out["fxx"][n0:]) and not np.any(out["fyy"][n0:])      # M-sites are never touched


def test_small_box_against_golden_and_oracle(oracle, synth, gold):
    box = synth.make_ion_box(216, 8)
    h = mbnrg.mbnrg_init(box["nttm2"], box["ion_code"])
    out = mbnrg.run_ion_box(h, box)
    cnt = mbnrg.mbnrg_counts(h)
    assert [cnt["npairs"], cnt["nactive"], cnt["nionion"]] == gold["box_counts"].tolist()
    assert abs(out["eng2"] - gold["box_scalars"][0]) <= 1e-10 * abs(gold["box_scalars"][0])
    assert abs(out["vir2"] - gold["box_scalars"][1]) <= 1e-10 * abs(gold["box_scalars"][1])
    assert np.abs(out["stress"] - gold["box_stress"]).max() <= 1e-10 * np.abs(gold["box_stress"]).max()
    for k in ("fxx", "fyy", "fzz"):
        assert rel_rms(out[k], gold["box_" + k]) <= 1e-9
    check_box(box, oracle, out, oracle.nrg_forces(box), cnt)
    # accumulation semantics: a second call adds the same contribution again; engsrp/virsrp pass through
    f0 = out["fxx"].copy()
    r2 = mbnrg.mbnrg_forces(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["nions"], box["listions"], box["xxx"],
                            box["yyy"], box["zzz"], out["fxx"], out["fyy"], out["fzz"], out["stress"], engsrp=5.0, virsrp=-2.0)
    assert np.abs(out["fxx"] - 2 * f0).max() <= 1e-9 * np.abs(f0).max()
    assert abs(r2["engsrp"] - 5.0 - r2["eng2"]) <= 1e-9 * abs(r2["eng2"]) and abs(r2["eng2"] - out["eng2"]) <= 1e-10 * abs(out["eng2"])
    with pytest.raises(mb.MBPolError):
        mbnrg.mbnrg_forces(h, 7, box["cell"], box["nttm2"], box["listttm2"], box["nions"], box["listions"], box["xxx"], box["yyy"],
                           box["zzz"], out["fxx"], out["fyy"], out["fzz"], out["stress"])
    mbnrg.mbnrg_fini(h)


def test_nacl_box_config4(oracle, synth):
    """BASELINE config 4: 4032 H2O + 64 Na/Cl."""
    box = synth.make_ion_box(4096, 64)
    h = mbnrg.mbnrg_init(box["nttm2"], box["ion_code"])
    out = mbnrg.run_ion_box(h, box)
    ref = oracle.nrg_forces(box, use_ref=oracle.have_nrg_reference)
    check_box(box, oracle, out, ref, mbnrg.mbnrg_counts(h))
    assert abs(sum(out[k].sum() for k in ("fxx", "fyy", "fzz"))) <= 1e-7 * np.abs(out["fxx"]).max()
    # shards: partial sums of 3 ranks add up to the whole; ion-ion pairs on rank 0 only
    tot = {k: np.zeros_like(out[k]) for k in ("fxx", "fyy", "fzz", "stress")}
    e = v = 0.0
    npairs = 0
    for rank in range(3):
        mbnrg.mbnrg_set_shard(h, rank, 3)
        o = mbnrg.run_ion_box(h, box)
        c = mbnrg.mbnrg_counts(h)
        npairs += c["npairs"]
        assert c["nionion"] == (ref["nionion"] if rank == 0 else 0)
        ro = oracle.nrg_forces(box, rank=rank, nranks=3)
        assert abs(o["eng2"] - ro["eng2"]) <= 1e-10 * abs(ro["eng2"])
        e += o["eng2"]
        v += o["vir2"]
        for k in tot:
            tot[k] += o[k]
    assert npairs == ref["npairs"] and abs(e - out["eng2"]) <= 1e-11 * abs(e) and abs(v - out["vir2"]) <= 1e-10 * abs(v)
    assert rel_rms(tot["fxx"], out["fxx"]) <= 1e-11
    mbnrg.mbnrg_fini(h)


def test_other_ions_and_cells(oracle, synth):
    """Every ion species at once (the 29 fitted ion pairs get exercised) in an orthorhombic and a triclinic cell."""
    box = synth.make_ion_box(343, 27, ions=("f", "cl", "br", "i", "li", "na", "k", "rb", "cs"), seed=77)
    for imcon in (2, 3):
        b = dict(box)
        cell = box["cell"].copy()
        if imcon == 3:
            cell[3] = 0.08 * box["L"]          # mild shear: b = (0.08 L, L, 0)
            cell[7] = -0.05 * box["L"]
        b["cell"], b["imcon"] = cell, imcon
        h = mbnrg.mbnrg_init(b["nttm2"], b["ion_code"])
        out = mbnrg.run_ion_box(h, b)
        ref = oracle.nrg_forces(b, use_ref=oracle.have_nrg_reference)
        assert ref["nionion"] > 20
        check_box(b, oracle, out, ref, mbnrg.mbnrg_counts(h))
        mbnrg.mbnrg_fini(h)


def test_degenerate_sizes():
    h = mbnrg.mbnrg_init(0, [6, 2])                      # no water at all: only the ion pair
    f = [np.zeros(2) for _ in range(3)]
    st = np.zeros(9)
    x = np.array([0.0, 3.0]); y = np.zeros(2); z = np.zeros(2)
    cell = np.diag([30.0, 30.0, 30.0]).ravel()
    r = mbnrg.mbnrg_forces(h, 1, cell, 0, np.zeros(0, dtype=np.int32), 2, np.array([1, 2], dtype=np.int32), x, y, z, f[0], f[1], f[2], st)
    assert mbnrg.mbnrg_counts(h) == dict(npairs=0, nactive=0, nionion=1) and r["eng2"] != 0.0 and abs(f[0][0] + f[0][1]) <= 1e-12 * abs(f[0][0])
    mbnrg.mbnrg_fini(h)
    h = mbnrg.mbnrg_init(9, [])                          # no ions: nothing to do
    f = [np.zeros(3) for _ in range(3)]
    r = mbnrg.mbnrg_forces(h, 1, cell, 9, np.array([1, 2, 3] * 3, dtype=np.int32)[:9], 0, np.zeros(0, dtype=np.int32), np.arange(3.0), np.zeros(3),
                           np.zeros(3), f[0], f[1], f[2], st)
    assert r["eng2"] == 0.0 and not np.any(f[0])
    mbnrg.mbnrg_fini(h)


# ---- 3B ion-water-water ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gold3():
    return np.load(os.path.join(ROOT, "tests", "golden", "mbnrg3.npz"))


def test_3b_golden_vectors(handle, gold3):
    """mbnrg_b200_eval_h2o_h2o_ion against the reference objects' outputs for all nine ion species (2e-10 of max(|E|,|g|): the
    fitted coefficients cancel heavily, see tests/test_mbnrg.py)."""
    for ion in range(1, 10):
        sel = gold3["tri_ion"] == ion
        E, g = mbnrg.mbnrg_eval_h2o_h2o_ion(handle, ion, gold3["tri_w"][sel])
        sc = np.maximum(np.abs(gold3["tri_E"][sel]), np.abs(gold3["tri_g"][sel]).max(1))
        assert np.all(np.abs(E - gold3["tri_E"][sel]) <= 2e-10 * sc), ion
        assert np.all(np.abs(g - gold3["tri_g"][sel]).max(1) <= 2e-10 * sc), ion


def _check_3b(out, ref, cnt):
    assert cnt["ntrimers"] == ref["ntrimers"]
    assert abs(out["eng3"] - ref["eng3"]) <= 1e-10 * abs(ref["eng3"])
    assert abs(out["vir3"] - ref["vir3"]) <= 1e-9 * abs(ref["vir3"])
    for k in ("fxx", "fyy", "fzz"):
        assert rel_rms(out[k], ref[k]) <= 1e-9, k
    assert np.abs(out["stress"] - ref["stress"]).max() <= 1e-9 * np.abs(ref["stress"]).max()


def test_3b_small_box_against_golden(oracle, synth, gold3):
    box = synth.make_ion_box(216, 8)
    h = mbnrg.mbnrg_init(box["nttm2"], box["ion_code"])
    out = mbnrg.run_ion_box_3b(h, box)
    cnt = mbnrg.mbnrg_counts_3b(h)
    mbnrg.mbnrg_fini(h)
    ref = dict(eng3=gold3["box_scalars"][0], vir3=gold3["box_scalars"][1], stress=gold3["box_stress"], fxx=gold3["box_fxx"],
               fyy=gold3["box_fyy"], fzz=gold3["box_fzz"], ntrimers=int(gold3["box_counts"][0]))
    _check_3b(out, ref, cnt)


def test_3b_nacl_box_config4(oracle, synth):
    """BASELINE config 4 with the 3B term: 4032 H2O + 64 Na/Cl, ~70 k ion-water-water trimers against the oracle driver around the
    reference's evaluator objects; three shards add up to the whole."""
    box = synth.make_ion_box(4096, 64)
    h = mbnrg.mbnrg_init(box["nttm2"], box["ion_code"])
    out = mbnrg.run_ion_box_3b(h, box)
    cnt = mbnrg.mbnrg_counts_3b(h)
    ref = oracle.nrg_forces3(box, use_ref=oracle.have_nrg_reference_3b)
    assert ref["ntrimers"] > 30000
    _check_3b(out, ref, cnt)
    assert abs(sum(out[k].sum() for k in ("fxx", "fyy", "fzz"))) <= 1e-7 * np.abs(out["fxx"]).max()
    tot = {k: np.zeros_like(out[k]) for k in ("fxx", "fyy", "fzz", "stress")}
    e, n = 0.0, 0
    for rank in range(3):
        mbnrg.mbnrg_set_shard(h, rank, 3)
        o = mbnrg.run_ion_box_3b(h, box)
        n += mbnrg.mbnrg_counts_3b(h)["nshard"]
        e += o["eng3"]
        for k in tot:
            tot[k] += o[k]
    mbnrg.mbnrg_fini(h)
    assert n == ref["ntrimers"] and abs(e - out["eng3"]) <= 1e-11 * abs(e) and rel_rms(tot["fxx"], out["fxx"]) <= 1e-11
