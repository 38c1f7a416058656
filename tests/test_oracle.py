"""CPU tests: pin the oracle (oracle/mbpol_oracle.cpp) against the reference's golden vectors.

Golden data come from the reference's own evaluator objects (tests/golden/make_golden.py); when oracle/_ref is
present (this container, and the GPU box via the shipped .so) the restatement is additionally compared live."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_rms, sort_rows

OFF6 = ~np.eye(6, dtype=bool)
OFF9 = ~np.eye(9, dtype=bool)


def test_cutoffs(oracle):
    assert oracle.cutoffs() == (6.5, 4.5)          # mbpol-2b-poly.cpp:1863, mbpol-3b-poly.cpp:1489


def test_appendix_a_known_answers(oracle):
    """SURVEY.md Appendix A (values produced by the reference objects)."""
    kat = json.load(open(os.path.join(GOLDEN, "kat_appendix_a.json")))
    E, g1, g2, fm = oracle.dimer(kat["w1"], kat["w2"])
    assert abs(E - 5.361320870419604) < 1e-10 * abs(E)
    assert abs(E - kat["2b"]["E"]) < 1e-10 * abs(E)
    np.testing.assert_allclose(np.r_[g1, g2], np.r_[kat["2b"]["g1"], kat["2b"]["g2"]], rtol=0, atol=2e-10)
    np.testing.assert_allclose(fm[OFF6], np.array(kat["2b"]["fm"])[OFF6], rtol=0, atol=2e-10)
    np.testing.assert_allclose(fm[0, 3], [10.028612512800663, 0.99006855251045423, -0.46523833105513823], atol=1e-9)
    np.testing.assert_allclose(fm[3, 0], [-7.075500562773243, -0.2445330328729618, -0.4811900687080648], atol=1e-9)
    # the pair matrix is asymmetric by design but its rows sum to the gradient
    np.testing.assert_allclose(fm.sum(1).ravel(), np.r_[g1, g2], atol=1e-10)
    E, g1, g2, g3, fm = oracle.trimer(kat["w1"], kat["w2"], kat["w3"])
    assert abs(E - (-0.37951111966318934)) < 1e-10 * abs(E)
    np.testing.assert_allclose(np.r_[g1, g2, g3], np.r_[kat["3b"]["g1"], kat["3b"]["g2"], kat["3b"]["g3"]], atol=2e-10)
    np.testing.assert_allclose(fm[OFF9], np.array(kat["3b"]["fm"])[OFF9], atol=2e-10)
    np.testing.assert_allclose(fm[0, 3], [0.39206505476806081, 0.02751333717670602, -0.01375666858835301], atol=1e-10)
    np.testing.assert_allclose(fm, -fm.transpose(1, 0, 2), atol=1e-14)      # 3B pair matrix IS antisymmetric


def test_polynomial_tables_against_reference_values(oracle):
    """Extracted monomial tables reproduce poly_2b_v6x::eval / poly_3b_v2x::eval (poly-*.cpp) on random inputs."""
    g = np.load(os.path.join(GOLDEN, "poly.npz"))
    for tag in ("2b", "3b"):
        a = oracle.coef(tag)
        for x, E, gr in zip(g["x" + tag], g["E" + tag], g["g" + tag]):
            e, gg = oracle.poly(tag, a, x)
            assert abs(e - E) <= 1e-11 * max(abs(E), np.abs(gr).max())
            assert np.abs(gg - gr).max() <= 1e-11 * np.abs(gr).max()


def test_golden_clusters(oracle, golden_clusters):
    c = golden_clusters
    for w, E, g, fm in zip(c["dimer_w"], c["dimer_E"], c["dimer_g"], c["dimer_fm"]):
        e, g1, g2, f = oracle.dimer(w[:9], w[9:])
        scale = max(abs(E), np.abs(g).max())
        assert abs(e - E) <= 1e-10 * scale
        assert np.abs(np.r_[g1, g2] - g).max() <= 1e-10 * scale
        assert np.abs(f - fm)[OFF6].max() <= 1e-10 * max(scale, np.abs(fm[OFF6]).max())
    for w, E, g, fm in zip(c["trimer_w"], c["trimer_E"], c["trimer_g"], c["trimer_fm"]):
        e, g1, g2, g3, f = oracle.trimer(w[:9], w[9:18], w[18:])
        scale = max(abs(E), np.abs(g).max())
        assert abs(e - E) <= 1e-10 * scale
        assert np.abs(np.r_[g1, g2, g3] - g).max() <= 1e-10 * scale
        assert np.abs(f - fm)[OFF9].max() <= 1e-10 * scale


def test_live_reference_objects(oracle, golden_clusters):
    if not oracle.have_reference:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    c = golden_clusters
    for w, E in zip(c["trimer_w"][:8], c["trimer_E"][:8]):
        e = oracle.trimer(w[:9], w[9:18], w[18:], ref=True)[0]
        assert e == E           # the reference objects are deterministic: bit-for-bit
    for w, E in zip(c["dimer_w"][:8], c["dimer_E"][:8]):
        assert oracle.dimer(w[:9], w[9:], ref=True)[0] == E


def test_driver_listed_matches_golden_faithful(oracle, box64, golden_box64):
    """Restated driver + restated evaluators (listed mode) vs reference evaluators in the reference's own
    O(N^2)/O(N^3) in-place-mutating enumeration (faithful mode, golden)."""
    out = oracle.forces(box64, mode="listed", use_ref=False, nthreads=4)
    g = golden_box64
    assert np.array_equal(sort_rows(out["dimers"]), sort_rows(g["dimers"]))       # bit-exact membership
    assert np.array_equal(sort_rows(out["trimers"]), sort_rows(g["trimers"]))
    sc = g["scalars"]
    for k, v in zip(("eng2", "vir2", "eng3", "vir3"), sc):
        assert abs(out[k] - v) <= 1e-10 * abs(v), k
    for k in ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"):
        assert rel_rms(out[k], g[k]) <= 1e-10, k


def test_driver_faithful_mode_and_bead_striding(oracle, synth):
    """mbpol.f90:214-226, 306-314: the flat pair/triple index is dealt round-robin over bead ranks; partial results
    add up to the single-rank result (what gdsum does, mbpol.f90:395-431)."""
    box = synth.make_water_box(27, seed=5)
    one = oracle.forces(box, mode="faithful")
    parts = [oracle.forces(box, mode="faithful", bead_rank=r, bead_size=3) for r in range(3)]
    assert sum(p["ndimer"] for p in parts) == one["ndimer"] and sum(p["ntrimer"] for p in parts) == one["ntrimer"]
    for k in ("eng2", "vir2", "eng3", "vir3"):
        assert abs(sum(p[k] for p in parts) - one[k]) <= 1e-10 * abs(one[k])
    for k in ("fxx", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"):
        assert rel_rms(sum(p[k] for p in parts), one[k]) <= 1e-10
    lst = oracle.forces(box, mode="listed")
    assert np.array_equal(sort_rows(one["trimers"]), sort_rows(lst["trimers"]))
    assert np.array_equal(sort_rows(one["dimers"]), sort_rows(lst["dimers"]))
    assert rel_rms(lst["fxx"], one["fxx"]) <= 1e-10


def test_heat_accumulator_identities(oracle, synth):
    """heatcurrent.f90: sum_i e_i = E; rows of force_matrix sum to the force (check_forces, :841-852); the compressed
    tensor W reproduces compute_heat_flux2's jpot (:703-728) computed from the full N x N x 3 matrix."""
    box = synth.make_water_box(27, seed=5)
    n = box["mxatms"]
    out = oracle.forces(box, mode="faithful", want_matrix=True)
    assert abs(out["atomic_energy"].sum() - (out["eng2"] + out["eng3"])) <= 1e-10 * abs(out["eng2"] + out["eng3"])
    F = out["force_matrix"].reshape(3, n, n)           # (c, j, i)
    for c, k in enumerate(("fxx", "fyy", "fzz")):
        np.testing.assert_allclose(F[c].sum(0), out[k], atol=1e-8 * np.abs(out[k]).max())
    jm = oracle.heat_flux2_matrix(n, out["force_matrix"], box["xxx"], box["yyy"], box["zzz"], box["vxx"], box["vyy"], box["vzz"])
    jw = oracle.heat_flux2_fdecomp(n, out["atomic_fdecomp"], box["vxx"], box["vyy"], box["vzz"])
    np.testing.assert_allclose(jw, jm, rtol=1e-10)
    assert np.abs(out["fxx"]).sum() > 0 and abs(out["fxx"].sum()) < 1e-8 * np.abs(out["fxx"]).sum()   # momentum conservation
    assert np.allclose(out["stress"].reshape(3, 3), out["stress"].reshape(3, 3).T)


def test_imaging(oracle):
    """image_vectors (mbpol.f90:1305-1361): nint = round half away from zero; imcon 1/2/3 agree on a cubic cell."""
    L = 10.0
    cell = np.diag([L, L, L]).ravel()
    dx = np.array([5.0, -5.0, 4.999, 15.2, -14.9, 0.3])
    r1 = oracle.image_vectors(1, cell, dx)
    np.testing.assert_allclose(r1, [-5.0, 5.0, 4.999, -4.8, -4.9, 0.3], atol=1e-12)
    np.testing.assert_array_equal(oracle.image_vectors(2, cell, dx), r1)
    np.testing.assert_allclose(oracle.image_vectors(3, cell, dx), r1, atol=1e-12)
    with pytest.raises(ValueError):
        oracle.image_vectors(0, cell, dx)


def test_cell_types_agree(oracle, synth):
    box = synth.make_water_box(27, seed=5)
    a = oracle.forces(box, mode="listed")
    for imcon in (2, 3):
        b = dict(box); b["imcon"] = imcon
        r = oracle.forces(b, mode="listed")
        assert np.array_equal(r["trimers"], a["trimers"])
        assert abs(r["eng3"] - a["eng3"]) <= 1e-10 * abs(a["eng3"])


def test_translation_and_relabelling_invariance(oracle, synth):
    box = synth.make_water_box(27, seed=5)
    a = oracle.forces(box, mode="listed")
    b = dict(box)
    L = box["L"]
    b["xxx"] = box["xxx"] + 0.37 * L; b["yyy"] = box["yyy"] - 2 * L; b["zzz"] = box["zzz"] + 0.11
    r = oracle.forces(b, mode="listed")
    assert r["ntrimer"] == a["ntrimer"]
    assert abs((r["eng2"] + r["eng3"]) - (a["eng2"] + a["eng3"])) <= 1e-9 * abs(a["eng2"] + a["eng3"])
    assert rel_rms(r["fxx"], a["fxx"]) <= 1e-9


def test_switches_vanish_at_the_cutoffs(oracle):
    w1 = np.array([0, 0, 0, 0.9572, 0, 0, -0.24, 0.9266, 0.0])
    for r, expect_zero in ((6.4999999, False), (6.5, False), (6.5000001, True)):
        w2 = w1 + np.array([r, 0, 0] * 3)
        E = oracle.dimer(w1, w2)[0]
        assert (E == 0.0) == expect_zero or abs(E) < 1e-12
    w2 = w1 + np.array([3.0, 0, 0] * 3)
    w3 = w1 + np.array([0, 4.6, 0] * 3)       # |ab| = 3.0, |ac| = 4.6 > 4.5, |bc| = 5.49 > 4.5 -> s = 0
    E, g1, g2, g3, _ = oracle.trimer(w1, w2, w3)
    assert E == 0.0 and not np.any(g1) and not np.any(g3)


def test_triclinic_cell_faithful_vs_listed(oracle, synth):
    """imcon = 3 (invert.f path): the reference's O(N^3) enumeration and the list-driven mode agree."""
    box = dict(synth.make_water_box(27, seed=5))
    L = box["L"]
    cell = np.array([[L, 0.0, 0.0], [0.2 * L, 1.05 * L, 0.0], [0.1 * L, -0.15 * L, 0.97 * L]])
    box["imcon"], box["cell"] = 3, cell.ravel().copy()
    a = oracle.forces(box, mode="faithful")
    b = oracle.forces(box, mode="listed")
    assert np.array_equal(sort_rows(a["trimers"]), sort_rows(b["trimers"]))
    assert np.array_equal(sort_rows(a["dimers"]), sort_rows(b["dimers"]))
    assert abs(a["eng3"] - b["eng3"]) <= 1e-10 * abs(a["eng3"]) and rel_rms(b["fxx"], a["fxx"]) <= 1e-10
