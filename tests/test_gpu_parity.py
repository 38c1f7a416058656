"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the oracle on
the same seeded inputs, against the committed golden fixtures, and -- at BASELINE sizes -- through size-independent
properties.  Tolerances are the north_star's: 1e-10 relative on total 2B+3B energy, 1e-9 relative RMS on forces and
per-atom tensors, bit-exact dimer/trimer list membership."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_rms, sort_rows

pytestmark = pytest.mark.gpu
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

E_TOL = 1e-10
F_TOL = 1e-9
OFF6 = ~np.eye(6, dtype=bool)
OFF9 = ~np.eye(9, dtype=bool)


@pytest.fixture(scope="module")
def mb():
    import mbpol_b200
    mbpol_b200.load_library()
    return mbpol_b200


def test_legacy_symbols_known_answers(mb):
    """mbpol_2b_poly_ / mbpol_3b_poly_ (the reference's own C symbols) on SURVEY.md Appendix A."""
    kat = json.load(open(os.path.join(GOLDEN, "kat_appendix_a.json")))
    E, g1, g2, fm = mb.mbpol_2b_poly(kat["w1"], kat["w2"])
    assert abs(E - 5.361320870419604) <= E_TOL * abs(E)
    np.testing.assert_allclose(np.r_[g1, g2], np.r_[kat["2b"]["g1"], kat["2b"]["g2"]], atol=F_TOL * 18)
    np.testing.assert_allclose(fm[OFF6], np.array(kat["2b"]["fm"])[OFF6], atol=F_TOL * 18)
    E, g1, g2, g3, fm = mb.mbpol_3b_poly(kat["w1"], kat["w2"], kat["w3"])
    assert abs(E - (-0.37951111966318934)) <= E_TOL * abs(E)
    np.testing.assert_allclose(np.r_[g1, g2, g3], np.r_[kat["3b"]["g1"], kat["3b"]["g2"], kat["3b"]["g3"]], atol=F_TOL)
    np.testing.assert_allclose(fm[OFF9], np.array(kat["3b"]["fm"])[OFF9], atol=F_TOL)
    r = mb.mbpol_cutoffs()
    assert r == (6.5, 4.5)


def test_golden_clusters_batched(mb, golden_clusters):
    c = golden_clusters
    E, g, fm = mb.mbpol_eval_dimers(c["dimer_w"])
    scale = np.maximum(np.abs(c["dimer_E"]), np.abs(c["dimer_g"]).max(1))
    assert np.all(np.abs(E - c["dimer_E"]) <= E_TOL * scale)
    assert np.all(np.abs(g - c["dimer_g"]).max(1) <= F_TOL * scale)
    assert np.abs(fm - c["dimer_fm"])[:, OFF6].max() <= F_TOL * np.abs(c["dimer_fm"][:, OFF6]).max()
    E, g, fm = mb.mbpol_eval_trimers(c["trimer_w"])
    scale = np.maximum(np.abs(c["trimer_E"]), np.abs(c["trimer_g"]).max(1))
    assert np.all(np.abs(E - c["trimer_E"]) <= E_TOL * scale)
    assert np.all(np.abs(g - c["trimer_g"]).max(1) <= F_TOL * scale)
    assert np.abs(fm - c["trimer_fm"])[:, OFF9].max() <= F_TOL * np.abs(c["trimer_fm"][:, OFF9]).max()


def compare(out, ref, keys):
    e, er = out["eng2"] + out["eng3"], ref["eng2"] + ref["eng3"]
    assert abs(e - er) <= E_TOL * abs(er)
    for k in ("eng2", "vir2", "eng3", "vir3"):
        assert abs(out[k] - ref[k]) <= 1e-9 * abs(ref[k]), k
    for k in keys:
        assert rel_rms(out[k], ref[k]) <= F_TOL, k


def test_box64_against_golden_faithful_reference(mb, box64, golden_box64):
    """64-water box vs the reference evaluators in the reference's own enumeration order (golden fixture)."""
    h = mb.mbpol_init(box64["nttm2"], flags=7)
    out = mb.run_box(h, box64, want_stress=True)
    d, t = mb.mbpol_lists(h)
    mb.mbpol_fini(h)
    g = golden_box64
    assert np.array_equal(sort_rows(d), sort_rows(g["dimers"]))
    assert np.array_equal(sort_rows(t), sort_rows(g["trimers"]))
    ref = {k: g[k] for k in ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress")}
    ref.update(dict(zip(("eng2", "vir2", "eng3", "vir3"), g["scalars"])))
    compare(out, ref, ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"))


@pytest.mark.parametrize("nw", [27, 256, 700])
def test_box_against_oracle(mb, oracle, synth, nw):
    box = synth.make_water_box(nw, seed=11 + nw)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    out = mb.run_box(h, box, want_stress=True)
    d, t = mb.mbpol_lists(h)
    c = mb.mbpol_counts(h)
    mb.mbpol_fini(h)
    ref = oracle.forces(box, mode="listed", use_ref=oracle.have_reference, nthreads=os.cpu_count())
    assert c["dimers"] == ref["ndimer"] and c["trimers"] == ref["ntrimer"]
    assert np.array_equal(sort_rows(d), sort_rows(ref["dimers"]))
    assert np.array_equal(sort_rows(t), sort_rows(ref["trimers"]))
    compare(out, ref, ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"))
    n = box["mxatms"]
    j1, j2 = oracle.heat_flux(n, ref["atomic_energy"], ref["atomic_stress"], box["vxx"], box["vyy"], box["vzz"])
    j3 = oracle.heat_flux2_fdecomp(n, ref["atomic_fdecomp"], box["vxx"], box["vyy"], box["vzz"])
    jr = np.r_[j1, j2, j3]
    assert np.abs(out["jflux"] - jr).max() <= F_TOL * np.abs(jr).max()


def test_default_flags_leave_atomic_stress_alone(mb, box64, oracle):
    """Reference default build: HEAT_STRESS off -> atomic_stress_mbpol stays zero."""
    h = mb.mbpol_init(box64["nttm2"])            # DEFAULT_FLAGS = energy + force decomposition
    out = mb.run_box(h, box64, want_stress=False)
    with pytest.raises(mb.MBPolError):
        mb.run_box(h, box64, want_stress=True)     # asking for it without the flag is an error, not garbage
    mb.mbpol_fini(h)
    ref = oracle.forces(box64, mode="listed", use_ref=oracle.have_reference, want_stress=False)
    compare(out, ref, ("fxx", "fyy", "fzz", "atomic_energy", "atomic_fdecomp"))


def test_accumulate_semantics(mb, box64):
    """Outputs are added into the caller's arrays (fxx = fxx - g, stress = stress + strs, mbpol.f90:247-255, 582-599)."""
    h = mb.mbpol_init(box64["nttm2"])
    a = mb.run_box(h, box64)
    n = box64["mxatms"]
    fx, fy, fz = np.full(n, 1.5), np.full(n, -2.0), np.zeros(n)
    st = np.arange(9.0)
    r = mb.mbpol_forces(h, box64["imcon"], box64["cell"], box64["nttm2"], box64["listttm2"], box64["xxx"], box64["yyy"],
                        box64["zzz"], fx, fy, fz, engsrp=10.0, virsrp=-3.0, stress=st)
    mb.mbpol_fini(h)
    assert rel_rms(fx - 1.5, a["fxx"]) < 1e-11 and rel_rms(fy + 2.0, a["fyy"]) < 1e-11
    assert rel_rms(st - np.arange(9.0), a["stress"]) < 1e-11
    assert abs(r["engsrp"] - 10.0 - (a["eng2"] + a["eng3"])) < 1e-9 * abs(a["eng2"])


def test_accumulate_semantics_pinned_arrays(mb, box64):
    """Caller arrays in page-locked memory take the fused end of mbpol_b200_forces (values uploaded, added on the device, copied
    straight back): same "+=" semantics, for every per-atom array, also when the call is repeated on the same arrays."""
    h = mb.mbpol_init(box64["nttm2"], flags=7)
    a = mb.run_box(h, box64, want_stress=True)
    n = box64["mxatms"]
    arrs = {"fxx": np.full(n, 1.5), "fyy": np.full(n, -2.0), "fzz": np.zeros(n), "ae": np.full(n, 0.25), "as": np.full(9 * n, -1.0), "afd": np.full(9 * n, 3.0)}
    for v in arrs.values():
        mb.host_register(v)
    try:
        for rep in (1, 2):
            st = np.arange(9.0)
            r = mb.mbpol_forces(h, box64["imcon"], box64["cell"], box64["nttm2"], box64["listttm2"], box64["xxx"], box64["yyy"], box64["zzz"],
                                arrs["fxx"], arrs["fyy"], arrs["fzz"], engsrp=10.0, stress=st, atomic_energy=arrs["ae"], atomic_stress=arrs["as"],
                                atomic_fdecomp=arrs["afd"], vxx=box64["vxx"], vyy=box64["vyy"], vzz=box64["vzz"])
            assert rel_rms(arrs["fxx"] - 1.5, rep * a["fxx"]) < 1e-11 and rel_rms(arrs["fyy"] + 2.0, rep * a["fyy"]) < 1e-11
            assert rel_rms(arrs["fzz"], rep * a["fzz"]) < 1e-11
            assert rel_rms(arrs["ae"] - 0.25, rep * a["atomic_energy"]) < 1e-11
            assert rel_rms(arrs["as"] + 1.0, rep * a["atomic_stress"]) < 1e-11
            assert rel_rms(arrs["afd"] - 3.0, rep * a["atomic_fdecomp"]) < 1e-11
            assert rel_rms(st - np.arange(9.0), a["stress"]) < 1e-11
            assert abs(r["engsrp"] - 10.0 - (a["eng2"] + a["eng3"])) < 1e-9 * abs(a["eng2"])
            assert rel_rms(r["jflux"], a["jflux"]) < 1e-10
    finally:
        for v in arrs.values():
            mb.host_unregister(v)
        mb.mbpol_fini(h)


def test_cell_types(mb, synth):
    """imcon 1/2/3 (mbpol.f90:1319-1358) on the same cubic cell give the same lists and energies; imcon 0 is rejected."""
    box = synth.make_water_box(216, seed=3)
    res = []
    for imcon in (1, 2, 3):
        b = dict(box); b["imcon"] = imcon
        h = mb.mbpol_init(b["nttm2"])
        out = mb.run_box(h, b)
        res.append((out, sort_rows(mb.mbpol_lists(h)[1])))
        mb.mbpol_fini(h)
    for out, tl in res[1:]:
        assert np.array_equal(tl, res[0][1])
        assert abs(out["eng3"] - res[0][0]["eng3"]) <= 1e-10 * abs(res[0][0]["eng3"])
        assert rel_rms(out["fxx"], res[0][0]["fxx"]) <= F_TOL
    b = dict(box); b["imcon"] = 0
    h = mb.mbpol_init(b["nttm2"])
    with pytest.raises(mb.MBPolError):
        mb.run_box(h, b)
    mb.mbpol_fini(h)


def test_shards_sum_to_the_whole(mb, synth):
    """(rank, nranks) shards of the work lists add up to the single-rank result -- what gdsum / NCCL completes."""
    box = synth.make_water_box(256)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    full = mb.run_box(h, box, want_stress=True)
    d_all, t_all = mb.mbpol_lists(h)
    acc = None
    dl, tl = [], []
    for r in range(3):
        mb.mbpol_set_shard(h, r, 3)
        p = mb.run_box(h, box, want_stress=True)
        d, t = mb.mbpol_lists(h)                       # this rank's shard: the lists it built itself
        sh = mb.mbpol_shard(h)
        assert np.all((d[:, 0] >= sh["first_water"]) & (d[:, 0] < sh["first_water"] + sh["n_waters"]))
        dl.append(d); tl.append(t)
        if acc is None:
            acc = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in p.items()}
        else:
            for k, v in p.items():
                acc[k] = acc[k] + v
    # the shards partition the lists of the whole box exactly
    assert np.array_equal(sort_rows(np.concatenate(dl)), sort_rows(d_all)) and np.array_equal(sort_rows(np.concatenate(tl)), sort_rows(t_all))
    mb.mbpol_fini(h)
    compare(acc, full, ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress", "jflux"))


def test_full_size_properties_4096(mb, oracle, synth):
    """BASELINE config 2 (4096 H2O): size-independent properties + sampled clusters against the oracle."""
    box = synth.make_water_box(4096)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    out = mb.run_box(h, box, want_stress=True)
    d, t = mb.mbpol_lists(h)
    assert len(np.unique(t, axis=0)) == len(t) and np.all(t[:, 0] < t[:, 1]) and np.all(t[:, 1] < t[:, 2])
    assert len(np.unique(d, axis=0)) == len(d) and np.all(d[:, 0] < d[:, 1])
    e = out["eng2"] + out["eng3"]
    assert abs(out["atomic_energy"].sum() - e) <= 1e-10 * abs(e)                     # sum_i e_i = E
    for k in ("fxx", "fyy", "fzz"):
        assert abs(out[k].sum()) <= 1e-9 * np.abs(out[k]).sum()                       # no net force
        assert not np.any(out[k][3 * 4096:])                                          # M-sites untouched
    assert np.allclose(out["stress"].reshape(3, 3), out["stress"].reshape(3, 3).T, rtol=0, atol=1e-9 * np.abs(out["stress"]).max())
    assert abs(np.trace(out["stress"].reshape(3, 3)) + (out["vir2"] + out["vir3"])) <= 1e-9 * abs(out["vir2"])   # tr(stress) = -virial
    # rigid translation (re-wrapped) + relabelling of the waters leave the physics unchanged
    L = box["L"]
    perm = np.random.default_rng(0).permutation(4096)
    b = dict(box)
    for k, s in (("xxx", 0.4 * L), ("yyy", -1.3), ("zzz", 7.7)):
        v = box[k][:3 * 4096].reshape(4096, 3)[perm].ravel() + s
        b[k] = np.concatenate([v - L * np.round(v / L), box[k][3 * 4096:]])
    for k in ("vxx", "vyy", "vzz"):
        b[k] = np.concatenate([box[k][:3 * 4096].reshape(4096, 3)[perm].ravel(), box[k][3 * 4096:]])
    out2 = mb.run_box(h, b, want_stress=True)
    assert mb.mbpol_counts(h)["trimers"] == len(t)
    assert abs((out2["eng2"] + out2["eng3"]) - e) <= 1e-9 * abs(e)
    f1 = out["fxx"][:3 * 4096].reshape(4096, 3)[perm].ravel()
    assert rel_rms(out2["fxx"][:3 * 4096], f1) <= F_TOL
    assert np.abs(out2["jflux"][:6] - out["jflux"][:6]).max() <= 1e-8 * np.abs(out["jflux"][:6]).max()
    mb.mbpol_fini(h)
    # sampled clusters of the full-size lists against the oracle evaluators
    nw = 4096
    mono = np.stack([box["xxx"][:3 * nw], box["yyy"][:3 * nw], box["zzz"][:3 * nw]], 1).reshape(nw, 3, 3)
    mono[:, 1:] = mono[:, :1] + (mono[:, 1:] - mono[:, :1]) - L * np.round((mono[:, 1:] - mono[:, :1]) / L)
    sel = t[np.random.default_rng(1).choice(len(t), 64, replace=False)]

    def near(wa, wb):
        dd = wa[0] - wb[0]
        return wa - wa[0] + wb[0] + (dd - L * np.round(dd / L))
    w = np.array([np.r_[near(mono[i], mono[j]).ravel(), mono[j].ravel(), near(mono[k], mono[j]).ravel()] for i, j, k in sel])
    E, g, _ = mb.mbpol_eval_trimers(w, want_fm=False)
    for q in range(len(sel)):
        ref = oracle.trimer(w[q, :9], w[q, 9:18], w[q, 18:], ref=oracle.have_reference)
        sc = max(abs(ref[0]), np.abs(np.r_[ref[1], ref[2], ref[3]]).max())
        assert abs(E[q] - ref[0]) <= E_TOL * sc and np.abs(g[q] - np.r_[ref[1], ref[2], ref[3]]).max() <= F_TOL * sc


def _whole_box_against_oracle(mb, oracle, synth, nw):
    """The WHOLE box against the oracle's listed mode driven by the reference's own evaluator objects when they are built
    (mbpol.f90:98-449 restated, list-driven): lists bit-exact, E 1e-10, every scalar 1e-9, forces / e_i / W / S / stress 1e-9."""
    box = synth.make_water_box(nw)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    out = mb.run_box(h, box, want_stress=True)
    d, t = mb.mbpol_lists(h)
    mb.mbpol_fini(h)
    ref = oracle.forces(box, mode="listed", use_ref=oracle.have_reference, nthreads=os.cpu_count())
    assert len(d) == ref["ndimer"] and len(t) == ref["ntrimer"]
    assert np.array_equal(sort_rows(d), sort_rows(ref["dimers"]))
    assert np.array_equal(sort_rows(t), sort_rows(ref["trimers"]))
    compare(out, ref, ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"))
    n = box["mxatms"]
    j1, j2 = oracle.heat_flux(n, ref["atomic_energy"], ref["atomic_stress"], box["vxx"], box["vyy"], box["vzz"])
    j3 = oracle.heat_flux2_fdecomp(n, ref["atomic_fdecomp"], box["vxx"], box["vyy"], box["vzz"])
    jr = np.r_[j1, j2, j3]
    assert np.abs(out["jflux"] - jr).max() <= F_TOL * np.abs(jr).max()
    return len(d), len(t)


def test_whole_box_4096_against_oracle(mb, oracle, synth):
    """BASELINE config 2 (4096 H2O, heat flux on): every list entry and every per-atom output against the oracle."""
    nd, nt = _whole_box_against_oracle(mb, oracle, synth, 4096)
    assert nd > 60000 and nt > 250000


def test_whole_box_32768_against_oracle(mb, oracle, synth):
    """BASELINE config 3, the headline box (32 768 H2O, 2.3 M trimers): the same whole-box comparison (a few seconds of host
    time with the reference objects on all cores)."""
    nd, nt = _whole_box_against_oracle(mb, oracle, synth, 32768)
    assert nd > 500000 and nt > 2000000


def _sheared_box(synth, nw, imcon, seed=21):
    """Non-cubic cells: the oxygens of a cubic box are mapped affinely into the new cell, the molecules stay rigid."""
    box = dict(synth.make_water_box(nw, seed=seed))
    L = box["L"]
    if imcon == 2:
        cell = np.diag([L, 1.12 * L, 0.93 * L])
    else:
        cell = np.array([[L, 0.0, 0.0], [0.21 * L, 1.05 * L, 0.0], [0.08 * L, -0.17 * L, 0.96 * L]])   # rows = cell vectors
    n3 = 3 * nw
    pos = np.stack([box["xxx"], box["yyy"], box["zzz"]], 1)
    mol = pos[:n3].reshape(nw, 3, 3)
    rel = mol - mol[:, :1]
    rel -= L * np.round(rel / L)                       # un-split molecules of the cubic box
    frac = mol[:, 0] / L                               # fractional oxygen coordinates in the cube
    newO = frac @ cell
    pos[:n3] = (newO[:, None, :] + rel).reshape(n3, 3)
    box["imcon"] = imcon
    box["cell"] = cell.ravel().copy()
    for k, c in zip(("xxx", "yyy", "zzz"), range(3)):
        box[k] = np.ascontiguousarray(pos[:, c])
    return box


@pytest.mark.parametrize("imcon", [2, 3])
def test_non_cubic_cells_against_oracle(mb, oracle, synth, imcon):
    """Orthorhombic (imcon 2) and triclinic (imcon 3, invert.f) cells: lists bit-exact, values within tolerance."""
    box = _sheared_box(synth, 343, imcon)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    out = mb.run_box(h, box, want_stress=True)
    d, t = mb.mbpol_lists(h)
    mb.mbpol_fini(h)
    ref = oracle.forces(box, mode="listed", use_ref=oracle.have_reference, nthreads=os.cpu_count())
    assert len(t) > 1000
    assert np.array_equal(sort_rows(d), sort_rows(ref["dimers"]))
    assert np.array_equal(sort_rows(t), sort_rows(ref["trimers"]))
    compare(out, ref, ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"))


@pytest.mark.parametrize("nb,nw", [(5, 64), (32, 512)])
def test_bead_batched_against_oracle(mb, oracle, synth, nb, nw):
    """PIMD-style batch (BASELINE config 5 is 32 beads x 512 H2O, multibead.f90:25-93): all beads in
    ONE call, every bead compared with the ORACLE run on that bead alone (lists, scalars, forces, per-atom tensors, flux)."""
    boxes = [synth.make_water_box(nw, seed=100 + b) for b in range(nb)]      # same L (same density), different coordinates
    n = boxes[0]["mxatms"]
    stack = lambda k: np.ascontiguousarray(np.stack([b[k] for b in boxes]))
    X, Y, Z, VX, VY, VZ = (stack(k) for k in ("xxx", "yyy", "zzz", "vxx", "vyy", "vzz"))
    F = [np.zeros((nb, n)) for _ in range(3)]
    ae, ast, afd = np.zeros((nb, n)), np.zeros((nb, 9 * n)), np.zeros((nb, 9 * n))
    h = mb.mbpol_init_batched(boxes[0]["nttm2"], nb, flags=7)
    r = mb.mbpol_forces_batched(h, 1, boxes[0]["cell"], boxes[0]["nttm2"], boxes[0]["listttm2"], X, Y, Z, F[0], F[1], F[2],
                                atomic_energy=ae, atomic_stress=ast, atomic_fdecomp=afd, vxx=VX, vyy=VY, vzz=VZ)
    c = mb.mbpol_counts(h)
    d, t = mb.mbpol_lists(h)              # water indices of all beads: bead b owns [b*nw, (b+1)*nw)
    mb.mbpol_fini(h)
    ntri = ndim = 0
    for b in range(nb):
        ref = oracle.forces(boxes[b], mode="listed", use_ref=oracle.have_reference, nthreads=os.cpu_count())
        ntri += ref["ntrimer"]
        ndim += ref["ndimer"]
        tb = t[(t[:, 0] >= b * nw) & (t[:, 0] < (b + 1) * nw)] - b * nw
        db = d[(d[:, 0] >= b * nw) & (d[:, 0] < (b + 1) * nw)] - b * nw
        assert np.array_equal(sort_rows(tb), sort_rows(ref["trimers"])) and np.array_equal(sort_rows(db), sort_rows(ref["dimers"])), b
        e, er = r["eng2"][b] + r["eng3"][b], ref["eng2"] + ref["eng3"]
        assert abs(e - er) <= E_TOL * abs(er), b
        for k in ("eng2", "vir2", "eng3", "vir3"):
            assert abs(r[k][b] - ref[k]) <= 1e-9 * abs(ref[k]), (b, k)
        for got, k in ((F[0][b], "fxx"), (F[1][b], "fyy"), (F[2][b], "fzz"), (ae[b], "atomic_energy"), (ast[b], "atomic_stress"),
                       (afd[b], "atomic_fdecomp"), (r["stress"][b], "stress")):
            assert rel_rms(got, ref[k]) <= F_TOL, (b, k)
        j1, j2 = oracle.heat_flux(n, ref["atomic_energy"], ref["atomic_stress"], boxes[b]["vxx"], boxes[b]["vyy"], boxes[b]["vzz"])
        j3 = oracle.heat_flux2_fdecomp(n, ref["atomic_fdecomp"], boxes[b]["vxx"], boxes[b]["vyy"], boxes[b]["vzz"])
        jr = np.r_[j1, j2, j3]
        assert np.abs(r["jflux"][b] - jr).max() <= F_TOL * np.abs(jr).max(), b
    assert c["trimers"] == ntri and c["dimers"] == ndim


@pytest.mark.parametrize("engine", ["tape", "fat", "sym"])
def test_every_polynomial_engine(engine):
    """The three selectable polynomial engines (MBPOL_B200_ENGINE) give the same physics: run the 64-water golden
    comparison in a fresh process per engine (the choice is read when a handle is created)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import mbpol_b200 as mb
from mbpol_b200 import synth
from conftest import GOLDEN, rel_rms, sort_rows
import os
g = np.load(os.path.join(GOLDEN, "box64.npz"))
box = synth.make_water_box(64)
h = mb.mbpol_init(box["nttm2"], flags=7)
out = mb.run_box(h, box, want_stress=True)
d, t = mb.mbpol_lists(h)
assert np.array_equal(sort_rows(t), sort_rows(g["trimers"])) and np.array_equal(sort_rows(d), sort_rows(g["dimers"]))
e, er = out["eng2"] + out["eng3"], g["scalars"][0] + g["scalars"][2]
assert abs(e - er) <= 1e-10 * abs(er), (e, er)
for k in ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"):
    assert rel_rms(out[k], g[k]) <= 1e-9, k
mb.mbpol_fini(h)
print("engine ok")
''' % (ROOT_DIR, os.path.join(ROOT_DIR, "tests"))
    env = dict(os.environ, MBPOL_B200_ENGINE=engine)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "engine ok" in r.stdout, r.stdout + r.stderr


def small_system(positions, L=30.0):
    """nw rigid waters with their oxygens at `positions` (a hand-made cluster in a big cubic box)."""
    w0 = np.array([[0.0, 0.0, 0.0], [0.9572, 0.0, 0.0], [-0.24, 0.9266, 0.0]])
    nw = len(positions)
    xyz = np.zeros((4 * max(nw, 1), 3))
    for m, p in enumerate(positions):
        xyz[3 * m:3 * m + 3] = w0 + np.asarray(p, dtype=float)
    cell = np.zeros(9)
    cell[0] = cell[4] = cell[8] = L
    return {"nw": nw, "mxatms": len(xyz), "nttm2": 3 * nw, "imcon": 1, "cell": cell, "L": L,
            "listttm2": np.arange(1, 3 * nw + 1, dtype=np.int32), "xxx": np.ascontiguousarray(xyz[:, 0]),
            "yyy": np.ascontiguousarray(xyz[:, 1]), "zzz": np.ascontiguousarray(xyz[:, 2]),
            "vxx": np.zeros(len(xyz)), "vyy": np.zeros(len(xyz)), "vzz": np.zeros(len(xyz))}


@pytest.mark.parametrize("positions,nd,nt", [
    ([], 0, 0),                                                        # empty system
    ([(0, 0, 0)], 0, 0),                                               # one water: no cluster at all
    ([(0, 0, 0), (2.9, 0.3, 0.1)], 1, 0),                              # one dimer, no trimer
    ([(0, 0, 0), (8.0, 0, 0)], 0, 0),                                  # beyond the 2B cutoff
    ([(0, 0, 0), (2.9, 0.3, 0.1), (1.2, 2.7, 0.4)], 3, 1),             # closed triangle: one trimer
    ([(0, 0, 0), (3.0, 0, 0), (7.0, 0.2, 0)], 2, 1),                   # chain: 0-2 is 7 A apart (no dimer), trimer via the centre water
    ([(0, 0, 0), (4.4, 0, 0), (9.1, 0, 0)], 2, 0),                     # 1-2 at 4.7 A: inside r2f, outside r3f -> no trimer
])
def test_degenerate_systems(mb, oracle, positions, nd, nt):
    """Empty and tiny inputs (the reference's loops simply do not execute): counts, energies and forces against the oracle."""
    box = small_system(positions)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    out = mb.run_box(h, box, want_stress=True)
    c = mb.mbpol_counts(h)
    mb.mbpol_fini(h)
    assert (c["dimers"], c["trimers"]) == (nd, nt)
    if box["nw"] == 0:
        assert out["eng2"] == 0.0 and out["eng3"] == 0.0 and not np.any(out["fxx"])
        return
    ref = oracle.forces(box, mode="listed", use_ref=oracle.have_reference, nthreads=1)
    assert (ref["ndimer"], ref["ntrimer"]) == (nd, nt)
    for k in ("eng2", "vir2", "eng3", "vir3"):
        assert abs(out[k] - ref[k]) <= E_TOL * max(abs(ref[k]), 1e-300), k
    if nd:      # planar clusters: a component can vanish by symmetry, so scale by the largest force / tensor entry
        fmax = max(np.abs(ref[k]).max() for k in ("fxx", "fyy", "fzz"))
        for k in ("fxx", "fyy", "fzz"):
            assert np.abs(out[k] - ref[k]).max() <= F_TOL * fmax, k
        for k in ("atomic_energy", "atomic_fdecomp", "atomic_stress", "stress"):
            assert np.abs(out[k] - ref[k]).max() <= F_TOL * np.abs(ref[k]).max(), k
    else:
        assert not np.any(out["fxx"]) and not np.any(out["atomic_energy"])


def test_unphysical_density_is_reported_not_truncated(mb):
    """More neighbours than the per-water row capacity: the call must fail with ECAPACITY, never drop clusters silently."""
    rng = np.random.default_rng(0)
    pos = rng.uniform(0.0, 4.0, size=(300, 3))             # 300 waters inside a 4 A cube
    box = small_system(pos, L=40.0)
    h = mb.mbpol_init(box["nttm2"], flags=3)
    with pytest.raises(mb.MBPolError) as e:
        mb.run_box(h, box)
    assert "status 3" in str(e.value)
    mb.mbpol_fini(h)


def test_config_and_field_files_feed_the_path(mb, synth, tmp_path):
    """Row f4: a box written as DL_POLY CONFIG + FIELD (centroid_write_config.f:38-87 format, ttm2list.f numbering), read back and
    evaluated, reproduces the result of the in-memory synthetic box bit for bit (the files carry 17 significant digits)."""
    from mbpol_b200 import config_io
    box = synth.make_water_box(256)
    config_io.write_config(str(tmp_path / "CONFIG"), box)
    config_io.write_field(str(tmp_path / "FIELD"), box)
    b2 = config_io.read_config(str(tmp_path / "CONFIG"))
    fld = config_io.read_field(str(tmp_path / "FIELD"))
    assert fld["natoms"] == b2["mxatms"] and np.array_equal(fld["listttm2"], b2["listttm2"]) and np.array_equal(fld["listttm2"], box["listttm2"])
    b2["listttm2"], b2["nttm2"] = fld["listttm2"], fld["nttm2"]                 # the topology comes from FIELD, as in the reference
    h = mb.mbpol_init(box["nttm2"], flags=7)
    a = mb.run_box(h, box, want_stress=True)
    d1, t1 = mb.mbpol_lists(h)
    b = mb.run_box(h, b2, want_stress=True)
    d2, t2 = mb.mbpol_lists(h)
    mb.mbpol_fini(h)
    assert np.array_equal(sort_rows(d1), sort_rows(d2)) and np.array_equal(sort_rows(t1), sort_rows(t2))
    for k in ("xxx", "yyy", "zzz", "vxx"):
        assert np.array_equal(box[k], b2[k]), k                                # the file round trip is exact
    for k in ("eng2", "eng3", "vir2", "vir3"):
        assert abs(a[k] - b[k]) <= 1e-12 * abs(a[k]), k                        # (sums run in a different order from call to call)
    for k in ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp"):
        assert rel_rms(b[k], a[k]) <= 1e-12, k
