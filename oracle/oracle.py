"""ctypes binding of the CPU oracle (oracle/mbpol_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmbpol_oracle.so")
REF_O2 = os.path.join(HERE, "_ref", "libmbpol_ref.so")
REF_O0 = os.path.join(HERE, "_ref", "libmbpol_ref_O0.so")
REF_NRG = os.path.join(HERE, "_ref", "libmbnrg_ref.so")
ION_CODES = {"f": 1, "cl": 2, "br": 3, "i": 4, "li": 5, "na": 6, "k": 7, "rb": 8, "cs": 9}      # mbnrg.f90:541-565

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the restatement (always) -- building the checker is not using it."""
    srcs = [os.path.join(HERE, f) for f in ("mbpol_oracle.cpp", "mbnrg_oracle.cpp")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", HERE, "libmbpol_oracle.so"])


def build_reference(fast=True):
    """Compile the reference's own evaluator sources in place (only where /root/reference exists)."""
    if not os.path.isdir("/root/reference/SOURCE"):
        return False
    subprocess.check_call(["make", "-C", HERE, "ref-fast" if fast else "ref", "ref-nrg"])
    return True


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self, load_reference=True):
        build()
        self.lib = C.CDLL(LIB)
        L = self.lib
        L.oracle_poly2b.restype = C.c_double
        L.oracle_poly3b.restype = C.c_double
        L.oracle_coef2b.restype = _dp
        L.oracle_coef3b.restype = _dp
        L.oracle_load_reference.argtypes = [C.c_char_p]
        self.ref_path = None
        if load_reference:
            for p in (REF_O2, REF_O0):
                if os.path.exists(p) and L.oracle_load_reference(p.encode()) == 0:
                    self.ref_path = p
                    break

        L.oracle_mbnrg_polyhi.restype = C.c_double
        L.oracle_mbnrg_coef_hi.restype = _dp
        L.oracle_mbnrg_load_reference.argtypes = [C.c_char_p]
        self.have_nrg_reference = bool(load_reference and os.path.exists(REF_NRG) and L.oracle_mbnrg_load_reference(REF_NRG.encode()) == 0)

    @property
    def have_reference(self):
        return self.ref_path is not None

    # ---- MB-nrg ion-water / ion-ion 2B (oracle/mbnrg_oracle.cpp) ----
    def nrg_polyhi(self, a, x):
        a, x, g = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(x, dtype=np.float64), np.zeros(8)
        return self.lib.oracle_mbnrg_polyhi(_p(a), _p(x), _p(g)), g

    def nrg_coef_hi(self, ion):
        return np.ctypeslib.as_array(self.lib.oracle_mbnrg_coef_hi(int(ion)), shape=(429,)).copy()

    def nrg_h2o_ion(self, ion, w, x, ref=False):
        w, x = np.ascontiguousarray(w, dtype=np.float64), np.ascontiguousarray(x, dtype=np.float64)
        E, g1, g2 = C.c_double(), np.zeros(9), np.zeros(3)
        self.lib.oracle_mbnrg_h2o_ion(int(ref), int(ion), _p(w), _p(x), C.byref(E), _p(g1), _p(g2))
        return E.value, g1, g2

    def nrg_ion_ion(self, a, b, w, x, ref=False):
        w, x = np.ascontiguousarray(w, dtype=np.float64), np.ascontiguousarray(x, dtype=np.float64)
        E, g1, g2 = C.c_double(), np.zeros(3), np.zeros(3)
        have = self.lib.oracle_mbnrg_ion_ion(int(ref), int(a), int(b), _p(w), _p(x), C.byref(E), _p(g1), _p(g2))
        return bool(have), E.value, g1, g2

    def nrg_poly3b(self, a, x):
        a, x, g = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(x, dtype=np.float64), np.zeros(21)
        self.lib.oracle_mbnrg_poly3b.restype = C.c_double
        return self.lib.oracle_mbnrg_poly3b(_p(a), _p(x), _p(g)), g

    def nrg_coef_3b(self, ion):
        self.lib.oracle_mbnrg_coef_3b.restype = _dp
        return np.ctypeslib.as_array(self.lib.oracle_mbnrg_coef_3b(int(ion)), shape=(1831,)).copy()

    @property
    def have_nrg_reference_3b(self):
        return bool(self.have_nrg_reference and self.lib.oracle_mbnrg_have_reference_3b())

    def nrg_h2o_h2o_ion(self, ion, w1, w2, x, ref=False):
        """mbnrg_3b_h2o_h2o_<ion>_poly_ (restatement, or the reference object with ref=True)."""
        w1, w2, x = (np.ascontiguousarray(v, dtype=np.float64) for v in (w1, w2, x))
        E, g1, g2, g3 = C.c_double(), np.zeros(9), np.zeros(9), np.zeros(3)
        self.lib.oracle_mbnrg_h2o_h2o_ion(int(ref), int(ion), _p(w1), _p(w2), _p(x), C.byref(E), _p(g1), _p(g2), _p(g3))
        return E.value, g1, g2, g3

    def nrg_forces3(self, box, use_ref=False, rank=0, nranks=1):
        """The 3B ion-water-water part of mbnrg_forces (disabled in the reference build, mbnrg.f90:4) on a synth.make_ion_box dict."""
        n = box["mxatms"]
        f = [np.zeros(n) for _ in range(3)]
        e, v, st, cnt = C.c_double(0.0), C.c_double(0.0), np.zeros(9), np.zeros(2, dtype=np.int64)
        lt = np.ascontiguousarray(box["listttm2"], dtype=np.int32)
        li = np.ascontiguousarray(box["listions"], dtype=np.int32)
        ic = np.ascontiguousarray(box["ion_code"], dtype=np.int32)
        cell = np.ascontiguousarray(box["cell"], dtype=np.float64)
        rc = self.lib.oracle_mbnrg_forces3(int(use_ref), int(rank), int(nranks), int(box["imcon"]), _p(cell), int(box["nttm2"]),
                                           lt.ctypes.data_as(_ip), int(box["nions"]), li.ctypes.data_as(_ip), ic.ctypes.data_as(_ip),
                                           _p(box["xxx"]), _p(box["yyy"]), _p(box["zzz"]), _p(f[0]), _p(f[1]), _p(f[2]),
                                           C.byref(e), C.byref(v), _p(st), cnt.ctypes.data_as(C.POINTER(C.c_longlong)))
        if rc:
            raise RuntimeError("oracle_mbnrg_forces3 failed (%d)" % rc)
        return dict(eng3=e.value, vir3=v.value, stress=st, fxx=f[0], fyy=f[1], fzz=f[2], ntrimers=int(cnt[0]), nshard=int(cnt[1]))

    def nrg_forces(self, box, use_ref=False, rank=0, nranks=1):
        """mbnrg_forces on a box dict of synth.make_ion_box: returns eng2, vir2, stress, forces and the pair counts."""
        n = box["mxatms"]
        f = [np.zeros(n) for _ in range(3)]
        e, v, st, cnt = C.c_double(0.0), C.c_double(0.0), np.zeros(9), np.zeros(3, dtype=np.int64)
        lt = np.ascontiguousarray(box["listttm2"], dtype=np.int32)
        li = np.ascontiguousarray(box["listions"], dtype=np.int32)
        ic = np.ascontiguousarray(box["ion_code"], dtype=np.int32)
        cell = np.ascontiguousarray(box["cell"], dtype=np.float64)
        rc = self.lib.oracle_mbnrg_forces(int(use_ref), int(rank), int(nranks), int(box["imcon"]), _p(cell), int(box["nttm2"]),
                                          lt.ctypes.data_as(_ip), int(box["nions"]), li.ctypes.data_as(_ip), ic.ctypes.data_as(_ip),
                                          _p(box["xxx"]), _p(box["yyy"]), _p(box["zzz"]), _p(f[0]), _p(f[1]), _p(f[2]),
                                          C.byref(e), C.byref(v), _p(st), cnt.ctypes.data_as(C.POINTER(C.c_longlong)))
        if rc:
            raise RuntimeError("oracle_mbnrg_forces failed (%d)" % rc)
        return dict(eng2=e.value, vir2=v.value, stress=st, fxx=f[0], fyy=f[1], fzz=f[2], npairs=int(cnt[0]), nactive=int(cnt[1]),
                    nionion=int(cnt[2]))

    def ref_flags(self):
        if self.ref_path is None:
            return None
        return "g++ -O2 -ffp-contract=off (x86-64 baseline)" if self.ref_path == REF_O2 else \
            "g++ -O2 drivers / -O0 polynomial bodies, -ffp-contract=off"

    def cutoffs(self):
        a, b = C.c_double(), C.c_double()
        self.lib.oracle_cutoffs(C.byref(a), C.byref(b))
        return a.value, b.value

    def coef(self, tag):
        n = 1153 if tag == "2b" else 1163
        p = self.lib.oracle_coef2b() if tag == "2b" else self.lib.oracle_coef3b()
        return np.ctypeslib.as_array(p, (n,)).copy()

    def poly(self, tag, a, x):
        nv = 31 if tag == "2b" else 36
        g = np.zeros(nv)
        f = self.lib.oracle_poly2b if tag == "2b" else self.lib.oracle_poly3b
        a = np.ascontiguousarray(a, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        return f(_p(a), _p(x), _p(g)), g

    def dimer(self, w1, w2, ref=False):
        w1 = np.ascontiguousarray(w1, dtype=np.float64).ravel()
        w2 = np.ascontiguousarray(w2, dtype=np.float64).ravel()
        E = C.c_double()
        g1, g2, fm = np.zeros(9), np.zeros(9), np.zeros(108)
        f = self.lib.oracle_ref_2b if ref else self.lib.oracle_mbpol_2b
        f(_p(w1), _p(w2), C.byref(E), _p(g1), _p(g2), _p(fm))
        return E.value, g1, g2, fm.reshape(6, 6, 3)

    def trimer(self, w1, w2, w3, ref=False):
        w = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (w1, w2, w3)]
        E = C.c_double()
        g1, g2, g3, fm = np.zeros(9), np.zeros(9), np.zeros(9), np.zeros(243)
        f = self.lib.oracle_ref_3b if ref else self.lib.oracle_mbpol_3b
        f(_p(w[0]), _p(w[1]), _p(w[2]), C.byref(E), _p(g1), _p(g2), _p(g3), _p(fm))
        return E.value, g1, g2, g3, fm.reshape(9, 9, 3)

    def image_vectors(self, imcon, cell, dx):
        dx = np.ascontiguousarray(dx, dtype=np.float64).copy()
        cell = np.ascontiguousarray(cell, dtype=np.float64)
        rc = self.lib.oracle_image_vectors(int(imcon), _p(cell), dx.size // 3, _p(dx))
        if rc:
            raise ValueError("imcon %d not supported" % imcon)
        return dx

    def forces(self, box, mode="listed", use_ref=False, nthreads=1, bead_rank=0, bead_size=1, want_stress=True,
               want_fdecomp=True, want_matrix=False, want_lists=True, max_clusters=0):
        """Run the restated mbpol_forces on a `box` dict (see synth.make_water_box); returns a dict of fresh outputs."""
        n = int(box["mxatms"])
        out = {k: np.zeros(n) for k in ("fxx", "fyy", "fzz", "atomic_energy")}
        out["atomic_stress"] = np.zeros(9 * n) if want_stress else None
        out["atomic_fdecomp"] = np.zeros(9 * n) if want_fdecomp else None
        out["force_matrix"] = np.zeros(3 * n * n) if want_matrix else None
        out["stress"] = np.zeros(9)
        sc = [C.c_double(0.0) for _ in range(4)]
        nd, nt = C.c_longlong(0), C.c_longlong(0)
        nw = int(box["nttm2"]) // 3
        dcap = 64 * nw if want_lists else 0
        tcap = 400 * nw if want_lists else 0
        dl = np.zeros(2 * dcap, dtype=np.int32) if want_lists else None
        tl = np.zeros(3 * tcap, dtype=np.int32) if want_lists else None
        timing = np.zeros(3)
        cell = np.ascontiguousarray(box["cell"], dtype=np.float64)
        lst = np.ascontiguousarray(box["listttm2"], dtype=np.int32)
        rc = self.lib.oracle_mbpol_forces(
            C.c_int(0 if mode == "faithful" else 1), C.c_int(int(use_ref)), C.c_int(nthreads), C.c_int(bead_rank),
            C.c_int(bead_size), C.c_int(int(box["imcon"])), _p(cell), C.c_int(int(box["nttm2"])),
            lst.ctypes.data_as(_ip), C.c_int(n), _p(box["xxx"]), _p(box["yyy"]), _p(box["zzz"]),
            _p(out["fxx"]), _p(out["fyy"]), _p(out["fzz"]), C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]),
            C.byref(sc[3]), _p(out["stress"]), _p(out["atomic_energy"]), _p(out["atomic_stress"]),
            _p(out["atomic_fdecomp"]), _p(out["force_matrix"]), C.byref(nd), C.byref(nt),
            None if dl is None else dl.ctypes.data_as(_ip), C.c_longlong(dcap),
            None if tl is None else tl.ctypes.data_as(_ip), C.c_longlong(tcap), C.c_longlong(max_clusters), _p(timing))
        if rc:
            raise RuntimeError("oracle_mbpol_forces failed: %d" % rc)
        out["eng2"], out["vir2"], out["eng3"], out["vir3"] = (s.value for s in sc)
        out["ndimer"], out["ntrimer"] = nd.value, nt.value
        if want_lists:
            assert nd.value <= dcap and nt.value <= tcap, "list capacity exceeded"
            out["dimers"] = dl[:2 * nd.value].reshape(-1, 2).copy()
            out["trimers"] = tl[:3 * nt.value].reshape(-1, 3).copy()
        out["timing"] = {"list_s": timing[0], "t2b_s": timing[1], "t3b_s": timing[2]}
        return out

    def heat_flux(self, n, aeng, astress, vx, vy, vz):
        j1, j2 = np.zeros(3), np.zeros(3)
        self.lib.oracle_heat_flux(C.c_int(n), _p(aeng), _p(astress), _p(vx), _p(vy), _p(vz), _p(j1), _p(j2))
        return j1, j2

    def heat_flux2_matrix(self, n, fmat, x, y, z, vx, vy, vz):
        j = np.zeros(3)
        self.lib.oracle_heat_flux2_matrix(C.c_int(n), _p(fmat), _p(x), _p(y), _p(z), _p(vx), _p(vy), _p(vz), _p(j))
        return j

    def heat_flux2_fdecomp(self, n, afd, vx, vy, vz):
        j = np.zeros(3)
        self.lib.oracle_heat_flux2_fdecomp(C.c_int(n), _p(afd), _p(vx), _p(vy), _p(vz), _p(j))
        return j
