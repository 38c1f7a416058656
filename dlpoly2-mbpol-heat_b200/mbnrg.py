"""Host-side mirror of the reference's MB-nrg call surface (mbnrg.f90) over libmbpol_b200.so (include/mbnrg_b200.h).

    reference (Fortran, mbnrg.f90)                                   here
    ---------------------------------------------------------------  ----------------------------------------------------
    call mbnrg_init(nrite,nttm2,nions,mxtmls,key,nummols)   :40-157  h = mbnrg_init(nttm2, ion_code)      (ion_code = the
                                                                         key of every ion in list order, selectpoly :68)
    call mbnrg_forces(imcon,cell,nttm2,listttm2,nions,listions,      r = mbnrg_forces(h, imcon, cell, nttm2, listttm2, nions,
         xxx,yyy,zzz,fxx,fyy,fzz,engsrp,virsrp,stress)     :170-485       listions, xxx, yyy, zzz, fxx, fyy, fzz, stress)
    call mbnrg_fini()                                      :161-166  mbnrg_fini(h)

No CPU fallback: every entry point raises when the CUDA library or an sm_100 GPU is missing.
"""
import ctypes as C

import numpy as np

from . import MBPolError, _p, load_library

ION_CODES = {"f": 1, "cl": 2, "br": 3, "i": 4, "li": 5, "na": 6, "k": 7, "rb": 8, "cs": 9}      # mbnrg.f90:541-565
_ip = C.POINTER(C.c_int)


class NrgHandle:
    def __init__(self, ptr, nttm2, nions):
        self.ptr, self.nttm2, self.nions = ptr, nttm2, nions


def _lib():
    lib = load_library()
    if not getattr(lib, "_nrg_ready", False):
        lib.mbnrg_b200_last_error.restype = C.c_char_p
        lib.mbnrg_b200_last_error.argtypes = [C.c_void_p]
        lib.mbnrg_b200_init.argtypes = [C.c_int, C.c_int, _ip, C.c_int, C.POINTER(C.c_void_p)]
        lib.mbnrg_b200_fini.argtypes = [C.c_void_p]
        lib.mbnrg_b200_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.mbnrg_b200_counts.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        lib.mbnrg_b200_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_float), _ip]
        lib.mbnrg_b200_counts_3b.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_float)]
        lib._nrg_ready = True
    return lib


def _check(h, rc, what):
    if rc:
        msg = _lib().mbnrg_b200_last_error(h.ptr if h is not None and h.ptr else None)
        raise MBPolError("%s failed with status %d: %s" % (what, rc, msg.decode() if msg else ""))


def mbnrg_init(nttm2, ion_code, device=-1):
    ic = np.ascontiguousarray(ion_code, dtype=np.int32)
    ptr = C.c_void_p()
    rc = _lib().mbnrg_b200_init(int(nttm2), len(ic), ic.ctypes.data_as(_ip), int(device), C.byref(ptr))
    h = NrgHandle(ptr, int(nttm2), len(ic))
    if rc:
        try:
            _check(h, rc, "mbnrg_b200_init")
        finally:
            if ptr:
                _lib().mbnrg_b200_fini(ptr)
    return h


def mbnrg_fini(h):
    if h.ptr:
        _check(None, _lib().mbnrg_b200_fini(h.ptr), "mbnrg_b200_fini")
        h.ptr = None


def mbnrg_set_shard(h, rank, nranks):
    _check(h, _lib().mbnrg_b200_set_shard(h.ptr, int(rank), int(nranks)), "mbnrg_b200_set_shard")


def mbnrg_forces(h, imcon, cell, nttm2, listttm2, nions, listions, xxx, yyy, zzz, fxx, fyy, fzz, stress, engsrp=0.0, virsrp=0.0):
    """One call of mbnrg_forces: forces and stress are accumulated in place; returns the updated engsrp / virsrp and the
    2B parts (the reference adds eng2 + eng3 with eng3 == 0: the 3B term is compiled out, mbnrg.f90:4)."""
    lt = np.ascontiguousarray(listttm2, dtype=np.int32)
    li = np.ascontiguousarray(listions, dtype=np.int32)
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    e, v = C.c_double(0.0), C.c_double(0.0)
    rc = _lib().mbnrg_b200_forces(h.ptr, int(imcon), _p(cell), int(nttm2), lt.ctypes.data_as(_ip), int(nions), li.ctypes.data_as(_ip),
                                  len(xxx), _p(xxx), _p(yyy), _p(zzz), _p(fxx), _p(fyy), _p(fzz), C.byref(e), C.byref(v), _p(stress))
    _check(h, rc, "mbnrg_b200_forces")
    return dict(eng2=e.value, vir2=v.value, engsrp=engsrp + e.value, virsrp=virsrp + v.value)


def mbnrg_counts(h):
    c = (C.c_longlong * 3)()
    _check(h, _lib().mbnrg_b200_counts(h.ptr, c), "mbnrg_b200_counts")
    return dict(npairs=int(c[0]), nactive=int(c[1]), nionion=int(c[2]))


def mbnrg_kernel_times(h):
    ms, n = (C.c_float * 4)(), C.c_int()
    _check(h, _lib().mbnrg_b200_kernel_times(h.ptr, ms, C.byref(n)), "mbnrg_b200_kernel_times")
    return dict(pairs_ms=ms[0], h2o_ion_ms=ms[1], ion_ion_ms=ms[2], total_ms=ms[3], launches=n.value)


def mbnrg_eval_h2o_ion(h, ion, w, x):
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 9)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    n = len(w)
    E, g1, g2 = np.zeros(n), np.zeros((n, 9)), np.zeros((n, 3))
    _check(h, _lib().mbnrg_b200_eval_h2o_ion(h.ptr, int(ion), n, _p(w), _p(x), _p(E), _p(g1), _p(g2)), "mbnrg_b200_eval_h2o_ion")
    return E, g1, g2


def mbnrg_eval_ion_ion(h, a, b, w, x):
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 3)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    n = len(w)
    E, g1, g2 = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
    _check(h, _lib().mbnrg_b200_eval_ion_ion(h.ptr, int(a), int(b), n, _p(w), _p(x), _p(E), _p(g1), _p(g2)), "mbnrg_b200_eval_ion_ion")
    return E, g1, g2


def run_ion_box(h, box):
    """Convenience for tests/bench: one mbnrg_forces call on a synth.make_ion_box box with zeroed accumulators."""
    n = int(box["mxatms"])
    out = {k: np.zeros(n) for k in ("fxx", "fyy", "fzz")}
    out["stress"] = np.zeros(9)
    out.update(mbnrg_forces(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["nions"], box["listions"], box["xxx"],
                            box["yyy"], box["zzz"], out["fxx"], out["fyy"], out["fzz"], out["stress"]))
    return out


# ---- 3B ion-water-water (optional: compiled out in the reference build, mbnrg.f90:4) -----------------------------------------
def mbnrg_forces_3b(h, imcon, cell, nttm2, listttm2, nions, listions, xxx, yyy, zzz, fxx, fyy, fzz, stress, engsrp=0.0, virsrp=0.0):
    """The 3B part of mbnrg_forces (mbnrg.f90:354-427): accumulates into fxx/fyy/fzz/stress, returns eng3 / vir3 and the updated sums."""
    e, v = C.c_double(0.0), C.c_double(0.0)
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    rc = _lib().mbnrg_b200_forces_3b(h.ptr, C.c_int(int(imcon)), _p(cell), C.c_int(int(nttm2)), _p(listttm2, np.int32), C.c_int(int(nions)),
                                     _p(listions, np.int32), C.c_int(len(xxx)), _p(xxx), _p(yyy), _p(zzz), _p(fxx), _p(fyy), _p(fzz),
                                     C.byref(e), C.byref(v), _p(stress))
    _check(h, rc, "mbnrg_b200_forces_3b")
    return {"eng3": e.value, "vir3": v.value, "engsrp": engsrp + e.value, "virsrp": virsrp + v.value, "stress": stress}


def mbnrg_counts_3b(h):
    c, ms = (C.c_longlong * 2)(), (C.c_float * 3)()
    _check(h, _lib().mbnrg_b200_counts_3b(h.ptr, c, ms), "mbnrg_b200_counts_3b")
    return {"ntrimers": int(c[0]), "nshard": int(c[1]), "list_ms": float(ms[0]), "eval_ms": float(ms[1]), "total_ms": float(ms[2])}


def mbnrg_eval_h2o_h2o_ion(h, ion, w):
    """Batched mbnrg_3b_h2o_h2o_<ion>_poly_: w [n, 21] = (w1, w2, x) -> E[n] (kcal/mol), g[n, 21]."""
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 21)
    n = len(w)
    E, g = np.zeros(n), np.zeros((n, 21))
    _check(h, _lib().mbnrg_b200_eval_h2o_h2o_ion(h.ptr, C.c_int(int(ion)), C.c_int(n), _p(w), _p(E), _p(g)), "mbnrg_b200_eval_h2o_h2o_ion")
    return E, g


def run_ion_box_3b(h, box):
    n = int(box["mxatms"])
    out = {k: np.zeros(n) for k in ("fxx", "fyy", "fzz")}
    st = np.zeros(9)
    r = mbnrg_forces_3b(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["nions"], box["listions"], box["xxx"], box["yyy"],
                        box["zzz"], out["fxx"], out["fyy"], out["fzz"], st)
    out.update(r)
    return out
