"""mbpol_b200 -- host-side mirror of the reference's MB-pol call surface over libmbpol_b200.so.

The product is the C-ABI library (include/mbpol_b200.h, csrc/); this module is the thin Python binding the
tests and bench.py use, named after the reference's own entry points:

    reference (Fortran, mbpol.f90)                    here
    ------------------------------------------------  ---------------------------------------------
    call mbpol_init(nrite, nttm2)          :52-85     h = mbpol_init(nttm2, flags=..., device=...)
    call mbpol_forces(imcon,cell,nttm2,listttm2,      out = mbpol_forces(h, imcon, cell, nttm2, listttm2,
         wat_monomer_ener,xxx,yyy,zzz,fxx,fyy,fzz,          xxx, yyy, zzz, fxx, fyy, fzz, engsrp, virsrp, stress, ...)
         engsrp,virsrp,stress)             :98-116
    call mbpol_fini()                      :89-94     mbpol_fini(h)
    mbpol_2b_poly_/mbpol_3b_poly_ (C ABI)             mbpol_2b_poly / mbpol_3b_poly (same arguments, arrays in/out)

There is NO CPU fallback: importing works anywhere (so the CPU test-suite can check the ABI), but every compute
entry point raises if the CUDA library cannot be loaded or no sm_100 GPU is present.
(The directory name contains '-', so import it through the `mbpol_b200` shim module at the repository root.)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MBPOL_B200_LIB", os.path.join(HERE, "libmbpol_b200.so"))   # override: A/B builds only
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "mbpol_b200.h")

WANT_ATOMIC_ENERGY, WANT_FDECOMP, WANT_ATOMIC_STRESS = 1, 2, 4
DEFAULT_FLAGS = 3
ENGUNIT = 418.4
FLOP_PER_DIMER, FLOP_PER_TRIMER = 30601, 111818     # counted flops of the reference bodies

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


class MBPolError(RuntimeError):
    pass


def load_library():
    """dlopen libmbpol_b200.so (built by __graft_entry__.build() / csrc/Makefile).  Fails loudly when missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MBPolError("libmbpol_b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.mbpol_b200_last_error.restype = C.c_char_p
        lib.mbpol_b200_stream.restype = C.c_void_p
        lib.mbpol_b200_init.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        for name in ("mbpol_b200_fini", "mbpol_b200_compute", "mbpol_b200_last_error", "mbpol_b200_stream"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.mbpol_b200_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.mbpol_b200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.mbpol_b200_comm_mode.argtypes = [C.c_void_p, C.c_int]
        lib.mbpol_b200_phase_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.mbpol_b200_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mbpol_b200_host_register.argtypes = [C.c_void_p, C.c_size_t]
        lib.mbpol_b200_host_unregister.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _p(a, dtype=np.float64):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, "arrays must be contiguous %s" % dtype
    return a.ctypes.data_as(_dp if dtype == np.float64 else _ip)


class Handle:
    def __init__(self, ptr, nttm2, flags):
        self.ptr, self.nttm2, self.flags = ptr, nttm2, flags


def _check(h, rc, what):
    if rc:
        msg = load_library().mbpol_b200_last_error(h.ptr if h else None)
        raise MBPolError("%s failed with status %d: %s" % (what, rc, msg.decode() if msg else ""))


def mbpol_cutoffs():
    a, b = C.c_double(), C.c_double()
    load_library().mbpol_b200_cutoffs(C.byref(a), C.byref(b))
    return a.value, b.value


def mbpol_init(nttm2, flags=DEFAULT_FLAGS, device=-1):
    """mbpol.f90:52-85.  nttm2 = 3 * number of waters."""
    lib = load_library()
    ptr = C.c_void_p()
    rc = lib.mbpol_b200_init(int(nttm2), int(device), int(flags), C.byref(ptr))
    if rc:
        raise MBPolError("mbpol_b200_init failed with status %d (4 = no sm_100 GPU, 2 = CUDA error)" % rc)
    return Handle(ptr, int(nttm2), int(flags))


def mbpol_init_batched(nttm2, nbeads, flags=DEFAULT_FLAGS, device=-1):
    """Bead-batched handle (PIMD/CMD: nbeads independent replicas of the same topology, multibead.f90:25-93)."""
    lib = load_library()
    ptr = C.c_void_p()
    rc = lib.mbpol_b200_init_batched(C.c_int(int(nttm2)), C.c_int(int(nbeads)), C.c_int(int(device)), C.c_int(int(flags)), C.byref(ptr))
    if rc:
        raise MBPolError("mbpol_b200_init_batched failed with status %d (4 = no sm_100 GPU, 2 = CUDA error)" % rc)
    h = Handle(ptr, int(nttm2), int(flags))
    h.nbeads = int(nbeads)
    return h


def mbpol_forces_batched(h, imcon, cell, nttm2, listttm2, xxx, yyy, zzz, fxx, fyy, fzz, atomic_energy=None, atomic_stress=None,
                         atomic_fdecomp=None, vxx=None, vyy=None, vzz=None):
    """All beads in one call: coordinate / force / per-atom arrays are [nbeads, mxatms] (C order); returns per-bead scalars."""
    nb = h.nbeads
    assert xxx.shape[0] == nb and xxx.flags.c_contiguous
    mxatms = xxx.shape[1]
    sc = [np.zeros(nb) for _ in range(4)]
    stress, jflux = np.zeros((nb, 9)), np.zeros((nb, 9))
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    rc = load_library().mbpol_b200_forces_batched(
        h.ptr, C.c_int(nb), C.c_int(int(imcon)), _p(cell), C.c_int(int(nttm2)), _p(listttm2, np.int32), C.c_int(mxatms),
        _p(xxx), _p(yyy), _p(zzz), _p(fxx), _p(fyy), _p(fzz), _p(sc[0]), _p(sc[1]), _p(sc[2]), _p(sc[3]), _p(stress),
        _p(atomic_energy), _p(atomic_stress), _p(atomic_fdecomp), _p(vxx), _p(vyy), _p(vzz), _p(jflux if vxx is not None else None))
    _check(h, rc, "mbpol_b200_forces_batched")
    return {"eng2": sc[0], "vir2": sc[1], "eng3": sc[2], "vir3": sc[3], "stress": stress, "jflux": jflux}


def mbpol_fini(h):
    """mbpol.f90:89-94."""
    if h.ptr:
        load_library().mbpol_b200_fini(h.ptr)
        h.ptr = None


def mbpol_set_shard(h, rank, nranks):
    _check(h, load_library().mbpol_b200_set_shard(h.ptr, int(rank), int(nranks)), "mbpol_b200_set_shard")


# ---- multi-GPU: NCCL inside the library (include/mbpol_b200.h) --------------------------------------------------------
UNIQUE_ID_BYTES = 128


def mbpol_comm_unique_id():
    """ncclGetUniqueId through the library: call on ONE rank and ship the bytes to the others (MPI_Bcast, torch.distributed, ...)."""
    buf = (C.c_ubyte * UNIQUE_ID_BYTES)()
    rc = load_library().mbpol_b200_comm_unique_id(buf)
    if rc:
        raise MBPolError("mbpol_b200_comm_unique_id failed with status %d" % rc)
    return bytes(buf)


def mbpol_comm_init(h, rank, nranks, uid):
    """The handle joins the communicator (and takes shard `rank` of `nranks`); every later step ends with the in-library reduction."""
    assert len(uid) == UNIQUE_ID_BYTES
    buf = (C.c_ubyte * UNIQUE_ID_BYTES).from_buffer_copy(uid)
    _check(h, load_library().mbpol_b200_comm_init(h.ptr, C.c_int(int(rank)), C.c_int(int(nranks)), buf), "mbpol_b200_comm_init")


def mbpol_comm_mode(h, mode):
    """0 no reduction, 1 all-reduce, 2 reduce-scatter (default with a communicator): rank r returns slice r of the per-atom arrays."""
    _check(h, load_library().mbpol_b200_comm_mode(h.ptr, C.c_int(int(mode))), "mbpol_b200_comm_mode")


def shard_bounds(nw, nranks, weight=None):
    """Water ranges of the ranks: bounds[r] .. bounds[r+1] (uniform, or balanced on the previous step's per-water weights)."""
    b = np.zeros(nranks + 1, dtype=np.int32)
    wp = None
    if weight is not None:
        weight = np.ascontiguousarray(weight, dtype=np.int32)
        assert len(weight) == nw
        wp = weight.ctypes.data_as(_ip)
    lib = load_library()
    lib.mbpol_b200_shard_bounds.restype = None
    lib.mbpol_b200_shard_bounds(C.c_int(int(nw)), C.c_int(int(nranks)), wp, b.ctypes.data_as(_ip))
    return b


def packed_layout(mxatms, nranks, nbeads=1):
    """Host mirror of the packed device buffer: (chunk, scalar offset, total count).  The per-atom part [fxx|fyy|fzz|e|S(9)|W(9)]
    of 22*mxatms doubles is padded to nranks equal even chunks (the unit of the reduce-scatter), then 32 scalars per bead."""
    chunk = (22 * mxatms + nranks - 1) // nranks
    chunk += chunk & 1
    return chunk, chunk * nranks, chunk * nranks + 32 * nbeads


def mbpol_shard(h):
    a, b, o = C.c_int(), C.c_int(), C.c_longlong()
    _check(h, load_library().mbpol_b200_shard(h.ptr, C.byref(a), C.byref(b), C.byref(o)), "mbpol_b200_shard")
    return {"first_water": a.value, "n_waters": b.value, "scalar_offset": o.value}


def mbpol_phase_times(h):
    ms = (C.c_float * 11)()
    _check(h, load_library().mbpol_b200_phase_times(h.ptr, ms, C.c_int(11)), "mbpol_b200_phase_times")
    keys = ("lists_ms", "k2b_ms", "k3b_ms", "total_ms", "k_trimers_ms", "gather_ms", "reduce_ms", "incidence_ms",
            "host_upload_ms", "host_compute_ms", "host_download_ms")
    return {k: float(ms[i]) for i, k in enumerate(keys)}


def host_register(a):
    """Pin a caller array once (cudaHostRegister) so that the per-step copies run at PCIe speed."""
    rc = load_library().mbpol_b200_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes))
    if rc:
        raise MBPolError("mbpol_b200_host_register failed with status %d" % rc)


def host_unregister(a):
    load_library().mbpol_b200_host_unregister(C.c_void_p(a.ctypes.data))


def mbpol_forces(h, imcon, cell, nttm2, listttm2, xxx, yyy, zzz, fxx, fyy, fzz, engsrp=0.0, virsrp=0.0, stress=None,
                 atomic_energy=None, atomic_stress=None, atomic_fdecomp=None, vxx=None, vyy=None, vzz=None):
    """The drop-in for mbpol_forces (mbpol.f90:98-449): accumulates into fxx/fyy/fzz/stress and the per-atom arrays
    in place and returns a dict with the updated engsrp/virsrp, the four partial scalars and the flux partials."""
    lib = load_library()
    sc = [C.c_double(0.0) for _ in range(4)]
    if stress is None:
        stress = np.zeros(9)
    jflux = np.zeros(9)
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    rc = lib.mbpol_b200_forces(h.ptr, C.c_int(int(imcon)), _p(cell), C.c_int(int(nttm2)), _p(listttm2, np.int32),
                               C.c_int(len(xxx)), _p(xxx), _p(yyy), _p(zzz), _p(fxx), _p(fyy), _p(fzz),
                               C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), _p(stress),
                               _p(atomic_energy), _p(atomic_stress), _p(atomic_fdecomp), _p(vxx), _p(vyy), _p(vzz),
                               _p(jflux if vxx is not None else None))
    _check(h, rc, "mbpol_b200_forces")
    eng2, vir2, eng3, vir3 = (s.value for s in sc)
    return {"eng2": eng2, "vir2": vir2, "eng3": eng3, "vir3": vir3, "engsrp": engsrp + eng2 + eng3,
            "virsrp": virsrp + vir2 + vir3, "stress": stress, "jflux": jflux}


def mbpol_upload(h, imcon, cell, nttm2, listttm2, xxx, yyy, zzz, vxx=None, vyy=None, vzz=None):
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    rc = load_library().mbpol_b200_upload(h.ptr, C.c_int(int(imcon)), _p(cell), C.c_int(int(nttm2)), _p(listttm2, np.int32),
                                          C.c_int(len(xxx)), _p(xxx), _p(yyy), _p(zzz), _p(vxx), _p(vyy), _p(vzz))
    _check(h, rc, "mbpol_b200_upload")


def mbpol_compute(h):
    _check(h, load_library().mbpol_b200_compute(h.ptr), "mbpol_b200_compute")


def mbpol_download(h, fxx, fyy, fzz, stress=None, atomic_energy=None, atomic_stress=None, atomic_fdecomp=None):
    sc = [C.c_double(0.0) for _ in range(4)]
    if stress is None:
        stress = np.zeros(9)
    jflux = np.zeros(9)
    rc = load_library().mbpol_b200_download(h.ptr, _p(fxx), _p(fyy), _p(fzz), C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]),
                                            C.byref(sc[3]), _p(stress), _p(atomic_energy), _p(atomic_stress), _p(atomic_fdecomp),
                                            _p(jflux))
    _check(h, rc, "mbpol_b200_download")
    eng2, vir2, eng3, vir3 = (s.value for s in sc)
    return {"eng2": eng2, "vir2": vir2, "eng3": eng3, "vir3": vir3, "stress": stress, "jflux": jflux}


def mbpol_device_outputs(h):
    """(device pointer, element count) of the packed FP64 output buffer -- for an NCCL all-reduce."""
    ptr, cnt = C.c_void_p(), C.c_longlong()
    _check(h, load_library().mbpol_b200_device_outputs(h.ptr, C.byref(ptr), C.byref(cnt)), "mbpol_b200_device_outputs")
    return ptr.value, cnt.value


def mbpol_stream(h):
    return load_library().mbpol_b200_stream(h.ptr)


def mbpol_counts(h):
    v = [C.c_longlong() for _ in range(4)]
    _check(h, load_library().mbpol_b200_counts(h.ptr, *[C.byref(x) for x in v]), "mbpol_b200_counts")
    return {"dimers": v[0].value, "trimers": v[1].value, "dimers_shard": v[2].value, "trimers_shard": v[3].value}


def mbpol_lists(h):
    """This rank's shard of the dimer / trimer lists (the whole lists on a single rank), 0-based water indices."""
    c = mbpol_counts(h)
    nd, nt = c["dimers_shard"], c["trimers_shard"]
    d = np.zeros((max(nd, 1), 2), dtype=np.int32)
    t = np.zeros((max(nt, 1), 3), dtype=np.int32)
    rc = load_library().mbpol_b200_get_lists(h.ptr, _p(d, np.int32), C.c_longlong(nd), _p(t, np.int32), C.c_longlong(nt))
    _check(h, rc, "mbpol_b200_get_lists")
    return d[:nd], t[:nt]


def mbpol_kernel_times(h):
    ms = (C.c_float * 4)()
    n = C.c_longlong()
    _check(h, load_library().mbpol_b200_kernel_times(h.ptr, ms, C.byref(n)), "mbpol_b200_kernel_times")
    return {"lists_ms": ms[0], "k2b_ms": ms[1], "k3b_ms": ms[2], "total_ms": ms[3], "launches": n.value}


def measure_fp64_peak():
    """Measured DFMA throughput of the current GPU in TFLOP/s (2 flop per DFMA)."""
    t, ms = C.c_double(), C.c_double()
    rc = load_library().mbpol_b200_measure_fp64_peak(C.byref(t), C.byref(ms))
    if rc:
        raise MBPolError("mbpol_b200_measure_fp64_peak failed with status %d" % rc)
    return t.value


def mbpol_eval_dimers(w, want_fm=True):
    """Batched mbpol_2b_poly_: w [n,18] imaged coordinates -> E[n], g[n,18], fm[n,6,6,3] (kcal/mol)."""
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 18)
    n = len(w)
    E, g = np.zeros(n), np.zeros((n, 18))
    fm = np.zeros((n, 108)) if want_fm else None
    rc = load_library().mbpol_b200_eval_dimers(C.c_int(n), _p(w), _p(E), _p(g), _p(fm))
    if rc:
        raise MBPolError("mbpol_b200_eval_dimers failed with status %d" % rc)
    return E, g, (fm.reshape(n, 6, 6, 3) if want_fm else None)


def mbpol_eval_trimers(w, want_fm=True):
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 27)
    n = len(w)
    E, g = np.zeros(n), np.zeros((n, 27))
    fm = np.zeros((n, 243)) if want_fm else None
    rc = load_library().mbpol_b200_eval_trimers(C.c_int(n), _p(w), _p(E), _p(g), _p(fm))
    if rc:
        raise MBPolError("mbpol_b200_eval_trimers failed with status %d" % rc)
    return E, g, (fm.reshape(n, 9, 9, 3) if want_fm else None)


def mbpol_2b_poly(w1, w2):
    """The legacy symbol itself, mbpol-2b-poly.cpp:1888-1893."""
    w1 = np.ascontiguousarray(w1, dtype=np.float64); w2 = np.ascontiguousarray(w2, dtype=np.float64)
    E = C.c_double()
    g1, g2, fm = np.zeros(9), np.zeros(9), np.zeros(108)
    load_library().mbpol_2b_poly_(_p(w1), _p(w2), C.byref(E), _p(g1), _p(g2), _p(fm))
    return E.value, g1, g2, fm.reshape(6, 6, 3)


def mbpol_3b_poly(w1, w2, w3):
    """The legacy symbol itself, mbpol-3b-poly.cpp:1526-1531 (gradients are accumulated into zeroed arrays)."""
    w = [np.ascontiguousarray(v, dtype=np.float64) for v in (w1, w2, w3)]
    E = C.c_double()
    g1, g2, g3, fm = np.zeros(9), np.zeros(9), np.zeros(9), np.zeros(243)
    load_library().mbpol_3b_poly_(_p(w[0]), _p(w[1]), _p(w[2]), C.byref(E), _p(g1), _p(g2), _p(g3), _p(fm))
    return E.value, g1, g2, g3, fm.reshape(9, 9, 3)


def run_box(h, box, want_stress=False, want_fdecomp=True, with_velocities=True):
    """Convenience for tests/bench: one mbpol_forces call on a synth box with fresh (zeroed) accumulators."""
    n = int(box["mxatms"])
    out = {k: np.zeros(n) for k in ("fxx", "fyy", "fzz", "atomic_energy")}
    out["atomic_stress"] = np.zeros(9 * n) if want_stress else None
    out["atomic_fdecomp"] = np.zeros(9 * n) if want_fdecomp else None
    v = (box["vxx"], box["vyy"], box["vzz"]) if with_velocities else (None, None, None)
    r = mbpol_forces(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["xxx"], box["yyy"], box["zzz"],
                     out["fxx"], out["fyy"], out["fzz"], stress=np.zeros(9), atomic_energy=out["atomic_energy"],
                     atomic_stress=out["atomic_stress"], atomic_fdecomp=out["atomic_fdecomp"], vxx=v[0], vyy=v[1], vzz=v[2])
    out.update(r)
    return out
