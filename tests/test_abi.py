"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/mbpol_b200.h declares,
and fails loudly (status code / exception, never a CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mbpol_b200 as mb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mbpol_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mbpol_(?:b200_\w+|[23]b_\w+_))\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = mb.load_library()
    names = declared_symbols()
    assert len(names) >= 20 and "mbpol_b200_forces" in names and "mbpol_3b_poly_" in names
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/mbpol_b200.h but not exported" % n


def test_no_oracle_in_product():
    """The product path must not link, import or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dlpoly2-mbpol-heat_b200")):
        if "/build" in dirpath or "/gen" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() and "mbpol_ref" not in txt, (dirpath, f)      # not even a mention
    import subprocess
    out = subprocess.run(["ldd", mb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "mbpol_ref" not in out


def test_cutoffs_need_no_gpu():
    assert mb.mbpol_cutoffs() == (6.5, 4.5)
    r = C.c_double()
    mb.load_library().mbpol_2b_cutoff_(C.byref(r))
    assert r.value == 6.5


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MBPolError):
        mb.mbpol_init(3 * 64)
    with pytest.raises(mb.MBPolError):
        mb.mbpol_eval_trimers(np.zeros((1, 27)))


def test_init_argument_checks():
    lib = mb.load_library()
    p = C.c_void_p()
    assert lib.mbpol_b200_init(-3, -1, 3, C.byref(p)) == 1       # EINVAL: negative nttm2 (0 waters is a legal, empty system)
    assert lib.mbpol_b200_init(10, -1, 3, C.byref(p)) == 1       # EINVAL: not a multiple of 3
    assert lib.mbpol_b200_fini(None) == 1
