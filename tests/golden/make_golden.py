#!/usr/bin/env python3
"""Generate the golden fixtures from the REFERENCE's own evaluator objects (oracle/_ref, compiled unmodified from
/root/reference/SOURCE by oracle/Makefile).  Run once in the build container; the outputs are committed:

  kat_appendix_a.json   the known-answer vectors of SURVEY.md Appendix A, re-generated and checked bit-for-bit
  clusters.npz          32 dimers + 32 trimers taken from a 64-water box (imaged as do_dimer/do_trimer pass them):
                        w, E, g, force_matrix from mbpol_2b_poly_ / mbpol_3b_poly_
  poly.npz              poly_2b_v6x::eval / poly_3b_v2x::eval for random (x) with the fitted coefficients: E, g
  box64.npz             the restated mbpol_forces driver in FAITHFUL mode (O(N^2)/O(N^3) loops, in-place re-imaging)
                        around the reference evaluators for the 64-water synthetic box: lists, scalars, forces,
                        per-atom energy / stress / force-decomposition tensor
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import mbpol_b200  # noqa: E402
from mbpol_b200 import synth  # noqa: E402
from oracle.oracle import Oracle, REF_O2, REF_O0  # noqa: E402

o = Oracle()
assert o.have_reference, "build oracle/_ref first (make -C oracle ref-fast)"

w1 = [0.00, 0.00, 0.00, 0.9572, 0.00, 0.00, -0.2400, 0.9266, 0.00]
w2 = [2.85, 0.20, -0.10, 3.3500, 0.95, 0.20, 3.4000, -0.5500, -0.3500]
w3 = [1.30, 2.60, 0.40, 1.1000, 3.10, 1.20, 2.1500, 2.9000, 0.0500]
E2, g1, g2, fm2 = o.dimer(w1, w2, ref=True)
E3, h1, h2, h3, fm3 = o.trimer(w1, w2, w3, ref=True)
assert E2 == 5.361320870419604 and E3 == -0.37951111966318934, (E2, E3)   # SURVEY.md Appendix A, bit for bit
kat = {"w1": w1, "w2": w2, "w3": w3,
       "2b": {"E": E2, "g1": g1.tolist(), "g2": g2.tolist(), "fm": fm2.tolist()},
       "3b": {"E": E3, "g1": h1.tolist(), "g2": h2.tolist(), "g3": h3.tolist(), "fm": fm3.tolist()},
       "source": "mbpol_2b_poly_/mbpol_3b_poly_ of the reference, g++ -O2 -ffp-contract=off x86-64 (oracle/Makefile)"}
json.dump(kat, open(os.path.join(HERE, "kat_appendix_a.json"), "w"), indent=1)

box = synth.make_water_box(64)
full = o.forces(box, mode="faithful", use_ref=True)
np.savez_compressed(os.path.join(HERE, "box64.npz"), **{k: v for k, v in full.items() if isinstance(v, np.ndarray)},
                    scalars=np.array([full["eng2"], full["vir2"], full["eng3"], full["vir3"]]))

# imaged clusters exactly as the driver hands them to the evaluators
nw = box["nw"]
mono = np.stack([box["xxx"][:3 * nw], box["yyy"][:3 * nw], box["zzz"][:3 * nw]], 1).reshape(nw, 3, 3)
L = box["L"]
for m in range(nw):
    for hh in (1, 2):
        d = mono[m, hh] - mono[m, 0]
        mono[m, hh] = mono[m, 0] + (d - L * np.round(d / L))
rng = np.random.default_rng(7)
dsel = full["dimers"][rng.choice(len(full["dimers"]), 32, replace=False)]
tsel = full["trimers"][rng.choice(len(full["trimers"]), 32, replace=False)]


def image_to(wa, wb):
    d = wa[0] - wb[0]
    d = d - L * np.round(d / L)
    out = wa - wa[0] + (wb[0] + d)
    return out


dw, dE, dg, dfm = [], [], [], []
for i, j in dsel:
    a, b = image_to(mono[i], mono[j]), mono[j]
    E, ga, gb, fm = o.dimer(a.ravel(), b.ravel(), ref=True)
    dw.append(np.r_[a.ravel(), b.ravel()]); dE.append(E); dg.append(np.r_[ga, gb]); dfm.append(fm)
tw, tE, tg, tfm = [], [], [], []
for i, j, k in tsel:
    a, b, c = image_to(mono[i], mono[j]), mono[j], image_to(mono[k], mono[j])
    E, ga, gb, gc, fm = o.trimer(a.ravel(), b.ravel(), c.ravel(), ref=True)
    tw.append(np.r_[a.ravel(), b.ravel(), c.ravel()]); tE.append(E); tg.append(np.r_[ga, gb, gc]); tfm.append(fm)
np.savez_compressed(os.path.join(HERE, "clusters.npz"), dimer_w=np.array(dw), dimer_E=np.array(dE), dimer_g=np.array(dg),
                    dimer_fm=np.array(dfm), trimer_w=np.array(tw), trimer_E=np.array(tE), trimer_g=np.array(tg),
                    trimer_fm=np.array(tfm))

lib = C.CDLL(REF_O2 if os.path.exists(REF_O2) else REF_O0)
res = {}
for tag, sym, nv in (("2b", "_ZN3x2o11poly_2b_v6x4evalEPKdS2_Pd", 31), ("3b", "_ZN3x2o11poly_3b_v2x4evalEPKdS2_Pd", 36)):
    f = getattr(lib, sym)
    f.restype = C.c_double
    a = o.coef(tag)
    xs = rng.uniform(0.05, 1.2, size=(8, nv))
    Es, gs = [], []
    for x in xs:
        g = np.zeros(nv)
        Es.append(f(a.ctypes.data_as(C.c_void_p), np.ascontiguousarray(x).ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p)))
        gs.append(g)
    res["x" + tag], res["E" + tag], res["g" + tag] = xs, np.array(Es), np.array(gs)
np.savez_compressed(os.path.join(HERE, "poly.npz"), **res)
print("golden fixtures written:", sorted(os.listdir(HERE)))
