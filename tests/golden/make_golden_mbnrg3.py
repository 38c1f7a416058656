#!/usr/bin/env python3
"""Golden fixtures of the MB-nrg 3B ion-water-water term from the REFERENCE's own evaluator objects (oracle/_ref/libmbnrg_ref.so:
mbnrg-3b-h2o-h2o-<ion>-poly.cpp + poly-3b-h2o-ion-v1x.cpp compiled unmodified by `make -C oracle ref-nrg`).  Run once in the build
container; the output is committed:

  mbnrg3.npz  tri_{ion,w,E,g}: 6 clusters per ion species (9 species) through mbnrg_3b_h2o_h2o_<ion>_poly_; w = (w1[9], w2[9], x[3]),
              g = gradient on the 7 sites (gradient arrays zeroed first, as do_trimer does)
              box_*: the restated 3B driver (oracle/mbnrg_oracle.cpp: mbnrg.f90:354-427 + do_trimer) around the reference evaluators
              for the 208-water + 8-ion synthetic NaCl box (synth.make_ion_box(216, 8)): eng3, vir3, stress, forces, trimer count
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import mbpol_b200  # noqa: E402,F401
from mbpol_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

o = Oracle()
assert o.have_nrg_reference_3b, "build oracle/_ref first (make -C oracle ref-nrg)"
rng = np.random.default_rng(2025)
w0 = np.array([0.00, 0.00, 0.00, 0.9572, 0.00, 0.00, -0.2400, 0.9266, 0.00])


def water(centre):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])
    return ((w0.reshape(3, 3) + rng.normal(scale=0.02, size=(3, 3))) @ R.T + centre).ravel()


out = {k: [] for k in ("tri_ion", "tri_w", "tri_E", "tri_g")}
for ion in range(1, 10):
    for (ra, rb) in ((2.4, 2.5), (2.8, 3.4), (3.2, 4.6), (2.6, 5.8), (4.0, 6.6), (6.2, 6.8)):      # inside and across the 0 -> 7 A switch
        x = rng.normal(size=3)
        da, db = rng.normal(size=3), rng.normal(size=3)
        da /= np.linalg.norm(da)
        db -= 0.3 * da
        db /= np.linalg.norm(db)
        w1, w2 = water(x + ra * da), water(x - rb * db)
        E, g1, g2, g3 = o.nrg_h2o_h2o_ion(ion, w1, w2, x, ref=True)
        out["tri_ion"].append(ion)
        out["tri_w"].append(np.r_[w1, w2, x])
        out["tri_E"].append(E)
        out["tri_g"].append(np.r_[g1, g2, g3])
out = {k: np.array(v) for k, v in out.items()}
box = synth.make_ion_box(216, 8)
r = o.nrg_forces3(box, use_ref=True)
out.update(box_scalars=np.array([r["eng3"], r["vir3"]]), box_stress=r["stress"], box_fxx=r["fxx"], box_fyy=r["fyy"], box_fzz=r["fzz"],
           box_counts=np.array([r["ntrimers"]]))
np.savez_compressed(os.path.join(HERE, "mbnrg3.npz"), **out)
print("wrote mbnrg3.npz:", {k: v.shape for k, v in out.items()}, "E range", out["tri_E"].min(), out["tri_E"].max())
