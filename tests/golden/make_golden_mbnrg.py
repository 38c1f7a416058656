#!/usr/bin/env python3
"""Golden fixtures of the MB-nrg ion-water / ion-ion 2B path from the REFERENCE's own evaluator objects
(oracle/_ref/libmbnrg_ref.so: mbnrg-2b-h2o-<ion>-poly.cpp, mbnrg-2b-<ion>-<ion>-poly.cpp, poly-2b-*-v1x.cpp compiled
unmodified by `make -C oracle ref-nrg`).  Run once in the build container; the output is committed:

  mbnrg.npz   h2o_ion_{ion,w,x,E,g1,g2}: 8 clusters per ion species (9 species) through mbnrg_2b_h2o_<ion>_poly_
              ion_ion_{a,b,w,x,E,g1,g2}: 4 separations for each of the 29 fitted ion pairs through mbnrg_2b_<a>_<b>_poly_
              box_*: the restated mbnrg_forces driver (oracle/mbnrg_oracle.cpp) around the reference evaluators for the
              208-water + 8-ion synthetic NaCl box (synth.make_ion_box(216, 8)): scalars, stress, forces, pair counts
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
assert o.have_nrg_reference, "build oracle/_ref first (make -C oracle ref-nrg)"
rng = np.random.default_rng(2024)
w0 = np.array([0.00, 0.00, 0.00, 0.9572, 0.00, 0.00, -0.2400, 0.9266, 0.00])
out = {k: [] for k in ("h2o_ion_ion", "h2o_ion_w", "h2o_ion_x", "h2o_ion_E", "h2o_ion_g1", "h2o_ion_g2",
                       "ion_ion_a", "ion_ion_b", "ion_ion_w", "ion_ion_x", "ion_ion_E", "ion_ion_g1", "ion_ion_g2")}
for ion in range(1, 10):
    for r in (2.2, 2.6, 3.0, 3.6, 4.4, 5.4, 5.9, 6.4):          # inside, and across the 5.5 -> 6.5 A switch
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        w = w0 + rng.normal(scale=0.03, size=9) + np.repeat(rng.normal(size=3), 1).tolist() * 3
        x = w[:3] + d * r
        E, g1, g2 = o.nrg_h2o_ion(ion, w, x, ref=True)
        for k, v in zip(("ion", "w", "x", "E", "g1", "g2"), (ion, w, x, E, g1, g2)):
            out["h2o_ion_" + k].append(v)
for a in range(1, 10):
    for b in range(a, 10):
        for r in (2.4, 3.5, 6.0, 7.6):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            y = rng.normal(size=3)
            have, E, g1, g2 = o.nrg_ion_ion(a, b, y + d * r, y, ref=True)
            if have:
                for k, v in zip(("a", "b", "w", "x", "E", "g1", "g2"), (a, b, y + d * r, y, E, g1, g2)):
                    out["ion_ion_" + k].append(v)
out = {k: np.array(v) for k, v in out.items()}
box = synth.make_ion_box(216, 8)
r = o.nrg_forces(box, use_ref=True)
out.update(box_scalars=np.array([r["eng2"], r["vir2"]]), box_stress=r["stress"], box_fxx=r["fxx"], box_fyy=r["fyy"], box_fzz=r["fzz"],
           box_counts=np.array([r["npairs"], r["nactive"], r["nionion"]]))
np.savez_compressed(os.path.join(HERE, "mbnrg.npz"), **out)
print("wrote mbnrg.npz:", {k: v.shape for k, v in out.items()})
