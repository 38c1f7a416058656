import sys, numpy as np
sys.path.insert(0, '.')
import mbpol_b200 as mb
from mbpol_b200 import synth, mbnrg
box = synth.make_water_box(27)
h = mb.mbpol_init(box["nttm2"], flags=7)
out = mb.run_box(h, box, want_stress=True)
print("mbpol", out["eng2"], out["eng3"], mb.mbpol_counts(h))
mb.mbpol_fini(h)
ib = synth.make_ion_box(64, 4)
hn = mbnrg.mbnrg_init(ib["nttm2"], ib["ion_code"])
o = mbnrg.run_ion_box(hn, ib)
print("mbnrg", o["eng2"], mbnrg.mbnrg_counts(hn))
mbnrg.mbnrg_fini(hn)
