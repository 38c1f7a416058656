"""Like ab_engines.py but prints the fine phase split (k_trimers alone, gather, incidence lists) per library build."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np
import mbpol_b200 as mb
from mbpol_b200 import synth
box = synth.make_water_box(int(sys.argv[1]))
h = mb.mbpol_init(box["nttm2"], flags=int(sys.argv[2]))
mb.mbpol_upload(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["xxx"], box["yyy"], box["zzz"], box["vxx"], box["vyy"], box["vzz"])
ts = []
for it in range(8):
    mb.mbpol_compute(h)
    if it >= 3: ts.append(mb.mbpol_phase_times(h))
n = int(box["mxatms"]); f = [np.zeros(n) for _ in range(3)]
r = mb.mbpol_download(h, *f)
print(json.dumps({k: float(np.median([t[k] for t in ts])) for k in ts[0]} | {"E": r["eng2"] + r["eng3"], "fsum": float(np.abs(f[0]).sum())}))
''' % ROOT
nw, flags = 32768, 7
res = {}
for rep in range(2):
    for v in sys.argv[1:]:
        label, lib = v.split("=")
        p = subprocess.run([sys.executable, "-c", CHILD, str(nw), str(flags)], env=dict(os.environ, MBPOL_B200_LIB=os.path.join(ROOT, lib)), capture_output=True, text=True, timeout=600)
        if p.returncode: print(label, "FAILED", p.stderr[-300:]); continue
        res.setdefault(label, []).append(json.loads(p.stdout.strip().splitlines()[-1]))
for label, rs in res.items():
    print("%-8s k_trimers %s gather %s inc %.3f k2b %.2f lists %.3f total %s  E %.12e fsum %.9e" % (label, " ".join("%.2f" % r["k_trimers_ms"] for r in rs),
          " ".join("%.2f" % r["gather_ms"] for r in rs), rs[0]["incidence_ms"], rs[0]["k2b_ms"], rs[0]["lists_ms"], " ".join("%.2f" % r["total_ms"] for r in rs), rs[0]["E"], rs[0]["fsum"]))
