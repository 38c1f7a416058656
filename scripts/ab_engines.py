"""A/B timing of library builds / polynomial engines on one GPU (run under gpurun).

    python scripts/ab_engines.py [--nw 32768] [--reps 3] [--steps 5] label=lib.so:engine[:engine2b] ...

Every variant runs in a fresh process (the engine is read when a handle is created), variants are interleaved `reps` times
and the median kernel times are printed -- builds must be compared interleaved, never from separate runs."""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np
import mbpol_b200 as mb
from mbpol_b200 import synth
nw, steps, flags = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
box = synth.make_water_box(nw)
h = mb.mbpol_init(box["nttm2"], flags=flags)
mb.mbpol_upload(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["xxx"], box["yyy"], box["zzz"], box["vxx"], box["vyy"], box["vzz"])
ts = []
for it in range(steps + 2):
    mb.mbpol_compute(h)
    if it >= 2: ts.append(mb.mbpol_kernel_times(h))
n = int(box["mxatms"])
f = [np.zeros(n) for _ in range(3)]
r = mb.mbpol_download(h, *f)
c = mb.mbpol_counts(h)
med = lambda k: float(np.median([t[k] for t in ts]))
print(json.dumps({"k3b_ms": med("k3b_ms"), "k2b_ms": med("k2b_ms"), "lists_ms": med("lists_ms"), "total_ms": med("total_ms"),
                  "E": r["eng2"] + r["eng3"], "fsum": float(np.abs(f[0]).sum()), "trimers": c["trimers"]}))
''' % ROOT

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nw", type=int, default=32768)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--flags", type=int, default=7)
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    res = {}
    for rep in range(a.reps):
        for v in a.variants:
            label, spec = v.split("=")
            parts = spec.split(":")
            env = dict(os.environ, MBPOL_B200_LIB=os.path.join(ROOT, parts[0]))
            if len(parts) > 1 and parts[1]: env["MBPOL_B200_ENGINE"] = parts[1]
            if len(parts) > 2 and parts[2]: env["MBPOL_B200_ENGINE_2B"] = parts[2]
            p = subprocess.run([sys.executable, "-c", CHILD, str(a.nw), str(a.steps), str(a.flags)], env=env, capture_output=True, text=True, timeout=600)
            if p.returncode:
                print(label, "FAILED", p.stderr[-400:]); continue
            res.setdefault(label, []).append(json.loads(p.stdout.strip().splitlines()[-1]))
    import numpy as np
    for label, rs in res.items():
        print("%-14s k3b %s  k2b %s  lists %.2f  total %.2f   E %.12e fsum %.10e" % (
            label, " ".join("%.2f" % r["k3b_ms"] for r in rs), " ".join("%.2f" % r["k2b_ms"] for r in rs),
            np.median([r["lists_ms"] for r in rs]), np.median([r["total_ms"] for r in rs]), rs[0]["E"], rs[0]["fsum"]))

if __name__ == "__main__":
    main()
