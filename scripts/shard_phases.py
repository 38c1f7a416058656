"""Device-time phases of ONE shard of an N-rank run, emulated on a single GPU (mbpol_b200_set_shard: communicator-free shards, uniform
boundaries): what a rank of an 8-GPU step costs apart from the collectives.  usage: shard_phases.py [nranks] [rank ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mbpol_b200 as mb
from mbpol_b200 import synth
nranks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = [int(a) for a in sys.argv[2:]] or [0, nranks // 2]
box = synth.make_water_box(32768)
for rank in ranks:
    h = mb.mbpol_init(box["nttm2"], flags=7)
    mb.mbpol_set_shard(h, rank, nranks)
    mb.mbpol_upload(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["xxx"], box["yyy"], box["zzz"], box["vxx"], box["vyy"], box["vzz"])
    ts = []
    for it in range(8):
        mb.mbpol_compute(h)
        if it >= 3: ts.append(mb.mbpol_phase_times(h))
    c = mb.mbpol_counts(h)
    print("rank %d/%d trimers %d: " % (rank, nranks, c["trimers_shard"]) + " ".join("%s %.3f" % (k[:-3], float(np.median([t[k] for t in ts]))) for k in ts[0] if not k.startswith("host")))
    mb.mbpol_fini(h)
