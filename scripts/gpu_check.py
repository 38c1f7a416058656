"""Quick end-to-end parity check of the CUDA path against the oracle (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mbpol_b200 as mb
from mbpol_b200 import synth
from oracle.oracle import Oracle

o = Oracle()
print("oracle reference:", o.ref_path)
w1 = np.array([0.00,0.00,0.00, 0.9572,0.00,0.00, -0.2400,0.9266,0.00])
w2 = np.array([2.85,0.20,-0.10, 3.3500,0.95,0.20, 3.4000,-0.5500,-0.3500])
w3 = np.array([1.30,2.60,0.40, 1.1000,3.10,1.20, 2.1500,2.9000,0.0500])
E, g1, g2, fm = mb.mbpol_2b_poly(w1, w2)
Eo, g1o, g2o, fmo = o.dimer(w1, w2, ref=o.have_reference)
print("2B KAT  E gpu %.15g oracle %.15g rel %.2e  g %.2e fm %.2e" % (E, Eo, abs(E-Eo)/abs(Eo), np.abs(np.r_[g1-g1o, g2-g2o]).max(), np.abs(fm-fmo)[~np.eye(6,dtype=bool)].max()))
E, g1, g2, g3, fm = mb.mbpol_3b_poly(w1, w2, w3)
Eo, g1o, g2o, g3o, fmo = o.trimer(w1, w2, w3, ref=o.have_reference)
print("3B KAT  E gpu %.15g oracle %.15g rel %.2e  g %.2e fm %.2e" % (E, Eo, abs(E-Eo)/abs(Eo), np.abs(np.r_[g1-g1o, g2-g2o, g3-g3o]).max(), np.abs(fm-fmo)[~np.eye(9,dtype=bool)].max()))

def srt(x): return x[np.lexsort(x.T[::-1])] if len(x) else x
for nw in [int(a) for a in sys.argv[1:]] or [64, 256]:
    box = synth.make_water_box(nw)
    h = mb.mbpol_init(box["nttm2"], flags=7)
    t = time.time(); out = mb.run_box(h, box, want_stress=True); tg = time.time() - t
    t = time.time(); ref = o.forces(box, mode="listed", use_ref=o.have_reference, nthreads=os.cpu_count()); tc = time.time() - t
    d, tr = mb.mbpol_lists(h)
    print("nw %d: dimers %d/%d trimers %d/%d  lists equal: %s %s   gpu call %.3fs (kernels %s)  oracle %.2fs" % (
        nw, len(d), ref["ndimer"], len(tr), ref["ntrimer"], np.array_equal(srt(d), srt(ref["dimers"])), np.array_equal(srt(tr), srt(ref["trimers"])), tg, mb.mbpol_kernel_times(h), tc))
    for k in ("eng2", "vir2", "eng3", "vir3"):
        print("   %-5s gpu %.12e oracle %.12e rel %.2e" % (k, out[k], ref[k], abs(out[k]-ref[k])/abs(ref[k])))
    et = out["eng2"]+out["eng3"]; er = ref["eng2"]+ref["eng3"]
    print("   total E rel %.2e" % (abs(et-er)/abs(er)))
    for k in ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp", "stress"):
        a, b = out[k], ref[k]
        print("   %-15s rel RMS %.2e  max %.2e" % (k, np.sqrt(np.mean((a-b)**2)/np.mean(b**2)), np.abs(a-b).max()/np.abs(b).max()))
    n = box["mxatms"]
    j1, j2 = o.heat_flux(n, ref["atomic_energy"], ref["atomic_stress"], box["vxx"], box["vyy"], box["vzz"])
    j3 = o.heat_flux2_fdecomp(n, ref["atomic_fdecomp"], box["vxx"], box["vyy"], box["vzz"])
    jr = np.r_[j1, j2, j3]
    print("   jflux rel", np.abs(out["jflux"]-jr)/np.abs(jr).max())
    mb.mbpol_fini(h)
