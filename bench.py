#!/usr/bin/env python3
"""bench.py -- MB-pol 2B+3B force + per-atom heat-flux decomposition step on B200(s), next to the CPU reference.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU evaluators on the host cores)

One "step" = one complete mbpol_forces-equivalent evaluation (mbpol.f90:98-449) of a synthetic water box:
GPU cell-list build of the dimer/trimer lists, 2B and 3B polynomial energies + gradients, scatter of forces,
per-atom energy and force-decomposition tensor, heat-flux contractions, and (N>1) one NCCL all-reduce of the packed
output buffer.  `value` = accepted trimers of the box / step time with inputs resident in HBM (CUDA events on the
library's stream, max over ranks); `e2e` = the same through the public mbpol_forces call with host buffers
(H2D of coordinates/velocities and D2H of forces + per-atom arrays inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mbpol_2b3b_trimers_per_s"
UNIT = "trimers/s"


# ------------------------------------------------------------------------------------------------------------
# host-side helpers shared with the CPU tests (tests/test_shard.py)
# ------------------------------------------------------------------------------------------------------------
def shard_range(total, rank, nranks):
    """Contiguous shard [lo, hi) of a work list -- identical to the rule inside mbpol_b200_compute."""
    return total * rank // nranks, total * (rank + 1) // nranks


def pack_outputs(out, mxatms, nranks=1):
    """Host mirror of the packed device buffer layout (include/mbpol_b200.h): per-atom arrays, padding to nranks chunks, 32 scalars."""
    n = mxatms
    chunk = (22 * n + nranks - 1) // nranks
    chunk += chunk & 1
    buf = np.zeros(chunk * nranks + 32)
    buf[0:n], buf[n:2 * n], buf[2 * n:3 * n] = out["fxx"], out["fyy"], out["fzz"]
    buf[3 * n:4 * n] = out["atomic_energy"]
    if out.get("atomic_stress") is not None:
        buf[4 * n:13 * n] = out["atomic_stress"]
    if out.get("atomic_fdecomp") is not None:
        buf[13 * n:22 * n] = out["atomic_fdecomp"]
    s = buf[chunk * nranks:]
    s[0], s[1], s[2], s[3] = out["eng2"], out["vir2"], out["eng3"], out["vir3"]
    s[4:13] = out["stress"]
    return buf


def allreduce_packed(t):
    """The single collective of a multi-GPU step: sum the packed per-rank partial outputs in place."""
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class DeviceArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        load = [s for s in sm if mx and s > 0.5 * mx[0]] or sm
        def col(k):
            v = []
            for r in self.rows:
                try:
                    v.append(float(r[k]))
                except (ValueError, IndexError):
                    pass
            return float(np.median(v)) if v else None
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "mem_mhz": col(7), "power_w": col(2)}


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU evaluators (oracle/_ref) under the restated list-driven driver
# ------------------------------------------------------------------------------------------------------------
def reference_sample(box, threads, n_trimers_sample, flags=7):
    """One bounded CPU step: S trimers plus the proportional number of dimers.  Returns (seconds, trimers, info)."""
    from oracle.oracle import Oracle
    o = Oracle()
    kind = "reference" if o.have_reference else "port"
    out = o.forces(box, mode="listed", use_ref=o.have_reference, nthreads=threads, want_stress=bool(flags & 4), want_fdecomp=bool(flags & 2),
                   want_lists=False, max_clusters=n_trimers_sample)
    nt, nd = out["ntrimer"], out["ndimer"]
    st, sd = min(nt, n_trimers_sample), min(nd, n_trimers_sample)
    t3, t2 = out["timing"]["t3b_s"], out["timing"]["t2b_s"]
    # a proportional step evaluates st trimers and st*nd/nt dimers
    t_step = t3 + t2 * (st * nd / max(nt, 1)) / max(sd, 1)
    info = {"kind": kind, "flags": o.ref_flags() or "restated evaluators, g++ -O2 -ffp-contract=off",
            "us_per_trimer_per_thread": 1e6 * t3 * threads / max(st, 1), "us_per_dimer_per_thread": 1e6 * t2 * threads / max(sd, 1),
            "list_s_excluded": out["timing"]["list_s"], "n_dimers": nd, "n_trimers": nt}
    return t_step, st, info


def run_reference(args, box, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # default: the WHOLE box per step (5-7 s at 32 768 H2O on 16 threads), so that ms_per_step is measured, not extrapolated, and the
    # per-dimer / per-trimer figures come from 0.5 M / 2.3 M clusters; --cpu-sample bounds the step for bigger boxes
    sample = args.cpu_sample or (1 << 40)
    vals = []
    info = {}
    for it in range(args.warmup + args.steps):
        t, st, info = reference_sample(box, threads, sample, args.flags)
        if it >= args.warmup:
            vals.append((t, st))
    tsum = sum(v[0] for v in vals)
    tri = sum(v[1] for v in vals)
    value = tri / tsum
    sample_txt = ("the whole %d-H2O box per step: %d dimers + %d trimers (lists excluded)" % (box["nw"], info["n_dimers"], info["n_trimers"])
                  if vals[0][1] >= info["n_trimers"] else
                  "%d trimers + proportional dimers of the %d-H2O box per step (lists excluded)" % (vals[0][1], box["nw"]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * info["n_trimers"] / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(box, args, info["n_dimers"], info["n_trimers"]),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": info["kind"], "sample": sample_txt,
                             "flags": info["flags"], "us_per_trimer_per_thread": info["us_per_trimer_per_thread"],
                             "us_per_dimer_per_thread": info["us_per_dimer_per_thread"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's own evaluator objects (oracle/_ref) under the restated list-driven driver, all host threads; the "
                    "reference's own O(N^3) enumeration (mbpol.f90:306-374) is NOT included -- the evaluators are driven from a neighbour list"}
    print(json.dumps(line))


def workload_config(box, args, nd, nt):
    return {"workload": "%d H2O MB-pol box (cubic, L=%.2f A, seed %d): 2B+3B energies, forces, per-atom energy and "
                        "force-decomposition tensor%s, heat-flux contractions" % (box["nw"], box["L"], box["seed"], " and per-atom stress" if args.flags & 4 else ""),
            "n_waters": box["nw"], "n_dimers": int(nd), "n_trimers": int(nt), "flags": args.flags,
            "parallelism": "water-range shards x%d (each rank builds and evaluates its own lists), replicated coordinates, in-library NCCL reduce-scatter" % args.gpus,
            "host_threads": os.cpu_count(),
            "l2": "flushed between timed steps (256 MiB memset)"}


def run_beads(args, mb, synth):
    """BASELINE config 5: nbeads replicas of an nw-water box in ONE batched call (PIMD/CMD, multibead.f90:25-93)."""
    import torch
    nb = args.beads
    boxes = [synth.make_water_box(args.nw, seed=1005 + b) for b in range(nb)]
    n = boxes[0]["mxatms"]
    stack = lambda k: np.ascontiguousarray(np.stack([b[k] for b in boxes]))
    X, Y, Z, VX, VY, VZ = (stack(k) for k in ("xxx", "yyy", "zzz", "vxx", "vyy", "vzz"))
    F = [np.zeros((nb, n)) for _ in range(3)]
    ae = np.zeros((nb, n))
    afd = np.zeros((nb, 9 * n)) if args.flags & 2 else None
    h = mb.mbpol_init_batched(boxes[0]["nttm2"], nb, flags=args.flags)
    call = lambda: mb.mbpol_forces_batched(h, 1, boxes[0]["cell"], boxes[0]["nttm2"], boxes[0]["listttm2"], X, Y, Z, F[0], F[1], F[2],
                                           atomic_energy=ae, atomic_fdecomp=afd, vxx=VX, vyy=VY, vzz=VZ)
    for _ in range(max(args.warmup, 3)):
        call()
    t, dev = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        call()
        t.append(time.perf_counter() - t0)
        dev.append(mb.mbpol_kernel_times(h))
    c = mb.mbpol_counts(h)
    ms_dev = float(np.mean([d["total_ms"] for d in dev]))
    line = {"metric": METRIC, "value": c["trimers"] / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d beads x %d H2O MB-pol, bead-batched 2B+3B in one call" % (nb, args.nw), "n_dimers": c["dimers"],
                       "n_trimers": c["trimers"], "flags": args.flags},
            "e2e": {"value": c["trimers"] / float(np.mean(t)), "unit": UNIT, "ms_per_step": 1e3 * float(np.mean(t)),
                    "h2d_bytes_per_step": 8 * nb * n * 6, "d2h_bytes_per_step": 8 * nb * (n * (4 + (9 if afd is not None else 0)) + 32)},
            "gpu_launches": int(dev[-1]["launches"]) * args.steps,
            "phases_ms": {k: float(np.mean([q[k] for q in dev])) for k in ("lists_ms", "k2b_ms", "k3b_ms", "total_ms")}}
    print(json.dumps(line))
    mb.mbpol_fini(h)


def run_ions(args, mb, synth):
    """BASELINE config 4 (next-row f1): NaCl(aq), --nw waters + --ions Na/Cl through mbnrg_forces (ion-water + ion-ion 2B; the
    3B term is compiled out in the reference, mbnrg.f90:4).  A "step" is one mbnrg_forces call; the unit is evaluated ion-water
    pairs per second.  The CPU baseline is the same call through the oracle driver around the reference's evaluator objects."""
    from mbpol_b200 import mbnrg
    box = synth.make_ion_box(args.nw, args.ions)
    if args.impl == "reference":        # the reference's own evaluator objects (or the oracle port) on one host core: the
        from oracle.oracle import Oracle     # reference path is serial per rank (mbnrg.f90:255-337)
        o = Oracle()
        for _ in range(args.warmup):
            o.nrg_forces(box, use_ref=o.have_nrg_reference)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = o.nrg_forces(box, use_ref=o.have_nrg_reference)
            r3 = o.nrg_forces3(box, use_ref=o.have_nrg_reference_3b)      # the 3B term BASELINE config 4 names (off in the reference build)
        dt = (time.perf_counter() - t0) / args.steps
        v = r["nactive"] / dt
        print(json.dumps({"impl": "reference", "metric": "mbnrg_ion_water_pairs_per_second", "value": v, "unit": "pairs/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "NaCl(aq) MB-nrg 2B + 3B: %d H2O + %d Na/Cl" % (box["nw"], box["nions"]),
                                     "pairs_within_10A": r["npairs"], "pairs_evaluated": r["nactive"], "ion_ion_pairs": r["nionion"],
                                     "trimers_3b": r3["ntrimers"]},
                          "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": 1, "kind": "reference" if o.have_nrg_reference else "port",
                                           "sample": "the whole step"},
                          "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    h = mbnrg.mbnrg_init(box["nttm2"], box["ion_code"])
    n = box["mxatms"]
    F = [np.zeros(n) for _ in range(3)]
    st = np.zeros(9)
    def call():      # the 2B call and the optional 3B call, as mbnrg_forces does with ENABLE_X2B and ENABLE_X3B (mbnrg.f90:255-427)
        mbnrg.mbnrg_forces(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["nions"], box["listions"],
                           box["xxx"], box["yyy"], box["zzz"], F[0], F[1], F[2], st)
        mbnrg.mbnrg_forces_3b(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], box["nions"], box["listions"],
                              box["xxx"], box["yyy"], box["zzz"], F[0], F[1], F[2], st)
    for _ in range(max(args.warmup, 3)):
        call()
    t, dev = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        call()
        t.append(time.perf_counter() - t0)
        dev.append(dict(mbnrg.mbnrg_kernel_times(h), **{"k3b_" + k: v for k, v in mbnrg.mbnrg_counts_3b(h).items() if k.endswith("_ms")}))
    c = mbnrg.mbnrg_counts(h)
    c3 = mbnrg.mbnrg_counts_3b(h)
    ms_dev = float(np.mean([d["total_ms"] + d["k3b_total_ms"] for d in dev]))
    ms_k = float(np.mean([d["h2o_ion_ms"] for d in dev]))
    flop_pair = 3888                                     # multiplies + adds of the reference polynomial body per evaluated pair
    ms_3b = float(np.mean([d["k3b_eval_ms"] for d in dev]))
    line = {"metric": "mbnrg_ion_water_pairs_per_second", "value": c["nactive"] / (ms_dev * 1e-3), "unit": "pairs/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "NaCl(aq) MB-nrg 2B + 3B: %d H2O + %d Na/Cl, ion-water + ion-ion + ion-water-water polynomials (the 3B term is "
                                   "compiled out in the reference build, mbnrg.f90:4)" % (box["nw"], box["nions"]),
                       "pairs_within_10A": c["npairs"], "pairs_evaluated": c["nactive"], "ion_ion_pairs": c["nionion"], "trimers_3b": c3["ntrimers"],
                       "l2": "inputs re-uploaded every step"},
            "e2e": {"value": c["nactive"] / float(np.mean(t)), "unit": "pairs/s", "ms_per_step": 1e3 * float(np.mean(t)),
                    "h2d_bytes_per_step": 8 * 3 * n + 4 * (box["nttm2"] + box["nions"]), "d2h_bytes_per_step": 8 * (3 * n + 16)},
            "gpu_launches": (int(dev[-1]["launches"]) + 4) * args.steps,
            "phases_ms": {k: float(np.mean([q[k] for q in dev])) for k in ("pairs_ms", "h2o_ion_ms", "ion_ion_ms", "total_ms", "k3b_list_ms", "k3b_eval_ms", "k3b_total_ms")},
            "roofline": {"bound": "latency", "achieved": c["nactive"] * flop_pair / (ms_k * 1e-3) / 1e12, "peak": None, "unit": "TFLOP/s",
                         "frac": None, "traffic": None,
                         "note": "%d pairs x 3888 flops = %.1f MFLOP per launch: %d CTAs of 128 threads cannot fill 148 SMs; the kernel time is "
                                 "launch + one dependent pass through a 2.4k-instruction body.  3B: %d trimers x 25538 FP64 instructions of our "
                                 "schedule in %.3f ms = %.2f TFLOP/s (2 flop per instruction)" % (
                                     c["nactive"], c["nactive"] * flop_pair / 1e6, (c["nactive"] + 127) // 128, c3["ntrimers"], ms_3b,
                                     2 * 25538 * c3["ntrimers"] / (ms_3b * 1e-3) / 1e12)}}
    if not args.no_cpu_baseline:
        from oracle.oracle import Oracle
        o = Oracle()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 5.0:
            r = o.nrg_forces(box, use_ref=o.have_nrg_reference)
            o.nrg_forces3(box, use_ref=o.have_nrg_reference_3b)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = {"value": r["nactive"] / dt, "unit": "pairs/s", "cores": 1, "kind": "reference" if o.have_nrg_reference else "port",
                                "sample": "the whole step (%d calls of the oracle driver around the reference evaluator objects, %.1f ms each)"
                                          % (reps, 1e3 * dt)}
    print(json.dumps(line))
    mbnrg.mbnrg_fini(h)


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nw", type=int, default=32768, help="waters in the synthetic box (BASELINE configs: 4096, 32768)")
    ap.add_argument("--flags", type=int, default=7,
                    help="1 atomic energy | 2 force decomposition | 4 atomic stress (HEAT_STRESS); default all three: forces + per-atom energy, "
                         "W and S (the north_star target); the reference default build corresponds to 3")
    ap.add_argument("--cpu-sample", type=int, default=0, help="trimers per CPU-baseline sample (default 4000 per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default=None, help="read the box from a DL_POLY CONFIG file instead of the synthetic generator")
    ap.add_argument("--beads", type=int, default=1, help="bead-batched evaluation (BASELINE config 5: --beads 32 --nw 512), 1 GPU")
    ap.add_argument("--ions", type=int, default=0, help="MB-nrg NaCl(aq) workload (BASELINE config 4: --nw 4096 --ions 64), 1 GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import mbpol_b200 as mb
    from mbpol_b200 import synth
    if args.ions > 0:
        return run_ions(args, mb, synth)
    if args.config:
        from mbpol_b200 import config_io
        box = config_io.read_config(args.config)
        args.nw = box["nw"]
    else:
        box = synth.make_water_box(args.nw)
    if args.beads > 1:
        return run_beads(args, mb, synth)

    if args.impl == "reference":
        run_reference(args, box, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = box["mxatms"]
    h = mb.mbpol_init(box["nttm2"], flags=args.flags, device=local)
    if world > 1:      # the library owns the communicator of the data path: unique id from rank 0, shipped with torch.distributed
        uid = torch.zeros(mb.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(mb.mbpol_comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        mb.mbpol_comm_init(h, rank, world, bytes(uid.cpu().numpy().tobytes()))
    stream = torch.cuda.ExternalStream(mb.mbpol_stream(h), device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    want_e, want_w, want_s = bool(args.flags & 1), bool(args.flags & 2), bool(args.flags & 4)

    def pinned(a):      # the caller's arrays live in pinned host memory (registered once, as a host code would do at start-up)
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        v = t.numpy()
        v[...] = a
        keep.append(t)
        return v
    keep = []
    X, Y, Z, VX, VY, VZ = (pinned(box[k]) for k in ("xxx", "yyy", "zzz", "vxx", "vyy", "vzz"))
    vel = (VX, VY, VZ)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: list build + 2B + 3B + flux + (N > 1) the in-library NCCL reduction ---------------------------
    mb.mbpol_upload(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], X, Y, Z, *vel)
    mb.mbpol_compute(h)
    k3 = []

    def steps(count):
        """`count` steps, each bracketed by its own event pair on the library's stream; the L2 flush of the NEXT step is launched
        right behind a step, so that this rank's bookkeeping (phase times) runs under it and never delays the other ranks at the
        next collective.  Returns the per-step device times in ms."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
        with torch.cuda.stream(stream):
            flush.zero_()
            for i, (e0, e1) in enumerate(ev):
                e0.record(stream)
                mb.mbpol_compute(h)          # (synchronises the stream at its end)
                e1.record(stream)
                if i + 1 < count:
                    flush.zero_()
                k3.append(mb.mbpol_phase_times(h))
        torch.cuda.synchronize()
        return [e0.elapsed_time(e1) for e0, e1 in ev]

    steps(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    k3.clear()
    ms = steps(args.steps)
    barrier()
    counts = mb.mbpol_counts(h)
    t_dev = torch.tensor([sum(ms) / len(ms)], dtype=torch.float64, device="cuda")
    t3 = torch.tensor([np.mean([k["k_trimers_ms"] for k in k3])], dtype=torch.float64, device="cuda")
    nts = torch.tensor([counts["trimers_shard"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        both = torch.cat([t3, nts])
        gathered = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(gathered, both)
        slow = max(gathered, key=lambda g: float(g[0]))      # the roofline kernel of the slowest rank
        t3, nts = slow[0:1], slow[1:2]
    ms_step = float(t_dev.item())
    launches = int(mb.mbpol_kernel_times(h)["launches"]) * args.steps

    # ---- end-to-end through the public call (host buffers, copies inside the timed region) ------------------
    outs = {k: pinned(np.zeros(n)) for k in ("fxx", "fyy", "fzz")}
    ae = pinned(np.zeros(n)) if want_e else None
    afd = pinned(np.zeros(9 * n)) if want_w else None
    ast = pinned(np.zeros(9 * n)) if want_s else None
    stress = np.zeros(9)

    def e2e_step():
        t0 = time.perf_counter()
        r = mb.mbpol_forces(h, box["imcon"], box["cell"], box["nttm2"], box["listttm2"], X, Y, Z,
                            outs["fxx"], outs["fyy"], outs["fzz"], stress=stress, atomic_energy=ae, atomic_stress=ast,
                            atomic_fdecomp=afd, vxx=vel[0], vyy=vel[1], vzz=vel[2])
        return time.perf_counter() - t0, r

    for _ in range(2):
        e2e_step()
    for v in list(outs.values()) + [a for a in (ae, afd, ast) if a is not None]:
        v[...] = 0.0
    stress[...] = 0.0
    barrier()
    te, hp = [], []
    for _ in range(args.steps):
        dt, res = e2e_step()
        te.append(dt)
        hp.append(mb.mbpol_phase_times(h))
    host_ms = {k[5:-3]: float(np.mean([q[k] for q in hp])) for k in ("host_upload_ms", "host_compute_ms", "host_download_ms")}
    barrier()
    clocks = sampler.stop() if rank == 0 else None      # sampled over the device-timed AND the end-to-end region
    t_e2e = torch.tensor([float(np.mean(te))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    narr = 3 + (1 if want_e else 0) + (9 if want_w else 0) + (9 if want_s else 0)
    h2d = 8 * n * 6 + 4 * box["nttm2"]
    d2h = 8 * (n * narr // world + 32)       # reduce-scatter: every rank returns its 1/N slice of the requested arrays

    # the e2e accumulators of all ranks, summed like the host's gdsum would: `steps` times the result of one step
    acc = np.concatenate([outs["fxx"], outs["fyy"], outs["fzz"]] + [a for a in (ae, afd, ast) if a is not None] +
                         [stress, [res["eng2"], res["eng3"], res["vir2"], res["vir3"]]])
    if world > 1:
        tacc = torch.from_numpy(acc).cuda()
        dist.all_reduce(tacc)
        acc = tacc.cpu().numpy()

    if rank == 0:
        nt, nd = counts["trimers"], counts["dimers"]
        peak = mb.measure_fp64_peak()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r2_k_trimers_traffic.json")
        if args.nw == 32768 and os.path.exists(tpath):      # ncu --set full capture of this build's k_trimers launch (profiles/)
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            if traffic is not None and world > 1:
                traffic = traffic / world
        k3ms = float(t3.item())
        achieved = float(nts.item()) * mb.FLOP_PER_TRIMER / (k3ms * 1e-3) * 1e-12
        step_flops = (nd * mb.FLOP_PER_DIMER + nt * mb.FLOP_PER_TRIMER)
        line = {"metric": METRIC, "value": nt / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(box, args, nd, nt),
                "e2e": {"value": nt / float(t_e2e.item()), "unit": UNIT, "ms_per_step": 1e3 * float(t_e2e.item()),
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_ms_rank0": host_ms,
                        "note": "mbpol_b200_forces with pinned host arrays; per rank: H2D of coordinates + velocities, step, D2H + '+=' of "
                                "this rank's slice of the reduced per-atom arrays"},
                "gpu_launches": launches,
                "roofline": {"bound": "fp64", "kernel": "k_trimers", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": traffic, "kernel_ms": k3ms,
                             "peak_source": "measured live: independent DFMA chains (mbpol_b200_measure_fp64_peak) at the SM clock in "
                                            "`clocks`; MEASURED_PEAKS.json has no FP64 entry",
                             "algorithmic_flop_per_trimer": mb.FLOP_PER_TRIMER,
                             "step_frac": step_flops / (ms_step * 1e-3) * 1e-12 / peak / world},
                "phases_ms": {k: float(np.mean([q[k] for q in k3])) for k in k3[0] if not k.startswith("host_")},
                "clocks": clocks}
        check = {"eng2": res["eng2"], "eng3": res["eng3"]}
        if world == 1 and not args.no_cpu_baseline:
            # the CPU leg: the reference's own evaluator objects (oracle/_ref) under the restated list-driven driver, on ALL trimers and
            # dimers of this box with every host thread -- it is the cpu_baseline AND the checker of the e2e result above
            from oracle.oracle import Oracle
            o = Oracle()
            threads = os.cpu_count() or 1
            t0 = time.perf_counter()
            ref = o.forces(box, mode="listed", use_ref=o.have_reference, nthreads=threads, want_stress=want_s, want_fdecomp=want_w)
            t_cpu = time.perf_counter() - t0
            tm = ref["timing"]
            t_eval = tm["t2b_s"] + tm["t3b_s"]
            line["cpu_baseline"] = {"value": ref["ntrimer"] / t_eval, "unit": UNIT, "cores": threads,
                                    "kind": "reference" if o.have_reference else "port",
                                    "sample": "the whole box: %d dimers + %d trimers (%.2f s of evaluator time, list build %.2f s excluded)"
                                              % (ref["ndimer"], ref["ntrimer"], t_eval, tm["list_s"]),
                                    "flags": o.ref_flags() or "restated evaluators, g++ -O2 -ffp-contract=off", "run_flags": args.flags,
                                    "us_per_trimer_per_thread": 1e6 * tm["t3b_s"] * threads / max(ref["ntrimer"], 1),
                                    "us_per_dimer_per_thread": 1e6 * tm["t2b_s"] * threads / max(ref["ndimer"], 1),
                                    "wall_s": t_cpu}
            # the reference's real enumeration is O(N^3) (mbpol.f90:306-374): time the faithful mode at BASELINE config 1 (256 H2O)
            b256 = synth.make_water_box(256)
            t0 = time.perf_counter()
            f256 = o.forces(b256, mode="faithful", use_ref=o.have_reference, nthreads=1, want_stress=want_s, want_fdecomp=want_w)
            t256 = time.perf_counter() - t0
            t0 = time.perf_counter()
            o.forces(b256, mode="listed", use_ref=o.have_reference, nthreads=1, want_stress=want_s, want_fdecomp=want_w)
            t256l = time.perf_counter() - t0
            enum_s = max(t256 - t256l, 0.0)
            line["cpu_baseline"]["faithful_256"] = {
                "s_per_step_1_thread": t256, "list_driven_s_per_step_1_thread": t256l, "enumeration_s": enum_s, "n_trimers": f256["ntrimer"],
                "triple_tests": 256 * 255 * 254 // 6,
                "extrapolated_enumeration_s_this_box_1_thread": enum_s * (box["nw"] / 256.0) ** 3,
                "note": "BASELINE config 1 (256 H2O): the reference driver's own O(N^3) triple loop (mbpol.f90:306-374, oracle faithful mode) "
                        "against the list-driven mode on one thread; their difference is the enumeration cost, and its N^3 scaling to this "
                        "box is an EXTRAPOLATION, not a measurement"}
            e_gpu, e_ref = acc[-4] + acc[-3], ref["eng2"] + ref["eng3"]
            k = args.steps
            rr = lambda x, y: float(np.sqrt(np.mean((x / k - y) ** 2) / np.mean(y ** 2)))
            e_gpu = res["eng2"] + res["eng3"]       # scalars are returned per call (fresh zero accumulators), arrays accumulate over k steps
            check.update({"rel_err": abs(e_gpu - e_ref) / abs(e_ref), "force_rel_rms": rr(acc[:3 * n], np.concatenate([ref["fxx"], ref["fyy"], ref["fzz"]])),
                          "against": "oracle listed mode with the reference evaluator objects, whole box, outside the timed region",
                          "n_trimers_equal": int(ref["ntrimer"]) == int(nt), "n_dimers_equal": int(ref["ndimer"]) == int(nd)})
        line["check"] = check
        print(json.dumps(line))
    mb.mbpol_fini(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
