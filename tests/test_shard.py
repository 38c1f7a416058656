"""CPU tests of the multi-GPU host logic: contiguous list shards and the one-collective reduction of the packed
per-rank outputs (world_size 2, gloo)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    sys.path.insert(0, ROOT)
    import bench
    for total in (0, 1, 7, 294963, 2341933):
        for n in (1, 2, 3, 4, 8):
            r = [bench.shard_range(total, k, n) for k in range(n)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(n - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_shard_bounds_uniform_and_weighted():
    """mbpol_b200_shard_bounds (host arithmetic inside the library, no GPU needed): contiguous water ranges that cover all waters;
    uniform without weights, balanced to within one water's weight with them."""
    sys.path.insert(0, ROOT)
    import mbpol_b200 as mb
    for nw in (0, 1, 7, 256, 32768):
        for n in (1, 2, 3, 8):
            b = mb.shard_bounds(nw, n)
            assert b[0] == 0 and b[-1] == nw and np.all(np.diff(b) >= 0)
            assert np.diff(b).max() - np.diff(b).min() <= 1
    rng = np.random.default_rng(3)
    w = rng.integers(100, 400, size=32768).astype(np.int32)
    w[:8192] *= 3                                        # a dense region: uniform ranges would be 2x off
    for n in (2, 4, 8):
        b = mb.shard_bounds(len(w), n, w)
        assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) > 0)
        loads = np.array([w[b[r]:b[r + 1]].sum() for r in range(n)], dtype=np.int64)
        assert loads.max() - loads.min() <= 2 * int(w.max())
        assert np.array_equal(b, mb.shard_bounds(len(w), n, w))      # deterministic: every rank cuts the same way
    assert np.array_equal(mb.shard_bounds(100, 4, np.zeros(100, dtype=np.int32)), mb.shard_bounds(100, 4))   # no weights yet -> uniform


def _worker_slices(rank, world, port, q):
    """The reduce-scatter contract of the library, emulated on the host: every rank adds ONLY its slice of the reduced per-atom buffer
    to its (zeroed) arrays and only rank 0 adds the scalars; the caller's own sum over the ranks (gdsum) then gives the full result."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import mbpol_b200 as mb
    from mbpol_b200 import synth
    from oracle.oracle import Oracle
    o = Oracle()
    box = synth.make_water_box(27, seed=5)
    n = box["mxatms"]
    part = o.forces(box, mode="listed", bead_rank=rank, bead_size=world)
    chunk, soff, total = mb.packed_layout(n, world)
    t = torch.from_numpy(bench.pack_outputs(part, n, world))
    assert len(t) == total
    dist.all_reduce(t)                                   # what ncclReduceScatter + the scalar all-reduce leave on the devices
    mine = np.zeros(total)
    lo, hi = rank * chunk, min((rank + 1) * chunk, 22 * n)
    mine[lo:hi] = t.numpy()[lo:hi]                       # this rank's slice only
    if rank == 0:
        mine[soff:] = t.numpy()[soff:]                   # scalars: rank 0 only
    g = torch.from_numpy(mine)
    dist.all_reduce(g)                                   # the caller's gdsum
    ref = bench.pack_outputs(o.forces(box, mode="listed"), n, world)
    q.put((rank, float(np.abs(g.numpy() - ref).max() / np.abs(ref).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slices_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker_slices, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert all(err < 1e-12 for _, err in res), res


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import mbpol_b200  # noqa: F401
    from mbpol_b200 import synth
    from oracle.oracle import Oracle
    o = Oracle()
    box = synth.make_water_box(27, seed=5)
    # each rank evaluates its slice of the work list (the oracle stands in for the device kernels here) ...
    part = o.forces(box, mode="listed", bead_rank=rank, bead_size=world)
    packed = bench.pack_outputs(part, box["mxatms"])
    t = torch.from_numpy(packed)
    bench.allreduce_packed(t)           # ... and ONE all-reduce over the packed buffer finishes the step
    full = o.forces(box, mode="listed")
    ref = bench.pack_outputs(full, box["mxatms"])
    err = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
    q.put((rank, err, float(t.numpy()[-1])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, _ in res:
        assert err < 1e-12, (rank, err)


def _worker_nrg(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mbpol_b200  # noqa: F401
    from mbpol_b200 import synth
    from oracle.oracle import Oracle
    o = Oracle()
    box = synth.make_ion_box(216, 8)
    # MB-nrg (row f1): water slices per rank, ion-ion pairs on rank 0 only; forces, energy, virial and stress of the ranks
    # add up to the single-rank result with ONE all-reduce of the packed partial sums
    part = o.nrg_forces(box, rank=rank, nranks=world)
    packed = np.concatenate([part["fxx"], part["fyy"], part["fzz"], [part["eng2"], part["vir2"]], part["stress"]])
    t = torch.from_numpy(packed)
    dist.all_reduce(t)
    full = o.nrg_forces(box)
    ref = np.concatenate([full["fxx"], full["fyy"], full["fzz"], [full["eng2"], full["vir2"]], full["stress"]])
    err = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
    q.put((rank, err, part["nionion"], part["npairs"], full["npairs"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_gloo_mbnrg():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker_nrg, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert all(err < 1e-12 for _, err, _, _, _ in res)
    assert res[0][2] > 0 and res[1][2] == 0                      # ion-ion pairs live on rank 0
    assert res[0][3] + res[1][3] == res[0][4]                    # the water slices partition the pair list
