"""Multi-GPU test of the in-library NCCL path (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
Two processes, one GPU each, driven through the C ABI only: mbpol_b200_comm_unique_id / _comm_init, then mbpol_b200_forces.  Every
rank returns its slice of the reduced per-atom arrays; their sum over the ranks (the host's gdsum) must equal the single-GPU result,
and the shards must partition the lists."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank(rank, world, q_uid, q_out, mode):
    sys.path.insert(0, ROOT)
    import mbpol_b200 as mb
    from mbpol_b200 import synth
    box = synth.make_water_box(700, seed=7)
    h = mb.mbpol_init(box["nttm2"], flags=7, device=rank)
    if rank == 0:
        uid = mb.mbpol_comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    mb.mbpol_comm_init(h, rank, world, uid)
    mb.mbpol_comm_mode(h, mode)
    outs = []
    for step in range(3):                    # step 0 sizes the lists by read-back, later steps run from the estimates and the weights
        outs.append(mb.run_box(h, box, want_stress=True))
    d, t = mb.mbpol_lists(h)
    c = mb.mbpol_counts(h)
    sh = mb.mbpol_shard(h)
    mb.mbpol_fini(h)
    q_out.put((rank, {k: v for k, v in outs[-1].items()}, d, t, c, sh))


@pytest.mark.parametrize("mode", [2, 1])
def test_two_gpus_reduce_in_library(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import mbpol_b200 as mb
    from mbpol_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import rel_rms, sort_rows
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=_rank, args=(r, 2, q_uid, q_out, mode)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q_out.get(timeout=600) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    box = synth.make_water_box(700, seed=7)
    h = mb.mbpol_init(box["nttm2"], flags=7, device=0)
    one = mb.run_box(h, box, want_stress=True)
    d1, t1 = mb.mbpol_lists(h)
    mb.mbpol_fini(h)
    # lists: the two shards partition the single-GPU lists; both ranks know the box totals (they rode through the reduction)
    assert np.array_equal(sort_rows(np.concatenate([res[0][2], res[1][2]])), sort_rows(d1))
    assert np.array_equal(sort_rows(np.concatenate([res[0][3], res[1][3]])), sort_rows(t1))
    for r in res:
        assert r[4]["trimers"] == len(t1) and r[4]["dimers"] == len(d1)
    assert res[0][5]["first_water"] == 0 and res[0][5]["n_waters"] + res[1][5]["n_waters"] == 700
    # values: sum over the ranks (mode 2: slices; mode 1: every rank returns everything -> take rank 0) == single GPU
    keys = ("fxx", "fyy", "fzz", "atomic_energy", "atomic_stress", "atomic_fdecomp")
    for k in keys:
        got = res[0][1][k] + res[1][1][k] if mode == 2 else res[0][1][k]
        assert rel_rms(got, one[k]) <= 1e-11, k
        if mode == 1:
            assert rel_rms(res[1][1][k], one[k]) <= 1e-11, k
    for k in ("eng2", "vir2", "eng3", "vir3"):      # scalars: rank 0 returns the totals, rank 1 zero
        assert abs(res[0][1][k] - one[k]) <= 1e-11 * abs(one[k]) and res[1][1][k] == 0.0, k
    assert rel_rms(res[0][1]["stress"], one["stress"]) <= 1e-11 and not np.any(res[1][1]["stress"])
    assert np.abs(res[0][1]["jflux"] - one["jflux"]).max() <= 1e-10 * np.abs(one["jflux"]).max()
