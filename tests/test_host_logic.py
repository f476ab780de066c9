"""Host-side ensemble logic: sharding, synthetic initial conditions, spectrum assembly; 2-rank gloo all-reduce."""
import os
import socket

import numpy as np
import pytest

from qcxms_b200 import ensemble_setup as es
from qcxms_b200.api import AUTOEV, FSTOAU, KB, load_molecule


def test_shards_partition_the_ensemble():
    for ntraj, g in ((1000, 8), (7, 2), (5, 8)):
        parts = [es.shard_indices(ntraj, g, r) for r in range(g)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(ntraj))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_initial_conditions_follow_the_global_trajectory_id():
    num, xyz, _ = load_molecule("caffeine")
    full = es.synthetic_initial_conditions(num, xyz, 6)
    part = es.synthetic_initial_conditions(num, xyz, 2, first_id=3)
    assert np.array_equal(full["xyz"][3:5], part["xyz"]) and np.array_equal(full["eimp"][3:5], part["eimp"])
    # mdinitu rule: every Cartesian component carries kT/2 at 500 K (reference src/mdinit.f90:10-52)
    ekin = 0.5 * (full["mass"][None, :, None] * full["velo"] ** 2).sum(axis=(1, 2))
    assert np.allclose(ekin, 1.5 * len(num) * KB * 500.0, rtol=1e-12)
    ev = full["eimp"] * AUTOEV
    assert ev.min() >= 1.0 and ev.max() <= 60.0 and np.all(full["tadd"] == 400.0 * FSTOAU) and np.all(full["velof"] == 1.0)


def test_spectrum_helpers():
    b = np.zeros(100); b[15] = 2; b[43] = 8
    s = es.spectrum_from_histogram(b)
    assert s[43] == 100.0 and s[15] == 25.0
    assert abs(es.cosine_similarity(b, 3 * b) - 1.0) < 1e-15 and es.cosine_similarity(b, np.roll(b, 1)) == 0.0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ntraj, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank histograms the (fake) fragment masses of its own shard, then ONE all-reduce combines them
    bins = torch.zeros(64, dtype=torch.float64)
    for t in es.shard_indices(ntraj, world, rank):
        bins[int(t) % 7 + 10] += 1.0
        bins[int(t) % 3 + 40] += 1.0
    es.allreduce_histogram(bins)
    if rank == 0:
        q.put(bins.numpy())
    dist.destroy_process_group()


def test_histogram_allreduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    ntraj, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ntraj, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros(64)
    for t in range(ntraj):
        want[t % 7 + 10] += 1; want[t % 3 + 40] += 1
    assert np.array_equal(got, want)
