"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank senses its own block of decision groups and the
occupancy exchange reassembles the capture's decisions in order.  The per-rank compute is stood in for by
the oracle port here (no GPU in this container); on the GPU box bench.py runs the same plumbing over NCCL
with the CUDA path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ngroups, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import crn_b200 as crn
        import oracle
        import importlib
        cdist = importlib.import_module("crn_b200.dist")
        cfg = crn.config_welch(1024, 4)
        gs = cfg.group_samples
        counts = [cdist.shard_groups(ngroups, world, r)[1] for r in range(world)]
        first, count = cdist.shard_groups(ngroups, world, rank)
        assert crn.shard_groups(ngroups, world, rank) == (first, count)
        sc = crn.synth_config(gs, dwell_groups=2, snr_db=10.0, seed=12)
        iq, _ = oracle.synth(sc, count * gs, first=first * gs)       # this rank's shard of ONE capture
        _, _, dec, _ = oracle.sense_port(cfg, iq)
        t = cdist.max_over_ranks(1.0 + rank)                            # slowest rank defines the step
        assert t == float(world)
        full = cdist.gather_occupancy(torch.from_numpy(dec), counts)
        hist = cdist.occupancy_histogram(torch.from_numpy(dec))
        dist.barrier()
        if rank == 0:
            q.put((full.numpy(), hist.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ngroups", [9, 8])
def test_two_ranks_reassemble_the_capture(crn, oracle, ngroups):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + ngroups
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ngroups, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, hist = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = crn.config_welch(1024, 4)
    sc = crn.synth_config(cfg.group_samples, dwell_groups=2, snr_db=10.0, seed=12)
    iq, _ = oracle.synth(sc, ngroups * cfg.group_samples)
    _, _, dec, _ = oracle.sense_port(cfg, iq)
    assert np.array_equal(full, dec)
    assert np.array_equal(hist, np.bincount(dec, minlength=4))
