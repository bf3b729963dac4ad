"""The N>1 host path on CPU: two gloo processes, band sharding, ordered merge, max-over-ranks.

The device is not available here, so each rank fills its bands from the oracle (allowed: tests/
is where the oracle may be used); what is under test is the band plan of the C ABI
(frc_plan_bands), the merge back into IterPairs order, and the collective plumbing."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_samples, q):
    import torch.distributed as dist

    from frackyfrac_b200 import dist as fdist
    from frackyfrac_b200 import engine, synth
    from oracle import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree = synth.random_tree(60, 5)
    rp, col, val = synth.random_table(tree, n_samples, 0.1, 6)
    ot, tab = orc.Tree.from_flat(tree.parent, tree.length), orc.Table.from_csr(rp, col, val)
    full = orc.unifrac(tab, ot, True, 1, 1)
    first, count = engine.plan_bands(n_samples, rank, world, band_rows=128)
    chunks = [(int(f), full[f:f + c].copy()) for f, c in zip(first, count)]  # this rank's ordered stream
    merged = fdist.gather_distances(chunks, n_samples, dst=0)
    slowest = fdist.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((np.array_equal(merged, full), slowest, len(first)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_sharding_merges_in_order(built):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 700, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, slowest, nb = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and slowest == 2.0 and nb == 3
