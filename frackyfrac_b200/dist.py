"""Multi-process plumbing around the C ABI: one process per GPU, torch.distributed.

The embedding is built sample-sharded and exchanged with ONE NCCL all-gather of its compact
form inside libfrcfrc_cuda (FRC_FLAG_SHARD_EMBED; the communicator is created here from an id
that torch.distributed carries between the ranks).  The pair stage then shards by tile bands
with no further exchange.  Host-side: the per-rank ordered streams are merged back into
IterPairs order, and timings are reduced as the max over ranks.
"""
from __future__ import annotations

import numpy as np

from . import engine


def env_rank_world():
    import os

    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_comm(ctx: "engine.Context", rank: int, world: int, device=None) -> None:
    """Creates the engine's NCCL communicator on `ctx`: rank 0 draws the id, torch.distributed
    (any backend) broadcasts its 128 bytes, every rank joins (collective)."""
    import torch
    import torch.distributed as dist

    if world <= 1:
        return
    if dist.get_backend() == "nccl" and device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    t = torch.zeros(engine.COMM_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        t = torch.tensor(list(engine.comm_unique_id()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    ctx.comm_init(bytes(t.cpu().tolist()), rank, world)


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def band_owner_table(n_samples: int, world: int, band_rows: int = 0):
    """[(first_index, count, owner)] for every band of the triangle, in flat-index order."""
    rows = []
    for r in range(world):
        f, c = engine.plan_bands(n_samples, r, world, band_rows)
        rows += [(int(a), int(b), r) for a, b in zip(f, c)]
    rows.sort()
    return rows


def gather_distances(local_chunks, n_samples: int, dst: int = 0):
    """Assemble the full flat vector on rank `dst` from each rank's (first_index, array) chunks.

    Bands are disjoint, so a SUM reduce of the zero-filled local vectors is an exact merge.
    (For the sizes where the full vector does not fit one host, ranks write their own files.)
    """
    import torch
    import torch.distributed as dist

    total = n_samples * (n_samples - 1) // 2 if n_samples >= 2 else 0
    full = np.zeros(total, np.float64)
    for first, a in local_chunks:
        full[first:first + len(a)] = a
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(full)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
        full = t.cpu().numpy()
    return full
