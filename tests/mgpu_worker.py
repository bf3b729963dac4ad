"""Worker of tests/test_multi_gpu.py, launched by torchrun with one rank per GPU.

Every rank runs its bands of the triangle with the embedding built sample-sharded and
all-gathered (FRC_FLAG_SHARD_EMBED); rank 0 also runs the whole job alone.  The merged
multi-rank stream must equal the single-rank one byte for byte and stay within
1e-5 of the oracle.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frackyfrac_b200 import dist as fdist, engine, synth  # noqa: E402


def main():
    rank, world, local = fdist.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = engine.Context(local)
    fdist.init_comm(ctx, rank, world)
    tree = synth.random_tree(1500, 301)
    rp, col, val = synth.random_table(tree, 1100, 0.03, 302)
    n = 1100
    ok = True
    for weighted, uw_flags in ((False, 0), (False, engine.FLAG_UW_BF16), (False, engine.FLAG_UW_BITS), (True, 0)):
        chunks = []
        with engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=ctx,
                        rank=rank, world=world, band_rows=128, flags=uw_flags | engine.FLAG_SHARD_EMBED) as job:
            for first, a in job.chunks():
                chunks.append((first, a))
            info = job.info()
        assert info.gather_bytes > 0, "the embedding was not all-gathered"
        full = fdist.gather_distances(chunks, n)
        if rank == 0:
            alone = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, path=engine.PATH_FAST, ctx=ctx,
                                   band_rows=128, flags=uw_flags)
            same = np.array_equal(full, alone, equal_nan=True)
            from oracle import oracle as orc
            want = orc.unifrac(orc.Table.from_csr(rp, col, val), orc.Tree.from_flat(tree.parent, tree.length), weighted, 1, 8)
            err = np.max(np.abs(full - want) / np.maximum(np.abs(want), 1e-12))
            print(f"weighted={weighted} flags={uw_flags} world={world} identical={same} max_rel_err={err:.2e} "
                  f"gather_bytes={info.gather_bytes}", flush=True)
            ok = ok and same and err < 1e-5
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
