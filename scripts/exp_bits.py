"""Timing attribution of the bits-fed pair kernel (FRC_BITS_DEBUG: wrong results when non-zero)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)
j = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx,
               band_rows=1 << 20, flags=engine.FLAG_NO_D2H | engine.FLAG_UW_BITS)
j.drain()
for dbg in (0, 1, 2, 3):
    os.environ["FRC_BITS_DEBUG"] = str(dbg)
    ms = []
    for _ in range(5):
        j.restart(); j.drain(); ms.append(j.info().pairs_ms)
    print(f"bits-fed dbg={dbg}: pairs_ms {np.median(ms):.4f} embed_ms {j.info().embed_ms:.4f}", flush=True)
