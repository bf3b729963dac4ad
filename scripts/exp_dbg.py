"""Timing attribution of the pair kernel (results are garbage for dbg != 0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)
for flags, name in ((0, "u8"), (engine.FLAG_UW_BF16, "bf16")):
    for gb in ("3", "16"):
        if name == "bf16" and gb == "16": continue
        os.environ["FRC_U8_GROUP_BINADES"] = gb
        j = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx,
                       band_rows=1 << 20, flags=engine.FLAG_NO_D2H | flags)
        j.drain()
        for dbg in (0, 1, 2, 4, 5, 6):
            os.environ["FRC_TC_DEBUG"] = str(dbg)
            ms = []
            for _ in range(5):
                j.restart(); j.drain(); ms.append(j.info().pairs_ms)
            print(f"{name} gb={gb} dbg={dbg}: pairs_ms {np.median(ms):.4f}", flush=True)
        os.environ["FRC_TC_DEBUG"] = "0"
        j.close()
