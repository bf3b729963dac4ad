"""GPU experiment: accuracy of the fast unweighted path vs K-chunk length and distance magnitude."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def case(name, tree, csr, ctx):
    rp, col, val = csr
    want = orc.unifrac(orc.Table.from_csr(rp, col, val), orc.Tree.from_flat(tree.parent, tree.length), False, 1, os.cpu_count())
    for chunk in (1, 4, 16, 64, 100000):
        os.environ["FRC_TC_CHUNK_KBLOCKS"] = str(chunk)
        with engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx) as j:
            got = np.concatenate([a for _, a in j.chunks()])
            info = j.info()
        ok = np.isfinite(want) & (want > 0)
        rel = (got[ok] - want[ok]) / want[ok]
        bins = [(0, 0.125), (0.125, 0.25), (0.25, 0.5), (0.5, 1.01)]
        parts = []
        for lo, hi in bins:
            m = (want[ok] >= lo) & (want[ok] < hi)
            if m.any():
                parts.append(f"[{lo},{hi}): n={m.sum()} max|e|={np.abs(rel[m]).max():.2e} mean e={rel[m].mean():+.2e}")
        print(f"{name:14s} chunk={chunk:6d} flagged={info.flagged_pairs:7d} pairs_ms={info.pairs_ms:.3f} | " + " | ".join(parts), flush=True)


ctx = engine.Context(0)
t = synth.random_tree(700, 71, shape="caterpillar")
case("caterpillar700", t, synth.random_table(t, 200, 0.03, 72), ctx)
t = synth.random_tree(10000, 1002)
case("cfg2-600", t, synth.random_table(t, 600, 0.02, 2002), ctx)
t = synth.random_tree(20000, 5, shape="caterpillar")
case("caterpillar20k", t, synth.random_table(t, 300, 0.01, 6), ctx)
t = synth.random_tree(50000, 1003)
case("cfg3-400", t, synth.random_table(t, 400, 0.02, 2003), ctx)
