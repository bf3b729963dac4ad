"""GPU experiment: CTA-pair tensor-core kernel vs single-CTA kernel (parity + time)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
from oracle import oracle as orc
ctx = engine.Context(0)
def run(tree, csr, ctas, band_rows=0, flags=0):
    os.environ["FRC_TC_CTAS"] = str(ctas)
    rp, col, val = csr
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx, band_rows=band_rows) as j:
        got = np.concatenate([a for _, a in j.chunks()])
        info = j.info()
    return got, info
for n_leaves, n_samples in ((500, 130), (1000, 300), (2000, 700), (3000, 1000)):
    tree = synth.random_tree(n_leaves, 21)
    csr = synth.random_table(tree, n_samples, 0.02, 22)
    want = orc.unifrac(orc.Table.from_csr(*csr), orc.Tree.from_flat(tree.parent, tree.length), False, 1, 8)
    a, ia = run(tree, csr, 1)
    b, ib = run(tree, csr, 2)
    ea = np.nanmax(np.abs(a - want) / np.maximum(np.abs(want), 1e-12)); eb = np.nanmax(np.abs(b - want) / np.maximum(np.abs(want), 1e-12))
    print(f"{n_leaves}x{n_samples}: err 1cta {ea:.2e} 2cta {eb:.2e} identical {np.array_equal(a, b)} pairs_ms {ia.pairs_ms:.3f} vs {ib.pairs_ms:.3f}", flush=True)
tree = synth.random_tree(10000, 1002)
csr = synth.random_table(tree, 5000, 0.02, 2002)
for ctas in (1, 2):
    os.environ["FRC_TC_CTAS"] = str(ctas)
    j = engine.Job(tree.parent, tree.length, *csr, weighted=False, path=engine.PATH_FAST, ctx=ctx, band_rows=1 << 20, flags=engine.FLAG_NO_D2H)
    j.drain()
    ms = []
    for _ in range(5):
        j.restart(); j.drain(); ms.append(j.info().pairs_ms)
    j.close()
    print(f"cfg2 single launch ctas={ctas}: kernel {np.median(ms):.4f} ms", flush=True)
