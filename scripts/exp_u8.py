"""Experiment: u8 kernel timing / accuracy vs group width and chunking (cfg2)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
from oracle import oracle as orc

leaves, samples = int(os.environ.get("LEAVES", 10000)), int(os.environ.get("SAMPLES", 5000))
tree = synth.random_tree(leaves, 1002)
rp, col, val = synth.random_table(tree, samples, 0.02, 2002)
ctx = engine.Context(0)
ot, tab = orc.Tree.from_flat(tree.parent, tree.length), orc.Table.from_csr(rp, col, val)
r0 = samples - 200
want, _, _ = orc.unifrac_rows(tab, ot, False, 1, os.cpu_count(), r0, samples)
first = r0 * (r0 - 1) // 2

def run(label, env, flags=0):
    for k, v in env.items():
        os.environ[k] = v
    j = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx,
                   band_rows=1 << 20, flags=engine.FLAG_NO_D2H | flags)
    j.drain()
    ms, em = [], []
    for _ in range(5):
        j.restart(); j.drain(); i = j.info(); ms.append(i.pairs_ms); em.append(i.embed_ms)
    i = j.info(); j.close()
    got = engine.unifrac(tree.parent, tree.length, rp, col, val, False, path=engine.PATH_FAST, ctx=ctx, flags=flags)[first:]
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12))
    print(f"{label:28s} kp {i.n_nodes_padded:6d} pairs_ms {np.median(ms):.4f} embed_ms {np.median(em):.4f} max_rel_err {err:.2e}", flush=True)
    for k in env:
        os.environ.pop(k, None)

run("bf16", {}, engine.FLAG_UW_BF16)
run("u8 bits-fed (in-kernel expand)", {}, engine.FLAG_UW_BITS)
for gb in (1, 2, 3, 4, 6, 8, 16):
    run(f"u8 group_binades={gb}", {"FRC_U8_GROUP_BINADES": str(gb)})
