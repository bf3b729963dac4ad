"""Writes the synthetic golden fixtures `synth_*.{tree,sparse,dense,*.want}`.

    python tests/golden/make_golden.py          # from the repo root; needs only the CPU oracle

The answers come from the oracle (oracle/unifrac_oracle.c through oracle/oracle.py), which is itself pinned on
the reference's fixtures next to these files.  They exist so that (a) a change to the oracle cannot move the
target silently and (b) the GPU box checks the CUDA paths against committed bytes, not only against a checker
built in the same run.  Cases (seeded, tiny):

  synth_a   150-leaf random-join tree, 40 samples; sample 7 repeats sample 3 (d = 0), sample 11 is sample 3 with one
            count changed (a small distance); sample 20 lists one species only
  synth_b   caterpillar of 60 leaves with integer lengths, 24 samples, a zero-length branch, a species name carried by
            TWO leaves (the value goes to both, unifrac.go:38-43) and one that only names an internal node (ignored, A2)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from frackyfrac_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def write(name, text):
    with open(os.path.join(HERE, name), "w") as f:
        f.write(text)


def answers(stem, tree_text, table_text, sparse):
    tree = orc.Tree.parse(tree_text)
    tab = orc.Table.parse(table_text, sparse)
    orc.validate_species(tab, tree)
    # wl = flag -l as CODED in the reference (normalize 0: post-order lists, unifrac.go:57,108-110) -- what `frcfrc -w -l`
    # prints; wl2 = flag -l as documented (normalize 2: sorted lists, raw values)
    for tag, weighted, normalize in (("uw", False, 1), ("w", True, 1), ("wl", True, 0), ("wl2", True, 2)):
        d = orc.unifrac(tab, tree, weighted, normalize, 1)
        write(f"{stem}.{tag}.want", "".join(orc.format_go(float(v)) + "\n" for v in d))


def case_a():
    tree = synth.random_tree(150, 9001)
    rp, col, val = synth.random_table(tree, 40, 0.08, 9002)
    m = rp[1]
    col[7 * m:8 * m] = col[3 * m:4 * m]
    val[7 * m:8 * m] = val[3 * m:4 * m]
    col[11 * m:12 * m] = col[3 * m:4 * m]
    val[11 * m:12 * m] = val[3 * m:4 * m]
    val[11 * m] += 1.0
    sparse = synth.to_sparse_text(tree, rp, col, val).split("\n")
    sparse[20] = sparse[20].split("\t")[0]
    sparse = "\n".join(sparse)
    write("synth_a.tree", synth.to_newick(tree) + "\n")
    write("synth_a.sparse", sparse)
    answers("synth_a", synth.to_newick(tree), sparse, True)
    # the same table in the dense format (sample 20 keeps its single species there too)
    names = [f"L{k}" for k in range(150)]
    rows = []
    for line in sparse.strip("\n").split("\n"):
        m_ = dict(tok.split(":") for tok in line.split("\t"))
        rows.append("\t".join(m_.get(n, "0") for n in names))
    dense = "\t".join(names) + "\n" + "\n".join(rows) + "\n"
    write("synth_a.dense", dense)
    t = orc.Tree.parse(synth.to_newick(tree))
    ds = orc.unifrac(orc.Table.parse(dense, False), t, True, 1, 1)
    ss = orc.unifrac(orc.Table.parse(sparse, True), t, True, 1, 1)
    assert np.array_equal(ds, ss, equal_nan=True), "dense and sparse renderings must agree"


def case_b():
    rng = np.random.default_rng(9003)
    n = 60
    lens = rng.integers(1, 9, 2 * n).tolist()
    # caterpillar: (((L0,L1)I1,L2)I2,...) with leaf L5 duplicated as a second leaf named L5 and branch 17 of length 0
    text = f"(L0:{lens[0]},L1:{lens[1]})I1:{lens[2]}"
    for k in range(2, n):
        leaf = "L5" if k == 33 else f"L{k}"
        ll = 0 if k == 17 else lens[2 * k]
        text = f"({text},{leaf}:{ll})I{k}:{lens[2 * k + 1]}"
    tree_text = text.rsplit(":", 1)[0] + ";"
    rows = []
    for s in range(24):
        pick = sorted(rng.choice(n, size=7, replace=False).tolist())
        toks = [f"L{k}:{int(rng.integers(1, 50))}" for k in pick if k != 33]
        if s % 5 == 0:
            toks.append("I9:4")          # names an internal node only: validated, then ignored
        if s % 6 == 1:
            toks.append("L5:3") if not any(t.startswith("L5:") for t in toks) else None
        rows.append("\t".join(toks))
    sparse = "\n".join(rows) + "\n"
    write("synth_b.tree", tree_text + "\n")
    write("synth_b.sparse", sparse)
    answers("synth_b", tree_text, sparse, True)


if __name__ == "__main__":
    case_a()
    case_b()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.startswith("synth_")))
