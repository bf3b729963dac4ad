"""Synthetic trees and abundance tables for tests and bench.py.

Trees: repeated uniform random joins of two current roots (B = 2n-1 nodes),
leaf names L0..L{n-1}, branch lengths Exp(mean 0.05) rounded to 6 significant
digits, root length 0.  Tables: every sample picks ceil(rho*n) distinct leaves
uniformly, counts = ceil(LogNormal(3, 1.5)).  Everything is seeded.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Tree:
    parent: np.ndarray    # int32 [B], pre-order ids
    length: np.ndarray    # float64 [B]
    leaf_ids: np.ndarray  # int32 [n_leaves]: pre-order id of leaf k (named "L{k}")
    names: list           # str per node ("" for internal)

    @property
    def n_nodes(self) -> int:
        return len(self.parent)


def _round_sig(x: np.ndarray, sig: int = 6) -> np.ndarray:
    x = np.asarray(x, np.float64)
    out = np.zeros_like(x)
    nz = x != 0
    mag = np.floor(np.log10(np.abs(x[nz])))
    # through text so that the Newick file and the binary array agree bit for bit
    out[nz] = np.array([float(f"{v:.{sig - 1}e}") for v in x[nz]]) if nz.sum() < 2_000_000 else \
        np.round(x[nz] / 10.0 ** (mag - sig + 1)) * 10.0 ** (mag - sig + 1)
    return out


def _preorder(children: list, root: int, n_total: int):
    """children[v] = tuple of child ids (creation ids); returns order, parent_pre."""
    order = np.empty(n_total, np.int64)
    pre_of = np.empty(n_total, np.int64)
    parent_pre = np.empty(n_total, np.int32)
    stack = [(root, -1)]
    k = 0
    while stack:
        v, p = stack.pop()
        order[k] = v
        pre_of[v] = k
        parent_pre[k] = p
        ch = children[v]
        for c in reversed(ch):
            stack.append((c, k))
        k += 1
    return order, pre_of, parent_pre


def random_tree(n_leaves: int, seed: int, shape: str = "random", mean_len: float = 0.05,
                integer_lengths: bool = False) -> Tree:
    rng = np.random.default_rng(seed)
    n = int(n_leaves)
    total = 2 * n - 1 if n > 1 else 1
    children = [()] * total
    if n == 1:
        root = 0
    elif shape == "random":
        roots = list(range(n))
        u1 = rng.random(n - 1)
        u2 = rng.random(n - 1)
        for k in range(n - 1):
            m = len(roots)
            i = int(u1[k] * m)
            j = int(u2[k] * (m - 1))
            if j >= i:
                j += 1
            new = n + k
            children[new] = (roots[i], roots[j])
            roots[i] = new
            roots[j] = roots[-1]
            roots.pop()
        root = roots[0]
    elif shape == "caterpillar":
        cur = 0
        for k in range(n - 1):
            new = n + k
            children[new] = (cur, k + 1)
            cur = new
        root = cur
    elif shape == "balanced":
        level = list(range(n))
        nxt = n
        while len(level) > 1:
            up = []
            for a in range(0, len(level) - 1, 2):
                children[nxt] = (level[a], level[a + 1])
                up.append(nxt)
                nxt += 1
            if len(level) % 2:
                up.append(level[-1])
            level = up
        root = level[0]
    else:
        raise ValueError(shape)
    order, pre_of, parent_pre = _preorder(children, root, total)
    if integer_lengths:
        length = rng.integers(1, 10, total).astype(np.float64)
    else:
        length = _round_sig(rng.exponential(mean_len, total))
    length[0] = 0.0
    names = [""] * total
    for k in range(n):
        names[int(pre_of[k])] = f"L{k}"
    return Tree(parent_pre, length, pre_of[:n].astype(np.int32), names)


def random_table(tree: Tree, n_samples: int, density: float, seed: int, integer_counts: bool = True):
    """CSR (row_ptr int64, col int32 = leaf node ids, val float64)."""
    rng = np.random.default_rng(seed)
    n = len(tree.leaf_ids)
    m = max(1, int(np.ceil(density * n)))
    row_ptr = np.arange(n_samples + 1, dtype=np.int64) * m
    col = np.empty(n_samples * m, np.int32)
    for s in range(n_samples):
        pick = rng.choice(n, size=m, replace=False, shuffle=False)
        col[s * m:(s + 1) * m] = tree.leaf_ids[pick]
    val = np.ceil(rng.lognormal(3.0, 1.5, n_samples * m))
    if not integer_counts:
        val = val * rng.random(len(val)) + 1e-3
    return row_ptr, col, val.astype(np.float64)


def to_newick(tree: Tree) -> str:
    B = tree.n_nodes
    kids = [[] for _ in range(B)]
    for v in range(1, B):
        kids[tree.parent[v]].append(v)
    out = []
    # iterative: emit "(" on entry, children separated by ",", ")name:len" on exit
    stack = [(0, 0)]
    while stack:
        v, i = stack.pop()
        if i == 0 and kids[v]:
            out.append("(")
        if i < len(kids[v]):
            if i > 0:
                out.append(",")
            stack.append((v, i + 1))
            stack.append((kids[v][i], 0))
            continue
        if kids[v]:
            out.append(")")
        out.append(tree.names[v])
        if v != 0:
            out.append(":" + repr(float(tree.length[v])))
    out.append(";")
    return "".join(out)


def _fmt(v: float) -> str:
    return str(int(v)) if float(v).is_integer() and abs(v) < 1e15 else repr(float(v))


def to_sparse_text(tree: Tree, row_ptr, col, val) -> str:
    lines = []
    for s in range(len(row_ptr) - 1):
        b, e = row_ptr[s], row_ptr[s + 1]
        lines.append("\t".join(f"{tree.names[c]}:{_fmt(v)}" for c, v in zip(col[b:e], val[b:e])))
    return "\n".join(lines) + "\n"


def to_dense_text(tree: Tree, row_ptr, col, val) -> str:
    n = len(tree.leaf_ids)
    col_of = {int(v): k for k, v in enumerate(tree.leaf_ids)}
    lines = ["\t".join(f"L{k}" for k in range(n))]
    for s in range(len(row_ptr) - 1):
        row = ["0"] * n
        for c, v in zip(col[row_ptr[s]:row_ptr[s + 1]], val[row_ptr[s]:row_ptr[s + 1]]):
            row[col_of[int(c)]] = _fmt(v)
        lines.append("\t".join(row))
    return "\n".join(lines) + "\n"
