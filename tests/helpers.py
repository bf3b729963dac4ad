"""Shared helpers of the test-suite: run one case through the oracle and through the C ABI."""
import numpy as np

from frackyfrac_b200 import engine, synth
from oracle import oracle as orc


def oracle_flat(tree: synth.Tree, csr, weighted, normalize=1, threads=4):
    rp, col, val = csr
    ot = orc.Tree.from_flat(tree.parent, tree.length)
    tab = orc.Table.from_csr(rp, col, val)
    return orc.unifrac(tab, ot, weighted, normalize, threads)


def gpu_flat(tree: synth.Tree, csr, weighted, normalize=True, **kw):
    rp, col, val = csr
    return engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, normalize, **kw)


def rel_err(got, want):
    """Per-distance relative error with a 1e-12 absolute floor; NaN must match NaN."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert (nan_g == nan_w).all(), "NaN pattern differs"
    ok = ~nan_w
    return np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-12) if ok.any() else np.zeros(0)
