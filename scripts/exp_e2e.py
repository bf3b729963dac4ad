"""GPU experiment: where the end-to-end (host buffers -> pinned host distances) time goes."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth  # noqa: E402

tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)

# raw PCIe numbers
h = torch.empty(100_000_000 // 8, dtype=torch.float64).pin_memory()
d = torch.empty_like(h, device="cuda")
for direction, (dst, src) in {"D2H": (h, d), "H2D": (d, h)}.items():
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{direction} 100 MB pinned: {min(ts)*1e3:.2f} ms = {0.1/min(ts):.1f} GB/s")

for band_rows in (0, 128, 256, 1280, 1 << 20):
    rows = []
    for it in range(8):
        t0 = time.perf_counter()
        job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx, band_rows=band_rows)
        t1 = time.perf_counter()
        first = None
        n = 0
        while True:
            _, f, c = job.next_raw()
            if first is None:
                first = time.perf_counter()
            if c == 0:
                break
            n += c
        t2 = time.perf_counter()
        info = job.info()
        job.close()
        t3 = time.perf_counter()
        rows.append((t1 - t0, first - t1, t2 - first, t3 - t2, t3 - t0))
    r = np.array(rows[2:]) * 1e3
    m = r.mean(axis=0)
    print(f"band_rows={band_rows:8d} bands={info.n_bands_total:3d}: create {m[0]:.2f}  first-chunk {m[1]:.2f}  rest {m[2]:.2f}  destroy {m[3]:.2f}  total {m[4]:.2f} ms"
          f" | dev run {info.run_ms:.2f} embed {info.embed_ms:.2f} h2d {info.h2d_ms:.2f}")

os.environ["FRC_ZERO_COPY"] = "1"
for band_rows in (0, 1 << 20):
    rows = []
    for it in range(8):
        t0 = time.perf_counter()
        job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx, band_rows=band_rows)
        t1 = time.perf_counter()
        n = job.drain()
        t2 = time.perf_counter()
        info = job.info()
        job.close()
        rows.append((t1 - t0, t2 - t1, time.perf_counter() - t0))
    m = (np.array(rows[2:]) * 1e3).mean(axis=0)
    print(f"ZERO-COPY band_rows={band_rows:8d}: create {m[0]:.2f} drain {m[1]:.2f} total {m[2]:.2f} ms | dev run {info.run_ms:.2f} pairs_ms {info.pairs_ms:.2f}")
