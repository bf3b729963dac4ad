"""Runs one BASELINE config at full size through the C ABI (host buffers -> pinned host distances, streamed),
checks the last rows of the triangle against the oracle, prints throughput.

  python scripts/run_big.py cfg4 [--no-d2h]            one GPU
  python scripts/run_big.py cfg5h --devices 8          one process driving 8 GPUs (opts.n_devices): the weighted
                                                       panels exceed one GPU -> capacity mode (PanelMap)
"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from frackyfrac_b200 import engine, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--check-rows", type=int, default=16)
ap.add_argument("--no-d2h", action="store_true")
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--devices", type=int, default=1)
a = ap.parse_args()
mode, leaves, samples, density, ts, bs = bench.CONFIGS[a.config]
weighted = mode == "weighted"
t0 = time.perf_counter()
tree = synth.random_tree(leaves, ts)
rp, col, val = synth.random_table(tree, samples, density, bs)
print(f"{a.config}: {mode}, {leaves} leaves ({tree.n_nodes} nodes) x {samples} samples, nnz {len(col)}; generated in {time.perf_counter() - t0:.1f}s", flush=True)
import torch
t0 = time.perf_counter()
ctx = engine.Context(devices=list(range(a.devices))) if a.devices > 1 else engine.Context(0)
print(f"context over {a.devices} device(s) in {time.perf_counter() - t0:.1f}s", flush=True)
n = samples
total = n * (n - 1) // 2
r0 = n - a.check_rows
first_checked = r0 * (r0 - 1) // 2
tail = np.zeros(total - first_checked)
t0 = time.perf_counter()
job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=ctx,
                 flags=a.flags | (engine.FLAG_NO_D2H if a.no_d2h else 0))
t1 = time.perf_counter()
seen, chunks, expect = 0, 0, 0
if a.no_d2h:
    seen = job.drain()
else:
    for first, arr in job.chunks_f32(copy=False):   # fp32 bands straight from the pinned rings, in flat-index order
        assert first == expect
        expect += len(arr)
        seen += len(arr)
        chunks += 1
        lo = max(first, first_checked)
        if first + len(arr) > lo:
            tail[lo - first_checked: first + len(arr) - first_checked] = arr[lo - first:]
t2 = time.perf_counter()
info = job.info()
mem = [torch.cuda.mem_get_info(d) for d in range(a.devices)]
job.close()
assert seen == total, (seen, total)
print(f"create {t1 - t0:.3f}s  stream {t2 - t1:.3f}s ({chunks} chunks)  device: h2d {info.h2d_ms:.1f} ms embed {info.embed_ms:.1f} ms "
      f"run {info.run_ms:.1f} ms | devices {info.n_devices} bands {info.n_bands_total} flagged {info.flagged_pairs} kp {info.n_nodes_padded} "
      f"| h2d {info.h2d_bytes / 2**30:.2f} GiB d2h {info.d2h_bytes / 2**30:.2f} GiB exchanged between devices {info.gather_bytes / 2**30:.2f} GiB "
      f"({info.gather_bytes / 2**30 / max(1, info.n_devices):.2f} GiB per device) "
      f"| HBM in use per device {[round((t - f) / 2**30, 1) for f, t in mem]} GiB", flush=True)
print(f"pairs/s: device {total / (info.run_ms / 1e3):.3e}  end-to-end {total / (t2 - t0):.3e}", flush=True)
if weighted:
    lane = total * 2.0 * tree.n_nodes / (info.run_ms / 1e3) / 1e12
    print(f"FP32 lane-ops: {lane:.1f} T/s over {info.n_devices} device(s) = {lane / info.n_devices / 37.22:.3f} of 148 SM x 128 lanes x 1.965 GHz per device "
          f"(whole pass incl. embedding)", flush=True)
if not a.no_d2h:
    # A distance depends on its two samples and the tree only, so the oracle runs on a SUB-TABLE: the last
    # check_rows samples plus 48 random others (embedding all 1e5..2e5 samples on the CPU would take minutes).
    rng = np.random.default_rng(7)
    others = np.sort(rng.choice(r0, size=min(48, r0), replace=False))
    sub = np.concatenate([others, np.arange(r0, n)])
    srp = np.concatenate([[0], np.cumsum(rp[sub + 1] - rp[sub])]).astype(np.int64)
    idx = np.concatenate([np.arange(rp[k], rp[k + 1]) for k in sub])
    ot, tab = orc.Tree.from_flat(tree.parent, tree.length), orc.Table.from_csr(srp, col[idx], val[idx])
    t0 = time.perf_counter()
    want = orc.unifrac(tab, ot, weighted, 1, os.cpu_count())
    m = len(sub)
    errs = []
    for a_ in range(len(others), m):          # rows of the tail
        i = int(sub[a_])
        for b_ in range(a_):
            j = int(sub[b_])
            got = tail[i * (i - 1) // 2 + j - first_checked]
            w = want[a_ * (a_ - 1) // 2 + b_]
            errs.append(abs(got - w) / max(abs(w), 1e-12))
    errs = np.array(errs)
    print(f"oracle on a {m}-sample sub-table ({len(errs)} pairs of rows [{r0},{n})) in {time.perf_counter() - t0:.1f}s on {os.cpu_count()} cores; "
          f"max rel err {errs.max():.2e} (mean {errs.mean():.1e})", flush=True)
    assert errs.max() < 1e-5
ctx.close()
