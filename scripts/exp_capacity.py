"""Weighted pair stage with the panels resident on every device vs sharded over the devices (capacity mode:
column panels read from peers over NVLink).  Usage: python scripts/exp_capacity.py [config] [n_devices]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from frackyfrac_b200 import engine, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode, leaves, samples, density, ts, bs = bench.CONFIGS[cfg]
tree = synth.random_tree(leaves, ts)
rp, col, val = synth.random_table(tree, samples, density, bs)
total = samples * (samples - 1) // 2
ctx = engine.Context(devices=list(range(nd))) if nd > 1 else engine.Context(0)
for cap in (["0", "1"] if nd > 1 else ["0"]):
    os.environ["FRC_CAPACITY"] = cap
    job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=True, path=engine.PATH_FAST, ctx=ctx, flags=engine.FLAG_NO_D2H)
    job.drain()
    for rep in range(2):
        job.restart()
        t0 = time.perf_counter()
        job.drain()
        info = job.info()
        lane = total * 2.0 * tree.n_nodes / (info.run_ms / 1e3) / 1e12
        print(f"{cfg} x{nd} capacity={cap}: run {info.run_ms:.1f} ms embed {info.embed_ms:.1f} ms -> {total / info.run_ms * 1e3:.3e} pairs/s, "
              f"{lane:.1f} T lane-op/s = {lane / nd / 37.22:.3f} of the FP32 lane roofline per device; exchanged {info.gather_bytes / 2**30:.1f} GiB", flush=True)
    job.close()
ctx.close()
