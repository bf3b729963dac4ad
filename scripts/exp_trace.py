"""FRC_TRACE timeline of one end-to-end call (host stages of frc_create, per-band device events and host
delivery times).  Usage: python scripts/exp_trace.py [config] [n_devices]   (config: a bench.py CONFIGS name;
the sample count grows with sqrt(n_devices) like bench.py's weak scaling)."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from frackyfrac_b200 import engine, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode, leaves, samples, density, ts, bs = bench.CONFIGS[cfg]
samples = int(round(samples * math.sqrt(nd)))
tree = synth.random_tree(leaves, ts)
rp, col, val = synth.random_table(tree, samples, density, bs)
ctx = engine.Context(devices=list(range(nd))) if nd > 1 else engine.Context(0)
for it in range(6):
    if it == 5:
        os.environ["FRC_TRACE"] = "1"
    t0 = time.perf_counter()
    job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=mode == "weighted", path=engine.PATH_FAST, ctx=ctx)
    t1 = time.perf_counter()
    job.drain()
    t2 = time.perf_counter()
    info = job.info()
    job.close()
    t3 = time.perf_counter()
    print(f"[call {it}] {cfg} x{nd} devices, {samples} samples: create {1e3 * (t1 - t0):.3f} ms, drain {1e3 * (t2 - t1):.3f} ms, "
          f"destroy {1e3 * (t3 - t2 if False else t3 - t2):.3f} ms, total {1e3 * (t3 - t0):.3f} ms; device: h2d {info.h2d_ms:.3f} embed {info.embed_ms:.3f} "
          f"run {info.run_ms:.3f} ms; h2d {info.h2d_bytes} B d2h {info.d2h_bytes} B gather {info.gather_bytes} B", flush=True)
ctx.close()
