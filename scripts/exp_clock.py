"""Is the pair kernel clock/power limited?  Cold (after idle) vs hot-loop timings + SM clock samples."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)
samples = []
stop = False
def poll():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.005)
th = threading.Thread(target=poll, daemon=True); th.start()
for flags, name in ((0, "u8"), (engine.FLAG_UW_BF16, "bf16")):
    j = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx,
                   band_rows=1 << 20, flags=engine.FLAG_NO_D2H | flags)
    j.drain()
    cold = []
    for _ in range(5):
        time.sleep(0.3)
        j.restart(); j.drain(); cold.append(j.info().pairs_ms)
    t0 = time.perf_counter(); hot = []
    for _ in range(1500):
        j.restart(); j.drain(); hot.append(j.info().pairs_ms)
    t1 = time.perf_counter()
    clk = [c for t, c, p in samples if t0 + 0.2 < t < t1]
    pw = [p for t, c, p in samples if t0 + 0.2 < t < t1]
    print(f"{name}: cold {np.round(cold, 4)}  hot first5 {np.round(hot[:5], 4)} median {np.median(hot):.4f} last5 {np.round(hot[-5:], 4)}"
          f" | loop {t1 - t0:.2f}s sm clock median {np.median(clk) if clk else None} min {min(clk) if clk else None} power median {np.median(pw) if pw else None:.0f} W max {max(pw) if pw else None:.0f} W", flush=True)
    j.close()
stop = True
