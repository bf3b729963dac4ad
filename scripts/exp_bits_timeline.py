import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)
L = engine.lib()
j = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx,
               band_rows=1 << 20, flags=engine.FLAG_NO_D2H | engine.FLAG_UW_BITS)
j.drain()
buf = (C.c_ulonglong * (148 * 8))()
L.frc_debug_bits_counters(buf, 148 * 8, 1)
j.restart(); j.drain()
ms = j.info().pairs_ms
L.frc_debug_bits_counters(buf, 148 * 8, 1)
a = np.array(buf[:], dtype=np.float64).reshape(148, 8)
names = ["loader: wait bits-stage free", "mma: wait tmem", "mma: wait operands", "prod: wait bits", "prod: wait operand stage", "prod: work"]
print(f"bits-fed pairs_ms {ms:.4f} ({ms * 1965:.0f} kcycles)")
for k, n in enumerate(names):
    src = a[0::2] if k in (1, 2) else a
    print(f"   {n:30s} mean {src[:, k].mean() / 1e3:9.1f}k  min {src[:, k].min() / 1e3:9.1f}k  max {src[:, k].max() / 1e3:9.1f}k")
