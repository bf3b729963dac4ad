"""Per-CTA cycle accounting of the pair kernel (FRC_TC_DEBUG=8)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)
L = engine.lib()
names = ["mma:wait_tmem", "mma:wait_operands", "mma:total", "prod:wait_stage", "epi:wait_acc", "epi:drain", "epi:ratio", "epi:total"]
for flags, name in ((0, "u8"), (engine.FLAG_UW_BF16, "bf16")):
    for dbg in (0, 4):
        os.environ["FRC_TC_DEBUG"] = str(dbg)
        j = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx,
                       band_rows=1 << 20, flags=engine.FLAG_NO_D2H | flags)
        j.drain(); j.restart(); j.drain()
        ms = j.info().pairs_ms
        buf = (C.c_ulonglong * (148 * 8))()
        L.frc_debug_tc_counters(buf, 148 * 8)
        a = np.array(buf[:], dtype=np.float64).reshape(148, 8)
        lead, peer = a[0::2], a[1::2]
        print(f"{name} dbg={dbg} pairs_ms {ms:.4f} ({ms * 1965:.0f} kcycles)")
        for k, n in enumerate(names):
            src = lead if k < 3 else a
            print(f"   {n:20s} mean {src[:, k].mean() / 1e3:9.1f}k  min {src[:, k].min() / 1e3:9.1f}k  max {src[:, k].max() / 1e3:9.1f}k")
        j.close()
