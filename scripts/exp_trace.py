import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import engine, synth
tree = synth.random_tree(10000, 1002)
rp, col, val = synth.random_table(tree, 5000, 0.02, 2002)
ctx = engine.Context(0)
for it in range(4):
    if it == 3:
        os.environ["FRC_TRACE"] = "1"
    job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=ctx)
    job.drain(); print(job.info().run_ms); job.close()
