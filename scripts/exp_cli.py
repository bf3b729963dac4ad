"""End-to-end time of the frcfrc stand-in CLI on cfg2 (files on disk -> text distances on disk)."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frackyfrac_b200 import hostlib, synth
n_leaves, n_samples = int(os.environ.get("LEAVES", 10000)), int(os.environ.get("SAMPLES", 5000))
tree = synth.random_tree(n_leaves, 1002)
rp, col, val = synth.random_table(tree, n_samples, 0.02, 2002)
d = "/tmp/frc_cli"; os.makedirs(d, exist_ok=True)
open(f"{d}/t.tree", "w").write(synth.to_newick(tree))
open(f"{d}/t.dense", "w").write(synth.to_dense_text(tree, rp, col, val))
open(f"{d}/t.sparse", "w").write(synth.to_sparse_text(tree, rp, col, val))
print("files:", {f: os.path.getsize(f"{d}/{f}") for f in os.listdir(d)}, flush=True)
nproc = os.cpu_count()
for args in (["-i", f"{d}/t.dense"], ["-s", "-i", f"{d}/t.sparse"], ["-w", "-s", "-i", f"{d}/t.sparse"]):
    for p in (1, nproc):
        for rep in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([hostlib.CLI_PATH, "-t", f"{d}/t.tree", "-o", f"{d}/out.txt", "-p", str(p)] + args,
                               capture_output=True, text=True, env=dict(os.environ, FRC_CLI_TIMING="1"))
            dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        pairs = n_samples * (n_samples - 1) // 2
        print(f"frcfrc {' '.join(a for a in args if not a.startswith('/'))} -p {p}: {dt:.2f}s wall ({pairs / dt:.3e} pairs/s), "
              f"output {os.path.getsize(f'{d}/out.txt') / 1e6:.0f} MB", flush=True)
        print("    " + "\n    ".join(l for l in r.stderr.splitlines() if l.startswith("[timing]") or l.startswith("Took")), flush=True)
