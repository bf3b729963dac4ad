#!/usr/bin/env python
"""UniFrac sample-pairs/sec on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

A step = one full pass of the hot path (branch embedding + every pair of the
lower triangle) over one synthetic batch.  Workload at N=1 is BASELINE.json
configs[1]: unweighted UniFrac, synthetic 10k-leaf tree x 5k samples
(SURVEY.md §8d generator, seeds 1002 / 2002).  For N>1 the sample count grows
with sqrt(N) so that pairs per GPU stay fixed ("weak"); tile bands of the
triangle are dealt to the ranks, no data-path collective (the embedding is
<1 % of a step and is rebuilt per rank).

  value  pairs/s with inputs resident in HBM (frc_restart + drain, distances stay in HBM)
  e2e    pairs/s through the C ABI from HOST buffers to pinned HOST distances
         (frc_create -> frc_next* -> frc_destroy every step)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (mode, leaves, samples, density, tree_seed, table_seed)
    "cfg2": ("unweighted", 10_000, 5_000, 0.02, 1002, 2002),
    "cfg2w": ("weighted", 10_000, 5_000, 0.02, 1002, 2002),
    "cfg3": ("weighted", 50_000, 20_000, 0.02, 1003, 2003),
    "cfg3u": ("unweighted", 50_000, 20_000, 0.02, 1003, 2003),
    "cfg4": ("unweighted", 100_000, 100_000, 0.02, 1004, 2004),
    "cfg4s": ("unweighted", 100_000, 30_000, 0.02, 1004, 2004),
    "cfg4w": ("weighted", 100_000, 100_000, 0.02, 1004, 2004),
    # 150 GB of u8 operands: more than one GPU holds -> the pair kernel expands its tiles from the bit rows
    "cfg4x": ("unweighted", 100_000, 250_000, 0.02, 1004, 2004),
    "tiny": ("unweighted", 1_000, 512, 0.02, 1009, 2009),
}


def load_traffic(config: str, kernel: str, world: int):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this config (1 GPU)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get(f"{config}:{kernel}")
    return e["bytes"] if e else None


def measure_int8_peak():
    """Dense int8 tensor peak measured the way MEASURED_PEAKS.json measures bf16: a cuBLASLt GEMM
    (torch._int_mm, 8192^3, best of 10, CUDA events).  None when this torch build cannot run it."""
    try:
        import torch

        n = 8192
        a = torch.randint(-4, 4, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-4, 4, (n, n), dtype=torch.int8, device="cuda")
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        return 2.0 * n ** 3 / (best / 1e3) / 1e12
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.windows = []  # (t0, t1) of timed regions, perf_counter

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [r for t, r in self.rows if any(a - 0.05 <= t <= b + 0.15 for a, b in self.windows)]
        window = "timed regions"
        if len(inside) < 3:  # regions shorter than the sampling period: use the whole loaded phase
            inside, window = [r for _, r in self.rows], "warm-up + timed regions"
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def make_workload(name: str, world: int, fixed: bool = False):
    from frackyfrac_b200 import synth

    mode, leaves, samples, density, ts, bs = CONFIGS[name]
    if world > 1 and not fixed:  # weak scaling: pairs per GPU constant
        samples = int(round(samples * math.sqrt(world)))
    tree = synth.random_tree(leaves, ts)
    rp, col, val = synth.random_table(tree, samples, density, bs)
    return mode, tree, (rp, col, val), samples, leaves, density


def cpu_baseline(tree, csr, weighted, target_s=12.0, threads=None):
    """The oracle (a port of the reference algorithm) on the host cores, bounded sample."""
    from oracle import oracle as orc

    threads = threads or os.cpu_count() or 1
    rp, col, val = csr
    n = len(rp) - 1
    ot = orc.Tree.from_flat(tree.parent, tree.length)
    tab = orc.Table.from_csr(rp, col, val)
    # calibrate on the last few rows, then size the sample for ~target_s of pair work
    r0 = max(1, n - max(2, threads // 2))
    _, emb_s, pair_s = orc.unifrac_rows(tab, ot, weighted, 1, threads, r0, n)
    cal_pairs = n * (n - 1) // 2 - r0 * (r0 - 1) // 2
    rate = cal_pairs / max(pair_s, 1e-6)
    want_pairs = min(n * (n - 1) // 2, int(rate * target_s))
    rows = max(1, min(n - 1, int(round(want_pairs / max(n - 1, 1)))))
    rb = n - rows
    _, emb_s, pair_s = orc.unifrac_rows(tab, ot, weighted, 1, threads, rb, n)
    pairs = n * (n - 1) // 2 - rb * (rb - 1) // 2
    total_pairs = n * (n - 1) // 2
    # whole-job rate implied by the sample: embedding charged in proportion
    t = pair_s + emb_s * pairs / total_pairs
    return {"value": pairs / t, "unit": "sample-pairs/s", "cores": threads, "kind": "port",
            "sample": f"rows [{rb},{n}) of the triangle = {pairs} pairs in {pair_s:.2f}s pair stage + "
                      f"{emb_s:.2f}s embedding of all {n} samples (charged pro rata)",
            "embed_s": emb_s, "pair_s": pair_s, "pairs": pairs}


def run_reference(args, world, rank):
    if rank != 0:
        return
    mode, tree, csr, samples, leaves, density = make_workload(args.config, world, args.fixed_size)
    weighted = mode == "weighted"
    vals, last = [], None
    t_all0 = time.perf_counter()
    # every step is a bounded sample of the workload, sized so that the W + K
    # steps the driver asked for end within ~2.5 minutes of CPU work
    per_step = min(args.ref_seconds, max(0.25, 150.0 / max(1, args.steps + args.warmup)))
    n_warm, n_steps = args.warmup, args.steps
    for it in range(n_warm + n_steps):
        last = cpu_baseline(tree, csr, weighted, target_s=per_step)
        if it >= n_warm:
            vals.append(last)
    ms = 1e3 * np.mean([v["pair_s"] + v["embed_s"] * v["pairs"] / (samples * (samples - 1) // 2) for v in vals])
    v = float(np.mean([x["value"] for x in vals]))
    out = {"impl": "reference", "metric": f"UniFrac sample-pairs/sec ({mode})", "value": v, "unit": "sample-pairs/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args.config, mode, leaves, samples, density, world),
           "cpu_baseline": {"value": v, "unit": "sample-pairs/s", "cores": last["cores"], "kind": "port",
                            "sample": last["sample"]},
           "e2e": {"value": v, "unit": "sample-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "Go toolchain absent: oracle/ (C port of frcfrc/unifrac.go, pthreads on all host cores) stands in "
                   "for `frcfrc -p $(nproc)`; each step is a bounded sample of the workload; "
                   f"{n_steps} timed steps of ~{per_step:.2f} s CPU work each after {n_warm} warm-up steps",
           "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(out), flush=True)


def workload_config(name, mode, leaves, samples, density, world, shard=False):
    return {"workload": f"{name}: {mode} UniFrac, synthetic {leaves}-leaf random-join tree ({2 * leaves - 1} nodes) x "
                        f"{samples} samples, leaf density {density}, lognormal counts",
            "leaves": leaves, "nodes": 2 * leaves - 1, "samples": samples, "pairs": samples * (samples - 1) // 2,
            "parallelism": (f"tile-band sharding x{world}, embedding " +
                            ("sample-sharded + NCCL all-gather" if shard else "rebuilt per rank")) if world > 1 else "single GPU",
            "l2": "operands (3 x samples x nodes, u8 or bf16 = 0.3 / 0.6 GB at cfg2) exceed the 126 MB L2 and "
                  "every step rewrites them: no L2 flush needed"}


def run_ours(args, world, rank, local_rank):
    import torch

    from frackyfrac_b200 import engine

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode, tree, csr, samples, leaves, density = make_workload(args.config, world, args.fixed_size)
    weighted = mode == "weighted"
    rp, col, val = csr
    total_pairs = samples * (samples - 1) // 2
    ctx = engine.Context(local_rank)
    peaks, peaks_kind = load_peaks()
    uw_flags = engine.FLAG_UW_BF16 if args.uw_kernel == "bf16" else 0
    # sample-sharded embedding + NCCL all-gather of its compact form (FRC_FLAG_SHARD_EMBED): pays when
    # the shardable part of the embedding (bit columns + row sums, ~O(samples x nodes)) outweighs the
    # latency of an all-gather (measured ~140 us at 8 ranks).  At the cfg2-derived sizes of the scaling
    # run it does not (N=8: 0.738 ms per step sharded vs 0.688 ms rebuilt per rank), so the default
    # is size-based; --shard-embed / --no-shard-embed force either.
    shard = world > 1 and not args.no_shard_embed and (args.shard_embed or samples * (2 * leaves - 1) >= 2_000_000_000)
    if shard:
        from frackyfrac_b200 import dist as fdist
        fdist.init_comm(ctx, rank, world)
        uw_flags |= engine.FLAG_SHARD_EMBED

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------- value: inputs resident in HBM
    job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=ctx,
                     rank=rank, world=world, flags=engine.FLAG_NO_D2H | uw_flags)
    sampler = ClockSampler(local_rank)
    sampler.start()
    job.drain()
    for _ in range(args.warmup):
        job.restart(); job.drain()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(args.steps):
        job.restart(); job.drain()
        info = job.info()
        dev_ms += info.run_ms
        launches += info.kernel_launches
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark(t0, t0 + wall)
    dev_s = maxreduce(dev_ms / 1e3)   # CUDA events: embedding start -> last band done
    wall_s = maxreduce(wall)
    info = job.info()
    job.close()
    value = total_pairs * args.steps / dev_s

    # ---------------------------------------------- e2e: host buffers through the C ABI
    class _NoInfo:
        h2d_bytes = d2h_bytes = 0
        h2d_ms = 0.0

    def e2e_step():
        with engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=ctx,
                        rank=rank, world=world, flags=uw_flags) as j:
            n = j.drain()   # native format of the fast paths: fp32 bands straight from the pinned ring (frc_next_f32)
            i = j.info()
        return n, i

    ei = _NoInfo()
    if not args.no_e2e:
        for _ in range(args.warmup):
            e2e_step()
    barrier()
    t0 = time.perf_counter()
    if not args.no_e2e:
        for _ in range(args.steps):
            _, ei = e2e_step()
    barrier()
    sampler.mark(t0, time.perf_counter())
    e2e_s = max(maxreduce(time.perf_counter() - t0), 1e-9)
    clocks = sampler.stop()
    e2e = None if args.no_e2e else {"value": total_pairs * args.steps / e2e_s, "unit": "sample-pairs/s",
           "h2d_bytes_per_step": int(ei.h2d_bytes), "d2h_bytes_per_step": int(ei.d2h_bytes),
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "timed": "host wall clock around frc_create..frc_next*..frc_destroy (includes host validation, "
                    "pinned staging, H2D, kernels, D2H of every distance)"}

    # ---------------------------------------------- roofline of the dominant kernel (rank 0, timed alone)
    roofline, cpu = None, None
    if rank == 0 and not args.no_roofline:
        rj = engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=ctx,
                        band_rows=1 << 20, flags=engine.FLAG_NO_D2H | uw_flags)
        rj.drain()
        ms = []
        for _ in range(max(3, args.steps)):
            rj.restart(); rj.drain()
            ms.append(rj.info().pairs_ms)
        ri = rj.info()
        rj.close()
        k_ms = float(np.median(ms))
        B = tree.n_nodes
        if weighted:
            # 2 FP32 lane-ops per (pair, node); peak = SMs * 128 lanes * clock
            peak = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            ach = total_pairs * 2.0 * B / (k_ms / 1e3) / 1e12
            roofline = {"bound": "fp32", "kernel": "k_weighted_tiles", "achieved": ach, "peak": peak,
                        "unit": "T lane-op/s", "frac": ach / peak,
                        "traffic": load_traffic(args.config, "k_weighted_tiles", world), "kernel_ms": k_ms,
                        "peak_source": "derived: 148 SM x 128 lanes x sm_max_mhz"}
        else:
            # algorithmic work = one multiply-add per (pair, node) = 2*B flops per pair.
            # bf16 hi/lo kernel: executes 2 bf16 planes on the kind::f16 pipe -> peak = measured bf16 burst.
            # u8 kernel: executes 2 u8 planes on the kind::i8 pipe, whose rate is 2x the bf16 one
            # (no measured int8 figure in MEASURED_PEAKS.json: derived as 2 x bf16_tflops, stated).
            i8 = ri.operand_kind >= 2
            int8_measured = measure_int8_peak() if i8 else None
            # the larger of the two estimates of the int8 peak, so that frac is not flattered
            peak = max(int8_measured or 0.0, peaks["bf16_tflops"] * 2.0) if i8 else peaks["bf16_tflops"]
            ach = total_pairs * 2.0 * B / (k_ms / 1e3) / 1e12
            n_tiles = sum(t + 1 for t in range((samples + 127) // 128))
            executed = n_tiles * 128 * 128 * ri.n_nodes_padded * 2 * 2.0 / (k_ms / 1e3) / 1e12
            kname = "k_unweighted_tc2<u8>" if i8 else "k_unweighted_tc2<bf16>"
            roofline = {"bound": "tensor", "kernel": kname,
                        "achieved": ach, "peak": peak,
                        "unit": "TOP/s" if i8 else "TFLOP/s", "frac": ach / peak,
                        "traffic": load_traffic(args.config, kname, world), "kernel_ms": k_ms,
                        "executed_tflops": executed, "executed_frac": executed / peak,
                        "achieved_vs_bf16_peak": ach / peaks["bf16_tflops"],
                        "int8_peak_measured_cublaslt": int8_measured,
                        "peak_source": (f"max(2 x {peaks_kind} bf16_tflops [kind::i8 issues at twice the kind::f16 rate], "
                                        "int8 GEMM measured live with torch._int_mm 8192^3 best of 10)" if i8 else
                                        f"{peaks_kind} bf16_tflops") + " (burst; kernel timed alone, one launch over all tiles)",
                        "note": "algorithmic ops = 2*B per pair; the kernel executes 2 operand planes and full "
                                "diagonal tiles over the padded contraction length, reported as executed_* (the "
                                "CTA-pair kernel also runs one masked tile above each second diagonal tile, not counted)"}
    if rank == 0:
        if not args.no_cpu:
            cpu = cpu_baseline(tree, csr, weighted, target_s=args.ref_seconds)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        out = {"metric": f"UniFrac sample-pairs/sec ({mode})", "value": value, "unit": "sample-pairs/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
               "higher_is_better": True, "scaling": "strong" if args.fixed_size else "weak", "vs_baseline": None,
               "dtype": "f32 numerator tiles, f64 embedding/denominators/output" if weighted else
                        ("bf16 x bf16 -> f32 (tcgen05 kind::f16), f64 row sums/epilogue/output" if info.operand_kind == 1 else
                         "u8 x u8 -> s32 (tcgen05 kind::i8, exact), s64 chunk / row sums, f32 ratio, f64 output"),
               "data": "synthetic",
               "config": workload_config(args.config, mode, leaves, samples, density, world, shard),
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "roofline": roofline, "cpu_baseline": cpu,
               "stages_ms": {"h2d": ei.h2d_ms, "embed": info.embed_ms, "pairs_kernels_sum": info.pairs_ms,
                             "fixup_sum": info.fixup_ms, "run": info.run_ms},
               "timing": "value: CUDA events on the engine's streams (embedding start -> last band done), summed over "
                         "steps, max over ranks; wall clock of the same region = %.3f ms/step" % (1e3 * wall_s / args.steps),
               "flagged_pairs": int(info.flagged_pairs), "bands": int(info.n_bands_total),
               "tree_height": int(info.tree_height)}
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--ref-seconds", type=float, default=12.0, help="CPU work per reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (configs whose output exceeds host RAM)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the kernel-alone timing leg")
    ap.add_argument("--min-warmup", type=int, default=3)
    ap.add_argument("--fixed-size", action="store_true",
                    help="N>1: keep the config's sample count (strong scaling) instead of growing it with sqrt(N)")
    ap.add_argument("--shard-embed", action="store_true", help="N>1: force the sharded embedding + all-gather")
    ap.add_argument("--no-shard-embed", action="store_true",
                    help="N>1: every rank rebuilds the whole embedding (no NCCL all-gather)")
    ap.add_argument("--uw-kernel", default="u8", choices=["u8", "bf16"],
                    help="operand encoding of the unweighted tensor-core kernel (FRC_FLAG_UW_BF16)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        port = os.environ.get("MASTER_PORT", "29517")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    args.warmup = max(args.warmup, args.min_warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_ours(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
