"""The C-ABI library loads, exports what include/frcfrc_cuda.h declares, and its structs match
the Python binding (CPU; no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "frcfrc_cuda.h")


def test_exports_every_declared_symbol(built):
    from frackyfrac_b200 import engine

    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(frc_[a-z_0-9]+)\s*\(", src)))
    assert set(declared) == set(engine.EXPORTS)
    L = engine.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.frc_abi_version() == 2


def test_struct_layout_matches_binding(built, tmp_path):
    from frackyfrac_b200 import engine

    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "frcfrc_cuda.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(frc_tree_t), sizeof(frc_csr_t), sizeof(frc_opts_t),'
                    ' sizeof(frc_info_t), offsetof(frc_opts_t, band_rows), offsetof(frc_info_t, flagged_pairs));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()))
    want = [C.sizeof(engine._Tree), C.sizeof(engine._Csr), C.sizeof(engine._Opts), C.sizeof(engine._Info),
            engine._Opts.band_rows.offset, engine._Info.flagged_pairs.offset]
    assert got == want


def test_no_cpu_fallback(built):
    """Without an sm_100 device the product path must fail loudly, never compute."""
    import torch

    from frackyfrac_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    with pytest.raises(engine.FrcError):
        engine.Context(0)
    with pytest.raises(engine.FrcError):
        engine.unifrac(np.array([-1, 0, 0], np.int32), np.array([0, 1, 1.0]), np.array([0, 1, 2], np.int64),
                       np.array([1, 2], np.int32), np.array([1.0, 1.0]), False)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under frackyfrac_b200/ may reference it."""
    pkg = os.path.join(ROOT, "frackyfrac_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)


def test_band_plan_covers_the_triangle(built):
    from frackyfrac_b200 import engine

    for n in (0, 1, 2, 127, 128, 129, 1000, 5000, 14142):
        total = n * (n - 1) // 2 if n >= 2 else 0
        for world in (1, 2, 3, 8):
            for flags in (0, engine.FLAG_NO_D2H):
                seen = []
                for r in range(world):
                    f, c = engine.plan_bands(n, r, world, flags=flags)
                    assert (np.diff(f) > 0).all()
                    seen += list(zip(f.tolist(), c.tolist()))
                seen.sort()
                pos = 0
                for f, c in seen:
                    assert f == pos and c > 0
                    pos += c
                assert pos == total
    # bands hold (nearly) equal pair counts, so the ranks are balanced: within 10 % of the mean
    for flags in (0, engine.FLAG_NO_D2H):
        per = np.array([engine.plan_bands(14142, r, 8, flags=flags)[1].sum() for r in range(8)], float)
        assert per.max() / per.mean() < 1.10 and per.min() / per.mean() > 0.90, per
    # distances stay in HBM: one launch on one GPU, two per rank otherwise; with D2H about 8 per rank
    assert len(engine.plan_bands(5000, 0, 1, flags=engine.FLAG_NO_D2H)[0]) == 1
    assert sum(len(engine.plan_bands(14142, r, 8, flags=engine.FLAG_NO_D2H)[0]) for r in range(8)) == 16
    assert len(engine.plan_bands(5000, 0, 1)[0]) == 9   # 8 of equal pair count, the first cut again (early first D2H)
    # explicit band_rows: uniform bands of whole tiles
    f, c = engine.plan_bands(700, 0, 1, band_rows=128)
    assert len(f) == 6 and c[0] == 128 * 127 // 2


def test_host_pool_under_tsan(tmp_path):
    """The context's worker pool (csrc/host_pool.h: spinning workers, run / start / wait) drives frc_create's table
    staging and frc_next's widening; a race there would be a rare hang on the GPU box.  ThreadSanitizer harness,
    CPU only (tests/pool_tsan.cpp)."""
    import shutil

    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = tmp_path / "pool_tsan"
    r = subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread",
                        "-I", os.path.join(here, "..", "frackyfrac_b200", "csrc"), "-o", str(exe),
                        os.path.join(here, "pool_tsan.cpp")], capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in (r.stderr + r.stdout).lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr


def test_host_widening_of_fp32_bands(built):
    """The host half of the fp32 PCIe leg (csrc/wire.cu widen_band: AVX2 cvtps2pd, plain or non-temporal stores)
    must be exact for every float — subnormals, infinities, NaN, signed zeros — at any length and alignment."""
    from frackyfrac_b200 import engine

    L = engine.lib()
    L.frc_debug_widen.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    L.frc_debug_widen.restype = None
    rng = np.random.default_rng(77)
    bits = rng.integers(0, 2**32, 70_000, dtype=np.uint64).astype(np.uint32)
    special = np.array([0, 0x80000000, 1, 0x007FFFFF, 0x00800000, 0x7F7FFFFF, 0x7F800000, 0xFF800000, 0x7FC00000,
                        0x3F800000, 0x3DCCCCCD], np.uint32)
    src_all = np.concatenate([special, bits]).view(np.float32)
    for n in (0, 1, 7, 8, 9, 31, 1000, 16384, 16385, len(src_all) - 3):
        for s_off in (0, 1, 3):
            for d_off in (0, 1, 2, 3):
                for stream in (0, 1):
                    src = src_all[s_off:s_off + n]
                    buf = np.full(n + 8, -7.0)
                    dst = buf[d_off:d_off + n]
                    L.frc_debug_widen(src.ctypes.data, dst.ctypes.data, n, stream)
                    with np.errstate(invalid="ignore"):   # (signalling NaNs among the random bit patterns)
                        want = src.astype(np.float64)
                    assert np.array_equal(dst.view(np.uint64), want.view(np.uint64)), (n, s_off, d_off, stream)
                    assert (buf[:d_off] == -7.0).all() and (buf[d_off + n:] == -7.0).all()   # nothing outside the range


def test_argument_errors_come_before_the_device(built):
    """frc_create checks its options and the shape of its inputs before it touches CUDA: on a box without a GPU the
    caller still gets the precise FRC_ERR_ARG (1) for a bad call, and a device error only for a good one."""
    import torch

    from frackyfrac_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("runs on the CPU box (on a GPU the same checks are covered by test_bad_arguments_are_errors)")
    parent, length = np.array([-1, 0, 0], np.int32), np.array([0, 1, 1.0])
    rp, col, val = np.array([0, 1, 2], np.int64), np.array([1, 2], np.int32), np.array([1.0, 1.0])

    def code(**kw):
        a = dict(parent=parent, length=length, row_ptr=rp, col=col, val=val, weighted=False)
        a.update(kw)
        with pytest.raises(engine.FrcError) as e:
            engine.Job(a.pop("parent"), a.pop("length"), a.pop("row_ptr"), a.pop("col"), a.pop("val"), **a)
        return e.value.code, str(e.value)

    assert code(normalize=False)[0] == 1 and "-l can only be used with weighted" in code(normalize=False)[1]   # frcfrc.go:84-86
    assert code(path=7)[0] == 1
    assert code(rank=3, world=2)[0] == 1
    assert code(band_rows=-5)[0] == 1
    assert code(parent=np.array([0, 0, 0], np.int32))[0] == 1           # root must have parent -1
    assert code(row_ptr=np.array([0, 2, 1], np.int64), col=np.array([1], np.int32), val=np.array([1.0]))[0] == 1
    assert code(row_ptr=np.array([1, 1, 2], np.int64))[0] == 1          # row_ptr[0] != 0
    good = code()
    assert good[0] in (2, 5), good                                      # only now the missing device shows
