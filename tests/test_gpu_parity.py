"""Parity of the CUDA path against the oracle, through the C ABI (GPU box only)."""
import os
import subprocess

import numpy as np
import pytest

from tests import kat
from tests.helpers import gpu_flat, oracle_flat, rel_err

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _read(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return f.read()


# ------------------------------------------------------------------ known answers
@pytest.mark.parametrize("case", kat.UNIT, ids=[c["name"] for c in kat.UNIT])
def test_unit_kat_bit_exact(gpu_ctx, case):
    """frcfrc/unifrac_test.go: reflect.DeepEqual on float64, i.e. bit-exact."""
    from frackyfrac_b200 import engine, hostlib

    tree = hostlib.Tree(case["tree"])
    tab = hostlib.Table(kat.sparse_text(case["abnd"]), True)
    rp, col, val = tab.resolve(tree)
    got = engine.unifrac(tree.parent, tree.length, rp, col, val, case["weighted"], ctx=gpu_ctx)
    assert got.tolist() == case["want"]


@pytest.mark.parametrize("fixture,sparse,weighted", kat.CLI)
def test_cli_golden_through_abi(gpu_ctx, fixture, sparse, weighted):
    """testdata/run.sh: output text must equal *.want byte for byte."""
    from frackyfrac_b200 import engine, hostlib

    tree = hostlib.Tree(_read(fixture + ".tree"))
    tab = hostlib.Table(_read(fixture + (".sparse" if sparse else ".dense")), sparse)
    rp, col, val = tab.resolve(tree)
    got = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, ctx=gpu_ctx)
    text = "".join(hostlib.format_go(v) + "\n" for v in got)
    assert text == _read(fixture + ".want")


@pytest.mark.parametrize("fixture,sparse,weighted", kat.CLI)
def test_cli_binary_golden(built, tmp_path, fixture, sparse, weighted):
    """The C++ stand-in for the Go binary, run exactly as testdata/run.sh runs frcfrc."""
    from frackyfrac_b200 import hostlib

    out = tmp_path / "got"
    cmd = [hostlib.CLI_PATH, "-i", os.path.join(GOLDEN, fixture + (".sparse" if sparse else ".dense")),
           "-t", os.path.join(GOLDEN, fixture + ".tree"), "-o", str(out)]
    if sparse:
        cmd.insert(1, "-s")
    if weighted:
        cmd.insert(1, "-w")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == _read(fixture + ".want")


@pytest.mark.parametrize("fixture,suffix,sparse,tag,weighted,normalize", kat.SYNTH)
def test_synthetic_golden_through_the_gpu(built, gpu_ctx, tmp_path, fixture, suffix, sparse, tag, weighted, normalize):
    """Committed synthetic fixtures (tests/golden/make_golden.py): the CLI stand-in reproduces the bytes (these sizes
    take the exact fp64 kernels; `-l` prints what the reference prints, FRCFRC_L=documented the documented variant),
    and the fast kernels land within 1e-5 of the committed values."""
    from frackyfrac_b200 import engine, hostlib

    want_text = _read(f"{fixture}.{tag}.want")
    out = tmp_path / "got"
    cmd = [hostlib.CLI_PATH, "-t", os.path.join(GOLDEN, fixture + ".tree"), "-i", os.path.join(GOLDEN, fixture + suffix),
           "-o", str(out)]
    cmd += (["-s"] if sparse else []) + (["-w"] if weighted else []) + (["-l"] if normalize != 1 else [])
    env = dict(os.environ, FRCFRC_L="documented") if normalize == 2 else dict(os.environ)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == want_text
    tree = hostlib.Tree(_read(fixture + ".tree"))
    rp, col, val = hostlib.Table(_read(fixture + suffix), sparse=sparse).resolve(tree)
    want = np.array([float(x) for x in want_text.split()])
    # through the ABI: bit-identical on the exact path in every mode ...
    exact = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, normalize, path=engine.PATH_EXACT, ctx=gpu_ctx)
    assert "".join(hostlib.format_go(v) + "\n" for v in exact) == want_text
    if normalize == 0:
        # ... and the reference's -l (unsorted merge-join) has no fast kernel: asking for one is an error
        with pytest.raises(engine.FrcError, match="as coded"):
            engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, 0, path=engine.PATH_FAST, ctx=gpu_ctx)
        return
    got = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, normalize, path=engine.PATH_FAST, ctx=gpu_ctx)
    assert rel_err(got, want).max() < 1e-5
    assert (got[want == 0] == 0).all()


def test_cli_binary_compressed_files(built, tmp_path):
    """aio.Open / aio.Create pick the codec from the suffix (frcfrc.go:93,100-109): gzip'd tree and table in,
    gzip'd and zstd'd distances out; the decoded output equals the golden .want, also with a multi-block output."""
    import gzip

    from frackyfrac_b200 import hostlib, synth

    for name in ("wtd.tree", "wtd.sparse"):
        (tmp_path / (name + ".gz")).write_bytes(gzip.compress(open(os.path.join(GOLDEN, name), "rb").read()))
    for suffix, decode in ((".gz", lambda p: gzip.decompress(p.read_bytes())), (".zst", lambda p: hostlib.read_file(str(p)))):
        out = tmp_path / ("got" + suffix)
        r = subprocess.run([hostlib.CLI_PATH, "-w", "-s", "-i", str(tmp_path / "wtd.sparse.gz"),
                            "-t", str(tmp_path / "wtd.tree.gz"), "-o", str(out), "-p", "3"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert decode(out).decode() == _read("wtd.want")
    # 1500 samples -> 1.1 M lines (~20 MB of text): several compressed blocks from 4 workers
    tree = synth.random_tree(400, 7)
    rp, col, val = synth.random_table(tree, 1500, 0.05, 8)
    (tmp_path / "big.tree").write_text(synth.to_newick(tree))
    (tmp_path / "big.sparse.gz").write_bytes(gzip.compress(synth.to_sparse_text(tree, rp, col, val).encode()))
    outs = []
    for name in ("big.out", "big.out.gz"):
        r = subprocess.run([hostlib.CLI_PATH, "-s", "-i", str(tmp_path / "big.sparse.gz"), "-t", str(tmp_path / "big.tree"),
                            "-o", str(tmp_path / name), "-p", "4"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append((tmp_path / name).read_bytes())
    assert len(outs[0]) > 8_000_000 and gzip.decompress(outs[1]) == outs[0]


# ------------------------------------------------------------------ exact path
@pytest.mark.parametrize("weighted,normalize", [(False, 1), (True, 1), (True, 2), (True, 0)])
@pytest.mark.parametrize("shape", ["random", "caterpillar", "balanced"])
def test_exact_path_bit_identical(gpu_ctx, weighted, normalize, shape):
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(300, 11, shape=shape)
    csr = synth.random_table(tree, 40, 0.05, 12, integer_counts=False)
    want = oracle_flat(tree, csr, weighted, normalize)
    got = gpu_flat(tree, csr, weighted, normalize, path=engine.PATH_EXACT, ctx=gpu_ctx)
    assert np.array_equal(got, want), f"max rel err {rel_err(got, want).max()}"


def test_exact_multifurcating_and_shared_names(gpu_ctx):
    """Children are summed in file order; a name on several leaves feeds all of them (A6)."""
    from frackyfrac_b200 import engine, hostlib
    from oracle import oracle as orc

    nwk = "((a:0.1,b:0.25,c:0.7,a:0.3)x:0.5,(d:1.5,(e:0.125,f:0.3):0.2,x:0.9):0.05,g:2)r:0.75;"
    tab_text = "a:0.3 d:7 x:5\nb:1.1 e:2 g:0.01\nc:3 f:1e-3 a:2\n\nd:1 e:1 f:1 g:1 x:2\n"
    tree = hostlib.Tree(nwk)
    tab = hostlib.Table(tab_text, True)
    rp, col, val = tab.resolve(tree)
    otree, otab = orc.Tree.parse(nwk), orc.Table.parse(tab_text, True)
    for weighted in (False, True):
        want = orc.unifrac(otab, otree, weighted, 1, 1)
        got = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, path=engine.PATH_EXACT, ctx=gpu_ctx)
        assert np.array_equal(got, want, equal_nan=True)
        assert np.isnan(want).sum() == 0 and len(want) == 10


def test_empty_samples_give_nan(gpu_ctx):
    """Two empty samples: 0/0 = NaN (A8); empty vs non-empty = 1."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(50, 3)
    rp, col, val = synth.random_table(tree, 4, 0.1, 4)
    m = rp[1]
    rp = np.array([0, m, m, 2 * m, 2 * m], np.int64)  # samples 1 and 3 empty
    for weighted in (False, True):
        for path in (engine.PATH_EXACT, engine.PATH_FAST):
            got = engine.unifrac(tree.parent, tree.length, rp, col[:2 * m], val[:2 * m], weighted, path=path, ctx=gpu_ctx)
            want = oracle_flat(tree, (rp, col[:2 * m], val[:2 * m]), weighted)
            assert np.isnan(want[4]) and np.isnan(got[4])  # pair (3,1)
            assert rel_err(got, want).max() < 1e-5


# ------------------------------------------------------------------- fast paths
# the two operand encodings of the unweighted tensor-core kernel (include/frcfrc_cuda.h FRC_FLAG_UW_BF16)
UW_KERNELS = [("u8", 0, 2), ("bf16", 2, 1), ("bits", 8, 3)]  # (id, flags, info.operand_kind)
uw_kernels = pytest.mark.parametrize("uw", UW_KERNELS, ids=[k[0] for k in UW_KERNELS])


@pytest.mark.parametrize("weighted,normalize", [(False, 1), (True, 1), (True, 2)])
@uw_kernels
def test_fast_path_within_tolerance(gpu_ctx, weighted, normalize, uw):
    """north_star: every distance within 1e-5 relative of the reference."""
    from frackyfrac_b200 import engine, synth

    if weighted and uw[0] == "bf16":
        pytest.skip("operand encoding only concerns unweighted")
    tree = synth.random_tree(1000, 21)
    csr = synth.random_table(tree, 300, 0.02, 22)
    want = oracle_flat(tree, csr, weighted, normalize)
    got = gpu_flat(tree, csr, weighted, normalize, path=engine.PATH_FAST, ctx=gpu_ctx, flags=uw[1])
    e = rel_err(got, want)
    assert e.max() < 1e-5, f"max rel err {e.max():.3e}"


@uw_kernels
@pytest.mark.parametrize("lengths", ["heavy_tail", "zeros_and_tiny", "constant", "one_huge"])
def test_fast_unweighted_length_distributions(gpu_ctx, uw, lengths):
    """Branch lengths spanning many binades, exact zeros, a single dominating branch: the u8 block
    floating point must keep every length to ~2^-20 relative, the bf16 split to 2^-17."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(1500, 91)
    rng = np.random.default_rng(92)
    L = tree.length.copy()
    if lengths == "heavy_tail":
        L = np.exp(rng.normal(-3.0, 4.0, len(L)))          # ~ 2^-30 .. 2^20 relative spread
    elif lengths == "zeros_and_tiny":
        L[rng.random(len(L)) < 0.3] = 0.0
        L[rng.random(len(L)) < 0.1] *= 1e-9
    elif lengths == "constant":
        L[:] = 0.37
    elif lengths == "one_huge":
        L[int(tree.leaf_ids[5])] = 1e6
    L[0] = 0.0
    tree = synth.Tree(tree.parent, L, tree.leaf_ids, tree.names)
    csr = synth.random_table(tree, 260, 0.03, 93)
    want = oracle_flat(tree, csr, False)
    with engine.Job(tree.parent, tree.length, *csr, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx,
                    flags=uw[1]) as job:
        got = np.concatenate([a for _, a in job.chunks()])
        # (the bits-fed kernel needs the integer mode: trees whose chunk scales span > 2^16 use operands)
        assert job.info().operand_kind == uw[2] or (uw[0] == "bits" and job.info().operand_kind == 2)
    e = rel_err(got, want)
    # bf16 hi/lo planes with fp32 TMEM accumulation lose small addends next to huge ones: on the
    # heavy-tailed tree that encoding reaches 2e-5 (measured on B200), which is why u8 block floating
    # point with exact integer accumulation is the default; everywhere else both meet 1e-5.
    tol = 5e-5 if (uw[0] == "bf16" and lengths == "heavy_tail") else 1e-5
    assert e.max() < tol, f"{lengths}: max rel err {e.max():.3e}"


def test_fast_unweighted_u8_is_tighter_than_bf16(gpu_ctx):
    """The u8 path accumulates exactly (int32) and quantises lengths to ~2^-20: its worst error must
    sit well below the 1e-5 budget and not above the bf16 path's."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(4000, 95)
    csr = synth.random_table(tree, 384, 0.02, 96)
    want = oracle_flat(tree, csr, False)
    e8 = rel_err(gpu_flat(tree, csr, False, path=engine.PATH_FAST, ctx=gpu_ctx), want).max()
    e16 = rel_err(gpu_flat(tree, csr, False, path=engine.PATH_FAST, ctx=gpu_ctx, flags=engine.FLAG_UW_BF16), want).max()
    assert e8 < 2e-6 and e8 <= e16 * 1.5, (e8, e16)


@pytest.mark.parametrize("normalize", [1, 2])
def test_fast_weighted_small_distances(gpu_ctx, normalize):
    """Samples that differ by a relative 1e-2 .. 1e-7 in a few abundances: the fp32 L1 tiles lose
    6e-8 / d relative there, so pairs with d < 1/32 must come out of the fixed-point fix-up pass
    (normalize=2: flag -l as documented, raw counts)."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(1200, 85)
    rp, col, val = synth.random_table(tree, 160, 0.04, 86, integer_counts=False)
    m = rp[1]
    rng = np.random.default_rng(87)
    for k, eps in enumerate((1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7), start=1):
        col[k * m:(k + 1) * m] = col[:m]
        val[k * m:(k + 1) * m] = val[:m] * (1.0 + eps * rng.random(m) * (rng.random(m) < 0.2))
    want = oracle_flat(tree, (rp, col, val), True, normalize)
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=True, normalize=normalize,
                    path=engine.PATH_FAST, ctx=gpu_ctx) as job:
        got = np.concatenate([a for _, a in job.chunks()])
        flagged = job.info().flagged_pairs
    assert (want[:21] < 1 / 32).all() and want[:21].min() < 1e-6   # the seven near-copies, pairwise
    assert flagged >= 21
    e = rel_err(got, want)
    assert e.max() < 1e-5, f"max rel err {e.max():.3e}"
    # fixed-point recompute, then one fp32 rounding on the PCIe leg (wire.cu)
    assert (np.abs(got[:21] - want[:21]) <= 1.2e-7 * want[:21] + 1e-15).all()


def _f32_stream(job):
    """The job's float32 bands (frc_next_f32) with their exceptions applied, as one float64 vector."""
    out = []
    for first, a in job.chunks_f32():
        d = a.astype(np.float64)
        xi, xv = job.exceptions()
        d[xi - first] = xv
        out.append(d)
    return np.concatenate(out) if out else np.zeros(0)


@pytest.mark.parametrize("weighted", [False, True])
def test_f32_stream_equals_f64_stream(gpu_ctx, weighted):
    """Fast-path kernels store fp32; a band crosses PCIe as 4 bytes per pair.  frc_next_f32 hands the pinned
    slot out as it is, frc_next widens the same values on the host: both streams carry identical numbers, and
    neither depends on the band decomposition."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(1500, 301)
    rp, col, val = synth.random_table(tree, 1100, 0.03, 302)
    m = rp[1]
    col[m:2 * m] = col[:m]            # an identical pair (d = 0 exactly) and a near copy (fix-up territory)
    val[m:2 * m] = val[:m]
    col[2 * m:3 * m] = col[:m]
    val[2 * m:3 * m] = val[:m] * (1.0 + 1e-3 * (np.arange(m) % 5 == 0))

    def run(f32, band_rows=0):
        with engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=gpu_ctx,
                        band_rows=band_rows) as job:
            got = _f32_stream(job) if f32 else np.concatenate([a for _, a in job.chunks()])
            return got, job.info()

    narrow, ni = run(True)
    ragged, _ = run(True, band_rows=128)       # many small bands, odd lengths and alignments
    wide, wi = run(False)
    pairs = 1100 * 1099 // 2
    assert ni.d2h_bytes == 4 * pairs and wi.d2h_bytes == 4 * pairs and ni.value_bytes == 4
    assert narrow[0] == 0.0 and wide[0] == 0.0
    assert np.array_equal(narrow, ragged)
    assert np.array_equal(narrow, wide, equal_nan=True)
    assert np.array_equal(narrow, narrow.astype(np.float32).astype(np.float64))
    assert rel_err(narrow, oracle_flat(tree, (rp, col, val), weighted)).max() < 1e-5
    # the two calls cannot be mixed on one pass; the exact path is float64 only
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=weighted, path=engine.PATH_FAST, ctx=gpu_ctx) as job:
        job.next_raw()
        with pytest.raises(engine.FrcError, match="mixed"):
            job.next_raw_f32()
    with engine.Job(tree.parent, tree.length, rp[:4], col[:rp[3]], val[:rp[3]], weighted=weighted, path=engine.PATH_EXACT,
                    ctx=gpu_ctx) as job:
        assert job.info().value_bytes == 8
        with pytest.raises(engine.FrcError, match="exact path"):
            job.next_raw_f32()


def test_f64_stream_delivers_large_bands_in_pieces(gpu_ctx):
    """frc_next widens into a cache-sized buffer: a band larger than it (2 M pairs) arrives over several calls,
    contiguous, in order, and with the same values as the default band plan; frc_next_f32 hands it out whole."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(300, 311)
    rp, col, val = synth.random_table(tree, 2500, 0.05, 312)
    pairs = 2500 * 2499 // 2
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx,
                    band_rows=4096) as job:
        runs = [(first, a.copy()) for first, a in job.chunks(copy=False)]
        info = job.info()
    assert info.n_bands_mine == 1 and [len(a) for _, a in runs] == [2 << 20, pairs - (2 << 20)]
    assert [f for f, _ in runs] == [0, 2 << 20]
    one_band = np.concatenate([a for _, a in runs])
    assert np.array_equal(one_band, gpu_flat(tree, (rp, col, val), False, path=engine.PATH_FAST, ctx=gpu_ctx))
    assert rel_err(one_band, oracle_flat(tree, (rp, col, val), False)).max() < 1e-5
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx,
                    band_rows=4096) as job:
        whole = [(first, a) for first, a in job.chunks_f32()]
    assert len(whole) == 1 and whole[0][0] == 0 and np.array_equal(whole[0][1].astype(np.float64), one_band)


def test_values_below_fp32_range_travel_as_exceptions(gpu_ctx):
    """A distance below fp32's range must still arrive: the exact fix-up pass reports it beside the fp32 band
    (wire.cuh), frc_next patches it into the doubles and frc_chunk_exceptions hands it to fp32 consumers."""
    from frackyfrac_b200 import engine, hostlib

    tree = hostlib.Tree("((a:1e-60,b:1e-60):1,(c:1,d:2):0.5);")
    tab = hostlib.Table("a:1\tc:1\nb:1\tc:1\nc:1\td:1\na:1\td:3\n", sparse=True)
    rp, col, val = tab.resolve(tree)
    want = oracle_flat(tree, (rp, col, val), False)
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx) as job:
        got = np.concatenate([a for _, a in job.chunks()])
        info = job.info()
    assert 0 < want[0] < 1e-59
    assert abs(got[0] - want[0]) <= 1e-5 * want[0]
    assert rel_err(got, want).max() < 1e-5
    assert info.d2h_bytes == 4 * 6 and info.exceptions >= 1
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx) as job:
        gen = job.chunks_f32()
        first, a = next(gen)
        xi, xv = job.exceptions()      # (of the band just handed out: valid until the next call)
        assert next(gen, None) is None
    assert first == 0 and a[0] == 0.0 and 0 in xi.tolist() and abs(xv[xi.tolist().index(0)] - want[0]) <= 1e-5 * want[0]


def test_fast_weighted_fixups_in_concurrent_bands(gpu_ctx):
    """Near-duplicate samples spread over MANY bands: consecutive bands run on two streams, so their fix-up
    kernels overlap and must not share a workspace (each stream has its own)."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(900, 185)
    n = 1400
    rp, col, val = synth.random_table(tree, n, 0.04, 186, integer_counts=False)
    m = rp[1]
    rng = np.random.default_rng(187)
    # every 64th sample is a near copy of sample 0: flagged pairs (d < 1/32) in every band of 128 rows
    copies = list(range(64, n, 64))
    for k in copies:
        col[k * m:(k + 1) * m] = col[:m]
        val[k * m:(k + 1) * m] = val[:m] * (1.0 + 1e-4 * rng.random(m) * (rng.random(m) < 0.3))
    want = oracle_flat(tree, (rp, col, val), True)
    for rep in range(3):
        with engine.Job(tree.parent, tree.length, rp, col, val, weighted=True, path=engine.PATH_FAST, ctx=gpu_ctx,
                        band_rows=128) as job:
            got = np.concatenate([a for _, a in job.chunks()])
            info = job.info()
        assert info.n_bands_mine == 11 and info.flagged_pairs >= len(copies) * (len(copies) + 1) // 2
        e = rel_err(got, want)
        assert e.max() < 1e-5, f"rep {rep}: max rel err {e.max():.3e}"
        idx = [i * (i - 1) // 2 + j for i in copies for j in [0] + [c for c in copies if c < i]]
        assert (np.abs(got[idx] - want[idx]) <= 1.2e-7 * want[idx] + 1e-15).all()


def test_flagged_pairs_are_deterministic(gpu_ctx):
    """The exact-recompute threshold of the u8 path (flag_u) is summed in a fixed order on the device: the same
    pairs are flagged on every run and for every band / owner layout."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(1500, 91)
    L = np.exp(np.random.default_rng(92).normal(-3.0, 4.0, tree.n_nodes))   # heavy tail: merged small-length columns
    L[0] = 0.0
    tree = synth.Tree(tree.parent, L, tree.leaf_ids, tree.names)
    csr = synth.random_table(tree, 520, 0.03, 93)
    counts, outs = [], []
    for kw in (dict(), dict(), dict(band_rows=128), dict(band_rows=256)):
        with engine.Job(tree.parent, tree.length, *csr, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx, **kw) as job:
            outs.append(np.concatenate([a for _, a in job.chunks()]))
            counts.append(job.info().flagged_pairs)
    parts = []
    for r in range(2):
        with engine.Job(tree.parent, tree.length, *csr, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx, band_rows=128,
                        rank=r, world=2) as job:
            for _ in job.chunks():
                pass
            parts.append(job.info().flagged_pairs)
    assert len(set(counts)) == 1 and sum(parts) == counts[0], (counts, parts)
    assert all(np.array_equal(outs[0], o) for o in outs[1:])


@pytest.mark.parametrize("slab", [128, 256])
def test_fast_weighted_embedding_in_slabs(gpu_ctx, monkeypatch, slab):
    """The fp64 embedding of the fast weighted path is built slab by slab of samples (bounded memory):
    the slab width must not change a single byte."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(900, 81)
    csr = synth.random_table(tree, 700, 0.03, 82, integer_counts=False)
    whole = gpu_flat(tree, csr, True, path=engine.PATH_FAST, ctx=gpu_ctx)
    monkeypatch.setenv("FRC_WS_SLAB", str(slab))
    slabs = gpu_flat(tree, csr, True, path=engine.PATH_FAST, ctx=gpu_ctx)
    assert np.array_equal(whole, slabs)
    assert rel_err(slabs, oracle_flat(tree, csr, True)).max() < 1e-5


@pytest.mark.parametrize("shape,levels", [("caterpillar", 0), ("caterpillar", 1), ("balanced", 0), ("random", 1),
                                          ("caterpillar", 2), ("random", 2), ("random", 3)])
def test_fast_unweighted_embedding_variants(gpu_ctx, monkeypatch, shape, levels):
    """Both embedding implementations (fused single launch / one launch per level), deep and flat trees (H9)."""
    from frackyfrac_b200 import engine, synth

    monkeypatch.setenv("FRC_EMBED_LEVELS", str(min(levels, 1)))  # 2 = fused embedding feeding the bf16 kernel
    if levels == 2:
        monkeypatch.setenv("FRC_UW_KERNEL", "bf16")
    if levels == 3:  # u8 operands with the fp64 chunk accumulation (used when chunk scales span > 2^16)
        monkeypatch.setenv("FRC_EMBED_LEVELS", "0")
        monkeypatch.setenv("FRC_U8_ACC", "f64")
    tree = synth.random_tree(700, 71, shape=shape)
    csr = synth.random_table(tree, 200, 0.03, 72)
    want = oracle_flat(tree, csr, False)
    with engine.Job(tree.parent, tree.length, *csr, weighted=False, path=engine.PATH_FAST, ctx=gpu_ctx) as job:
        got = np.concatenate([a for _, a in job.chunks()])
        assert job.info().tree_height == {"caterpillar": 699, "balanced": 10, "random": job.info().tree_height}[shape]
    assert rel_err(got, want).max() < 1e-5


@pytest.mark.parametrize("weighted,flags", [(False, 0), (False, 2), (False, 8), (True, 0)])
def test_fast_path_identical_and_near_identical_samples(gpu_ctx, weighted, flags):
    """Identical samples must give exactly 0; near-identical ones stay within 1e-5 relative."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(800, 31)
    rp, col, val = synth.random_table(tree, 130, 0.03, 32)
    m = rp[1]
    col[m:2 * m] = col[:m]          # sample 1 == sample 0
    val[m:2 * m] = val[:m]
    col[2 * m:3 * m] = col[:m]      # sample 2 = sample 0 with one leaf swapped
    val[2 * m:3 * m] = val[:m]
    others = np.setdiff1d(tree.leaf_ids, col[:m])
    col[2 * m] = others[0]
    want = oracle_flat(tree, (rp, col, val), weighted)
    got = gpu_flat(tree, (rp, col, val), weighted, path=engine.PATH_FAST, ctx=gpu_ctx, flags=flags)
    assert want[0] == 0.0 and got[0] == 0.0
    assert rel_err(got, want).max() < 1e-5


@pytest.mark.parametrize("weighted,flags", [(False, 0), (False, 2), (False, 8), (True, 0)])
def test_fast_path_multi_band_and_sharded(gpu_ctx, weighted, flags):
    """Band streaming order, and world=2 band sharding: union of ranks == single rank, same bytes."""
    from frackyfrac_b200 import engine, synth
    import functools

    engine = type("E", (), {"unifrac": staticmethod(functools.partial(engine.unifrac, flags=flags)),
                            "PATH_FAST": engine.PATH_FAST})

    tree = synth.random_tree(500, 41)
    rp, col, val = synth.random_table(tree, 700, 0.02, 42)
    want = oracle_flat(tree, (rp, col, val), weighted)
    one = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, path=engine.PATH_FAST, band_rows=128, ctx=gpu_ctx)
    assert rel_err(one, want).max() < 1e-5
    parts = [engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, path=engine.PATH_FAST, band_rows=128,
                            rank=r, world=2, ctx=gpu_ctx) for r in range(2)]
    assert np.array_equal(parts[0] + parts[1], one)
    assert ((parts[0] != 0) & (parts[1] != 0)).sum() == 0
    whole = engine.unifrac(tree.parent, tree.length, rp, col, val, weighted, path=engine.PATH_FAST, band_rows=1024, ctx=gpu_ctx)
    assert np.array_equal(whole, one), "results must not depend on the band decomposition"


def test_stream_order_and_early_destroy(gpu_ctx):
    """frc_next yields strictly increasing contiguous runs; destroying mid-stream is legal."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(200, 51)
    rp, col, val = synth.random_table(tree, 600, 0.05, 52)
    with engine.Job(tree.parent, tree.length, rp, col, val, weighted=True, path=engine.PATH_FAST, band_rows=128, ctx=gpu_ctx) as job:
        expect = 0
        for first, a in job.chunks():
            assert first == expect
            expect += len(a)
        assert expect == 600 * 599 // 2
        info = job.info()
        assert info.n_bands_mine == 5 and info.kernel_launches > 0
    job = engine.Job(tree.parent, tree.length, rp, col, val, weighted=False, path=engine.PATH_FAST, band_rows=128, ctx=gpu_ctx)
    next(job.chunks())
    job.close()  # the consumer's `break`
    again = engine.unifrac(tree.parent, tree.length, rp, col, val, False, ctx=gpu_ctx)
    assert len(again) == 600 * 599 // 2


def test_fewer_than_two_samples_give_an_empty_stream(gpu_ctx):
    """A9: IterPairs over fewer than two samples yields nothing (common/common.go:21-31): the job is valid, the stream
    ends at once, on every path."""
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(40, 63)
    rp, col, val = synth.random_table(tree, 1, 0.2, 64)
    for n in (0, 1):
        for weighted in (False, True):
            for path in (engine.PATH_AUTO, engine.PATH_FAST, engine.PATH_EXACT):
                with engine.Job(tree.parent, tree.length, rp[:n + 1], col[:rp[n]], val[:rp[n]], weighted=weighted, path=path,
                                ctx=gpu_ctx) as job:
                    assert list(job.chunks()) == []
                    info = job.info()
                    assert info.n_pairs_total == 0 and info.n_bands_total == 0
                got = engine.unifrac(tree.parent, tree.length, rp[:n + 1], col[:rp[n]], val[:rp[n]], weighted, path=path, ctx=gpu_ctx)
                assert len(got) == 0


def test_bad_arguments_are_errors(gpu_ctx):
    from frackyfrac_b200 import engine, synth

    tree = synth.random_tree(20, 61)
    rp, col, val = synth.random_table(tree, 3, 0.2, 62)
    internal = int(np.setdiff1d(np.arange(tree.n_nodes), tree.leaf_ids)[0])
    bad_col = col.copy(); bad_col[0] = internal
    with pytest.raises(engine.FrcError):
        engine.unifrac(tree.parent, tree.length, rp, bad_col, val, False, ctx=gpu_ctx)
    bad_val = val.copy(); bad_val[1] = 0.0
    with pytest.raises(engine.FrcError):
        engine.unifrac(tree.parent, tree.length, rp, col, bad_val, True, ctx=gpu_ctx)
    bad_parent = tree.parent.copy(); bad_parent[3] = 7
    with pytest.raises(engine.FrcError):
        engine.unifrac(bad_parent, tree.length, rp, col, val, False, ctx=gpu_ctx)
    # parent < child everywhere but not a pre-order numbering: node 3 hangs under 1 after 2 closed that subtree
    with pytest.raises(engine.FrcError, match="pre-order"):
        engine.unifrac(np.array([-1, 0, 0, 1], np.int32), np.ones(4), np.array([0, 1, 2], np.int64),
                       np.array([2, 3], np.int32), np.ones(2), False, ctx=gpu_ctx)
    # a leaf listed twice in a row: an error where values are used
    dup_col = col.copy(); dup_col[1] = dup_col[0]
    with pytest.raises(engine.FrcError, match="twice"):
        engine.unifrac(tree.parent, tree.length, rp, dup_col, val, True, ctx=gpu_ctx)
    with pytest.raises(engine.FrcError):  # -l only with -w (frcfrc.go:84-86)
        engine.unifrac(tree.parent, tree.length, rp, col, val, False, normalize=False, ctx=gpu_ctx)
    ok = engine.unifrac(tree.parent, tree.length, rp, col, val, False, ctx=gpu_ctx)
    assert len(ok) == 3
