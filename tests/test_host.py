"""The C++ host side above the C ABI against the oracle's independent readers (CPU)."""
import os
import subprocess

import numpy as np
import pytest

from tests import kat

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_newick_flattening_matches_oracle(built):
    from frackyfrac_b200 import hostlib, synth
    from oracle import oracle as orc

    texts = ["(s2:3,s1:1,s3:5);", "((s1:1,s2:3):2,(s3:2,s4:5):1);", " ( a : 1.5 ,\n(b:2e-3,c)x:4 [comment] )root:0.25 ;",
             "('it''s':1,b:2);", "((a,b),(c,(d,e)));", "x;"]
    for shape in ("random", "caterpillar", "balanced"):
        texts.append(synth.to_newick(synth.random_tree(257, 9, shape=shape)))
    for txt in texts:
        h = hostlib.Tree(txt)
        p, l = orc.Tree.parse(txt).flatten()
        assert np.array_equal(h.parent, p) and np.array_equal(h.length, l), txt[:40]
        assert h.parent[0] == -1 and (h.parent[1:] < np.arange(1, len(p))).all()
    assert hostlib.Tree("('it''s':1,b:2);").names()[1] == "it's"
    # branch-length tokens: the from_chars fast path and the strtod path must agree with the oracle's strtod
    forms = ["0", "7", "0.1", ".5", "5.", "1e-3", "1E5", "2.5e+3", "-2", "-.25", "+3", "0x1p3", "1e999", "1e-999",
             "inf", "-inf", "Infinity", "nan", "0.30000000000000004", "123456789012345678901234567890",
             "4.9406564584124654e-324", "1.7976931348623157e308", "00012.5000"]
    txt = "(" + ",".join(f"n{k}:{f}" for k, f in enumerate(forms)) + ");"
    h, (p, l) = hostlib.Tree(txt), orc.Tree.parse(txt).flatten()
    assert np.array_equal(h.parent, p) and np.array_equal(h.length, l, equal_nan=True)
    assert h.length[1:].tolist()[:9] == [0.0, 7.0, 0.1, 0.5, 5.0, 1e-3, 1e5, 2500.0, -2.0]
    for bad_len in ["1d5", "1e", "--1", "1.2.3", "e5", "0x", "1_000"]:
        with pytest.raises(hostlib.HostError):
            hostlib.Tree(f"(a:{bad_len},b:1);")
        with pytest.raises(orc.OracleError):
            orc.Tree.parse(f"(a:{bad_len},b:1);")
    for bad in ["(a,b)", "(a,b));", "((a,b);", "(a:x,b);", ""]:
        with pytest.raises(hostlib.HostError):
            hostlib.Tree(bad)
        with pytest.raises(orc.OracleError):
            orc.Tree.parse(bad)


def test_table_readers_match_reference_tests_and_oracle(built):
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    dense = "   aa  bbbb    \n1\t2\n 3  \t  4 \t\n"            # parser_test.go:10
    sparse = "a:11 b:222  \n  b:32 c:7\n\nd:1\tc:4\ta:10\n"    # parser_test.go:29
    assert hostlib.Table(dense, False).maps() == [{"aa": 1, "bbbb": 2}, {"aa": 3, "bbbb": 4}]
    assert hostlib.Table(sparse, True).maps() == [{"a": 11, "b": 222}, {"b": 32, "c": 7}, {}, {"d": 1, "c": 4, "a": 10}]
    rng = np.random.default_rng(3)
    names = [f"sp{k}" for k in range(30)] + ["w:x", "dup", "dup"]
    for _ in range(20):
        rows = []
        for _ in range(rng.integers(1, 6)):
            rows.append(" ".join(f"{names[k]}:{rng.integers(1, 99)}" for k in rng.integers(0, len(names), rng.integers(0, 9))))
        txt = "\n".join(rows) + "\n"
        assert hostlib.Table(txt, True).maps() == orc.Table.parse(txt, True).maps()
        hdr = " ".join(names[:6] + ["dup", "dup"])
        body = "\n".join(" ".join(str(rng.integers(0, 4)) for _ in range(8)) for _ in range(4))
        txt = hdr + "\n" + body + "\r\n"
        assert hostlib.Table(txt, False).maps() == orc.Table.parse(txt, False).maps()
    for bad, sp in [("a\n", True), (":\n", True), ("a:0\n", True), ("a:-1\n", True), ("a:nan\n", True),
                    ("\n1 2\n", False), ("a b\n1\n", False), ("a b\n1 -2\n", False), ("a b\n\n", False)]:
        with pytest.raises(hostlib.HostError):
            hostlib.Table(bad, sp)


def test_resolve_semantics(built):
    """Names -> leaf ids: every leaf with the name (A6), internal-only names dropped (A2),
    unknown names are the reference's validation error (unifrac.go:85-88)."""
    from frackyfrac_b200 import hostlib

    tree = hostlib.Tree("((x:1,y:1)in:1,(x:1,z:1):1);")
    rp, col, val = hostlib.Table("x:3 in:9\nz:2\n\n", True).resolve(tree)
    assert rp.tolist() == [0, 2, 3, 3] and col.tolist() == [2, 5, 6] and val.tolist() == [3, 3, 2]
    with pytest.raises(hostlib.HostError, match=r'sample #2 has value 0.5 for species "nope" which is not in the tree'):
        hostlib.Table("x:1\nnope:0.5\n", True).resolve(tree)


def test_go_format_matches_oracle(built):
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    rng = np.random.default_rng(11)
    vals = np.concatenate([rng.random(3000), rng.random(1000) * 10.0 ** rng.integers(-12, 3, 1000),
                           [0.0, 1.0, np.nan, 1e-5, 1e-4, 2 / 3, 22 / 36, 0.1 + 0.2]])
    for v in vals:
        assert hostlib.format_go(float(v)) == orc.format_go(float(v))
    # a third, independent implementation: for 0 < v < 1 CPython's repr() is the same text as Go's %v
    # (shortest round-trip digits, exponent form below 1e-4 with at least two exponent digits)
    more = np.concatenate([rng.random(20000), rng.random(5000).astype(np.float32).astype(np.float64),
                           10.0 ** -rng.uniform(0, 300, 3000), np.nextafter(1e-4, [0.0, 1.0]), [5e-324, 2.2250738585072014e-308]])
    for v in more[(more > 0) & (more < 1)]:
        assert hostlib.format_go(float(v)) == repr(float(v)), repr(float(v))
    assert [hostlib.format_go(v) for v in (0.0, 1.0, float("nan"), 100000.0, 1e6, 123456789.0, -0.5)] == \
        ["0", "1", "NaN", "100000", "1e+06", "1.23456789e+08", "-0.5"]


def test_cli_flag_rules(built):
    """frcfrc.go:70-88 and common.go:13-18: usage exit 0 without args; ERROR: ... exit 2."""
    from frackyfrac_b200 import hostlib

    r = subprocess.run([hostlib.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "" and r.stderr.startswith("FrackyFrac calculates UniFrac")
    assert "  -t string\n    \tPath to tree file, required\n" in r.stderr
    for args, msg in [(["-w"], "ERROR: please provide a tree file with -t\n"),
                      (["-t", "x", "-p", "0"], "ERROR: bad number of threads: 0\n"),
                      (["-t", "x", "-l"], "ERROR: -l can only be used with weighted unifrac\n")]:
        r = subprocess.run([hostlib.CLI_PATH] + args, capture_output=True, text=True)
        assert r.returncode == 2 and r.stderr == msg and r.stdout == ""
    r = subprocess.run([hostlib.CLI_PATH, "-q"], capture_output=True, text=True)
    assert r.returncode == 2 and r.stderr.startswith("flag provided but not defined: -q\n")
    # validation errors surface as the reference's message with exit status 2, before any device work
    tree = os.path.join(GOLDEN, "wtd.tree")
    r = subprocess.run([hostlib.CLI_PATH, "-t", tree, "-s"], input="nope:1\n", capture_output=True, text=True)
    assert r.returncode == 2 and 'species "nope" which is not in the tree' in r.stderr


def test_parallel_writer_is_byte_identical(built):
    """N1: the ordered parallel writer must emit exactly the single-threaded bytes."""
    import time

    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    rng = np.random.default_rng(17)
    v = rng.random(300_000)
    v[::1000] = 0.0
    v[1::1000] = 1.0
    v[2::1000] = np.nan
    v[3::1000] = rng.random(300) * 1e-7
    one = hostlib.format_lines(v, 1)
    assert one == hostlib.format_lines(v, 7) == hostlib.format_lines(v, 64)
    lines = one.decode().split("\n")
    assert lines[-1] == "" and len(lines) == len(v) + 1
    for k in list(range(0, 4000, 7)) + [len(v) - 1]:
        assert lines[k] == orc.format_go(float(v[k]))
    t0 = time.perf_counter(); hostlib.format_lines(v, 1); t1 = time.perf_counter(); hostlib.format_lines(v, 8); t2 = time.perf_counter()
    assert (t2 - t1) < (t1 - t0) * 3 + 0.05  # never pathologically slower (loose: this box may have one core)


def test_compressed_files_round_trip_by_suffix(built, tmp_path):
    """aio.Open / aio.Create (frcfrc.go:93,100-106): the suffix picks the codec; what counts is the decoded bytes.
    Output is a run of independent gzip members / zstd frames written in order by several workers."""
    import gzip
    import shutil

    from frackyfrac_b200 import hostlib

    rng = np.random.default_rng(5)
    text = hostlib.format_lines(rng.random(1_200_000), 4)  # ~23 MB: several 4 MB blocks
    chunks = [text[:7], b"", text[7:9_000_001], text[9_000_001:]]
    plain = tmp_path / "d.txt"
    hostlib.write_file(str(plain), chunks, 3)
    assert plain.read_bytes() == text and hostlib.read_file(str(plain)) == text
    for threads in (1, 5):
        gz = tmp_path / f"d{threads}.txt.gz"
        hostlib.write_file(str(gz), chunks, threads)
        raw = gz.read_bytes()
        assert raw[:2] == b"\x1f\x8b" and len(raw) < len(text) * 0.6
        assert gzip.decompress(raw) == text          # an independent decoder reads every member
        assert hostlib.read_file(str(gz)) == text
    # input side: single-member and multi-member files from another encoder, and an empty payload
    other = tmp_path / "in.tsv.gz"
    other.write_bytes(gzip.compress(b"a\tb\n1\t2\n") + gzip.compress(b"3\t4\n", 9))
    assert hostlib.read_file(str(other)) == b"a\tb\n1\t2\n3\t4\n"
    empty = tmp_path / "e.gz"
    hostlib.write_file(str(empty), [], 2)
    assert gzip.decompress(empty.read_bytes()) == b"" and hostlib.read_file(str(empty)) == b""
    # damaged input is an error, never silently short data
    cut = tmp_path / "cut.gz"
    cut.write_bytes(gz.read_bytes()[:-9])
    with pytest.raises(hostlib.HostError, match="gzip"):
        hostlib.read_file(str(cut))
    junk = tmp_path / "junk.gz"
    junk.write_bytes(b"plain text, wrong suffix\n")
    with pytest.raises(hostlib.HostError, match="gzip"):
        hostlib.read_file(str(junk))
    with pytest.raises(hostlib.HostError, match="open .*missing.gz"):
        hostlib.read_file(str(tmp_path / "missing.gz"))
    # zstandard frames (libzstd is loaded at run time; the zstd CLI, when present, is the independent decoder)
    zst = tmp_path / "d.txt.zst"
    hostlib.write_file(str(zst), chunks, 4)
    assert zst.read_bytes()[:4] == b"\x28\xb5\x2f\xfd" and hostlib.read_file(str(zst)) == text
    if shutil.which("zstd"):
        r = subprocess.run(["zstd", "-dc", str(zst)], capture_output=True)
        assert r.returncode == 0 and r.stdout == text
    bad = tmp_path / "bad.zst"
    bad.write_bytes(zst.read_bytes()[:-5])
    with pytest.raises(hostlib.HostError, match="zstd"):
        hostlib.read_file(str(bad))


def test_parallel_table_parser_matches_single_thread_and_oracle(built):
    """N2: rows parsed by several workers (parser.go's ngoroutines) must give the same CSR, the same
    species order and the same first error as one worker, and the same maps as the oracle's parser."""
    from frackyfrac_b200 import hostlib, synth
    from oracle import oracle as orc

    tree = synth.random_tree(400, 5)
    rp, col, val = synth.random_table(tree, 900, 0.05, 6, integer_counts=False)
    ht = hostlib.Tree(synth.to_newick(tree))
    for sparse, text in ((False, synth.to_dense_text(tree, rp, col, val)), (True, synth.to_sparse_text(tree, rp, col, val))):
        assert len(text) > 5 * 65536  # enough for several chunks
        one = hostlib.Table(text, sparse, 1)
        many = hostlib.Table(text, sparse, 7)
        assert one.maps() == many.maps() == orc.Table.parse(text, sparse).maps()
        a, b = one.resolve(ht), many.resolve(ht)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(a[0], rp)
        # the first error in FILE order wins, whichever worker meets it
        lines = text.split("\n")
        bad_late, bad_early = len(lines) - 3, len(lines) // 2
        for where, tok in ((bad_late, "1e999" if not sparse else "L1:1e999"), (bad_early, "-3" if not sparse else "L2:-3")):
            row = lines[where].split("\t")
            row[3 if not sparse else 0] = tok
            lines[where] = "\t".join(row)
        broken = "\n".join(lines)
        msgs = []
        for th in (1, 7):
            with pytest.raises(hostlib.HostError) as ei:
                hostlib.Table(broken, sparse, th)
            msgs.append(str(ei.value))
        assert msgs[0] == msgs[1] and f"row #{bad_early + 1}:" in msgs[0], msgs
    # dense: a wrong token count is reported before a bad value of the same row (parser.go:61-64)
    with pytest.raises(hostlib.HostError) as ei:
        hostlib.Table("a b c\n1 x\n", False, 1)
    assert "has 2 values, expected 3" in str(ei.value)
    with pytest.raises(hostlib.HostError) as ei:
        hostlib.Table("a b c\n1 x 2\n", False, 1)
    assert "value #2" in str(ei.value) and "invalid syntax" in str(ei.value)


def test_float_token_forms_match_the_oracle_parser(built):
    """strconv.ParseFloat forms the fast path (std::from_chars) does not take itself: sign, hex floats,
    inf / nan spellings, underscores, out-of-range exponents.  Accept / reject must equal the oracle's."""
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    tokens = ["7", "0", "+5", "-0", ".5", "5.", "1e3", "1E-3", "0x1p-2", "1_000", "1e999", "1e-999", "inf", "Inf",
              "+Infinity", "nan", "NaN", "-1", "abc", "1.5e", "--1", "0x", "1e+400", "4.9e-324", "1.7976931348623157e308",
              "0x1A", "0X1.8P+1", "0x_1p0", "0xp1", "nan(1)", "NAN(abc)", "+nan", "1__0", "_1", "1_", "1_.5", "1._5", "1e1_0", "1e_5",
              "1e5e3", ".e5", ".", "+", "1p3", "0x1e3p1", "infinit", "1_0.2_5e0_1"]
    # strconv.ParseFloat's own rules (strconv/atof.go), independent of what the C library's strtod would take:
    # underscores may separate digits, a hex mantissa needs its p exponent, nan takes no payload
    go_accepts = {"1_000": 1000.0, "0x1p-2": 0.25, "0X1.8P+1": 3.0, "0x_1p0": 1.0, "1e1_0": 1e10, "0x1e3p1": 966.0,
                  "1_0.2_5e0_1": 102.5, "+5": 5.0, "5.": 5.0, ".5": 0.5}
    go_rejects = ["0x1A", "0xp1", "nan(1)", "NAN(abc)", "1__0", "_1", "1_", "1_.5", "1._5", "1e_5", "1e5e3", ".e5", ".", "+",
                  "1p3", "infinit", "0x", "1.5e", "1e999"]
    for tok, want in go_accepts.items():
        assert hostlib.Table(f"a:{tok}\n", True, 1).maps() == [{"a": want}], tok
        assert orc.Table.parse(f"a:{tok}\n", True).maps() == [{"a": want}], tok
    for tok in go_rejects:
        with pytest.raises(hostlib.HostError):
            hostlib.Table(f"a:{tok}\n", True, 1)
        with pytest.raises(orc.OracleError):
            orc.Table.parse(f"a:{tok}\n", True)
    for tok in tokens:
        outcomes = []
        for parse in (lambda t: hostlib.Table(t, True, 1).maps(), lambda t: orc.Table.parse(t, True).maps()):
            try:
                outcomes.append(parse(f"a:{tok}\n"))
            except Exception:
                outcomes.append("error")
        assert outcomes[0] == outcomes[1], (tok, outcomes)
        outcomes = []
        for parse in (lambda t: hostlib.Table(t, False, 1).maps(), lambda t: orc.Table.parse(t, False).maps()):
            try:
                outcomes.append(parse(f"a b\n1 {tok}\n"))
            except Exception:
                outcomes.append("error")
        assert outcomes[0] == outcomes[1], (tok, outcomes)


def test_sprspr_converter_follows_the_reference_test(built, tmp_path):
    """sprspr/sprspr_test.go:12-37: three dense inputs and their sparse renderings (fields compared after sorting,
    because the reference ranges over a Go map), through the library and through the stand-in binary; plus the
    round trip dense -> sparse -> maps on a bigger table and the reference's error exit."""
    from frackyfrac_b200 import hostlib, synth

    cases = [("s1\n1.3", "s1:1.3"),
             ("s1\ts2\n0\t4\n3\t0\n", "s2:4\ns1:3"),
             ("s1\ts2\ts3\n4\t3\t2\n5\t0\t8\n0\t0\t10", "s1:4\ts2:3\ts3:2\ns1:5\ts3:8\ns3:10")]

    def canon(text):
        rows = text.removesuffix("\n").split("\n")
        return "\n".join("\t".join(sorted(r.split("\t"))) for r in rows)

    for dense, want in cases:
        assert canon(hostlib.to_sparse(dense).decode()) == want
        r = subprocess.run([hostlib.SPRSPR_PATH], input=dense, capture_output=True, text=True)
        assert r.returncode == 0 and canon(r.stdout) == want
        assert r.stderr.startswith("SparseySparse converts dense format abundance tables to sparse format.\n")
    tree = synth.random_tree(60, 3)
    rp, col, val = synth.random_table(tree, 25, 0.2, 4, integer_counts=False)
    dense = synth.to_dense_text(tree, rp, col, val)
    sparse = hostlib.to_sparse(dense, threads=3)
    assert hostlib.Table(sparse, True).maps() == hostlib.Table(dense, False).maps()
    assert sparse == hostlib.to_sparse(dense, threads=1)
    r = subprocess.run([hostlib.SPRSPR_PATH], input="a b\n1\n", capture_output=True, text=True)
    assert r.returncode == 2 and r.stderr.rstrip().split("\n")[-1].startswith("ERROR: ")


def test_fp32_bands_format_like_their_widened_doubles(built):
    """The CLI formats fast-path distances straight from the fp32 bands (frc_next_f32): the text must be what Go's
    fmt.Fprintln prints for float64(f) -- i.e. the double formatter applied to the exactly widened value -- for any
    thread count, and a band's exceptions (distances below fp32's range) replace their fp32 stand-ins."""
    from frackyfrac_b200 import hostlib

    rng = np.random.default_rng(5)
    f = np.concatenate([rng.random(20_000, dtype=np.float32), np.array([0.0, 1.0, np.nan, 1e-5, 3e-39, 0.6111111], np.float32)])
    want = hostlib.format_lines(f.astype(np.float64), 1)
    for threads in (1, 3, 8):
        assert hostlib.format_lines_f32(f, threads=threads) == want
    assert b"0.6111111044883728\n" in want and want.splitlines()[-4] == b"NaN"
    # exceptions: flat indices relative to the band's first index; entries outside the band are ignored
    first = 1_000_000
    xi, xv = np.array([first + 3, first + len(f) - 5, first - 1, first + len(f)], np.int64), np.array([1e-60, 2.5e-45, 9.0, 9.0])
    d = f.astype(np.float64)
    d[3], d[len(f) - 5] = 1e-60, 2.5e-45
    for threads in (1, 4):
        assert hostlib.format_lines_f32(f, first, xi, xv, threads) == hostlib.format_lines(d, 1)


def test_float_token_grammar_differential(built):
    """Random tokens over the alphabet of Go's float syntax (digits, '.', '_', exponents, hex prefix, signs, inf / nan
    letters): the C++ host reader and the C oracle implement strconv.ParseFloat's grammar independently
    (hostlib.cpp go_float_syntax, unifrac_oracle.c go_float_token) and must agree on accept / reject and on the value."""
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    rng = np.random.default_rng(11)
    alphabet = list("0123456789") * 3 + list("..__eEpPxX+-") + list("abfnNiIt")
    accepted = 0
    for _ in range(6000):
        tok = "".join(rng.choice(alphabet, size=int(rng.integers(1, 9))))
        res = []
        for parse, err in ((lambda t: hostlib.Table(t, True, 1).maps(), hostlib.HostError),
                           (lambda t: orc.Table.parse(t, True).maps(), orc.OracleError)):
            try:
                res.append(parse(f"a:{tok}\n"))
            except err:
                res.append("error")
        assert res[0] == res[1], (tok, res)
        accepted += res[0] != "error"
    assert accepted > 300   # (the generator does produce plenty of valid spellings)
