"""The oracle pinned against the reference's own known answers (CPU, no GPU needed)."""
import os
import subprocess

import numpy as np
import pytest

from tests import kat

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _read(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return f.read()


@pytest.mark.parametrize("case", kat.UNIT, ids=[c["name"] for c in kat.UNIT])
def test_unit_kat_bit_exact(built, case):
    """frcfrc/unifrac_test.go:12-74 compares with reflect.DeepEqual: bit-exact float64."""
    from oracle import oracle as orc

    tree = orc.Tree.parse(case["tree"])
    tab = orc.Table.parse(kat.sparse_text(case["abnd"]), True)
    orc.validate_species(tab, tree)
    for threads in (1, 3):
        assert orc.unifrac(tab, tree, case["weighted"], 1, threads).tolist() == case["want"]


@pytest.mark.parametrize("fixture,sparse,weighted", kat.CLI)
def test_cli_golden(built, fixture, sparse, weighted):
    """testdata/run.sh:3-16: six byte-exact diffs against *.want."""
    from oracle import oracle as orc

    tree = orc.Tree.parse(_read(fixture + ".tree"))
    tab = orc.Table.parse(_read(fixture + (".sparse" if sparse else ".dense")), sparse)
    orc.validate_species(tab, tree)
    d = orc.unifrac(tab, tree, weighted, 1, 2)
    assert "".join(orc.format_go(v) + "\n" for v in d) == _read(fixture + ".want")
    cmd = [orc.CLI_PATH, "-t", os.path.join(GOLDEN, fixture + ".tree"),
           "-i", os.path.join(GOLDEN, fixture + (".sparse" if sparse else ".dense"))]
    cmd += (["-s"] if sparse else []) + (["-w"] if weighted else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == _read(fixture + ".want")


@pytest.mark.parametrize("fixture,suffix,sparse,tag,weighted,normalize", kat.SYNTH)
def test_synthetic_golden_bytes(built, fixture, suffix, sparse, tag, weighted, normalize):
    """The committed synthetic fixtures (tests/golden/make_golden.py): the oracle, through its library and its
    CLI, must keep producing exactly these bytes — a change to the checker cannot move the target unnoticed."""
    from oracle import oracle as orc

    tree = orc.Tree.parse(_read(fixture + ".tree"))
    tab = orc.Table.parse(_read(fixture + suffix), sparse)
    orc.validate_species(tab, tree)
    want = _read(f"{fixture}.{tag}.want")
    for threads in (1, 3):
        d = orc.unifrac(tab, tree, weighted, normalize, threads)
        assert "".join(orc.format_go(float(v)) + "\n" for v in d) == want
    if normalize == 2:
        return  # the oracle's CLI restates the reference's -l literally (unsorted lists, normalize = 0: the `wl`
                # fixtures); -l as documented (normalize = 2, `wl2`) is only reachable through the library
    cmd = [orc.CLI_PATH, "-t", os.path.join(GOLDEN, fixture + ".tree"), "-i", os.path.join(GOLDEN, fixture + suffix)]
    cmd += (["-s"] if sparse else []) + (["-w"] if weighted else []) + (["-l"] if normalize == 0 else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == want


def test_pair_order_is_iterpairs(built):
    """common/common_test.go:8-18: (2,1),(4,1),(4,2),(8,1),(8,2),(8,4) -> row-major lower triangle."""
    from oracle import oracle as orc

    # four samples with one private leaf each: d(i,j) identifies the pair through the lengths
    tree = orc.Tree.parse("(a:1,b:2,c:4,d:8);")
    tab = orc.Table.parse("a:1\nb:1\nc:1\nd:1\n", True)
    d = orc.unifrac(tab, tree, False)
    assert d.tolist() == [1.0] * 6
    tab = orc.Table.parse("a:1 b:1\nb:1\nc:1 b:1\nd:1 b:1\n", True)
    d = orc.unifrac(tab, tree, False)
    # pair (i,j): unique = private lengths, common = 2 (leaf b)
    want = [1 / 3, (1 + 4) / 7, 4 / 6, (1 + 8) / 11, 8 / 10, (4 + 8) / 14]
    assert np.allclose(d, want, rtol=0, atol=0)


def test_flat_index_math():
    """trtr/dist_test.go:9: flat index of (i,j), j<i is i(i-1)/2 + j."""
    k = 0
    for i in range(7):
        for j in range(i):
            assert i * (i - 1) // 2 + j == k
            k += 1


def test_go_float_format(built):
    """fmt.Fprintln(w, f) (frcfrc.go:59): shortest round-trip, %e below 1e-4."""
    from oracle import oracle as orc

    cases = {1.0: "1", 0.0: "0", 2 / 3: "0.6666666666666666", 22 / 36: "0.6111111111111112",
             19 / 28: "0.6785714285714286", 16 / 22: "0.7272727272727273", 1e-4: "0.0001", 1e-5: "1e-05",
             1.5e-7: "1.5e-07", 0.5: "0.5", float("nan"): "NaN", float("inf"): "+Inf", 123456.0: "123456",
             1e6: "1e+06", 0.1: "0.1", 5e-324: "5e-324"}
    for v, s in cases.items():
        assert orc.format_go(v) == s, (v, orc.format_go(v), s)
    rng = np.random.default_rng(5)
    for v in np.concatenate([rng.random(2000), rng.random(2000) * 1e-6]):
        s = orc.format_go(float(v))
        assert float(s) == v and len(s) <= len(repr(float(v))) + 1


def test_parsers_follow_reference_tests(built):
    """parser/parser_test.go:9-47."""
    from oracle import oracle as orc

    t = orc.Table.parse("   aa  bbbb    \n1\t2\n 3  \t  4 \t\n", False)
    assert t.maps() == [{"aa": 1, "bbbb": 2}, {"aa": 3, "bbbb": 4}]
    t = orc.Table.parse("a:11 b:222  \n  b:32 c:7\n\nd:1\tc:4\ta:10\n", True)
    assert t.maps() == [{"a": 11, "b": 222}, {"b": 32, "c": 7}, {}, {"d": 1, "c": 4, "a": 10}]
    # splitSparse (parser_test.go:48-80): last colon splits; empty name / no colon are errors
    assert orc.Table.parse("c:d:e::7\n", True).maps() == [{"c:d:e:": 7}]
    for bad in ["a\n", ":\n", "a:\n", ":5\n", "a:0\n", "a:-1\n", "a:nan\n", "a:inf\n", "a:x\n"]:
        with pytest.raises(orc.OracleError):
            orc.Table.parse(bad, True)
    for bad in ["\n1 2\n", "a b\n1\n", "a b\n1 2 3\n", "a b\n1 -2\n", "a b\n1 x\n", "a b\n\n"]:
        with pytest.raises(orc.OracleError):
            orc.Table.parse(bad, False)
    # zeros are dropped in dense rows, duplicates: last non-zero wins (parser.go:75-78)
    assert orc.Table.parse("a a b\n1 0 0\n2 3 0\n", False).maps() == [{"a": 1}, {"a": 3}]


def test_validate_species_message(built):
    """unifrac.go:80-93."""
    from oracle import oracle as orc

    tree = orc.Tree.parse("((a:1,b:1)ab:1,c:1);")
    orc.validate_species(orc.Table.parse("a:1 ab:2\n", True), tree)  # internal names validate
    with pytest.raises(orc.OracleError, match=r'sample #2 has value 2.5 for species "zz" which is not in the tree'):
        orc.validate_species(orc.Table.parse("a:1\nzz:2.5\n", True), tree)


def test_quirks(built):
    """Appendix A of SURVEY.md: A2 leaf-only abundance, A3 all-node normalisation, A5 root length, A6 shared names."""
    from oracle import oracle as orc

    tree = orc.Tree.parse("((a:1,b:1)ab:1,c:1);")
    # A2: a species named like an internal node is silently ignored
    d = orc.unifrac(orc.Table.parse("a:1 ab:50\na:1\n", True), tree, True)
    assert d.tolist() == [0.0]
    # A5: a root branch length is common to every non-empty pair
    t0 = orc.Tree.parse("(a:1,b:1);")
    t1 = orc.Tree.parse("(a:1,b:1):2;")
    tab = orc.Table.parse("a:1\nb:1\n", True)
    assert orc.unifrac(tab, t0, False).tolist() == [1.0]
    assert orc.unifrac(tab, t1, False).tolist() == [0.5]
    # A6: every leaf carrying the name gets the abundance
    t2 = orc.Tree.parse("((x:1,y:1):1,(x:1,z:1):1);")
    ids, vals = orc.flat_nodes(orc.Table.parse("x:3\n", True), t2, 2, 0)
    assert ids.tolist() == [0, 1, 2, 4, 5] and vals.tolist() == [6, 3, 3, 3, 3]
    # A3: the normaliser is the sum over ALL nodes (leaf value counted once per ancestor)
    ids, vals = orc.flat_nodes(orc.Table.parse("x:3\n", True), t2, 1, 0)
    assert vals.tolist() == [6 / 18, 3 / 18, 3 / 18, 3 / 18, 3 / 18]


def test_l_flag_reference_quirk(built):
    """With -l the reference never sorts its node lists (the sort lives inside
    normalizeFlatNodes, unifrac.go:57), so its merge-join runs over post-order
    lists.  The oracle restates that faithfully as normalize=0; normalize=2 is -l
    as documented (sorted lists, raw values).  The CUDA engine implements both (frc_opts_t.normalize 0 / 2;
    the CLI's -l is 0, like the reference).  This test pins the difference between the two."""
    from oracle import oracle as orc

    tree = orc.Tree.parse("((s1:1,s2:3):2,(s3:2,s4:5):1);")
    tab = orc.Table.parse("s1:4 s3:1\ns2:2 s3:3\n", True)
    faithful = orc.unifrac(tab, tree, True, 0)
    intent = orc.unifrac(tab, tree, True, 2)
    # intent: |4-0|*1 + |0-2|*3 + |4-2|*2 + |1-3|*2 + |1-3|*1 + root 0 over sums
    assert intent.tolist() == [(4 + 6 + 4 + 4 + 2) / (4 + 6 + 12 + 8 + 4)]
    assert faithful[0] != intent[0]
    # on a tree where post-order happens to be id-sorted per sample pair (star tree) they agree
    star = orc.Tree.parse("(a:1,b:2,c:3);")
    t2 = orc.Table.parse("a:1 b:2\nb:1 c:5\n", True)
    assert np.isfinite(orc.unifrac(t2, star, True, 0)).all()


def test_empty_and_degenerate(built):
    """A8/A9: N<2 -> no output; two empty samples -> NaN; disjoint -> 1; identical -> 0."""
    from oracle import oracle as orc

    tree = orc.Tree.parse("(a:1,b:2,c:3);")
    assert len(orc.unifrac(orc.Table.parse("a:1\n", True), tree, False)) == 0
    d = orc.unifrac(orc.Table.parse("\n\na:1\na:2\nb:1\n", True), tree, True)
    assert np.isnan(d[0]) and d[1] == 1.0 and d[5] == 0.0 and d[9] == 1.0
    assert orc.format_go(d[0]) == "NaN"


def _random_case(rng):
    """A small tree as text (shared leaf names, named internal nodes, zero and root lengths) and sample maps."""
    n_leaves = int(rng.integers(2, 14))
    pool = [f"s{k}" for k in range(max(2, n_leaves - 2))]          # fewer names than leaves: some repeat (A6)
    nodes = [f"{pool[int(rng.integers(len(pool)))]}:{int(rng.integers(0, 9)) / 2}" for _ in range(n_leaves)]
    k = 0
    while len(nodes) > 1:
        take = int(rng.integers(2, min(4, len(nodes)) + 1))
        idx = sorted(rng.choice(len(nodes), size=take, replace=False).tolist(), reverse=True)
        kids = [nodes.pop(i) for i in idx][::-1]
        name = f"i{k}" if rng.random() < 0.5 else ""
        nodes.append(f"({','.join(kids)}){name}:{int(rng.integers(0, 9)) / 4}")
        k += 1
    text = nodes[0] if rng.random() < 0.5 else nodes[0].rsplit(":", 1)[0]   # with / without a root length (A5)
    samples = []
    for _ in range(int(rng.integers(2, 7))):
        m = {}
        for nm in pool + [f"i{q}" for q in range(k)]:
            if rng.random() < 0.35:
                m[nm] = float(rng.integers(1, 40)) if rng.random() < 0.7 else float(rng.random() * 3 + 0.01)
        samples.append(m)
    if rng.random() < 0.3:
        samples[0] = {}                                               # an empty sample (A8: NaN rows)
    return text + ";", samples


def test_c_oracle_against_the_python_restatement(built):
    """oracle/py_restatement.py keeps the reference's own structures (maps, recursion, merge-join); the C oracle is
    dense and threaded.  Same inputs, same summation order -> identical bits, in every mode."""
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc
    from oracle import py_restatement as pyr

    rng = np.random.default_rng(2024)
    checked = 0
    for _ in range(150):
        text, samples = _random_case(rng)
        ht = hostlib.Tree(text)
        names = ht.names()
        known = set(names)
        samples = [{k: v for k, v in m.items() if k in known} for m in samples]
        tab_text = "".join("\t".join(f"{k}:{v!r}" for k, v in m.items()) + "\n" for m in samples)
        tree = orc.Tree.parse(text)
        tab = orc.Table.parse(tab_text, True)
        orc.validate_species(tab, tree)
        for weighted in (False, True):
            for normalize in ((1,) if not weighted else (0, 1, 2)):
                want = pyr.unifrac(samples, ht.parent.tolist(), ht.length.tolist(), names, weighted, normalize)
                got = orc.unifrac(tab, tree, weighted, normalize, 2)
                assert len(got) == len(want)
                for g, w in zip(got.tolist(), want):
                    assert g == w or (g != g and w != w), (text, tab_text, weighted, normalize, g, w)
                checked += len(want)
    assert checked > 3000
    # and the reference's three unit KATs through the restatement itself
    for case in kat.UNIT:
        ht = hostlib.Tree(case["tree"])
        got = pyr.unifrac([{k: float(v) for k, v in m.items()} for m in case["abnd"]], ht.parent.tolist(),
                          ht.length.tolist(), ht.names(), case["weighted"], 1)
        assert got == case["want"]
    # ... and the reference's six CLI fixtures, byte for byte
    for fixture, sparse, weighted in kat.CLI:
        ht = hostlib.Tree(_read(fixture + ".tree"))
        maps = hostlib.Table(_read(fixture + (".sparse" if sparse else ".dense")), sparse).maps()
        d = pyr.unifrac(maps, ht.parent.tolist(), ht.length.tolist(), ht.names(), weighted, 1)
        assert "".join(orc.format_go(v) + "\n" for v in d) == _read(fixture + ".want")
