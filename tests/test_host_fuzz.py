"""Differential fuzzing of the C++ host readers against the oracle's independent readers (CPU).

Both restate the reference's input semantics (third-party Newick reader used at frcfrc/frcfrc.go:109-114,
parser/parser.go:21-140) and were written separately (C++ state machine vs C recursive descent): on any text
they must either both fail or agree on every array.  Hypothesis drives the grammar, including the corners the
reference's own tests do not pin (odd numeric forms, quoted labels, comments, blank lines, CRLF, repeats)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

NUMS = st.sampled_from(["0", "1", "7", "0.5", ".25", "3.", "1e-3", "2E2", "-1", "+2", "1e400", "1e-400", "0x10", "inf",
                        "nan", "1.5e", "abc", "1,5", "", "00.10", "12345678901234567890", "4.9e-324"])
LABELS = st.sampled_from(["", "a", "b", "L10", "sp_1", "x.y", "a b", "'q l'", "'it''s'", "a[c]b", "7", "-"])
WS = st.sampled_from(["", "", "", " ", "\t", "\n", " \r\n "])


@st.composite
def newick(draw, depth=0):
    ws = draw(WS)
    if depth < 4 and draw(st.integers(0, 9)) < (6 if depth == 0 else 4):
        kids = draw(st.lists(newick(depth + 1), min_size=1, max_size=4))
        body = "(" + ",".join(kids) + ")"
    else:
        body = ""
    label = draw(LABELS)
    length = draw(st.one_of(st.just(None), NUMS))
    comment = draw(st.sampled_from(["", "", "[note]", "[&x=1]"]))
    out = ws + body + label + comment
    if length is not None:
        out += draw(WS) + ":" + draw(WS) + length
    return out + draw(WS)


@settings(max_examples=int(__import__("os").environ.get("FRC_FUZZ_EXAMPLES", "300")), deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(newick(), st.sampled_from([";", ";", ";\n", "", ";;", "; (a,b);"]))
def test_newick_readers_agree(built, body, tail):
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    text = body + tail
    try:
        o = orc.Tree.parse(text)
    except orc.OracleError:
        with pytest.raises(hostlib.HostError):
            hostlib.Tree(text)
        return
    h = hostlib.Tree(text)
    p, l = o.flatten()
    assert np.array_equal(h.parent, p), text
    assert np.array_equal(h.length, l, equal_nan=True), text


TOK_NAME = st.sampled_from(["a", "b", "c", "dd", "a:b", "x y"[0], "é", "7", ":", ""])
TOK_VAL = st.sampled_from(["1", "2", "10", "0.5", "1e3", "0", "-1", "nan", "inf", "+4", "1e999", "0x2", "", "x", "1e-999", "007"])
SEP = st.sampled_from([" ", "\t", "  ", " \t "])
EOL = st.sampled_from(["\n", "\n", "\r\n", "\n\n"])


@st.composite
def sparse_table(draw):
    rows = []
    for _ in range(draw(st.integers(0, 5))):
        toks = [draw(TOK_NAME) + draw(st.sampled_from([":", ":", ":", "", "::"])) + draw(TOK_VAL)
                for _ in range(draw(st.integers(0, 5)))]
        rows.append(draw(st.sampled_from(["", " "])) + draw(SEP).join(toks) + draw(st.sampled_from(["", " "])))
    return "".join(r + draw(EOL) for r in rows) + draw(st.sampled_from(["", "a:1", "b:2 "]))


@st.composite
def dense_table(draw):
    ncol = draw(st.integers(0, 4))
    hdr = draw(SEP).join(draw(st.sampled_from(["a", "b", "c", "a", "x:y", "9"])) for _ in range(ncol))
    rows = []
    for _ in range(draw(st.integers(0, 4))):
        k = ncol + draw(st.sampled_from([0, 0, 0, 0, -1, 1]))
        rows.append(draw(SEP).join(draw(TOK_VAL) for _ in range(max(k, 0))))
    return hdr + draw(EOL) + "".join(r + draw(EOL) for r in rows)


def _both(text, sparse):
    from frackyfrac_b200 import hostlib
    from oracle import oracle as orc

    try:
        want = orc.Table.parse(text, sparse).maps()
    except orc.OracleError:
        for threads in (1, 3):
            with pytest.raises(hostlib.HostError):
                hostlib.Table(text, sparse, threads)
        return
    for threads in (1, 3):
        got = hostlib.Table(text, sparse, threads).maps()
        assert len(got) == len(want), repr(text)
        for g, w in zip(got, want):
            assert g.keys() == w.keys(), repr(text)
            for k in w:
                assert g[k] == w[k] or (g[k] != g[k] and w[k] != w[k]), repr(text)


@settings(max_examples=int(__import__("os").environ.get("FRC_FUZZ_EXAMPLES", "300")), deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(sparse_table())
def test_sparse_table_readers_agree(built, text):
    _both(text, True)


@settings(max_examples=int(__import__("os").environ.get("FRC_FUZZ_EXAMPLES", "300")), deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(dense_table())
def test_dense_table_readers_agree(built, text):
    _both(text, False)
