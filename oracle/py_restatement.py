"""A second, independent restatement of the reference's UniFrac in plain Python (TEST INFRASTRUCTURE).

It keeps the reference's own data structures — per-sample name -> value maps, a pointer tree walked recursively,
lists of (node id, subtree sum), a merge-join per pair — so that the C oracle (oracle/unifrac_oracle.c, dense and
multithreaded) is cross-checked by something that does not share its design.  Python floats are IEEE doubles and
every sum below runs in the reference's order, so the two must agree bit for bit on small inputs
(tests/test_oracle_kat.py).  Pure-Python loops: small cases only.  Nothing outside tests/ may import this module.

Follows fluhus/frackyfrac:
  frcfrc/unifrac.go:32-53    abundanceToFlatNodes  -> _flat_nodes
  frcfrc/unifrac.go:56-67    normalizeFlatNodes    -> _normalize
  frcfrc/unifrac.go:127-133  enumerateNodes        -> ids are pre-order indices (the flat arrays already are)
  frcfrc/unifrac.go:144-171  unifracDistUnweighted -> _unweighted
  frcfrc/unifrac.go:174-205  unifracDistWeighted   -> _weighted
  frcfrc/unifrac.go:97-124   unifrac               -> unifrac
  common/common.go:21-31     IterPairs             -> the (i, j < i) double loop
"""
from __future__ import annotations

import math
import sys


def _flat_nodes(abnd: dict[str, float], v: int, children: list[list[int]], names: list[str], out: list) -> float:
    """Post-order subtree sums; only a childless node takes its species' value (unifrac.go:38-43)."""
    total = 0.0
    for c in children[v]:
        total += _flat_nodes(abnd, c, children, names, out)
    if not children[v]:
        a = abnd.get(names[v], 0.0)
        if a > 0:
            total += a
    if total > 0:
        out.append([v, total])
    return total


def _normalize(nodes: list) -> None:
    nodes.sort(key=lambda n: n[0])
    total = 0.0
    for n in nodes:
        total += n[1]
    for n in nodes:
        n[1] /= total


def _unweighted(a: list, b: list, dist: list[float]) -> float:
    result = common = 0.0
    i = j = 0
    while i < len(a) and j < len(b):
        if a[i][0] < b[j][0]:
            result += dist[a[i][0]]
            i += 1
        elif a[i][0] > b[j][0]:
            result += dist[b[j][0]]
            j += 1
        else:
            common += dist[a[i][0]]
            i += 1
            j += 1
    for x in a[i:]:
        result += dist[x[0]]
    for x in b[j:]:
        result += dist[x[0]]
    den = result + common
    return result / den if den != 0 else math.nan  # Go: 0/0 = NaN, x/0 cannot occur (den >= result >= 0)


def _weighted(a: list, b: list, dist: list[float]) -> float:
    numer = denom = 0.0
    i = j = 0
    while i < len(a) and j < len(b):
        if a[i][0] < b[j][0]:
            numer += dist[a[i][0]] * a[i][1]
            denom += dist[a[i][0]] * a[i][1]
            i += 1
        elif a[i][0] > b[j][0]:
            numer += dist[b[j][0]] * b[j][1]
            denom += dist[b[j][0]] * b[j][1]
            j += 1
        else:
            numer += dist[a[i][0]] * abs(a[i][1] - b[j][1])
            denom += dist[a[i][0]] * (a[i][1] + b[j][1])
            i += 1
            j += 1
    for x in a[i:]:
        numer += dist[x[0]] * x[1]
        denom += dist[x[0]] * x[1]
    for x in b[j:]:
        numer += dist[x[0]] * x[1]
        denom += dist[x[0]] * x[1]
    if denom == 0:
        return math.nan if numer == 0 or numer != numer else math.copysign(math.inf, numer)
    return numer / denom


def unifrac(samples: list[dict[str, float]], parent: list[int], length: list[float], names: list[str],
            weighted: bool, normalize: int = 1) -> list[float]:
    """normalize: 1 = default; 0 = the reference's -l (lists stay in post-order, unifrac.go:108-110);
    2 = -l as documented (sorted by id, not divided)."""
    n = len(parent)
    children: list[list[int]] = [[] for _ in range(n)]
    for v in range(1, n):
        children[parent[v]].append(v)  # ascending pre-order id = file order
    sys.setrecursionlimit(max(sys.getrecursionlimit(), n + 100))
    sets = []
    for abnd in samples:
        nodes: list = []
        _flat_nodes(abnd, 0, children, names, nodes)
        if normalize == 1:
            _normalize(nodes)
        elif normalize == 2:
            nodes.sort(key=lambda x: x[0])
        sets.append(nodes)
    fn = _weighted if weighted else _unweighted
    return [fn(sets[i], sets[j], length) for i in range(len(sets)) for j in range(i)]
