"""ctypes binding of the parity oracle (oracle/unifrac_oracle.c).

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module;
nothing under frackyfrac_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
CLI_PATH = os.path.join(_HERE, "_build", "frcfrc_oracle")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (idempotent)."""
    src = os.path.join(_HERE, "unifrac_oracle.c")
    stale = (not os.path.exists(_LIB_PATH) or not os.path.exists(CLI_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_tree_parse.restype = C.c_void_p
        L.orc_tree_parse.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_tree_num_nodes.restype = C.c_int64
        L.orc_tree_num_nodes.argtypes = [C.c_void_p]
        L.orc_table_parse.restype = C.c_void_p
        L.orc_table_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_table_num_samples.restype = C.c_int64
        L.orc_table_num_samples.argtypes = [C.c_void_p]
        L.orc_table_sample_size.restype = C.c_int64
        L.orc_table_sample_size.argtypes = [C.c_void_p, C.c_int64]
        L.orc_table_entry_name.restype = C.c_char_p
        L.orc_table_entry_name.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_table_entry_value.restype = C.c_double
        L.orc_table_entry_value.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_tree_from_flat.restype = C.c_void_p
        L.orc_tree_from_flat.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_table_from_csr.restype = C.c_void_p
        L.orc_table_from_csr.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_validate_species.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_unifrac.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_unifrac_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_flat_nodes.restype = C.c_int64
        L.orc_flat_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_void_p]
        L.orc_tree_flatten.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_format_go.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _err() -> str:
    return lib().orc_last_error().decode("utf-8", "replace")


class Tree:
    def __init__(self, handle):
        if not handle:
            raise OracleError(_err())
        self.h = handle

    @classmethod
    def parse(cls, text: str | bytes) -> "Tree":
        b = text.encode() if isinstance(text, str) else text
        return cls(lib().orc_tree_parse(b, len(b)))

    @classmethod
    def from_flat(cls, parent: np.ndarray, length: np.ndarray) -> "Tree":
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        length = np.ascontiguousarray(length, dtype=np.float64)
        return cls(lib().orc_tree_from_flat(len(parent), parent.ctypes.data, length.ctypes.data))

    @property
    def num_nodes(self) -> int:
        return lib().orc_tree_num_nodes(self.h)

    def flatten(self):
        n = self.num_nodes
        parent = np.empty(n, np.int32)
        length = np.empty(n, np.float64)
        lib().orc_tree_flatten(self.h, parent.ctypes.data, length.ctypes.data)
        return parent, length

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_tree_free(self.h)
            self.h = None


class Table:
    def __init__(self, handle):
        if not handle:
            raise OracleError(_err())
        self.h = handle

    @classmethod
    def parse(cls, text: str | bytes, sparse: bool) -> "Table":
        b = text.encode() if isinstance(text, str) else text
        return cls(lib().orc_table_parse(b, len(b), int(sparse)))

    @classmethod
    def from_csr(cls, row_ptr, leaf_id, val) -> "Table":
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        leaf_id = np.ascontiguousarray(leaf_id, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        return cls(lib().orc_table_from_csr(len(row_ptr) - 1, row_ptr.ctypes.data,
                                            leaf_id.ctypes.data, val.ctypes.data))

    @property
    def num_samples(self) -> int:
        return lib().orc_table_num_samples(self.h)

    def maps(self) -> list[dict[str, float]]:
        L = lib()
        out = []
        for s in range(self.num_samples):
            out.append({L.orc_table_entry_name(self.h, s, k).decode(): L.orc_table_entry_value(self.h, s, k)
                        for k in range(L.orc_table_sample_size(self.h, s))})
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_table_free(self.h)
            self.h = None


def validate_species(tab: Table, tree: Tree) -> None:
    if lib().orc_validate_species(tab.h, tree.h):
        raise OracleError(_err())


def unifrac(tab: Table, tree: Tree, weighted: bool, normalize: int = 1, nthreads: int = 1) -> np.ndarray:
    """All n(n-1)/2 distances in IterPairs order.  normalize: see unifrac_oracle.h."""
    n = tab.num_samples
    out = np.empty(max(n * (n - 1) // 2, 0), np.float64)
    if lib().orc_unifrac(tab.h, tree.h, int(weighted), int(normalize), nthreads, out.ctypes.data):
        raise OracleError(_err())
    return out


def unifrac_rows(tab: Table, tree: Tree, weighted: bool, normalize: int, nthreads: int,
                 row_begin: int, row_end: int):
    """Distances of pairs (i, j<i), i in [row_begin,row_end); returns (out, embed_s, pair_s)."""
    npairs = row_end * (row_end - 1) // 2 - row_begin * (row_begin - 1) // 2
    out = np.empty(max(npairs, 0), np.float64)
    te, tp = C.c_double(0), C.c_double(0)
    if lib().orc_unifrac_rows(tab.h, tree.h, int(weighted), int(normalize), nthreads,
                              row_begin, row_end, out.ctypes.data, C.byref(te), C.byref(tp)):
        raise OracleError(_err())
    return out, te.value, tp.value


def flat_nodes(tab: Table, tree: Tree, normalize: int, sample: int):
    cap = tree.num_nodes
    ids = np.empty(cap, np.int64)
    vals = np.empty(cap, np.float64)
    n = lib().orc_flat_nodes(tab.h, tree.h, int(normalize), sample, cap, ids.ctypes.data, vals.ctypes.data)
    return ids[:n].copy(), vals[:n].copy()


def format_go(v: float) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_format_go(float(v), buf, 64)
    return buf.value.decode()
