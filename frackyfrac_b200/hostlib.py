"""ctypes binding of libfrcfrc_host (the C++ host side above the C ABI).

Mirrors what frackyfrac keeps in Go: parser.ParseAbundance / ParseSparseAbundance
(parser/parser.go), the Newick reader, validateSpecies (frcfrc/unifrac.go:80-93),
enumerateNodes' flattening (:127-133) and the fmt.Fprintln float format.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libfrcfrc_host.so")
CLI_PATH = os.path.join(_HERE, "_build", "frcfrc")
SPRSPR_PATH = os.path.join(_HERE, "_build", "sprspr")

_lib = None


class HostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HostError(f"{LIB_PATH} is missing: run `python -m frackyfrac_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.frch_last_error.restype = C.c_char_p
        L.frch_tree_parse.restype = C.c_void_p
        L.frch_tree_parse.argtypes = [C.c_char_p, C.c_size_t]
        L.frch_tree_free.argtypes = [C.c_void_p]
        L.frch_tree_size.argtypes = [C.c_void_p]
        L.frch_tree_size.restype = C.c_int32
        L.frch_tree_parent.argtypes = [C.c_void_p]
        L.frch_tree_parent.restype = C.POINTER(C.c_int32)
        L.frch_tree_length.argtypes = [C.c_void_p]
        L.frch_tree_length.restype = C.POINTER(C.c_double)
        L.frch_tree_name.argtypes = [C.c_void_p, C.c_int32]
        L.frch_tree_name.restype = C.c_char_p
        L.frch_table_parse.restype = C.c_void_p
        L.frch_table_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        L.frch_table_free.argtypes = [C.c_void_p]
        L.frch_table_samples.argtypes = [C.c_void_p]
        L.frch_table_samples.restype = C.c_int64
        L.frch_table_row_ptr.argtypes = [C.c_void_p]
        L.frch_table_row_ptr.restype = C.POINTER(C.c_int64)
        L.frch_table_species_ids.argtypes = [C.c_void_p]
        L.frch_table_species_ids.restype = C.POINTER(C.c_int32)
        L.frch_table_values.argtypes = [C.c_void_p]
        L.frch_table_values.restype = C.POINTER(C.c_double)
        L.frch_table_species_name.argtypes = [C.c_void_p, C.c_int32]
        L.frch_table_species_name.restype = C.c_char_p
        L.frch_resolve.restype = C.c_void_p
        L.frch_resolve.argtypes = [C.c_void_p, C.c_void_p]
        L.frch_table_parse_mt.restype = C.c_void_p
        L.frch_table_parse_mt.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        L.frch_resolve_mt.restype = C.c_void_p
        L.frch_resolve_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.frch_csr_free.argtypes = [C.c_void_p]
        L.frch_csr_nnz.argtypes = [C.c_void_p]
        L.frch_csr_nnz.restype = C.c_int64
        L.frch_csr_row_ptr.argtypes = [C.c_void_p]
        L.frch_csr_row_ptr.restype = C.POINTER(C.c_int64)
        L.frch_csr_col.argtypes = [C.c_void_p]
        L.frch_csr_col.restype = C.POINTER(C.c_int32)
        L.frch_csr_val.argtypes = [C.c_void_p]
        L.frch_csr_val.restype = C.POINTER(C.c_double)
        L.frch_format_go.argtypes = [C.c_double, C.c_char_p]
        L.frch_format_lines.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_size_t)]
        L.frch_format_lines.restype = C.c_void_p
        L.frch_format_lines_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                            C.POINTER(C.c_size_t)]
        L.frch_format_lines_f32.restype = C.c_void_p
        L.frch_free.argtypes = [C.c_void_p]
        L.frch_to_sparse.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
        L.frch_to_sparse.restype = C.c_void_p
        L.frch_read_file.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
        L.frch_read_file.restype = C.c_void_p
        L.frch_write_file.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int]
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Tree:
    """Flattened tree: parent / length in pre-order ids, as the C ABI takes them."""

    def __init__(self, text: str | bytes):
        b = text.encode() if isinstance(text, str) else text
        self.h = lib().frch_tree_parse(b, len(b))
        if not self.h:
            raise HostError(lib().frch_last_error().decode())
        n = lib().frch_tree_size(self.h)
        self.parent = _arr(lib().frch_tree_parent(self.h), n, np.int32)
        self.length = _arr(lib().frch_tree_length(self.h), n, np.float64)

    def names(self):
        return [lib().frch_tree_name(self.h, v).decode() for v in range(len(self.parent))]

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.frch_tree_free(self.h)
            self.h = None


class Table:
    """Per-sample species maps (parser.ParseAbundance / ParseSparseAbundance)."""

    def __init__(self, text: str | bytes, sparse: bool, threads: int = 1):
        b = text.encode() if isinstance(text, str) else text
        self.threads = threads
        self.h = lib().frch_table_parse_mt(b, len(b), int(sparse), threads)
        if not self.h:
            raise HostError(lib().frch_last_error().decode())

    @property
    def n_samples(self) -> int:
        return lib().frch_table_samples(self.h)

    def maps(self) -> list[dict[str, float]]:
        L = lib()
        n = self.n_samples
        rp = _arr(L.frch_table_row_ptr(self.h), n + 1, np.int64)
        ids = _arr(L.frch_table_species_ids(self.h), int(rp[-1]), np.int32)
        vals = _arr(L.frch_table_values(self.h), int(rp[-1]), np.float64)
        return [{L.frch_table_species_name(self.h, int(ids[k])).decode(): float(vals[k])
                 for k in range(rp[s], rp[s + 1])} for s in range(n)]

    def resolve(self, tree: Tree):
        """validateSpecies + name -> leaf ids: (row_ptr, col, val) for the C ABI."""
        L = lib()
        c = L.frch_resolve_mt(self.h, tree.h, self.threads)
        if not c:
            raise HostError(L.frch_last_error().decode())
        try:
            nnz = L.frch_csr_nnz(c)
            rp = _arr(L.frch_csr_row_ptr(c), self.n_samples + 1, np.int64)
            return rp, _arr(L.frch_csr_col(c), nnz, np.int32), _arr(L.frch_csr_val(c), nnz, np.float64)
        finally:
            L.frch_csr_free(c)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.frch_table_free(self.h)
            self.h = None


def format_go(v: float) -> str:
    buf = C.create_string_buffer(48)
    lib().frch_format_go(float(v), buf)
    return buf.value.decode()


def format_lines(values, threads: int = 1) -> bytes:
    """One Go-%v formatted value per line (what frcfrc writes to stdout / -o)."""
    a = np.ascontiguousarray(values, np.float64)
    n = C.c_size_t()
    p = lib().frch_format_lines(a.ctypes.data, len(a), threads, C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        lib().frch_free(p)


def format_lines_f32(values, first_index: int = 0, ex_index=None, ex_value=None, threads: int = 1) -> bytes:
    """The same text from a float32 band of the fast paths (frc_next_f32) and its exceptions
    (frc_chunk_exceptions): every value widened exactly, as Go's float64(f), then printed with %v."""
    a = np.ascontiguousarray(values, np.float32)
    xi = np.ascontiguousarray(ex_index if ex_index is not None else [], np.int64)
    xv = np.ascontiguousarray(ex_value if ex_value is not None else [], np.float64)
    n = C.c_size_t()
    p = lib().frch_format_lines_f32(a.ctypes.data, len(a), first_index, xi.ctypes.data, xv.ctypes.data, len(xi), threads,
                                    C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        lib().frch_free(p)


def read_file(path: str) -> bytes:
    """aio.Open stand-in: the whole file, decoded by suffix (.gz, .zst)."""
    n = C.c_size_t()
    p = lib().frch_read_file(os.fsencode(path), C.byref(n))
    if not p:
        raise HostError(lib().frch_last_error().decode())
    try:
        return C.string_at(p, n.value)
    finally:
        lib().frch_free(p)


def write_file(path: str, chunks: list[bytes], threads: int = 1) -> None:
    """aio.Create stand-in: writes the chunks in order, encoded by suffix (.gz, .zst) on `threads` workers."""
    arr = (C.c_char_p * len(chunks))(*chunks)
    sizes = (C.c_size_t * len(chunks))(*[len(c) for c in chunks])
    if lib().frch_write_file(os.fsencode(path), arr, sizes, len(chunks), threads):
        raise HostError(lib().frch_last_error().decode())


def to_sparse(text: str | bytes, threads: int = 1) -> bytes:
    """sprspr: a dense table as sparse-format text (sprspr/sprspr.go:19-36)."""
    b = text.encode() if isinstance(text, str) else text
    n = C.c_size_t()
    p = lib().frch_to_sparse(b, len(b), threads, C.byref(n))
    if not p:
        raise HostError(lib().frch_last_error().decode())
    try:
        return C.string_at(p, n.value)
    finally:
        lib().frch_free(p)
