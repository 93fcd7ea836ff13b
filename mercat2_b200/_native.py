"""ctypes binding of include/mercat2_b200.h -- the only way the Python layer reaches the GPU.

There is no CPU fallback: if the shared library is missing or no CUDA device can be opened, the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmercat2_b200.so"

MC2_HOST, MC2_DEVICE = 0, 1

# every symbol include/mercat2_b200.h declares: (name, restype, argtypes)
_VP, _U64, _I64, _INT = C.c_void_p, C.c_uint64, C.c_int64, C.c_int
_PP = C.POINTER(C.c_void_p)
_PU64 = C.POINTER(C.c_uint64)
SYMBOLS = [
    ("mc2_engine_create", _INT, [_INT, _PP]),
    ("mc2_engine_destroy", None, [_VP]),
    ("mc2_last_error", C.c_char_p, []),
    ("mc2_version", C.c_char_p, []),
    ("mc2_engine_trim", _INT, [_VP]),
    ("mc2_engine_set_option", _INT, [_VP, C.c_char_p, _I64]),
    ("mc2_engine_get_stat", _I64, [_VP, C.c_char_p]),
    ("mc2_engine_profile", _INT, [_VP, _VP, _U64, _PU64]),
    ("mc2_count_text", _INT, [_VP, _VP, _U64, _INT, _INT, _I64, _PP]),
    ("mc2_count_symbols", _INT, [_VP, _VP, _U64, _INT, _INT, _I64, _PP]),
    ("mc2_count_batch", _INT, [_VP, _VP, _VP, C.c_uint32, _INT, _INT, _I64, _VP]),
    ("mc2_count_text_rows", _INT, [_VP, _VP, _U64, _INT, _INT, _I64, _VP, _U64, _PU64]),
    ("mc2_count_text_rows_split", _INT, [_VP, _VP, _U64, _INT, _INT, _I64, _VP, _VP, _U64, _PU64]),
    ("mc2_count_sample", _INT, [_VP, _VP, _U64, _INT, _INT, _I64, _U64, _PP, _PU64, _PU64, _U64]),
    ("mc2_chunk_offsets", _INT, [_VP, _VP, _U64, _INT, _U64, _PU64, _U64, _PU64]),
    ("mc2_sample_begin", _INT, [_VP, _INT, _I64, _PP]),
    ("mc2_sample_add_text", _INT, [_VP, _VP, _U64, _INT, _U64, _PU64]),
    ("mc2_sample_add_file", _INT, [_VP, C.c_char_p, _INT, _U64, _PU64, _PU64]),
    ("mc2_sample_add_rows", _INT, [_VP, _VP, _VP, _U64]),
    ("mc2_sample_finish", _INT, [_VP, _PP]),
    ("mc2_sample_abort", None, [_VP]),
    ("mc2_table_rows", _U64, [_VP]),
    ("mc2_table_k", _INT, [_VP]),
    ("mc2_table_total", _U64, [_VP]),
    ("mc2_table_export", _INT, [_VP, _VP, _VP]),
    ("mc2_table_write_tsv", _INT, [_VP, C.c_char_p, C.c_char_p]),
    ("mc2_table_tsv", _INT, [_VP, C.c_char_p, _VP, _U64, _PU64]),
    ("mc2_table_free", None, [_VP]),
    ("mc2_table_info", _INT, [_VP, C.POINTER(_INT), C.POINTER(_INT), _PU64, _PU64]),
    ("mc2_table_device_rows", _INT, [_VP, _PP, _PP, _PU64]),
    ("mc2_table_lower_bound", _INT, [_VP, _VP, _U64, _VP]),
    ("mc2_table_export_packed", _INT, [_VP, _VP, _VP, _U64, _PU64]),
    ("mc2_table_export_wide", _INT, [_VP, _VP, _VP]),
    ("mc2_table_from_rows", _INT, [_VP, _INT, _INT, _INT, _VP, _VP, _U64, _INT, _VP, _VP, _U64, _PP]),
    ("mc2_table_tsv_body", _INT, [_VP, _VP, _U64, _PU64]),
    ("mc2_sample_dense", _INT, [_VP, _PP, _PU64, C.POINTER(_INT)]),
    ("mc2_sample_dense_plan", _INT, [_VP, _INT]),
    ("mc2_device_copy", _INT, [_VP, _VP, _VP, _U64]),
    ("mc2_keys_open", _INT, [_VP, _VP, _U64, _INT, _INT, _PP]),
    ("mc2_keys_sample", _INT, [_VP, _PP, _PU64]),
    ("mc2_keys_partition", _INT, [_VP, C.c_uint32]),
    ("mc2_partition_keys", _INT, [_VP, _VP, _U64, _INT, _INT, C.c_uint32, _PP]),
    ("mc2_keys_info", _INT, [_VP, _PP, _VP, _PU64, _PU64, _VP]),
    ("mc2_keys_free", None, [_VP]),
    ("mc2_sample_add_keys", _INT, [_VP, _VP, _U64, _INT, _U64, _U64]),
    ("mc2_count_exceptions", _INT, [_VP, _VP, _U64, _INT, _INT, _PP]),
    ("mc2_table_from_tsv", _INT, [_VP, _VP, _U64, _INT, _PP]),
    ("mc2_merge_tables", _INT, [_VP, _VP, C.c_uint32, _PP]),
    ("mc2_matrix_rows", _U64, [_VP]),
    ("mc2_matrix_k", _INT, [_VP]),
    ("mc2_matrix_export", _INT, [_VP, _VP, _VP]),
    ("mc2_matrix_write_tsv", _INT, [_VP, C.c_char_p, C.c_char_p, _VP, _INT]),
    ("mc2_matrix_free", None, [_VP]),
    ("mc2_merge_tables_reference", _INT, [_VP, _VP, C.c_uint32, C.c_char_p, C.c_char_p, _VP]),
    ("mc2_table_export_counts", _INT, [_VP, _VP, _VP]),
    ("mc2_matrix_top_rows", _INT, [_VP, C.c_uint32, _VP, C.POINTER(C.c_uint32)]),
    ("mc2_fastq_to_fasta", _INT, [_VP, _VP, _U64, _INT, _PP]),
    ("mc2_text_info", _INT, [_VP, _PP, _PU64]),
    ("mc2_text_export", _INT, [_VP, _VP, _U64]),
    ("mc2_text_free", None, [_VP]),
    ("mc2_protein_metrics", _INT, [_VP, _VP, _U64, _INT, _PP]),
    ("mc2_sequence_metrics", _INT, [_VP, _VP, _PU64, _U64, _PP]),
    ("mc2_metrics_records", _U64, [_VP]),
    ("mc2_metrics_export", _INT, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    ("mc2_metrics_free", None, [_VP]),
]


class Mc2Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"mercat2_b200 error {code}: {message}")
        self.code = code


class NonAsciiError(Mc2Error, ValueError):
    """Input holds bytes >= 0x80 (the reference would decode UTF-8 and count code points)."""


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen the in-tree library (building it first if it is absent and nvcc exists)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            from . import build as _build
            _build.build()
        lib = C.CDLL(str(LIB_PATH))
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(lib, name)          # AttributeError here = header and library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def _check(lib, status):
    if status < 0:
        msg = lib.mc2_last_error().decode(errors="replace")
        raise (NonAsciiError if status == -3 else Mc2Error)(status, msg)
    return status


def _as_buffer(data):
    """-> (address, nbytes, memspace, keepalive) for bytes-like, numpy, CUDA torch tensors or a DeviceText."""
    if isinstance(data, DeviceText):
        return data.ptr, data.nbytes, MC2_DEVICE, data
    if hasattr(data, "is_cuda"):                      # torch tensor (uint8), host or device
        t = data.contiguous()
        if t.element_size() != 1:
            raise TypeError("text tensor must be uint8")
        if t.is_cuda:
            # the engine works on its own stream: whatever produced this tensor must have finished
            import torch
            torch.cuda.current_stream(t.device).synchronize()
        return t.data_ptr(), t.numel(), (MC2_DEVICE if t.is_cuda else MC2_HOST), t
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data).view(np.uint8)
        return a.ctypes.data, a.size, MC2_HOST, a
    if isinstance(data, str):
        data = data.encode("utf-8", errors="surrogateescape")       # (non-ASCII is fine inside header lines only)
    mv = memoryview(data)
    a = np.frombuffer(mv, dtype=np.uint8)
    return a.ctypes.data, a.size, MC2_HOST, (mv, a)


class DeviceText:
    """A text produced on the device (e.g. FASTA converted from FASTQ): usable as input of every counting call."""

    def __init__(self, engine, handle):
        self._engine, self._h = engine, handle
        ptr, n = C.c_void_p(), C.c_uint64(0)
        _check(engine._lib, engine._lib.mc2_text_info(handle, C.byref(ptr), C.byref(n)))
        self.ptr, self.nbytes = ptr.value or 0, int(n.value)

    def to_bytes(self) -> bytes:
        buf = bytearray(self.nbytes)
        if self.nbytes:
            ref = (C.c_char * self.nbytes).from_buffer(buf)
            _check(self._engine._lib, self._engine._lib.mc2_text_export(self._h, ref, self.nbytes))
            del ref
        return bytes(buf)

    def close(self):
        if getattr(self, "_h", None) and self._engine._h:
            self._engine._lib.mc2_text_free(self._h)
        self._h = None

    def __del__(self):
        self.close()


class Table:
    """A finished count table (rows sorted by k-mer text)."""

    def __init__(self, engine, handle):
        self._engine, self._h = engine, handle

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "_h", None) and self._engine._h:
            self._engine._lib.mc2_table_free(self._h)
        self._h = None

    @property
    def rows(self) -> int:
        return int(self._engine._lib.mc2_table_rows(self._h))

    @property
    def k(self) -> int:
        return int(self._engine._lib.mc2_table_k(self._h))

    @property
    def total(self) -> int:
        return int(self._engine._lib.mc2_table_total(self._h))

    def arrays(self):
        """(kmers as uint8[rows, k], counts as uint64[rows])"""
        rows, k = self.rows, self.k
        kmers = np.empty((rows, k), dtype=np.uint8)
        counts = np.empty(rows, dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_export(self._h, kmers.ctypes.data, counts.ctypes.data))
        return kmers, counts

    def to_dict(self) -> dict:
        kmers, counts = self.arrays()
        k = self.k
        blob = kmers.tobytes().decode("ascii")
        return {blob[i * k:(i + 1) * k]: int(c) for i, c in enumerate(counts.tolist())}

    def tsv_bytes(self, basename: str) -> bytes:
        lib = self._engine._lib
        size = C.c_uint64(0)
        _check(lib, lib.mc2_table_tsv(self._h, basename.encode(), None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        _check(lib, lib.mc2_table_tsv(self._h, basename.encode(), buf, size.value, C.byref(size)))
        return buf.raw[:size.value]

    def write_tsv(self, path, basename: str) -> bool:
        """Write the per-sample TSV; returns False (and writes nothing) for an empty table."""
        lib = self._engine._lib
        return _check(lib, lib.mc2_table_write_tsv(self._h, os.fsencode(str(path)), basename.encode())) == 0

    # ---- device-resident exchange (mercat2_b200.distributed) --------------------------------------
    def info(self) -> dict:
        enc, kind, nf, nw = C.c_int(0), C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_info(self._h, C.byref(enc), C.byref(kind), C.byref(nf), C.byref(nw)))
        return {"encoding": enc.value, "key_kind": kind.value, "packed_rows": int(nf.value), "wide_rows": int(nw.value)}

    def device_rows(self):
        """(keys_ptr, counts_ptr, rows): device addresses of the packed rows, sorted by key; valid while the table lives."""
        keys, counts, n = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_device_rows(self._h, C.byref(keys), C.byref(counts), C.byref(n)))
        return keys.value or 0, counts.value or 0, int(n.value)

    def packed_to_host(self, keys_addr: int, counts_addr: int, capacity: int) -> int:
        """Copy the packed rows (64-bit order-preserving key + count) to host buffers at raw addresses (pinned memory:
        one DMA per array); returns the rows copied."""
        n = C.c_uint64(0)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_export_packed(self._h, keys_addr or None, counts_addr or None, capacity, C.byref(n)))
        return int(n.value)

    def packed_arrays(self):
        """(keys uint64[rows], counts uint64[rows]) of the packed rows, sorted by key."""
        n = self.info()["packed_rows"]
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint64)
        self.packed_to_host(keys.ctypes.data, counts.ctypes.data, n)
        return keys, counts

    def counts_array(self) -> np.ndarray:
        """The count vector of the table (sorted k-mer order) without the k-mer text: what the alpha-diversity metrics read."""
        counts = np.empty(self.rows, dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_export_counts(self._h, counts.ctypes.data, None))
        return counts

    def count_spectrum(self) -> dict:
        """Reductions of the count vector computed on the device: observed k-mers, sum and sum of squares of the counts,
        largest count and the number of k-mers seen exactly 1..10 times."""
        sp = np.zeros(16, dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_export_counts(self._h, None, sp.ctypes.data))
        return {"observed": int(sp[0]), "total": int(sp[1]), "sum_squares": int(sp[2]) + (int(sp[3]) << 64), "max": int(sp[4]),
                "seen_exactly": {i + 1: int(sp[5 + i]) for i in range(10)}}

    def lower_bound(self, splitters) -> list:
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        cuts = np.zeros(len(sp), dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_lower_bound(self._h, sp.ctypes.data, len(sp), cuts.ctypes.data))
        return [int(x) for x in cuts]

    def wide_arrays(self):
        """The literal-byte rows: (kmers uint8[rows, k], counts uint64[rows])."""
        nw, k = self.info()["wide_rows"], self.k
        kmers = np.empty((nw, k), dtype=np.uint8)
        counts = np.empty(nw, dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_table_export_wide(self._h, kmers.ctypes.data, counts.ctypes.data))
        return kmers, counts

    def tsv_body(self) -> bytes:
        """The TSV rows without the header line."""
        lib = self._engine._lib
        size = C.c_uint64(0)
        _check(lib, lib.mc2_table_tsv_body(self._h, None, 0, C.byref(size)))
        buf = bytearray(size.value)
        if size.value:
            ref = (C.c_char * size.value).from_buffer(buf)
            _check(lib, lib.mc2_table_tsv_body(self._h, ref, size.value, C.byref(size)))
            del ref
        return bytes(buf)


class Matrix:
    """Sample x k-mer count matrix (merge_tsv): rows = sorted union of the samples' k-mers, one column per sample."""

    def __init__(self, engine, handle, samples):
        self._engine, self._h, self.samples = engine, handle, samples

    def close(self):
        if getattr(self, "_h", None) and self._engine._h:
            self._engine._lib.mc2_matrix_free(self._h)
        self._h = None

    def __del__(self):
        self.close()

    @property
    def rows(self) -> int:
        return int(self._engine._lib.mc2_matrix_rows(self._h))

    @property
    def k(self) -> int:
        return int(self._engine._lib.mc2_matrix_k(self._h))

    def arrays(self):
        """(kmers uint8[rows, k], counts uint64[rows, samples])"""
        rows, k = self.rows, self.k
        kmers = np.empty((rows, k), dtype=np.uint8)
        counts = np.empty((rows, self.samples), dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_matrix_export(self._h, kmers.ctypes.data, counts.ctypes.data))
        return kmers, counts

    def top_rows(self, top: int = 5) -> list:
        """Row indices of the `top` k-mers with the largest count sum (= mean) over the samples, largest first, earlier
        rows winning ties (the selection of lib/mercat2_figures.py:50-65)."""
        out = np.zeros(max(top, 1), dtype=np.uint64)
        found = C.c_uint32(0)
        lib = self._engine._lib
        _check(lib, lib.mc2_matrix_top_rows(self._h, top, out.ctypes.data, C.byref(found)))
        return [int(x) for x in out[:found.value]]

    def write_tsv(self, path, corner: str, names, transposed: bool = False):
        names = [str(n).encode() for n in names]
        if len(names) != self.samples:
            raise ValueError("one name per sample")
        arr = (C.c_char_p * len(names))(*names)
        lib = self._engine._lib
        _check(lib, lib.mc2_matrix_write_tsv(self._h, os.fsencode(str(path)), corner.encode(), arr, 1 if transposed else 0))


class Engine:
    """One CUDA device + stream.  Not thread-safe: one host thread per engine."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        handle = C.c_void_p()
        _check(self._lib, self._lib.mc2_engine_create(device, C.byref(handle)))
        self._h = handle
        self.device = device
        # tuning / test overrides: MERCAT2_B200_OPTIONS="name=value,name=value"
        for item in filter(None, os.environ.get("MERCAT2_B200_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            self.set_option(name.strip(), int(value))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mc2_engine_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """Release the device workspace the engine keeps between calls (and the pool's cached memory)."""
        _check(self._lib, self._lib.mc2_engine_trim(self._h))

    def set_option(self, name: str, value: int):
        _check(self._lib, self._lib.mc2_engine_set_option(self._h, name.encode(), int(value)))

    def stat(self, name: str) -> int:
        return int(self._lib.mc2_engine_get_stat(self._h, name.encode()))

    def profile(self) -> dict:
        """Per-kernel CUDA-event totals gathered while option 'profile' is on."""
        import json
        size = C.c_uint64(0)
        _check(self._lib, self._lib.mc2_engine_profile(self._h, None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value + 1)
        _check(self._lib, self._lib.mc2_engine_profile(self._h, buf, size.value, C.byref(size)))
        return json.loads(buf.raw[:size.value].decode())

    def count_text(self, data, k: int, min_count: int) -> Table:
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_count_text(self._h, addr, n, space, k, min_count, C.byref(out)))
        return Table(self, out)

    def count_text_rows(self, data, k: int, min_count: int, rows_addr: int, capacity: int) -> int:
        """count_text with the table delivered as packed rows (16 bytes: key, count; sorted) into host memory at
        `rows_addr` (room for `capacity` rows; pinned memory lets the download overlap the counting).  Returns the rows."""
        addr, n, space, keep = _as_buffer(data)
        rows = C.c_uint64(0)
        _check(self._lib, self._lib.mc2_count_text_rows(self._h, addr, n, space, k, min_count, rows_addr or None, capacity, C.byref(rows)))
        return int(rows.value)

    def count_text_rows_split(self, data, k: int, min_count: int, keys_addr: int, counts_addr: int, capacity: int) -> int:
        """count_text_rows with the rows as two host arrays: uint64 keys at `keys_addr`, uint32 counts at `counts_addr`
        (12 bytes per row over PCIe instead of 16).  Raises when a count needs more than 32 bits.  Returns the rows."""
        addr, n, space, keep = _as_buffer(data)
        rows = C.c_uint64(0)
        _check(self._lib, self._lib.mc2_count_text_rows_split(self._h, addr, n, space, k, min_count, keys_addr or None, counts_addr or None,
                                                               capacity, C.byref(rows)))
        return int(rows.value)

    def count_batch(self, texts, k: int, min_count: int) -> list:
        """One table per text (each text = one sample file smaller than the -s trigger), counted in a single pass over
        all of them; identical to [count_text(t, k, min_count) for t in texts]."""
        bufs = [_as_buffer(t) for t in texts]
        if not bufs:
            return []
        spaces = {b[2] for b in bufs}
        if len(spaces) > 1:
            raise ValueError("all texts of a batch must live in the same memory space")
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b[0] or None for b in bufs])
        sizes = (C.c_uint64 * n)(*[b[1] for b in bufs])
        outs = (C.c_void_p * n)()
        _check(self._lib, self._lib.mc2_count_batch(self._h, ptrs, sizes, n, spaces.pop(), k, min_count, outs))
        return [Table(self, C.c_void_p(outs[j])) for j in range(n)]

    def count_symbols(self, data, k: int, min_count: int = 1) -> Table:
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_count_symbols(self._h, addr, n, space, k, min_count, C.byref(out)))
        return Table(self, out)

    def count_sample(self, data, k: int, min_count: int, chunk_bytes: int = 0):
        """-> (Table, piece start offsets)"""
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        cap = 2 + (n // chunk_bytes if chunk_bytes else 0)
        offs = (C.c_uint64 * cap)()
        nchunks = C.c_uint64(0)
        _check(self._lib, self._lib.mc2_count_sample(self._h, addr, n, space, k, min_count, chunk_bytes,
                                                     C.byref(out), C.byref(nchunks), offs, cap))
        return Table(self, out), [int(offs[i]) for i in range(min(cap, nchunks.value))]

    def chunk_offsets(self, data, chunk_bytes: int) -> list:
        """Byte offsets where the reference's Chunker starts each piece of this text."""
        addr, n, space, keep = _as_buffer(data)
        cap = 2 + (n // chunk_bytes if chunk_bytes else 0)
        offs = (C.c_uint64 * cap)()
        npieces = C.c_uint64(0)
        _check(self._lib, self._lib.mc2_chunk_offsets(self._h, addr, n, space, chunk_bytes, offs, cap, C.byref(npieces)))
        return [int(offs[i]) for i in range(min(cap, npieces.value))]

    def table_from_rows(self, k: int, encoding: int, key_kind: int, keys_ptr: int, counts_ptr: int, rows: int, on_device: bool,
                        wide_kmers=None, wide_counts=None) -> Table:
        """Table from packed rows at raw addresses (equal keys are summed) plus optional literal-byte rows (numpy)."""
        out = C.c_void_p()
        nw = 0 if wide_counts is None else len(wide_counts)
        wk = np.ascontiguousarray(wide_kmers, dtype=np.uint8) if nw else None
        wc = np.ascontiguousarray(wide_counts, dtype=np.uint64) if nw else None
        _check(self._lib, self._lib.mc2_table_from_rows(self._h, k, encoding, key_kind, keys_ptr or None, counts_ptr or None, rows,
                                                        MC2_DEVICE if on_device else MC2_HOST, wk.ctypes.data if nw else None,
                                                        wc.ctypes.data if nw else None, nw, C.byref(out)))
        return Table(self, out)

    def table_from_tsv(self, data) -> Table:
        """Parse the bytes of a per-sample TSV file (header line, then '<k-mer>\\t<count>' rows) on the device."""
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_table_from_tsv(self._h, addr, n, space, C.byref(out)))
        return Table(self, out)

    def merge_tables(self, tables) -> Matrix:
        """Sample x k-mer matrix of several tables (columns in the given order)."""
        tables = list(tables)
        arr = (C.c_void_p * len(tables))(*[t._h for t in tables])
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_merge_tables(self._h, arr, len(tables), C.byref(out)))
        return Matrix(self, out, len(tables))

    def merge_tables_reference(self, tables, path, corner: str, names):
        """Write merge_tsv's combined table exactly as the reference does (lib/mercat2_report.py:98-160)."""
        tables = list(tables)
        names = [str(n).encode() for n in names]
        if len(names) != len(tables):
            raise ValueError("one name per table")
        arr = (C.c_void_p * len(tables))(*[t._h for t in tables])
        narr = (C.c_char_p * len(names))(*names)
        _check(self._lib, self._lib.mc2_merge_tables_reference(self._h, arr, len(tables), os.fsencode(str(path)), corner.encode(), narr))

    def partition_keys(self, data, k: int, groups: int) -> "Keys":
        """Order-preserving 64-bit keys of every window of a plain nucleotide FASTA text, grouped into `groups`
        ascending key ranges (boundaries from this text's own key sample)."""
        keys = self.open_keys(data, k)
        keys.partition(groups)
        return keys

    def open_keys(self, data, k: int) -> "Keys":
        """Parse a plain nucleotide FASTA text into packed symbols and sample its key prefixes; the caller may sum
        `Keys.sample_ptr` over ranks (all-reduce) before `Keys.partition(groups)`."""
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_keys_open(self._h, addr, n, space, k, C.byref(out)))
        return Keys(self, out)

    def fastq_to_fasta(self, data) -> DeviceText:
        """FASTQ text -> FASTA text on the device (lib/mercat2_fasta.py:175-198: `sed -n '1~4s/^@/>/p;2~4p'`)."""
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_fastq_to_fasta(self._h, addr, n, space, C.byref(out)))
        return DeviceText(self, out)

    def count_exceptions(self, data, k: int) -> Table:
        """Unfiltered table of the windows that hold a symbol outside ACGT (literal-byte rows)."""
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_count_exceptions(self._h, addr, n, space, k, C.byref(out)))
        return Table(self, out)

    def device_copy(self, dst_ptr: int, src_ptr: int, nbytes: int):
        _check(self._lib, self._lib.mc2_device_copy(self._h, dst_ptr, src_ptr, nbytes))

    def sample(self, k: int, min_count: int) -> "Sample":
        return Sample(self, k, min_count)

    def _metrics_arrays(self, handle):
        lib = self._lib
        try:
            r = int(lib.mc2_metrics_records(handle))
            res = {
                "header_off": np.empty(r, np.uint64), "header_len": np.empty(r, np.uint32),
                "length": np.empty(r, np.uint64), "pi": np.empty(r, np.float64), "mw": np.empty(r, np.float64),
                "hydro": np.empty(r, np.float64), "status": np.empty(r, np.uint8),
            }
            _check(lib, lib.mc2_metrics_export(handle, *[res[key].ctypes.data for key in
                                                         ("header_off", "header_len", "length", "pi", "mw", "hydro", "status")]))
        finally:
            lib.mc2_metrics_free(handle)
        return res

    def sequence_metrics(self, seqs):
        """Unrounded pI / MW / hydropathy for a list of raw sequences (str or bytes)."""
        blobs = [s.encode("latin-1", errors="replace") if isinstance(s, str) else bytes(s) for s in seqs]
        offsets = np.zeros(len(blobs) + 1, np.uint64)
        if blobs:
            offsets[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
        data = np.frombuffer(b"".join(blobs) + b"\0", dtype=np.uint8)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_sequence_metrics(self._h, data.ctypes.data, offsets.ctypes.data_as(_PU64),
                                                         len(blobs), C.byref(out)))
        return self._metrics_arrays(out)

    def protein_metrics(self, data):
        """-> dict of numpy arrays: header_off, header_len, length, pi, mw, hydro, status (unrounded)."""
        addr, n, space, keep = _as_buffer(data)
        out = C.c_void_p()
        _check(self._lib, self._lib.mc2_protein_metrics(self._h, addr, n, space, C.byref(out)))
        return self._metrics_arrays(out)


class Keys:
    """Packed symbols of one text on the device, then (after partition) its keys grouped by ascending key range."""

    def __init__(self, engine, handle):
        self._engine, self._h, self.groups = engine, handle, 0
        ptr, n = C.c_void_p(), C.c_uint64(0)
        lib = engine._lib
        _check(lib, lib.mc2_keys_sample(handle, C.byref(ptr), C.byref(n)))
        self.sample_ptr, self.sample_entries = ptr.value or 0, int(n.value)     # uint32[entries] on the device

    def partition(self, groups: int):
        lib = self._engine._lib
        _check(lib, lib.mc2_keys_partition(self._h, groups))
        self.groups = groups
        ptr, total, exc = C.c_void_p(), C.c_uint64(0), C.c_uint64(0)
        sizes = np.zeros(groups, dtype=np.uint64)
        bounds = np.zeros(groups + 1, dtype=np.uint64)
        _check(lib, lib.mc2_keys_info(self._h, C.byref(ptr), sizes.ctypes.data, C.byref(total), C.byref(exc), bounds.ctypes.data))
        self.ptr, self.total, self.exception_symbols = ptr.value or 0, int(total.value), int(exc.value)
        self.sizes = [int(x) for x in sizes]
        self.bounds = [int(x) for x in bounds]          # group g holds the keys with 32-bit prefix in [bounds[g], bounds[g+1])

    def close(self):
        if getattr(self, "_h", None) and self._engine._h:
            self._engine._lib.mc2_keys_free(self._h)
        self._h = None

    def __del__(self):
        self.close()


class Sample:
    """Several files of one sample: each add_text() is chunked and counted like one chunk-file list."""

    def __init__(self, engine: Engine, k: int, min_count: int):
        self._engine = engine
        self._h = C.c_void_p()
        _check(engine._lib, engine._lib.mc2_sample_begin(engine._h, k, min_count, C.byref(self._h)))

    def add_text(self, data, chunk_bytes: int = 0) -> int:
        addr, n, space, keep = _as_buffer(data)
        nchunks = C.c_uint64(0)
        lib = self._engine._lib
        _check(lib, lib.mc2_sample_add_text(self._h, addr, n, space, chunk_bytes, C.byref(nchunks)))
        return int(nchunks.value)

    def add_file(self, path, chunk_bytes: int = 0, gunzip=None) -> int:
        """Stream a file through the engine's own reader (pinned buffers, inflate for '.gz'); returns the pieces counted."""
        nchunks, nbytes = C.c_uint64(0), C.c_uint64(0)
        lib = self._engine._lib
        flag = -1 if gunzip is None else int(bool(gunzip))
        _check(lib, lib.mc2_sample_add_file(self._h, os.fsencode(str(path)), flag, chunk_bytes, C.byref(nchunks), C.byref(nbytes)))
        self.text_bytes = int(nbytes.value)
        return int(nchunks.value)

    def add_rows(self, kmers, counts):
        """Add counted rows: kmers uint8[rows, k], counts uint64[rows] (equal k-mers are summed at finish)."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint8)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        lib = self._engine._lib
        _check(lib, lib.mc2_sample_add_rows(self._h, kmers.ctypes.data, counts.ctypes.data, len(counts)))

    def add_keys(self, keys_ptr: int, n: int, on_device: bool = True, prefix_lo: int = 0, prefix_hi: int = 0):
        """Count n packed keys at a raw address: ALL occurrences of one key range of the current chunk (min_count applies
        to their totals).  prefix_lo / prefix_hi: the 32-bit prefix range they lie in (Keys.bounds), 0, 0 = unknown."""
        lib = self._engine._lib
        _check(lib, lib.mc2_sample_add_keys(self._h, keys_ptr or None, n, MC2_DEVICE if on_device else MC2_HOST, prefix_lo, prefix_hi))

    def dense(self):
        """(device address, bins, encoding) of the per-sample dense table; bins == 0 if the sample is not on the dense path."""
        ptr, bins, enc = C.c_void_p(), C.c_uint64(0), C.c_int(0)
        lib = self._engine._lib
        _check(lib, lib.mc2_sample_dense(self._h, C.byref(ptr), C.byref(bins), C.byref(enc)))
        return ptr.value or 0, int(bins.value), enc.value

    def dense_plan(self, encoding: int):
        lib = self._engine._lib
        _check(lib, lib.mc2_sample_dense_plan(self._h, encoding))

    def finish(self) -> Table:
        out = C.c_void_p()
        lib = self._engine._lib
        h, self._h = self._h, None
        _check(lib, lib.mc2_sample_finish(h, C.byref(out)))
        return Table(self._engine, out)

    def __del__(self):
        if getattr(self, "_h", None) and self._engine._h:
            self._engine._lib.mc2_sample_abort(self._h)
            self._h = None


_default_engines: dict = {}


def default_engine(device: int | None = None) -> Engine:
    """Process-wide engine for the drop-in module functions (device from MERCAT2_B200_DEVICE / LOCAL_RANK)."""
    if device is None:
        device = int(os.environ.get("MERCAT2_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    eng = _default_engines.get(device)
    if eng is None:
        eng = _default_engines[device] = Engine(device)
    return eng


def bind_to_gpu_numa(device: int) -> bool:
    """Pin the calling process to the CPUs next to GPU `device` (NVML's ideal affinity).  Host buffers allocated
    afterwards (pinned memory in particular) land on that NUMA node, so uploads do not cross sockets when several
    ranks feed several GPUs.  Opt-in: it changes the process's CPU affinity.  Returns False if NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device
        if visible:
            ids = [x.strip() for x in visible.split(",") if x.strip()]
            if device < len(ids) and ids[device].isdigit():
                index = int(ids[device])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False
