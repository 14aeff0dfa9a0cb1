"""ctypes bindings of the checkers: oracle/libvgoracle.so (our C restatement) and, when it has been
built in a container that holds /root/reference, oracle/_ref/libvgref.so (the unmodified reference).
TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libvgoracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libvgref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "varigraph_ref")
NOKMER = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u8(buf) -> np.ndarray:
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return np.frombuffer(buf, dtype=np.uint8)
    return np.ascontiguousarray(buf).view(np.uint8).reshape(-1)


def build_oracle() -> str:
    src = os.path.join(ORACLE_DIR, "vg_oracle.c")
    if not os.path.exists(ORACLE_LIB) or os.path.getmtime(src) > os.path.getmtime(ORACLE_LIB):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s", "-B", "oracle"], check=True)
    return ORACLE_LIB


class Oracle:
    def __init__(self):
        self.lib = L = ctypes.CDLL(build_oracle())
        L.vgo_nt4.restype = c_uint8
        L.vgo_nt4.argtypes = [c_uint8]
        L.vgo_hash64.restype = c_uint64
        L.vgo_hash64.argtypes = [c_uint64, c_uint64]
        L.vgo_sketch.restype = c_int64
        L.vgo_sketch.argtypes = [c_void_p, c_int64, c_uint32, c_void_p, c_int64]
        L.vgo_positions.restype = c_int64
        L.vgo_positions.argtypes = [c_void_p, c_int64, c_uint32, c_void_p]
        L.vgo_murmur3_x64_128_sum.restype = c_uint64
        L.vgo_murmur3_x64_128_sum.argtypes = [c_uint64, c_uint32]
        L.vgo_cbf_size.restype = c_uint64
        L.vgo_cbf_size.argtypes = [c_uint64, c_double]
        L.vgo_cbf_num_hashes.restype = c_uint32
        L.vgo_cbf_num_hashes.argtypes = [c_uint64, c_uint64]
        L.vgo_cbf_add.argtypes = [c_void_p, c_uint64, c_void_p, c_uint32, c_uint64]
        L.vgo_cbf_count.restype = c_uint8
        L.vgo_cbf_count.argtypes = [c_void_p, c_uint64, c_void_p, c_uint32, c_uint64]
        L.vgo_cbf_find.restype = c_int
        L.vgo_cbf_find.argtypes = [c_void_p, c_uint64, c_void_p, c_uint32, c_uint64]
        L.vgo_cbf_fill.restype = c_uint64
        L.vgo_cbf_fill.argtypes = [c_void_p, c_uint64, c_void_p, c_uint32, c_void_p, c_int64, c_uint32]
        L.vgo_index_create.restype = c_void_p
        L.vgo_index_create.argtypes = [c_void_p, c_uint64]
        L.vgo_index_destroy.argtypes = [c_void_p]
        L.vgo_count_lines.restype = c_uint64
        L.vgo_count_lines.argtypes = [c_void_p, c_void_p, c_int64, c_uint32, c_void_p, POINTER(c_uint64),
                                      POINTER(c_uint64)]
        L.vgo_fastq_to_lines.restype = c_int64
        L.vgo_fastq_to_lines.argtypes = [c_void_p, c_int64, c_void_p, c_int64, POINTER(c_uint64),
                                         POINTER(c_uint64), POINTER(c_int)]

    def hash64(self, x: int, mask: int) -> int:
        return int(self.lib.vgo_hash64(x, mask))

    def sketch(self, seq, k: int) -> np.ndarray:
        b = _u8(seq)
        out = np.empty(max(b.size, 1), dtype=np.uint64)
        n = self.lib.vgo_sketch(b.ctypes.data, b.size, k, out.ctypes.data, out.size)
        return out[:n].copy()

    def positions(self, buf, k: int) -> np.ndarray:
        """Per-byte keys of a staged chunk ('\\n' separates reads): key of the k-mer ending at each
        byte, or ~0 -- what vg_encode_positions must return."""
        b = _u8(buf)
        out = np.empty(b.size, dtype=np.uint64)
        self.lib.vgo_positions(b.ctypes.data, b.size, k, out.ctypes.data)
        return out

    def count_lines(self, keys: np.ndarray, buf, k: int, counts: np.ndarray | None = None):
        """-> (counts u8[n], positions, hits); pass `counts` to keep accumulating into it."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        b = _u8(buf)
        idx = self.lib.vgo_index_create(keys.ctypes.data, keys.size)
        if counts is None:
            counts = np.zeros(keys.size, dtype=np.uint8)
        hits = c_uint64(0)
        nreads = c_uint64(0)
        pos = self.lib.vgo_count_lines(idx, b.ctypes.data, b.size, k, counts.ctypes.data, byref(hits),
                                       byref(nreads))
        self.lib.vgo_index_destroy(idx)
        return counts, int(pos), int(hits.value)

    def fastq_to_lines(self, text: bytes):
        """-> (lines bytes, nreads, read_bases, status)"""
        cap = len(text) + 16
        out = np.empty(cap, dtype=np.uint8)
        nreads, bases, st = c_uint64(0), c_uint64(0), c_int(0)
        t = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(0, dtype=np.uint8)
        w = self.lib.vgo_fastq_to_lines(t.ctypes.data if t.size else None, t.size, out.ctypes.data, cap,
                                        byref(nreads), byref(bases), byref(st))
        return out[:w].tobytes(), int(nreads.value), int(bases.value), int(st.value)

    def cbf_fill(self, m: int, seeds: np.ndarray, seq, k: int, filt: np.ndarray | None = None):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        b = _u8(seq)
        if filt is None:
            filt = np.zeros(m, dtype=np.uint8)
        n = self.lib.vgo_cbf_fill(filt.ctypes.data, m, seeds.ctypes.data, seeds.size, b.ctypes.data, b.size, k)
        return filt, int(n)

    def cbf_count(self, filt, m, seeds, key):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        return int(self.lib.vgo_cbf_count(filt.ctypes.data, m, seeds.ctypes.data, seeds.size, int(key)))

    def cbf_find(self, filt, m, seeds, key):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        return int(self.lib.vgo_cbf_find(filt.ctypes.data, m, seeds.ctypes.data, seeds.size, int(key)))


class Reference:
    """The unmodified reference through oracle/ref_harness.cpp.  `available` is False on machines
    where oracle/_ref was never built (it is built wherever /root/reference exists)."""

    available = os.path.exists(REF_LIB)

    def __init__(self):
        if not self.available:
            raise RuntimeError("oracle/_ref/libvgref.so not built")
        self.lib = R = ctypes.CDLL(REF_LIB)
        R.ref_hash64.restype = c_uint64
        R.ref_hash64.argtypes = [c_uint64, c_uint64]
        R.ref_murmur3_x64_128_sum.restype = c_uint64
        R.ref_murmur3_x64_128_sum.argtypes = [c_uint64, c_uint32]
        R.ref_sketch.restype = c_int64
        R.ref_sketch.argtypes = [c_void_p, c_int64, c_uint32, c_void_p, c_int64]
        R.ref_cbf_create.restype = c_void_p
        R.ref_cbf_create.argtypes = [c_uint64, c_double, c_void_p, c_uint32]
        R.ref_cbf_destroy.argtypes = [c_void_p]
        R.ref_cbf_size.restype = c_uint64
        R.ref_cbf_size.argtypes = [c_void_p]
        R.ref_cbf_num_hashes.restype = c_uint32
        R.ref_cbf_num_hashes.argtypes = [c_void_p]
        R.ref_cbf_seeds.argtypes = [c_void_p, c_void_p]
        R.ref_cbf_filter.restype = POINTER(c_uint8)
        R.ref_cbf_filter.argtypes = [c_void_p]
        R.ref_cbf_fill.argtypes = [c_void_p, c_void_p, c_int64, c_uint32]
        R.ref_cbf_count.argtypes = [c_void_p, c_uint64]
        R.ref_cbf_find.argtypes = [c_void_p, c_uint64]
        R.ref_graph_load.restype = c_void_p
        R.ref_graph_load.argtypes = [c_char_p, c_uint32]
        R.ref_graph_destroy.argtypes = [c_void_p]
        R.ref_graph_num_kmers.restype = c_uint64
        R.ref_graph_num_kmers.argtypes = [c_void_p]
        R.ref_graph_kmer_len.restype = c_uint32
        R.ref_graph_kmer_len.argtypes = [c_void_p]
        R.ref_graph_keys.argtypes = [c_void_p, c_void_p]
        R.ref_graph_reset.argtypes = [c_void_p]
        R.ref_count_files.restype = c_double
        R.ref_count_files.argtypes = [c_void_p, POINTER(c_char_p), c_int, c_uint32, c_void_p, POINTER(c_uint64)]

        R.ref_map_create.restype = c_void_p
        R.ref_map_create.argtypes = [c_void_p, c_uint64, c_uint32]
        R.ref_map_destroy.argtypes = [c_void_p]
        R.ref_map_reset.argtypes = [c_void_p]
        R.ref_map_count_files.restype = c_double
        R.ref_map_count_files.argtypes = [c_void_p, POINTER(c_char_p), c_int, c_uint32, c_void_p, POINTER(c_uint64)]

    def map_create(self, keys: np.ndarray, k: int):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        return self.lib.ref_map_create(keys.ctypes.data, keys.size, k)

    def map_destroy(self, h) -> None:
        self.lib.ref_map_destroy(h)

    def map_count_files(self, h, n: int, files, threads: int, want_counts: bool = True):
        """FastqKmer::build_fastq_index over a caller-supplied key set -> (counts, read_bases, seconds)"""
        arr = (c_char_p * len(files))(*[os.fsencode(f) for f in files])
        counts = np.zeros(n, dtype=np.uint8) if want_counts else None
        rb = c_uint64(0)
        self.lib.ref_map_reset(h)
        sec = self.lib.ref_map_count_files(h, arr, len(files), threads,
                                           counts.ctypes.data if want_counts else None, byref(rb))
        return counts, int(rb.value), float(sec)

    def sketch(self, seq, k: int) -> np.ndarray:
        b = _u8(seq)
        out = np.empty(max(b.size, 1), dtype=np.uint64)
        n = self.lib.ref_sketch(b.ctypes.data, b.size, k, out.ctypes.data, out.size)
        return out[:n].copy()

    def graph_load(self, path: str, threads: int = 4):
        h = self.lib.ref_graph_load(os.fsencode(path), threads)
        n = int(self.lib.ref_graph_num_kmers(h))
        keys = np.empty(n, dtype=np.uint64)
        self.lib.ref_graph_keys(h, keys.ctypes.data)
        return h, keys, int(self.lib.ref_graph_kmer_len(h))

    def graph_destroy(self, h) -> None:
        self.lib.ref_graph_destroy(h)

    def count_files(self, h, n: int, files, threads: int):
        """-> (counts u8[n] in ref_graph_keys order, read_bases, seconds of build_fastq_index)"""
        arr = (c_char_p * len(files))(*[os.fsencode(f) for f in files])
        counts = np.zeros(n, dtype=np.uint8)
        rb = c_uint64(0)
        self.lib.ref_graph_reset(h)
        sec = self.lib.ref_count_files(h, arr, len(files), threads, counts.ctypes.data, byref(rb))
        return counts, int(rb.value), float(sec)
