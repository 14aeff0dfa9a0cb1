"""ctypes binding of libvgb200.so (include/vgb200.h) for the tests and bench.py.

The product's host side is C++ (varigraph_b200/host/, mirroring the reference's classes);
this module only lets Python drive the same C ABI.  There is no fallback: if the shared
library is missing, importing it raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VG_LIB") or os.path.join(_HERE, "libvgb200.so")  # VG_LIB: an A/B build of the same sources

VG_OK, VG_E_INVALID, VG_E_CUDA, VG_E_NOMEM, VG_E_IO, VG_E_STATE = 0, -1, -2, -3, -4, -5


class VgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvgb200 error {code}: {msg}")
        self.code = code


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). varigraph_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P = POINTER
    sig = {
        "vg_last_error": (c_char_p, []),
        "vg_version": (c_int, []),
        "vg_ctx_create": (c_int, [c_int, c_int, P(c_void_p)]),
        "vg_ctx_destroy": (c_int, [c_void_p]),
        "vg_ctx_device": (c_int, [c_void_p]),
        "vg_ctx_synchronize": (c_int, [c_void_p]),
        "vg_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
        "vg_probe_random_sectors": (c_int, [c_void_p, c_uint64, c_uint32, P(c_double), P(c_double)]),
        "vg_index_create": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_double, P(c_void_p)]),
        "vg_index_create_device": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_double, P(c_void_p)]),
        "vg_index_destroy": (c_int, [c_void_p]),
        "vg_index_size": (c_uint64, [c_void_p]),
        "vg_index_table_bytes": (c_uint64, [c_void_p]),
        "vg_index_partitions": (c_uint32, [c_void_p]),
        "vg_index_slices": (c_uint32, [c_void_p]),
        "vg_index_launches": (c_uint64, [c_void_p]),
        "vg_index_duplicates": (c_uint64, [c_void_p]),
        "vg_count_h2d_bytes": (c_uint64, [c_void_p]),
        "vg_index_set_timing": (c_int, [c_void_p, c_int]),
        "vg_index_timing": (c_int, [c_void_p, P(c_double), P(c_double), P(c_uint64), P(c_uint64)]),
        "vg_count_begin": (c_int, [c_void_p]),
        "vg_count_submit": (c_int, [c_void_p, c_void_p, c_uint64]),
        "vg_count_submit_device": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
        "vg_count_files": (c_int, [c_void_p, P(c_char_p), c_int, c_int, P(c_uint64)]),
        "vg_count_files_multi": (c_int, [P(c_void_p), c_int, P(c_char_p), c_int, c_int, P(c_uint64)]),
        "vg_count_flush": (c_int, [c_void_p]),
        "vg_index_fastq_blocks": (c_uint64, [c_void_p]),
        "vg_fastq_record_boundary": (ctypes.c_int64, [c_char_p, c_uint64, c_uint64]),
        "vg_fastq_strip_block": (ctypes.c_int64, [c_char_p, c_uint64, c_int, c_void_p, P(c_uint64), P(ctypes.c_int64)]),
        "vg_gunzip_parallel": (c_int, [c_char_p, c_int, c_uint64, P(c_void_p), P(c_uint64)]),
        "vg_gunzip_free": (None, [c_void_p]),
        "vg_crc32": (c_uint32, [c_uint32, c_char_p, c_uint64]),
        "vg_index_set_flags": (c_int, [c_void_p, c_void_p]),
        "vg_count_histogram": (c_int, [c_void_p, c_void_p]),
        "vg_count_end": (c_int, [c_void_p, c_void_p, P(c_uint64), P(c_uint64)]),
        "vg_count_extract_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
        "vg_index_slots": (c_uint64, [c_void_p]),
        "vg_index_slot_perm": (c_int, [c_void_p, c_void_p]),
        "vg_count_end_slots": (c_int, [c_void_p, c_void_p, P(c_uint64), P(c_uint64)]),
        "vg_count_slots_device": (c_int, [c_void_p, P(c_void_p)]),
        "vg_count_stats": (c_int, [c_void_p, P(c_uint64), P(c_uint64)]),
        "vg_count_keys": (c_uint64, [c_void_p]),
        "vg_encode_positions_device": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_void_p, c_void_p]),
        "vg_encode_positions": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_void_p]),
        "vg_cbf_create": (c_int, [c_void_p, c_uint64, c_uint32, c_void_p, P(c_void_p)]),
        "vg_cbf_destroy": (c_int, [c_void_p]),
        "vg_cbf_add_sequence": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, P(c_uint64)]),
        "vg_cbf_download": (c_int, [c_void_p, c_void_p]),
        "vg_cbf_query": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
        "vg_comm_create": (c_int, [c_void_p, c_int, c_int, c_uint64, P(c_void_p)]),
        "vg_comm_create_local": (c_int, [P(c_void_p), c_int, c_uint64, P(c_void_p)]),
        "vg_comm_handle": (c_int, [c_void_p, c_void_p]),
        "vg_comm_connect": (c_int, [c_void_p, c_void_p]),
        "vg_comm_destroy": (c_int, [c_void_p]),
        "vg_comm_rank": (c_int, [c_void_p]),
        "vg_comm_world": (c_int, [c_void_p]),
        "vg_comm_launches": (c_uint64, [c_void_p]),
        "vg_comm_barrier": (c_int, [c_void_p]),
        "vg_comm_check": (c_int, [c_void_p]),
        "vg_count_allreduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        "vg_index_replicate": (c_int, [c_void_p, c_int, c_void_p, P(c_void_p)]),
        "vg_count_allreduce_slots": (c_int, [c_void_p, c_void_p, c_void_p, P(c_void_p)]),
        "vg_index_create_sharded": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_double, c_uint64, P(c_void_p)]),
        "vg_index_create_sharded_device": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_double, c_uint64, P(c_void_p)]),
        "vg_index_own_keys": (c_uint64, [c_void_p]),
        "vg_count_room": (c_uint64, [c_void_p]),
        "vg_host_alloc": (c_int, [P(c_void_p), c_uint64]),
        "vg_host_free": (c_int, [c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _chk(rc: int) -> None:
    if rc != VG_OK:
        raise VgError(rc, lib.vg_last_error().decode("utf-8", "replace"))


def _ptr(a: np.ndarray) -> c_void_p:
    return c_void_p(a.ctypes.data)


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return np.frombuffer(buf, dtype=np.uint8)
    a = np.ascontiguousarray(buf)
    return a.view(np.uint8).reshape(-1)


class Context:
    """One CUDA device (vg_ctx)."""

    def __init__(self, device: int = 0, buffer_mb: int = 64):
        h = c_void_p()
        _chk(lib.vg_ctx_create(device, buffer_mb, byref(h)))
        self._h = h
        self.device = device

    def synchronize(self) -> None:
        _chk(lib.vg_ctx_synchronize(self._h))

    def set_stream(self, stream: int) -> None:
        _chk(lib.vg_ctx_set_stream(self._h, c_void_p(stream)))

    def probe_random_sectors(self, table_bytes: int, rounds: int = 64):
        """-> (GB/s, sectors/s) of uniform random 32-byte gathers over a table of table_bytes."""
        gb, sec = c_double(0), c_double(0)
        _chk(lib.vg_probe_random_sectors(self._h, table_bytes, rounds, byref(gb), byref(sec)))
        return float(gb.value), float(sec.value)

    def encode_positions(self, bases, k: int) -> np.ndarray:
        b = _as_u8(bases)
        out = np.empty(b.size, dtype=np.uint64)
        _chk(lib.vg_encode_positions(self._h, _ptr(b), b.size, k, _ptr(out)))
        return out

    def encode_positions_device(self, dev_ptr: int, nbytes: int, k: int, out_ptr: int, stream: int = 0) -> None:
        _chk(lib.vg_encode_positions_device(self._h, c_void_p(dev_ptr), nbytes, k, c_void_p(out_ptr),
                                            c_void_p(stream)))

    def close(self) -> None:
        if self._h:
            lib.vg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


HANDLE_BYTES = 64


class Comm:
    """One rank of a group of GPUs that map each other's memory over NVLink (vg_comm).
    `exchange(my_handle: bytes) -> [bytes] * world` carries the handles between the processes
    (torch.distributed.all_gather_object in the tests and bench.py)."""

    def __init__(self, ctx: Context, rank: int, world: int, arena_bytes: int, exchange=None):
        h = c_void_p()
        _chk(lib.vg_comm_create(ctx._h, rank, world, arena_bytes, byref(h)))
        self._h = h
        self.ctx = ctx
        self.rank, self.world = rank, world
        if world > 1:
            mine = ctypes.create_string_buffer(HANDLE_BYTES)
            _chk(lib.vg_comm_handle(self._h, mine))
            everyone = exchange(mine.raw)
            assert len(everyone) == world and all(len(x) == HANDLE_BYTES for x in everyone)
            _chk(lib.vg_comm_connect(self._h, ctypes.create_string_buffer(b"".join(everyone), world * HANDLE_BYTES)))

    @property
    def launches(self) -> int:
        return int(lib.vg_comm_launches(self._h))

    def barrier(self) -> None:
        _chk(lib.vg_comm_barrier(self._h))

    def check(self) -> None:
        _chk(lib.vg_comm_check(self._h))

    def allreduce_counts(self, index: "Index", want_host: bool = True, dev_out: int = 0):
        """COLLECTIVE: min(255, sum over ranks) of a replicated index's counts -> u8[n] (host) or None."""
        out = np.empty(index.n, dtype=np.uint8) if want_host else None
        _chk(lib.vg_count_allreduce(self._h, index._h, _ptr(out) if want_host else None, c_void_p(dev_out)))
        return out

    def replicate(self, root: int, index: "Index" = None) -> "Index":
        """COLLECTIVE: the root passes the index it built, the others None; everybody gets a member of the replica group."""
        h = c_void_p()
        _chk(lib.vg_index_replicate(self._h, root, index._h if index is not None else None, byref(h)))
        if index is not None:
            return index
        ix = Index.__new__(Index)
        ix._h, ix.ctx, ix.comm = h, self.ctx, None
        ix.n, ix.k = int(lib.vg_index_size(h)), None
        return ix

    def allreduce_slots(self, index: "Index", want_host: bool = True):
        """COLLECTIVE: slot-order count reduce of a replica group -> (u8[slots] host or None, device pointer)."""
        out = np.empty(index.slots, dtype=np.uint8) if want_host else None
        p = c_void_p()
        _chk(lib.vg_count_allreduce_slots(self._h, index._h, _ptr(out) if want_host else None, byref(p)))
        return out, int(p.value or 0)

    def close(self) -> None:
        if self._h:
            lib.vg_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """Device twin of mGraphKmerHashHapStrMap plus the read-coverage counters (vg_index).
    With `comm` the index is sharded over the group's GPUs (vg_index_create_sharded): flush() and
    end() are then collective and a round takes at most `round_bytes` of bases (see room())."""

    def __init__(self, ctx: Context, keys, k: int, load_factor: float = 0.0, comm: "Comm" = None,
                 round_bytes: int = 0):
        h = c_void_p()
        if isinstance(keys, tuple):  # (device pointer, n): keys resident in the memory of ctx's GPU
            dev_ptr, n = keys
            if comm is None:
                _chk(lib.vg_index_create_device(ctx._h, c_void_p(dev_ptr), n, k, load_factor, byref(h)))
            else:
                _chk(lib.vg_index_create_sharded_device(comm._h, c_void_p(dev_ptr), n, k, load_factor, round_bytes, byref(h)))
            self._h, self.ctx, self.comm, self.n, self.k = h, ctx, comm, int(n), k
            return
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        if comm is None:
            _chk(lib.vg_index_create(ctx._h, _ptr(keys), keys.size, k, load_factor, byref(h)))
        else:
            _chk(lib.vg_index_create_sharded(comm._h, _ptr(keys), keys.size, k, load_factor, round_bytes, byref(h)))
        self._h = h
        self.ctx = ctx
        self.comm = comm
        self.n = int(keys.size)
        self.k = k

    @property
    def own_keys(self) -> int:
        return int(lib.vg_index_own_keys(self._h))

    def room(self) -> int:
        return int(lib.vg_count_room(self._h))

    @property
    def table_bytes(self) -> int:
        return int(lib.vg_index_table_bytes(self._h))

    @property
    def partitions(self) -> int:
        return int(lib.vg_index_partitions(self._h))

    @property
    def slices(self) -> int:
        return int(lib.vg_index_slices(self._h))

    @property
    def launches(self) -> int:
        return int(lib.vg_index_launches(self._h))

    @property
    def duplicates(self) -> int:
        return int(lib.vg_index_duplicates(self._h))

    @property
    def h2d_bytes_last(self) -> int:
        return int(lib.vg_count_h2d_bytes(self._h))

    def set_timing(self, on: bool) -> None:
        _chk(lib.vg_index_set_timing(self._h, 1 if on else 0))

    def timing(self) -> dict:
        """Per-launch device milliseconds of the two phases since set_timing(True)."""
        sc, sw, ns, nw = c_double(0), c_double(0), c_uint64(0), c_uint64(0)
        _chk(lib.vg_index_timing(self._h, byref(sc), byref(sw), byref(ns), byref(nw)))
        return {"scatter_ms_total": sc.value, "scatter_launches": int(ns.value), "sweep_ms_total": sw.value,
                "sweeps": int(nw.value),
                "scatter_ms_per_sample": sc.value / max(1, int(nw.value)) if nw.value else sc.value,
                "sweep_ms_per_sweep": sw.value / max(1, int(nw.value))}

    @property
    def fastq_blocks(self) -> int:
        return int(lib.vg_index_fastq_blocks(self._h))

    def begin(self) -> None:
        _chk(lib.vg_count_begin(self._h))

    def submit(self, bases) -> None:
        b = _as_u8(bases)
        _chk(lib.vg_count_submit(self._h, _ptr(b), b.size))

    def submit_ptr(self, host_ptr: int, nbytes: int) -> None:
        _chk(lib.vg_count_submit(self._h, c_void_p(host_ptr), nbytes))

    def submit_device(self, dev_ptr: int, nbytes: int, stream: int = 0) -> None:
        _chk(lib.vg_count_submit_device(self._h, c_void_p(dev_ptr), nbytes, c_void_p(stream)))

    def flush(self) -> None:
        _chk(lib.vg_count_flush(self._h))

    def set_flags(self, flags) -> None:
        if flags is None:
            _chk(lib.vg_index_set_flags(self._h, None))
            return
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        assert f.size == self.n
        _chk(lib.vg_index_set_flags(self._h, _ptr(f)))

    def histogram(self) -> np.ndarray:
        out = np.zeros(256, dtype=np.uint64)
        _chk(lib.vg_count_histogram(self._h, _ptr(out)))
        return out

    def count_files(self, paths, threads: int = 4) -> int:
        arr = (c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        rb = c_uint64(0)
        _chk(lib.vg_count_files(self._h, arr, len(paths), threads, byref(rb)))
        return int(rb.value)

    def end(self, want_counts: bool = True):
        """-> (counts u8[n] or None, positions, hits)"""
        out = np.empty(self.n, dtype=np.uint8) if want_counts else None
        pos, hits = c_uint64(0), c_uint64(0)
        _chk(lib.vg_count_end(self._h, _ptr(out) if want_counts else None, byref(pos), byref(hits)))
        return out, int(pos.value), int(hits.value)

    @property
    def slots(self) -> int:
        """Length of the slot-order count vector (vg_index_slots)."""
        return int(lib.vg_index_slots(self._h))

    def slot_perm(self) -> np.ndarray:
        """perm[i] = position of keys[i] in the slot-order count vector (0xffffffff: never counted)."""
        out = np.empty(self.n, dtype=np.uint32)
        _chk(lib.vg_index_slot_perm(self._h, _ptr(out)))
        return out

    def end_slots(self, want_counts: bool = True):
        """-> (counts u8[slots] in slot order or None, positions, hits)"""
        out = np.empty(self.slots, dtype=np.uint8) if want_counts else None
        pos, hits = c_uint64(0), c_uint64(0)
        _chk(lib.vg_count_end_slots(self._h, _ptr(out) if want_counts else None, byref(pos), byref(hits)))
        return out, int(pos.value), int(hits.value)

    def slots_device(self) -> int:
        """Device pointer of the slot-order count vector after a flush (asynchronous on the context stream)."""
        p = c_void_p()
        _chk(lib.vg_count_slots_device(self._h, byref(p)))
        return int(p.value or 0)

    def stats(self):
        pos, hits = c_uint64(0), c_uint64(0)
        _chk(lib.vg_count_stats(self._h, byref(pos), byref(hits)))
        return int(pos.value), int(hits.value)

    @property
    def keys_scattered(self) -> int:
        """K-mers that passed the pre-filter, as of the last stats() / end() (vg_count_keys)."""
        return int(lib.vg_count_keys(self._h))

    def extract_device(self, dev_ptr: int, elem_bytes: int, stream: int = 0) -> None:
        _chk(lib.vg_count_extract_device(self._h, c_void_p(dev_ptr), elem_bytes, c_void_p(stream)))

    def close(self) -> None:
        if self._h:
            lib.vg_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CountingBloom:
    """Device twin of BloomFilter / BloomFilterKernel (vg_cbf)."""

    def __init__(self, ctx: Context, m: int, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        h = c_void_p()
        _chk(lib.vg_cbf_create(ctx._h, m, seeds.size, _ptr(seeds), byref(h)))
        self._h = h
        self.ctx = ctx
        self.m = int(m)

    def add_sequence(self, seq, k: int) -> int:
        b = _as_u8(seq)
        added = c_uint64(0)
        _chk(lib.vg_cbf_add_sequence(self._h, _ptr(b), b.size, k, byref(added)))
        return int(added.value)

    def download(self) -> np.ndarray:
        out = np.empty(self.m, dtype=np.uint8)
        _chk(lib.vg_cbf_download(self._h, _ptr(out)))
        return out

    def query(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        cnt = np.empty(keys.size, dtype=np.uint8)
        fnd = np.empty(keys.size, dtype=np.uint8)
        _chk(lib.vg_cbf_query(self._h, _ptr(keys), keys.size, _ptr(cnt), _ptr(fnd)))
        return cnt, fnd

    def close(self) -> None:
        if self._h:
            lib.vg_cbf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FastqKmerKernel:
    """Python mirror of the reference's FastqKmerKernel (include/fastq_kmer.cuh:10-49): same
    constructor arguments and method name, `mReadBase` as the public result; the map is stood in
    for by its key array and `c` comes back as a u8 vector in the same order."""

    def __init__(self, index: Index, fastqFileNameVec, kmerLen: int, threads: int, buffer: int = 100):
        if kmerLen != index.k:
            raise ValueError("kmerLen differs from the index's k")
        self.index = index
        self.fastqFileNameVec_ = list(fastqFileNameVec)
        self.threads_ = threads
        self.buffer_ = buffer
        self.mReadBase = 0
        self.c = None

    def build_fastq_index_kernel(self) -> None:
        if not self.fastqFileNameVec_:
            raise VgError(VG_E_INVALID, "Parameter error: -f")
        self.index.begin()
        self.mReadBase += self.index.count_files(self.fastqFileNameVec_, self.threads_)
        self.c, self.positions, self.hits = self.index.end()
