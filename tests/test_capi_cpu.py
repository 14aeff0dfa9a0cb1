"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/vgb200.h declares, and
refuses to compute without a GPU (no CPU fallback exists)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(vglib):
    syms = declared_symbols()
    assert len(syms) >= 24
    lib = ctypes.CDLL(vglib.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header(vglib):
    for s in declared_symbols():
        assert getattr(vglib.lib, s).argtypes is not None, s


def test_version_and_error_string(vglib):
    assert vglib.lib.vg_version() >= 100
    assert isinstance(vglib.lib.vg_last_error(), bytes)


def test_sass_is_sm100a_only(vglib):
    """The shipped binary carries sm_100a code and nothing else (no multi-arch dispatch)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", vglib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(vglib):
    with pytest.raises(vglib.VgError) as ei:
        vglib.Context(0)
    assert ei.value.code == vglib.VG_E_CUDA
    assert "no CPU path" in str(ei.value)


def test_product_never_touches_oracle():
    """Nothing under varigraph_b200/ may reference oracle/ (the oracle is a checker, not a fallback)."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "varigraph_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"vg_oracle|libvgoracle|libvgref|oracle_binding|vgo_", txt):
                    bad.append(f)
    assert not bad, bad


def test_fastq_record_boundary_heuristic(vglib, tmp_path):
    """Host half of the raw FASTQ road: block cuts land on record starts even when quality lines begin
    with '@' or '+', with CRLF line ends, and give up (-1) where no four-line record starts in the window."""
    recs = []
    for i in range(200):
        seq = b"ACGT" * (5 + i % 7)
        qual = (b"@" if i % 3 == 0 else b"+" if i % 3 == 1 else b"I") + b"#" * (len(seq) - 1)
        eol = b"\r\n" if i % 2 else b"\n"
        recs.append(b"@r%d x" % i + eol + seq + eol + b"+" + eol + qual + eol)
    data = b"".join(recs)
    p = tmp_path / "b.fq"
    p.write_bytes(data)
    starts = set(np.cumsum([0] + [len(r) for r in recs]).tolist())
    f = vglib.lib.vg_fastq_record_boundary
    assert f(str(p).encode(), 0, 4096) == 0
    assert f(str(p).encode(), len(data) + 5, 4096) == len(data)
    for at in range(1, len(data) - 400, 37):
        b = f(str(p).encode(), at, 4096)
        assert b in starts and b >= at and not any(at <= s < b for s in starts), (at, b)
    # a window too short to see a whole record, and a file that is not FASTQ at all
    assert f(str(p).encode(), 10, 8) == -1
    q = tmp_path / "x.txt"
    q.write_bytes(b"@not fastq\n" * 50)
    assert f(str(q).encode(), 5, 4096) == -1
    assert f(str(tmp_path / "missing").encode(), 5, 4096) == -2


def _strip(vglib, data: bytes, last=True):
    import ctypes
    out = np.empty(len(data) // 2 + 512, dtype=np.uint8)
    bases, bad = ctypes.c_uint64(0), ctypes.c_int64(0)
    w = vglib.lib.vg_fastq_strip_block(data, len(data), 1 if last else 0, out.ctypes.data, ctypes.byref(bases), ctypes.byref(bad))
    return out[:w].tobytes(), int(bases.value), int(bad.value)


def _no_empty(lines: bytes) -> bytes:
    """The staged form without empty reads (they emit no k-mer; the feeder does not stage them)."""
    return b"".join(x + b"\n" for x in lines.split(b"\n") if x)


def test_fastq_strip_block_vs_kseq_restatement(vglib, oracle):
    """Host half of the strip road: the vectorised four-line scanner accepts exactly the records kseq reads as four-line
    records.  On randomly damaged FASTQ, (what it wrote) + (kseq on the rest, from the record it stopped at) must equal
    kseq on the whole text -- sequences and seq.l sum -- and a clean text must pass untouched."""
    import random
    rng = random.Random(2026)
    recs = []
    for i in range(3000):
        n = rng.choice([0, 1, 1, 30, 75, 100, 150, 151, 191, 192, 193, 250, 400]) if i % 9 == 0 else 150
        seq = bytes(rng.choice(b"ACGTNacgt") for _ in range(n))
        qual = bytes(rng.choice(b"@+>IJ#") for _ in range(n))
        eol = b"\r\n" if i % 7 == 0 else b"\n"
        recs.append(b"@r%d some text" % i + eol + seq + eol + b"+" + (b"r%d" % i if i % 5 == 0 else b"") + eol + qual + eol)
    clean = b"".join(recs)
    lines, nreads, bases, status = oracle.fastq_to_lines(clean)
    lines = _no_empty(lines)
    w, b, bad = _strip(vglib, clean)
    assert bad == -1 and b == bases and w == lines
    w, b, bad = _strip(vglib, clean[:-1])  # no final newline at EOF
    assert bad == -1 and b == bases and w == lines
    w, b, bad = _strip(vglib, clean[:-1], last=False)  # ... which a block in the middle of a file may not have
    assert bad >= 0 and clean[bad - 1: bad] == b"\n"
    for trial in range(300):
        data = bytearray(clean[: rng.randrange(2000, len(clean))] if trial % 4 == 0 else clean)
        for _ in range(rng.randint(1, 3)):
            at = rng.randrange(len(data))
            kind = rng.randrange(7)
            if kind == 0:
                nl = data.find(b"\n", at)
                if nl >= 0:
                    del data[nl]
            elif kind == 1:
                data.insert(at, 10)
            elif kind == 2:
                data.insert(at, rng.choice(b"@+>\r\x00 "))
            elif kind == 3:
                data[at] = rng.choice(b"@+>\rN\x00")
            elif kind == 4:
                del data[max(at, len(data) // 2):]
            elif kind == 5:
                nl = data.find(b"\n", at)
                if nl >= 0:
                    data.insert(nl, 13)
            else:
                data[at:at] = b">fasta\nACGTACGT\n"
        data = bytes(data)
        want_lines, _, want_bases, _ = oracle.fastq_to_lines(data)
        want_lines = _no_empty(want_lines)
        w, b, bad = _strip(vglib, data)
        if bad < 0:
            assert (w, b) == (want_lines, want_bases), trial
            continue
        assert bad == 0 or data[bad - 1: bad] == b"\n", trial
        rest_lines, _, rest_bases, _ = oracle.fastq_to_lines(data[bad:])
        assert w + _no_empty(rest_lines) == want_lines and b + rest_bases == want_bases, (trial, bad)


# ---- the parallel inflater behind the gzip road (vg_gzip.cpp): bytes identical to zlib's ------------------------
def _gunzip(vglib, path, threads, chunk):
    import ctypes
    p, n = ctypes.c_void_p(), ctypes.c_uint64()
    rc = vglib.lib.vg_gunzip_parallel(str(path).encode(), threads, chunk, ctypes.byref(p), ctypes.byref(n))
    if rc != 0:
        return rc, None
    data = ctypes.string_at(p, n.value)
    vglib.lib.vg_gunzip_free(p)
    return 0, data


def _fastq_text(n, seed, L=150):
    import random
    r = random.Random(seed)
    return "".join("@read%d/1 lane:%d\n%s\n+\n%s\n" % (i, i % 7, "".join(r.choice("ACGTN") for _ in range(L)),
                                                          "".join(r.choice("FFFFFFFF:,#") for _ in range(L))) for i in range(n)).encode()


def _bgzf(data: bytes) -> bytes:
    """What bgzip writes: gzip members of <= 64 KiB of text each with a BC extra field, then the empty EOF member."""
    import zlib
    out = []
    for i in range(0, len(data), 65280):
        blk = data[i:i + 65280]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        cd = co.compress(blk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + (len(cd) + 25).to_bytes(2, "little") + cd +
                   zlib.crc32(blk).to_bytes(4, "little") + len(blk).to_bytes(4, "little"))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def test_parallel_gunzip_matches_zlib(vglib, tmp_path):
    """Every shape of gzip the feeder can meet, cut into chunks small enough that block starts are searched and windows
    are resolved many times over: single member at three levels, concatenated members, bgzip blocks, stored blocks
    (level 0), members so small they use the fixed code, sync-flushed streams (empty stored blocks), binary content (no
    block start is ever found: the first worker inflates everything), an empty member, trailing garbage, a flipped bit."""
    import gzip
    import random
    import zlib
    txt = _fastq_text(6000, 1)
    cases = []
    for lvl in (1, 6, 9):
        cases.append(("level%d" % lvl, txt, gzip.compress(txt, lvl)))
    parts = [_fastq_text(300, 10 + i) for i in range(25)]
    cases.append(("members", b"".join(parts), b"".join(gzip.compress(p, 6) for p in parts)))
    cases.append(("bgzf", txt, _bgzf(txt)))
    cases.append(("stored", txt[:300000], gzip.compress(txt[:300000], 0)))
    tiny = [b"@r\nACGT\n+\nFFFF\n"] * 200
    cases.append(("fixed", b"".join(tiny), b"".join(gzip.compress(p, 6) for p in tiny)))
    co, sf = zlib.compressobj(6, zlib.DEFLATED, 31), b""
    for i in range(0, len(txt), 70000):
        sf += co.compress(txt[i:i + 70000]) + co.flush(zlib.Z_SYNC_FLUSH)
    cases.append(("syncflush", txt, sf + co.flush()))
    r = random.Random(5)
    binary = bytes(r.getrandbits(8) for _ in range(150000)) + txt[:100000] + bytes(r.getrandbits(3) for _ in range(100000))
    cases.append(("binary", binary, gzip.compress(binary, 6)))
    cases.append(("empty", b"", gzip.compress(b"", 6)))
    cases.append(("trailing", txt[:100000], gzip.compress(txt[:100000], 6) + b"\0" * 100))
    for name, raw, gz in cases:
        if name != "trailing":
            assert gzip.decompress(gz) == raw
        path = tmp_path / (name + ".gz")
        path.write_bytes(gz)
        for threads, chunk in ((1, 1 << 20), (4, 4096), (8, 20000), (3, 1 << 16)):
            rc, got = _gunzip(vglib, path, threads, chunk)
            assert rc == 0 and got == raw, (name, threads, chunk, rc)
    bad = bytearray(gzip.compress(txt, 6))
    bad[len(bad) // 2] ^= 0x55
    (tmp_path / "bad.gz").write_bytes(bad)
    assert _gunzip(vglib, tmp_path / "bad.gz", 4, 1 << 16)[0] == -2
    for cut in (len(bad) // 3, len(bad) - 9, 30):  # truncated: must stop, not inflate the zeros behind the end for ever
        (tmp_path / "cut.gz").write_bytes(gzip.compress(txt, 6)[:cut])
        assert _gunzip(vglib, tmp_path / "cut.gz", 4, 1 << 16)[0] == -2
    (tmp_path / "plain.txt").write_bytes(txt[:1000])
    assert _gunzip(vglib, tmp_path / "plain.txt", 4, 1 << 16)[0] == -2
    assert _gunzip(vglib, tmp_path / "missing.gz", 4, 1 << 16)[0] == -1


def test_crc32_matches_zlib(vglib):
    import random
    import zlib
    r = random.Random(9)
    for n in list(range(0, 200)) + [1000, 4096, 65537, 1_000_003]:
        b = r.randbytes(n)
        for init in (0, 0x12345678):
            assert vglib.lib.vg_crc32(init, b, n) == zlib.crc32(b, init), (n, init)


def test_parallel_gunzip_fuzz(vglib, tmp_path):
    """Random payloads (text-like, low-entropy, binary, mixtures), every zlib strategy (default, filtered, Huffman only,
    RLE, fixed codes -- the last three leave no or few dynamic blocks to search for), levels 0-9, small windows, one or
    several members, sync flushes at random places; chunk sizes down to 1 KiB.  The bytes must be zlib's."""
    import random
    import zlib
    rng = random.Random(20261017)
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]

    def payload(n):
        kind = rng.random()
        if kind < 0.4:
            return _fastq_text(max(1, n // 330), rng.randrange(1 << 30))[:n]
        if kind < 0.55:
            return bytes(rng.choice(b"ACGT\n") for _ in range(n))
        if kind < 0.7:
            return bytes([rng.choice(b"AN")]) * n
        if kind < 0.85:
            return rng.randbytes(n)
        return b"".join(rng.choice([b"@read\n", b"ACGTACGT", b"+\n", b"FFFF:FFF\n", rng.randbytes(5)]) for _ in range(n // 6))

    for trial in range(60):
        raw_parts, gz = [], b""
        for _ in range(rng.choice([1, 1, 2, 5])):
            data = payload(rng.choice([0, 1, 200, 5000, 70_000, 300_000]))
            co = zlib.compressobj(rng.randrange(0, 10), zlib.DEFLATED, 16 + rng.choice([9, 12, 15]), rng.randrange(1, 10),
                                  rng.choice(strategies))
            cut = sorted(rng.randrange(0, len(data) + 1) for _ in range(rng.choice([0, 0, 1, 3])))
            prev = 0
            for c in cut:
                gz += co.compress(data[prev:c]) + co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))
                prev = c
            gz += co.compress(data[prev:]) + co.flush()
            raw_parts.append(data)
        raw = b"".join(raw_parts)
        path = tmp_path / ("fuzz%d.gz" % trial)
        path.write_bytes(gz)
        for threads, chunk in ((rng.choice([1, 2, 5, 8]), rng.choice([1024, 4096, 30_000, 1 << 20])) for _ in range(2)):
            rc, got = _gunzip(vglib, path, threads, chunk)
            assert rc == 0 and got == raw, (trial, threads, chunk, rc, len(raw))
