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
