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
