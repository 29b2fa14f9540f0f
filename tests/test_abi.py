"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2r.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

import b2r
from conftest import ROOT, have_gpu


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2r.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(b2r.ABI_SYMBOLS)


def test_flag_values_agree_with_the_header():
    """B2R_FLAG_* of include/b2r.h against the Python binding's FLAG_* constants (the C++ faces include the header itself)."""
    src = open(os.path.join(ROOT, "include", "b2r.h")).read()
    flags = {name: 1 << int(bit) for name, bit in re.findall(r"\bB2R_FLAG_([A-Z0-9_]+)\s*=\s*1u\s*<<\s*(\d+)", src)}
    assert len(flags) >= 12 and len(set(flags.values())) == len(flags)          # every flag its own bit
    for name, value in flags.items():
        assert getattr(b2r, "FLAG_" + name) == value, name


def test_library_exports_every_symbol():
    L = C.CDLL(b2r.LIB_PATH)
    for s in header_symbols():
        assert hasattr(L, s), s
    assert b2r.lib().b2r_abi_version() == 1


def test_struct_layouts_match_reference_pods():
    import scenes
    assert scenes.SPHERE_DTYPE.itemsize == 32 and scenes.MATERIAL_DTYPE.itemsize == 96 and scenes.NODE_DTYPE.itemsize == 32
    assert C.sizeof(b2r.Config) == 36


@pytest.mark.skipif(have_gpu(), reason="CPU-box behaviour")
def test_no_cpu_fallback():
    """Without a CUDA device the renderer refuses to exist: the product has no CPU path."""
    import scenes
    with pytest.raises(b2r.B2RError) as e:
        b2r.Renderer(scenes.default_scene(), 64, 64)
    assert e.value.code == b2r.ERR_CUDA


def test_reference_binding_builds_and_fails_loudly_without_gpu():
    """include/b2r_reference_binding.hpp compiles against the reference's own headers (tests/refbinding, built where /root/reference
    exists). On a CPU-only box the drop-in must fail loudly — no CPU fallback — while the reference's own Renderer<> in the same
    program is untouched."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "refbinding", "refbinding")
    if not os.path.exists(exe):
        pytest.skip("tests/refbinding/refbinding not present (built only where /root/reference exists)")
    if have_gpu():
        pytest.skip("GPU box: run by tests/test_gpu_parity.py")
    p = subprocess.run([exe, "5"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "libb2r" in (p.stderr + p.stdout)
