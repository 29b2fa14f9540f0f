import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cpu-raytracing-experiments_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once per session (no-op when up to date; the GPU box reuses the prebuilt files)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def hostcheck():
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    f, u = C.c_float, C.c_uint32
    lib.hc_hash_2d.restype = u; lib.hc_hash_2d.argtypes = [u, u]
    lib.hc_hash_u32.restype = u; lib.hc_hash_u32.argtypes = [u]
    lib.hc_sincos.argtypes = [f, C.POINTER(f), C.POINTER(f)]
    lib.hc_asin.restype = f; lib.hc_asin.argtypes = [f]
    lib.hc_atan2.restype = f; lib.hc_atan2.argtypes = [f, f]
    lib.hc_hemisphere.argtypes = [f, f, C.c_void_p]
    lib.hc_sample_sphere.argtypes = [C.c_void_p, f, f, f, f, f, C.c_void_p]
    lib.hc_sphere_pdf.restype = f; lib.hc_sphere_pdf.argtypes = [f, f]
    lib.hc_power.restype = f; lib.hc_power.argtypes = [f, f]
    lib.hc_power_over_f.restype = f; lib.hc_power_over_f.argtypes = [f, f]
    lib.hc_median5.restype = f; lib.hc_median5.argtypes = [C.c_void_p]
    lib.hc_median8.restype = f; lib.hc_median8.argtypes = [C.c_void_p]
    lib.hc_sphere_closest.restype = C.c_int; lib.hc_sphere_closest.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(f)]
    lib.hc_closest_scalar.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(f), C.POINTER(C.c_int32)]
    lib.hc_sphere_any.restype = C.c_int; lib.hc_sphere_any.argtypes = [C.c_void_p, C.c_void_p, f]
    lib.hc_pcg3.argtypes = [u, C.c_void_p, C.POINTER(u), u, C.POINTER(u)]
    return lib


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
