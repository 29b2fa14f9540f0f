"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports it. Builds oracle/liboracle*.so with `make -C oracle` when missing.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORC_BVH, ORC_SLOT_EXACT, ORC_NO_MIS, ORC_GGX = 1, 2, 4, 8

_f = C.c_float; _u = C.c_uint32; _i = C.c_int32; _p = C.c_void_p
_fp = C.POINTER(C.c_float); _up = C.POINTER(C.c_uint32)


def build(force=False):
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so", "liboracle_fast.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not all(os.path.exists(os.path.join(_HERE, "_ref", n)) for n in ("librefrng.so", "librefsampling.so", "librefbvh.so", "librefrenderer.so", "librefrenderer_ggx.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _load(fast=False):
    build()
    lib = C.CDLL(os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so"))
    lib.orc_create.restype = _p; lib.orc_create.argtypes = [_u] * 5
    lib.orc_destroy.argtypes = [_p]
    lib.orc_set_scene.restype = C.c_int; lib.orc_set_scene.argtypes = [_p, _p, _u, _p, _u, _fp, _fp, _i, _i]
    lib.orc_set_camera_lookat.argtypes = [_p, _fp, _fp, _f, _f]
    lib.orc_set_camera_raw.argtypes = [_p, _fp, _fp, _f, _f, _f, _f]
    lib.orc_get_camera_raw.argtypes = [_p, _fp]
    lib.orc_generate_ray_at.argtypes = [_p, _i, _i, _fp, _fp]
    lib.orc_reset.argtypes = [_p]
    lib.orc_set_accumulations.argtypes = [_p, _u]
    lib.orc_get_accumulations.restype = _u; lib.orc_get_accumulations.argtypes = [_p]
    lib.orc_accumulate.argtypes = [_p, _u, C.c_int]
    lib.orc_accumulate_tiles.argtypes = [_p, _up, _u, _u, C.c_int]
    lib.orc_read_buckets.argtypes = [_p, _fp]
    lib.orc_render.restype = C.c_int; lib.orc_render.argtypes = [_p, _fp, C.c_int]
    lib.orc_read_counters.argtypes = [_p, C.POINTER(C.c_uint64)]
    lib.orc_reset_counters.argtypes = [_p]
    lib.orc_bvh_node_count.restype = _u; lib.orc_bvh_node_count.argtypes = [_p]
    lib.orc_bvh_read.argtypes = [_p, _p, _p, _p]
    lib.orc_light_count.restype = _u; lib.orc_light_count.argtypes = [_p]
    lib.orc_read_lights.argtypes = [_p, _p]
    lib.orc_generate_rays.argtypes = [_p, _u, _fp]
    lib.orc_trace_closest.argtypes = [_p, _fp, _u, _fp, _p]
    lib.orc_trace_shadow.argtypes = [_p, _fp, _fp, _u, _p]
    lib.orc_trace_closest_scalar.argtypes = [_p, _fp, _u, _fp, _p]
    lib.orc_build_bvh.restype = _u; lib.orc_build_bvh.argtypes = [_p, _u, _p, _p, _p]
    for name, res, args in [
        ("orc_hash_u32", _u, [_u]), ("orc_hash_2d", _u, [_u, _u]), ("orc_pcg_generate", _u, [_up]),
        ("orc_rand_unit_float", _f, [_up]), ("orc_rand_bounded_int", _u, [_up, _u]), ("orc_make_unit_float", _f, [_u]),
        ("orc_bitreverse", _u, [_u]), ("orc_fast_sincos", None, [_f, _fp, _fp]), ("orc_fast_asin", _f, [_f]),
        ("orc_fast_atan2", _f, [_f, _f]), ("orc_median5", _f, [_fp]), ("orc_median_k", _f, [_fp, _u]),
        ("orc_hemisphere", None, [_f, _f, _fp]), ("orc_tangent_space", None, [_fp, _fp]), ("orc_to_local", None, [_fp, _fp, _fp]),
        ("orc_to_world", None, [_fp, _fp, _fp]), ("orc_orthonormal_basis", None, [_fp, _fp]),
        ("orc_sample_direction_to_sphere", None, [_fp, _f, _f, _f, _f, _f, _fp]), ("orc_sphere_pdf", _f, [_f, _f]),
        ("orc_power_heuristic", _f, [_f, _f]), ("orc_power_heuristic_over_f", _f, [_f, _f]), ("orc_tonemap", None, [_fp]),
        ("orc_quat_look_at", None, [_fp, _fp]),
    ]:
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        _libs[fast] = _load(fast)
    return _libs[fast]


def _fptr(a):
    return a.ctypes.data_as(_fp)


def farr(*v):
    return (C.c_float * len(v))(*v)


class Oracle:
    """One CPU Renderer (Renderer.hpp:28-479 restated) bound to a scene."""

    def __init__(self, width, height, max_bounces=16, K=5, flags=0, fast=False):
        self.L = lib(fast)
        self.h = self.L.orc_create(width, height, max_bounces, K, flags)
        if not self.h:
            raise ValueError("orc_create: width/height must be multiples of 16, 1<=K<=64")
        self.width, self.height, self.K, self.max_bounces = width, height, K, max_bounces
        self.npix = width * height

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h); self.h = None

    __del__ = close

    def set_scene(self, scene):
        geo = np.ascontiguousarray(scene["geometry"]); mat = np.ascontiguousarray(scene["material"])
        amb = farr(*scene["ambient"])
        hdri = scene.get("hdri")
        if hdri is not None:
            hdri = np.ascontiguousarray(hdri, np.float32)
            n = self.L.orc_set_scene(self.h, geo.ctypes.data, len(geo), mat.ctypes.data, len(mat), amb, _fptr(hdri), hdri.shape[1], hdri.shape[0])
        else:
            n = self.L.orc_set_scene(self.h, geo.ctypes.data, len(geo), mat.ctypes.data, len(mat), amb, None, 0, 0)
        cam = scene["camera"]
        self.L.orc_set_camera_lookat(self.h, farr(*cam["eye"]), farr(*cam["dir"]), cam["focal_length"], cam["exposure"])
        self.n_geom = len(geo)
        return n

    def camera_raw(self):
        out = (C.c_float * 11)(); self.L.orc_get_camera_raw(self.h, out); return np.array(out[:], np.float32)

    def reset(self):
        self.L.orc_reset(self.h)

    def set_accumulations(self, acc):
        self.L.orc_set_accumulations(self.h, acc)

    def accumulate(self, n=1, threads=None):
        self.L.orc_accumulate(self.h, n, threads or os.cpu_count() or 1)

    def accumulate_tiles(self, tiles, n=1, threads=None):
        t = np.ascontiguousarray(tiles, np.uint32)
        self.L.orc_accumulate_tiles(self.h, t.ctypes.data_as(_up), len(t), n, threads or os.cpu_count() or 1)

    def buckets(self):
        out = np.empty((self.K, 3, self.npix), np.float32); self.L.orc_read_buckets(self.h, _fptr(out)); return out

    def render(self, tonemap=True):
        out = np.zeros((self.height, self.width, 4), np.float32)
        rc = self.L.orc_render(self.h, _fptr(out), 1 if tonemap else 0)
        return rc, out

    def counters(self):
        out = (C.c_uint64 * 7)(); self.L.orc_read_counters(self.h, out)
        return dict(zip(["extension_rays", "shadow_rays", "shaded_hits", "terminated", "sphere_tests", "box_tests", "dropped"], out[:]))

    def reset_counters(self):
        self.L.orc_reset_counters(self.h)

    def bvh(self):
        from_dt = _scene_dtypes()
        n = self.L.orc_bvh_node_count(self.h)
        nodes = np.zeros(n, from_dt[2]); prims = np.zeros(self.n_geom, from_dt[0]); ids = np.zeros(self.n_geom, np.uint32)
        self.L.orc_bvh_read(self.h, nodes.ctypes.data, prims.ctypes.data, ids.ctypes.data)
        return nodes, prims, ids

    def lights(self):
        n = self.L.orc_light_count(self.h); out = np.zeros(n, np.int32); self.L.orc_read_lights(self.h, out.ctypes.data); return out

    def generate_rays(self, acc):
        out = np.empty((self.npix, 6), np.float32); self.L.orc_generate_rays(self.h, acc, _fptr(out)); return out

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32); n = len(rays)
        t = np.empty(n, np.float32); p = np.empty(n, np.int32)
        self.L.orc_trace_closest(self.h, _fptr(rays), n, _fptr(t), p.ctypes.data); return t, p

    def trace_shadow(self, rays, tfar):
        rays = np.ascontiguousarray(rays, np.float32); tfar = np.ascontiguousarray(tfar, np.float32); n = len(rays)
        o = np.empty(n, np.uint8); self.L.orc_trace_shadow(self.h, _fptr(rays), _fptr(tfar), n, o.ctypes.data); return o


def _scene_dtypes():
    import importlib.util
    p = os.path.join(os.path.dirname(_HERE), "cpu-raytracing-experiments_b200", "scenes.py")
    spec = importlib.util.spec_from_file_location("b2r_scenes", p); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m.SPHERE_DTYPE, m.MATERIAL_DTYPE, m.NODE_DTYPE


def tile_to_raster(buf_tileorder, width, height):
    """[..., npix] in tile order (t = tile*256 + ID) -> [..., height, width] raster."""
    ht, vt = width // 16, height // 16
    a = np.asarray(buf_tileorder).reshape(buf_tileorder.shape[:-1] + (vt, ht, 16, 16))
    a = np.moveaxis(a, -2, -3)  # vt,16,ht,16
    return a.reshape(buf_tileorder.shape[:-1] + (height, width))


# ---------------------------------------------------------------------------------------------- the reference itself
REF_RENDERER_PATH = os.path.join(_HERE, "_ref", "librefrenderer.so")
REF_RENDERER_GGX_PATH = os.path.join(_HERE, "_ref", "librefrenderer_ggx.so")  # the same sources with `#define BRDF 1` (ref_renderer_build.sh)
REF_MAX_BOUNCES = (1, 2, 4, 8, 16)  # Renderer<>'s max_bounces is a template argument: the values instantiated by ref_renderer_wrap.cpp
ORC_SLOT_EXACT = 2  # oracle flag: closest-hit SIMD blocks of 8 + scalar tail by stream slot, exactly as BVH.hpp:250-286


def have_reference_renderer(ggx=False):
    return os.path.exists(REF_RENDERER_GGX_PATH if ggx else REF_RENDERER_PATH)


_ref_renderer = {}


def _ref_renderer_lib(ggx=False):
    if ggx not in _ref_renderer:
        L = C.CDLL(REF_RENDERER_GGX_PATH if ggx else REF_RENDERER_PATH)
        L.ref_renderer_create.restype = _p
        L.ref_renderer_create.argtypes = [_p, _u, _p, _u, _p, _p, _f, _f, _p, _p, _i, _i, _u, _u, _u]
        L.ref_renderer_destroy.argtypes = [_p]; L.ref_renderer_accumulate.argtypes = [_p, _u]; L.ref_renderer_set_accumulations.argtypes = [_p, _u]
        L.ref_renderer_accumulations.restype = _u; L.ref_renderer_accumulations.argtypes = [_p]
        L.ref_renderer_read_buckets.argtypes = [_p, _p]; L.ref_renderer_render.restype = C.c_int; L.ref_renderer_render.argtypes = [_p, _p, _u]
        L.ref_renderer_light_count.restype = _u; L.ref_renderer_light_count.argtypes = [_p]
        _ref_renderer[ggx] = L
    return _ref_renderer[ggx]


class ReferenceRenderer:
    """The reference's OWN Renderer<> (Renderer.hpp) compiled from /root/reference by oracle/ref_renderer_build.sh — same calls as
    Oracle: accumulate(n), buckets() -> [5][3][npix] (tile order), render() -> (acted, RGBA32F raster frame). K is fixed at 5."""

    def __init__(self, scene, width, height, max_bounces=16, ggx=False):
        if max_bounces not in REF_MAX_BOUNCES:
            raise ValueError(f"max_bounces must be one of {REF_MAX_BOUNCES} (template instantiations)")
        self.L = _ref_renderer_lib(ggx); self.w, self.h = width, height
        sd = _scene_dtypes()
        self._geo = np.ascontiguousarray(scene["geometry"], dtype=sd[0]); self._mat = np.ascontiguousarray(scene["material"], dtype=sd[1])
        cam = scene["camera"]; eye = farr(*cam["eye"]); d = farr(*cam["dir"]); amb = farr(*scene["ambient"])
        hd = scene.get("hdri"); hp, hw, hh = None, 0, 0
        if hd is not None:
            self._hdri = np.ascontiguousarray(hd, dtype=np.float32); hp = self._hdri.ctypes.data; hh, hw = self._hdri.shape[:2]
        self.hnd = self.L.ref_renderer_create(self._geo.ctypes.data, len(self._geo), self._mat.ctypes.data, len(self._mat), C.cast(eye, _p), C.cast(d, _p),
                                              cam["focal_length"], cam["exposure"], C.cast(amb, _p), hp, hw, hh, width, height, max_bounces)
        if not self.hnd:
            raise RuntimeError("ref_renderer_create failed")

    def close(self):
        if self.hnd:
            self.L.ref_renderer_destroy(self.hnd); self.hnd = None

    def accumulate(self, n=1):
        self.L.ref_renderer_accumulate(self.hnd, n)

    def set_accumulations(self, acc):
        self.L.ref_renderer_set_accumulations(self.hnd, acc)

    def buckets(self):
        out = np.zeros((5, 3, self.w * self.h), np.float32); self.L.ref_renderer_read_buckets(self.hnd, out.ctypes.data); return out

    def render(self):
        fb = np.zeros((self.h, self.w, 4), np.float32)
        return bool(self.L.ref_renderer_render(self.hnd, fb.ctypes.data, fb.size)), fb
