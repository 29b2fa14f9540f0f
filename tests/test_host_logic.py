"""Host-side product code (libb2r.so, no GPU needed): BVH builder == oracle bit-for-bit, 128-byte flattening round-trips,
light list, camera set-up, argument checking."""
import ctypes as C

import numpy as np
import pytest

import b2r
import oracle_py
import scenes


@pytest.mark.parametrize("n", [1, 2, 9, 100, 5000])
def test_bvh_bit_exact_vs_oracle(n):
    sc = scenes.default_scene() if n == 9 else scenes.random_scene(n, light_every=10)
    nodes, prims, ids = b2r.build_bvh(sc["geometry"])
    o = oracle_py.Oracle(16, 16); o.set_scene(sc)
    on, op, oi = o.bvh()
    assert nodes.tobytes() == on.tobytes()      # node order + boxes, bit-exact (north_star parity gate)
    assert prims.tobytes() == op.tobytes()      # leaf order
    assert np.array_equal(ids, oi)


def test_bvh_ties_are_stable():
    """Equal centroids (default scene: spheres 0/4 share x and z) — ties keep the lower index (Q19 canonical rule)."""
    g = np.zeros(8, scenes.SPHERE_DTYPE); g["radius_sq"] = 1.0; g["position"][:, 1] = np.arange(8) % 2
    nodes, prims, ids = b2r.build_bvh(g)
    o = oracle_py.Oracle(16, 16)
    sc = scenes.Scene(geometry=g, material=scenes.default_scene()["material"], camera=scenes.default_scene()["camera"], ambient=(0, 0, 0), hdri=None)
    o.set_scene(sc)
    assert nodes.tobytes() == o.bvh()[0].tobytes() and np.array_equal(ids, o.bvh()[2])


def test_lights_and_camera_match_oracle():
    for sc in (scenes.default_scene(), scenes.random_scene(3000)):
        o = oracle_py.Oracle(1280, 720); o.set_scene(sc)
        assert np.array_equal(b2r.find_lights(sc["geometry"], sc["material"]), o.lights())
        cam = b2r.camera_lookat(sc["camera"]["eye"], sc["camera"]["dir"], 1280, 720, sc["camera"]["focal_length"], 1.0)
        assert cam.tobytes() == o.camera_raw().tobytes()
    rs = np.random.RandomState(5)
    L = oracle_py.lib()
    for d in rs.randn(200, 3).astype(np.float32):  # every quat_cast branch
        q = np.zeros(4, np.float32); L.orc_quat_look_at(d.ctypes.data_as(C.POINTER(C.c_float)), q.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(b2r.camera_lookat((0, 0, 0), d, 16, 16, 50.0)[3:7], q)


def test_abi_argument_errors():
    L = b2r.lib()
    assert L.b2r_bvh_build(None, 3, None, None, None, None) == b2r.ERR_ARG
    h = C.c_void_p(None)
    for w, hgt, k in [(100, 64, 5), (64, 0, 5), (64, 64, 0), (64, 64, 65)]:
        cfg = b2r.Config(w, hgt, 16, k, 0, 0, 0, 0, 0)
        assert L.b2r_create(C.byref(h), C.byref(cfg)) == b2r.ERR_ARG and not h.value
    assert L.b2r_accumulate(None, 1) == b2r.ERR_ARG and b"null" in L.b2r_last_error()
