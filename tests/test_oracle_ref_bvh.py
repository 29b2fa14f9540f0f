"""BVH node and leaf order (a north_star parity gate) and the three sphere-test formulas, against the reference's OWN source:
BVH.hpp:17-87,91-206 (Node, SplitHeuristic, the constructor) and :239-287 / :292-304 (bodies of intersect_prims and
intersect_prims_shadow) compiled verbatim into oracle/_ref/librefbvh.so (oracle/Makefile `ref`).

* tests/golden/bvh_kat.json holds that library's outputs (tests/gen_golden.py) and travels everywhere; the oracle AND the product's
  host builder (b2r_bvh_build in libb2r.so — no GPU needed) must reproduce them bit for bit;
* where the library itself is present the comparison is repeated live, up to the 100k-sphere scene of BASELINE config C3."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import b2r
import gen_golden
import oracle_py
import scenes

G = os.path.join(os.path.dirname(__file__), "golden")
HAVE_REF = os.path.exists(os.path.join(os.path.dirname(oracle_py.__file__), "_ref", "librefbvh.so"))


def oracle_build(geo):
    L = oracle_py.lib(); n = len(geo)
    nodes = np.zeros((max(2 * n - 1, 1), 32), np.uint8); prims = np.zeros(n, geo.dtype); ids = np.zeros(n, np.uint32)
    nn = L.orc_build_bvh(geo.ctypes.data, n, nodes.ctypes.data, prims.ctypes.data, ids.ctypes.data)
    return nodes[:nn].tobytes(), prims


def product_build(geo):
    nodes, prims, ids = b2r.build_bvh(geo)
    assert np.array_equal(prims["position"], geo["position"][ids])  # prim_ids is the permutation the reorder applied
    return nodes.tobytes(), prims


def test_builder_matches_reference_builder_golden():
    want = json.load(open(os.path.join(G, "bvh_kat.json")))["builds"]
    for name, geo in gen_golden.bvh_scenes():
        assert gen_golden.bvh_digest(*oracle_build(geo)) == want[name], f"oracle builder differs from the reference's on {name}"
        assert gen_golden.bvh_digest(*product_build(geo)) == want[name], f"b2r_bvh_build differs from the reference's on {name}"


def test_half_area_quirk_matches_reference():
    """Q17: Node::half_area() returns d.y*d.z only; checked through the builder above and directly here on the product's nodes."""
    want = json.load(open(os.path.join(G, "bvh_kat.json")))["half_area"]
    H = float.fromhex
    for lo, hi, area in want:
        d = np.array([H(v) for v in hi], np.float32) - np.array([H(v) for v in lo], np.float32)
        assert float(np.float32(d[1] * d[2])).hex() == area


def test_sphere_loops_match_reference_golden(hostcheck):
    want = json.load(open(os.path.join(G, "bvh_kat.json")))["intersect"]
    geos = dict(gen_golden.bvh_scenes()); L = oracle_py.lib(); ctx = {}
    for rec, (scene_name, kind, rays, tfar) in zip(want, gen_golden.intersect_inputs()):
        assert (rec["scene"], rec["kind"], rec["n"]) == (scene_name, kind, len(rays))
        if scene_name not in ctx:
            sc = scenes.default_scene() if scene_name == "default" else scenes.random_scene(100)
            assert sc["geometry"].tobytes() == geos[scene_name].tobytes()
            o = oracle_py.Oracle(16, 16); o.set_scene(sc); ctx[scene_name] = o
        o = ctx[scene_name]; n = len(rays)
        tf = np.zeros(n, np.float32); pid = np.zeros(n, np.int32)
        fn = L.orc_trace_closest if kind == "simd" else L.orc_trace_closest_scalar
        fn(o.h, oracle_py._fptr(rays), n, oracle_py._fptr(tf), pid.ctypes.data)
        occ = o.trace_shadow(rays, tfar)
        assert hashlib.sha256(tf.tobytes()).hexdigest() == rec["sha256_tfar"] and hashlib.sha256(pid.tobytes()).hexdigest() == rec["sha256_prim"], (scene_name, kind, n)
        assert hashlib.sha256(np.asarray(occ, np.uint8).tobytes()).hexdigest() == rec["sha256_occluded"]
        assert int((pid >= 0).sum()) == rec["hits"]
        if kind == "tail":  # the product's scalar-tail routine (B2R_FLAG_REFERENCE_EXACT) on the same rays
            nodes, prims, ids = o.bvh()
            sph = np.ascontiguousarray(np.concatenate([prims["position"], prims["radius_sq"][:, None]], axis=1).astype(np.float32))
            for i in range(n):
                best, bp = C.c_float(), C.c_int32()
                hostcheck.hc_closest_scalar(sph.ctypes.data, len(sph), rays[i].ctypes.data, C.byref(best), C.byref(bp))
                assert bp.value == pid[i] and (bp.value < 0 or np.float32(best.value) == tf[i])
        if kind == "simd":  # the product's closest-hit routine is the AVX2+FMA formula (DESIGN.md "Numerics"); its any-hit the shadow one
            nodes, prims, ids = o.bvh()
            sph = np.concatenate([prims["position"], prims["radius_sq"][:, None]], axis=1).astype(np.float32)
            for i in range(0, n, 9):
                best, bp = np.float32(3.4028234663852886e38), -1
                for j in range(len(sph)):
                    dd = C.c_float()
                    if hostcheck.hc_sphere_closest(sph[j].ctypes.data, rays[i].ctypes.data, C.byref(dd)) and dd.value < best:
                        best, bp = np.float32(dd.value), j
                assert bp == pid[i] and (bp < 0 or best == tf[i])
                assert any(hostcheck.hc_sphere_any(sph[j].ctypes.data, rays[i].ctypes.data, C.c_float(tfar[i])) for j in range(len(sph))) == bool(occ[i])
    for o in ctx.values():
        o.close()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/librefbvh.so not present (built only where /root/reference exists)")
def test_live_builder_up_to_c3_size():
    ref = gen_golden.ref_bvh_lib()
    for n, seed in ((17, 1), (257, 2), (4099, 3), (100000, 0x04D15A07)):
        geo = np.ascontiguousarray(scenes.random_scene(n, seed=seed)["geometry"])
        want = gen_golden.bvh_digest(*gen_golden.ref_bvh_build(ref, geo))
        assert gen_golden.bvh_digest(*oracle_build(geo)) == want, n
        assert gen_golden.bvh_digest(*product_build(geo)) == want, n


def test_split_heuristic_argument_matches_reference_golden():
    """BoundingVolumeHierarchy(primitives, SplitHeuristic{log_cluster_size, cost_ratio}) (BVH.hpp:70-83,:90) — the constructor's second
    argument, which the app leaves at its default: b2r_bvh_build_ex reproduces the reference's node arrays and leaf order for other
    values too (clustered leaf cost, cheaper / dearer "do not split" bound)."""
    want = json.load(open(os.path.join(G, "bvh_heuristic_kat.json")))
    geos = dict(gen_golden.bvh_scenes()); n = 0
    for key, digest in want.items():
        name, L, ratio = key.split("|")
        nodes, prims, ids = b2r.build_bvh(geos[name], int(L), float(ratio))
        assert gen_golden.bvh_digest(nodes.tobytes(), prims) == digest, key
        n += 1
    assert n == len(gen_golden.HEURISTICS) * 7
    assert len({d["sha256_nodes"] for k, d in want.items() if k.startswith("random1000|")}) > 3     # the argument does change the tree


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/librefbvh.so not present (built only where /root/reference exists)")
def test_split_heuristic_live_on_all_cores():
    ref = gen_golden.ref_bvh_lib()
    geo = np.ascontiguousarray(scenes.random_scene(30000, seed=0x5EED)["geometry"])      # above the 20 000-sphere threshold of the threaded build
    for L, ratio in ((0, 1.0), (2, 0.5), (0, 4.0)):
        want = gen_golden.bvh_digest(*gen_golden.ref_bvh_build_h(ref, geo, L, ratio))
        nodes, prims, ids = b2r.build_bvh(geo, L, ratio)
        assert gen_golden.bvh_digest(nodes.tobytes(), prims) == want, (L, ratio)
