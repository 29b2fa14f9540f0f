"""Regenerates tests/golden/*.json. Run in the build container (needs /root/reference for the reference-derived part):

    python tests/gen_golden.py

* rng_kat.json    — outputs of the reference's OWN Random.hpp/Bitmanip.hpp, compiled verbatim into oracle/_ref/librefrng.so
                    (oracle/Makefile `ref`): hash_u32, hash_2d, pcg streams, rand_bounded_int, make_unit_float, bitreverse.
                    These pin the oracle's RNG layer to the reference itself.
* sampling_kat.json — outputs of the reference's OWN Sampling.hpp, VectorMath.hpp:581-662 and Color.hpp:30-74, compiled verbatim into
                    oracle/_ref/librefsampling.so: fast_sincos/asin/atan2, median 3/5, hemisphere, orthonormal_basis, tangent_space,
                    to_local/to_world, conePdf/spherePdf, sample_direction_to_sphere, powerHeuristic(_over_f), ACES tonemapping
                    (scalar and Vec8f); and of Camera.hpp:5-59,81-87 (Projection, View, generate_ray). These pin the oracle's
                    sampling / scalar-math / tonemap / camera layer to the reference itself.
* bvh_kat.json    — outputs of the reference's OWN BVH builder (BVH.hpp:17-87,91-206) and sphere loops (:239-287 closest hit — AVX2+FMA block and
                    scalar tail — and :292-304 shadow), compiled verbatim into oracle/_ref/librefbvh.so: digests of node arrays and leaf
                    order for the default and random scenes, Node::half_area() values (Q17), hit distances / ids / occlusion flags.
* renderer_kat.json — outputs of the reference's OWN Renderer<>::Accumulate / Render (Renderer.hpp and everything it includes, compiled into
                    oracle/_ref/librefrenderer.so by oracle/ref_renderer_build.sh): digests of the five bucket-sum planes and of the
                    tonemapped RGBA32F frame for small renders of the default, random, white-furnace and sky/HDRI scenes.
* renderer_ggx_kat.json — the same from the reference's GGX build (`#define BRDF 1`, Renderer.hpp:70; oracle/_ref/librefrenderer_ggx.so, with the
                    all-zero gloss_decay_table that build has to supply).
* survey_kat.json — the known-answer table of SURVEY.md §8c (derived from the same reference file), transcribed.
* oracle_frames.json — outputs of the ORACLE (not the reference, which cannot be built: SURVEY §8c) on small inputs: bucket-sum
                    checksums, counters, BVH order. They pin the oracle against accidental change and give the GPU tests a
                    fixture that does not need the oracle's code path at all.
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "cpu-raytracing-experiments_b200")]
import oracle_py  # noqa: E402
import scenes  # noqa: E402


def gen_rng():
    oracle_py.build()
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librefrng.so"))
    u = C.c_uint32
    for n in ("ref_hash_u32", "ref_hash_2d", "ref_pcg_generate", "ref_rand_bounded_int", "ref_bitreverse"):
        getattr(ref, n).restype = u
    ref.ref_rand_unit_float.restype = C.c_float; ref.ref_make_unit_float.restype = C.c_float
    rs = np.random.RandomState(20261018)
    xs = [0, 1, 2, 5, 0xFFFFFFFF, 0x80000000, 12345678] + [int(v) for v in rs.randint(0, 2 ** 32, 57, dtype=np.uint64)]
    pairs = [(0, 0), (1, 0), (0, 1), (1, 33), (2, 33), (5, 8447), (64, (2088959 * 33) & 0xFFFFFFFF)] + \
            [(int(a), int(b)) for a, b in rs.randint(0, 2 ** 32, (57, 2), dtype=np.uint64)]
    out = {"hash_u32": [[x, ref.ref_hash_u32(u(x))] for x in xs],
           "hash_2d": [[a, b, ref.ref_hash_2d(u(a), u(b))] for a, b in pairs],
           "bitreverse": [[x, ref.ref_bitreverse(u(x))] for x in xs],
           "make_unit_float": [[x, float(ref.ref_make_unit_float(u(x))).hex()] for x in xs],
           "pcg": [], "bounded": []}
    for s0 in xs[:24]:
        st = u(s0); outs = []
        for _ in range(6):
            outs.append([ref.ref_pcg_generate(C.byref(st)), st.value])
        st = u(s0); fl = [float(ref.ref_rand_unit_float(C.byref(st))).hex() for _ in range(4)]
        out["pcg"].append({"state": s0, "out_next": outs, "floats": fl})
    for s0 in xs[:24]:
        for rng in (1, 2, 3, 100, 1000):
            st = u(s0); out["bounded"].append([s0, rng, ref.ref_rand_bounded_int(C.byref(st), u(rng))])
    json.dump(out, open(os.path.join(HERE, "golden", "rng_kat.json"), "w"), indent=0)


def ref_sampling_lib():
    oracle_py.build()
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librefsampling.so"))
    f = C.c_float
    for n in ("ref_fast_asin", "ref_fast_atan2", "ref_median3", "ref_median5", "ref_cone_pdf", "ref_sphere_pdf", "ref_power_heuristic", "ref_power_heuristic_over_f"):
        getattr(ref, n).restype = f
    ref.ref_fast_asin.argtypes = [f]; ref.ref_fast_atan2.argtypes = [f, f]; ref.ref_median3.argtypes = [f, f, f]; ref.ref_median5.argtypes = [C.c_void_p]
    ref.ref_cone_pdf.argtypes = [f]; ref.ref_sphere_pdf.argtypes = [f, f]; ref.ref_power_heuristic.argtypes = [f, f]; ref.ref_power_heuristic_over_f.argtypes = [f, f]
    ref.ref_fast_sincos.argtypes = [f, C.POINTER(f), C.POINTER(f)]
    ref.ref_hemisphere.argtypes = [f, f, C.c_void_p]
    for n in ("ref_orthonormal_basis", "ref_tangent_space"):
        getattr(ref, n).argtypes = [C.c_void_p, C.c_void_p]
    for n in ("ref_to_local", "ref_to_world"):
        getattr(ref, n).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    ref.ref_sample_direction_to_sphere.argtypes = [C.c_void_p, f, f, f, f, f, C.c_void_p]
    ref.ref_tonemap_scalar.argtypes = [C.c_void_p]; ref.ref_tonemap_vec8.argtypes = [C.c_void_p] * 3
    return ref


def sampling_inputs(seed=20261018, n=160):
    """The inputs of the sampling known-answer vectors: shared by the generator and by the live cross-check test."""
    rs = np.random.RandomState(seed); F = np.float32
    unit = rs.randn(n, 3); unit = (unit / np.linalg.norm(unit, axis=1, keepdims=True)).astype(F)
    near_pole = np.array([[1e-4, 0, -1], [0, 3e-4, -0.99999994], [0, 0, -1], [0, 0, 1], [1, 0, 0], [0, -1, 0], [2e-4, -1e-4, -0.9999999]], F)
    return {
        "sincos": np.concatenate([rs.uniform(0, 2 * np.pi, n), [0.0, np.pi / 2, np.pi, 2 * np.pi, 6.2831855, -1.0, 40.0, 1e-30]]).astype(F),
        "asin": np.concatenate([rs.uniform(-1.2, 1.2, n), [0.0, 1.0, -1.0, 1e-40]]).astype(F),
        "atan2": np.concatenate([rs.uniform(-2, 2, (n, 2)), [[0, 0], [0, 1], [1, 0], [-1, 0], [0, -1], [1, 1], [-0.0, -1.0]]]).astype(F),
        "median": rs.rand(n, 5).astype(F),
        "unit01": np.concatenate([rs.rand(n, 2), [[0, 0], [1, 1], [1, 0], [0, 1], [0.5, 0.25]]]).astype(F),
        "normals": np.concatenate([unit, near_pole]).astype(F),
        "vectors": rs.uniform(-1, 1, (n + len(near_pole), 3)).astype(F),
        # sample_direction_to_sphere: sin^2(theta_max) spread over both branches of the small-angle switch, centre distance, two uniforms
        "cone": np.stack([rs.rand(n) ** 6, rs.uniform(0.05, 400, n), rs.rand(n), rs.rand(n)], axis=1).astype(F),
        "pairs": np.concatenate([rs.rand(n, 2) * 5, [[0, 0], [1e-4, 1e-4], [3, 0]]]).astype(F),
        "rgb": np.concatenate([rs.rand(n, 3) * 4, rs.rand(16, 3) * 1e-3, [[0, 0, 0], [100, 50, 1]]]).astype(F),
    }


def eval_sampling(fn, inp):
    """fn: dict of callables with the ref_* signatures (reference library or oracle adapter). Returns hex-float records."""
    f = C.c_float; H = lambda a: [float(v).hex() for v in np.asarray(a, np.float32).ravel()]
    out = {k: [] for k in ("sincos", "asin", "atan2", "median3", "median5", "hemisphere", "onb", "tangent", "to_local", "to_world", "cone_pdf", "sphere_pdf",
                           "sample_sphere", "power", "power_over_f", "tonemap")}
    for x in inp["sincos"]:
        s, c = f(), f(); fn["sincos"](f(x), C.byref(s), C.byref(c)); out["sincos"].append(H([s.value, c.value]))
    for x in inp["asin"]:
        out["asin"].append(H([fn["asin"](f(x))]))
    for y, x in inp["atan2"]:
        out["atan2"].append(H([fn["atan2"](f(y), f(x))]))
    for v in inp["median"]:
        v = np.ascontiguousarray(v)
        if "median3" in fn:
            out["median3"].append(H([fn["median3"](f(v[0]), f(v[1]), f(v[2]))]))
        out["median5"].append(H([fn["median5"](v.ctypes.data)]))
    for t, s in inp["unit01"]:
        o = np.zeros(3, np.float32); fn["hemisphere"](f(t), f(s), o.ctypes.data); out["hemisphere"].append(H(o))
    for nrm, vec in zip(inp["normals"], inp["vectors"]):
        nrm = np.ascontiguousarray(nrm); vec = np.ascontiguousarray(vec)
        o6 = np.zeros(6, np.float32); fn["onb"](nrm.ctypes.data, o6.ctypes.data); out["onb"].append(H(o6))
        q = np.zeros(4, np.float32); fn["tangent"](nrm.ctypes.data, q.ctypes.data); out["tangent"].append(H(q))
        o = np.zeros(3, np.float32); fn["to_local"](q.ctypes.data, vec.ctypes.data, o.ctypes.data); out["to_local"].append(H(o))
        o = np.zeros(3, np.float32); fn["to_world"](q.ctypes.data, vec.ctypes.data, o.ctypes.data); out["to_world"].append(H(o))
    for (s2, cd, t, s), wc in zip(inp["cone"], inp["normals"]):
        wc = np.ascontiguousarray(wc); r2 = np.float32(s2 * cd * cd)
        o5 = np.zeros(5, np.float32); fn["sample_sphere"](wc.ctypes.data, f(s2), f(cd), f(r2), f(t), f(s), o5.ctypes.data); out["sample_sphere"].append(H(o5))
        out["sphere_pdf"].append(H([fn["sphere_pdf"](f(r2), f(np.float32(cd * cd)))]))
        if "cone_pdf" in fn:
            out["cone_pdf"].append(H([fn["cone_pdf"](f(t))]))
    for a, b in inp["pairs"]:
        out["power"].append(H([fn["power"](f(a), f(b))])); out["power_over_f"].append(H([fn["power_over_f"](f(a), f(b))]))
    for rgb in inp["rgb"]:
        o = np.ascontiguousarray(rgb).copy(); fn["tonemap"](o.ctypes.data); out["tonemap"].append(H(o))
    return out


def ref_sampling_fns(ref):
    def tonemap_vec8(p):  # lane 0 of the Vec8f overload Renderer::Render uses (Renderer.hpp:461-473), checked against the scalar overload's formula too
        rgb = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (3,))
        r, g, b = (np.full(8, rgb[i], np.float32) for i in range(3))
        ref.ref_tonemap_vec8(r.ctypes.data, g.ctypes.data, b.ctypes.data)
        rgb[0], rgb[1], rgb[2] = r[0], g[0], b[0]
    return {"sincos": ref.ref_fast_sincos, "asin": ref.ref_fast_asin, "atan2": ref.ref_fast_atan2, "median3": ref.ref_median3, "median5": ref.ref_median5,
            "hemisphere": ref.ref_hemisphere, "onb": ref.ref_orthonormal_basis, "tangent": ref.ref_tangent_space, "to_local": ref.ref_to_local,
            "to_world": ref.ref_to_world, "cone_pdf": ref.ref_cone_pdf, "sphere_pdf": ref.ref_sphere_pdf, "sample_sphere": ref.ref_sample_direction_to_sphere,
            "power": ref.ref_power_heuristic, "power_over_f": ref.ref_power_heuristic_over_f, "tonemap": tonemap_vec8}


def camera_inputs(seed=20261018, n=24):
    """(eye, dir, W, H, focal_mm) set-ups — the default scene's camera (Application.cpp:95-100), C3's, and random ones — each with
    a few pixels and sub-pixel samples."""
    rs = np.random.RandomState(seed); F = np.float32
    cams = [((-0.2, 0.3, 1.0), (0.1, -0.4, -1.0), 1280, 720, 40.0), ((0.0, 60.0, 300.0), (0.0, 0.0, -1.0), 1920, 1088, 50.0),
            ((1.0, 2.0, 3.0), (0.0, -1.0, 1e-3), 640, 368, 24.0), ((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 320, 192, 85.0)]
    for _ in range(n):
        d = rs.randn(3); cams.append((tuple(rs.uniform(-50, 50, 3)), tuple(d), int(rs.randint(1, 240)) * 16, int(rs.randint(1, 135)) * 16, float(rs.uniform(12, 200))))
    out = []
    for eye, d, w, h, fl in cams:
        px = [(0, 0), (w - 1, h - 1), (w // 2, h // 2)] + [(int(rs.randint(0, w)), int(rs.randint(0, h))) for _ in range(5)]
        out.append({"eye": [float(F(v)) for v in eye], "dir": [float(F(v)) for v in d], "w": w, "h": h, "focal": float(F(fl)),
                    "pixels": [[x, y, float(F(rs.rand())), float(F(rs.rand()))] for x, y in px]})
    return out


def eval_camera_ref(ref, cams):
    H = lambda a: [float(v).hex() for v in np.asarray(a, np.float32).ravel()]
    f = C.c_float; ref.ref_generate_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, f, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rows = []
    for c in cams:
        eye = np.array(c["eye"], np.float32); d = np.array(c["dir"], np.float32)
        for x, y, s0, s1 in c["pixels"]:
            smp = np.array([s0, s1], np.float32); o6 = np.zeros(6, np.float32); cam7 = np.zeros(7, np.float32)
            ref.ref_generate_ray(eye.ctypes.data, d.ctypes.data, c["w"], c["h"], f(c["focal"]), x, y, smp.ctypes.data, o6.ctypes.data, cam7.ctypes.data)
            rows.append(H(o6) + H(cam7))
    return rows


def eval_camera_oracle(cams):
    """the oracle's Camera restatement on the same inputs: ray (6 floats) + orient wxyz, half_width, half_height, z"""
    H = lambda a: [float(v).hex() for v in np.asarray(a, np.float32).ravel()]
    rows = []
    for c in cams:
        o = oracle_py.Oracle(c["w"], c["h"], max_bounces=2, K=1)
        o.L.orc_set_camera_lookat(o.h, oracle_py.farr(*c["eye"]), oracle_py.farr(*c["dir"]), c["focal"], 1.0)
        raw = o.camera_raw()
        for x, y, s0, s1 in c["pixels"]:
            o6 = (C.c_float * 6)(); o.L.orc_generate_ray_at(o.h, x, y, oracle_py.farr(s0, s1), o6)
            rows.append(H(o6[:]) + H(raw[3:10]))
        o.close()
    return rows


def gen_sampling():
    ref = ref_sampling_lib()
    out = eval_sampling(ref_sampling_fns(ref), sampling_inputs())
    out["camera"] = eval_camera_ref(ref, camera_inputs())
    out["_inputs"] = "tests/gen_golden.py sampling_inputs(seed=20261018, n=160)"
    json.dump(out, open(os.path.join(HERE, "golden", "sampling_kat.json"), "w"), indent=0)


# ---------------------------------------------------------------- reference BVH builder and sphere loops (oracle/_ref/librefbvh.so)
def camera_move_inputs(seed=20261018, n=40):
    """Camera{eye, dir} followed by a few {RotateLocal(angles), TranslateLocal(offset)} steps (the app's mouse-look and WASD handlers,
    Application.cpp:236-247,299): small and large angles, all three axes."""
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n):
        eye = rs.uniform(-5, 5, 3).astype(np.float32); d = rs.randn(3).astype(np.float32)
        k = int(rs.randint(1, 6))
        ang = (rs.uniform(-1, 1, (k, 3)) * (0.02 if i % 2 else 2.5)).astype(np.float32)
        if i % 3 == 0: ang[:, 2] = 0.0                                       # the app only sends pitch and yaw
        off = rs.uniform(-2, 2, (k, 3)).astype(np.float32)
        out.append((eye, d, np.ascontiguousarray(ang), np.ascontiguousarray(off)))
    return out


def eval_camera_move(fn, moves):
    """fn(eye, dir, angles, offsets, n, out7) -> list of hex-float strings [pos xyz, orient wxyz] per case"""
    rows = []
    for eye, d, ang, off in moves:
        out = np.zeros(7, np.float32)
        fn(C.c_void_p(eye.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(ang.ctypes.data), C.c_void_p(off.ctypes.data), len(ang), C.c_void_p(out.ctypes.data))
        rows.append([float(v).hex() for v in out])
    return rows


def gen_camera_move():
    json.dump({"camera_move": eval_camera_move(ref_sampling_lib().ref_camera_move, camera_move_inputs())},
              open(os.path.join(HERE, "golden", "camera_move_kat.json"), "w"), indent=0)


def ref_bvh_lib():
    oracle_py.build()
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librefbvh.so"))
    ref.ref_bvh_build.restype = C.c_uint32; ref.ref_bvh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    ref.ref_node_half_area.restype = C.c_float; ref.ref_node_half_area.argtypes = [C.c_void_p, C.c_void_p]
    ref.ref_intersect_closest.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    ref.ref_intersect_shadow.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    return ref


def bvh_scenes():
    """(name, geometry) pairs of the BVH known-answer vectors: no two spheres share a centroid coordinate except in `default`
    (whose nine spheres sit inside the insertion-sort range of both std::sort implementations)."""
    out = [("default", scenes.default_scene()["geometry"])]
    for n in (1, 2, 3, 4, 5, 9, 33, 100, 1000, 5000):
        out.append((f"random{n}", scenes.random_scene(n)["geometry"]))
    out.append(("random777_light_every_10", scenes.random_scene(777, light_every=10, seed=0xB0B)["geometry"]))
    return [(k, np.ascontiguousarray(g)) for k, g in out]


def bvh_digest(nodes_bytes, prims):
    """sha256 of the 32-byte node records and of the reordered spheres' named fields (the records' padding is not compared)."""
    fields = np.concatenate([prims["position"].astype(np.float32).view(np.uint32), prims["radius_sq"].astype(np.float32).view(np.uint32)[:, None],
                             prims["material_ID"].astype(np.int32).view(np.uint32)[:, None]], axis=1)
    return {"n_nodes": len(nodes_bytes) // 32, "sha256_nodes": hashlib.sha256(nodes_bytes).hexdigest(), "sha256_prims": hashlib.sha256(fields.tobytes()).hexdigest()}


def ref_bvh_build(ref, geo):
    n = len(geo); nodes = np.zeros((max(2 * n - 1, 1), 32), np.uint8); prims = np.zeros(n, geo.dtype)
    nn = ref.ref_bvh_build(geo.ctypes.data, n, nodes.ctypes.data, prims.ctypes.data)
    return nodes[:nn].tobytes(), prims


HEURISTICS = [(0, 1.0), (0, 0.25), (0, 4.0), (1, 1.0), (2, 0.5), (3, 2.0)]   # SplitHeuristic{log_cluster_size, cost_ratio}, BVH.hpp:70-83


def ref_bvh_build_h(ref, geo, log_cluster_size, cost_ratio):
    n = len(geo); nodes = np.zeros((max(2 * n - 1, 1), 32), np.uint8); prims = np.zeros(n, geo.dtype)
    ref.ref_bvh_build_h.restype = C.c_uint32
    nn = ref.ref_bvh_build_h(C.c_void_p(geo.ctypes.data), C.c_uint32(n), C.c_uint32(log_cluster_size), C.c_float(cost_ratio), C.c_void_p(nodes.ctypes.data), C.c_void_p(prims.ctypes.data))
    return nodes[:nn].tobytes(), prims


def gen_bvh_heuristics():
    """the reference's constructor with non-default SplitHeuristic arguments -> tests/golden/bvh_heuristic_kat.json"""
    ref = ref_bvh_lib(); out = {}
    for name, geo in bvh_scenes():
        if name in ("random1", "random2", "random3", "random4", "random5000"): continue
        for L, ratio in HEURISTICS:
            out[f"{name}|{L}|{ratio}"] = bvh_digest(*ref_bvh_build_h(ref, geo, L, ratio))
    json.dump(out, open(os.path.join(HERE, "golden", "bvh_heuristic_kat.json"), "w"), indent=0)


def intersect_inputs(seed=20261018):
    """seeded rays for the sphere-loop vectors, against the `default` and `random100` scenes in the reference's BVH leaf order"""
    rs = np.random.RandomState(seed); out = []
    for scene_name, span in (("default", 3.0), ("random100", 150.0)):
        for n in (256, 64, 8):      # multiples of 8: every ray goes through the AVX2+FMA block (BVH.hpp:250-268)
            rays = np.zeros((n, 6), np.float32); rays[:, :3] = rs.uniform(-span, span, (n, 3)); d = rs.randn(n, 3); rays[:, 3:] = d / np.linalg.norm(d, axis=1, keepdims=True)
            out.append((scene_name, "simd", rays, rs.uniform(0.5, 2 * span, n).astype(np.float32)))
        for n in (7, 5, 1):         # fewer than 8: every ray goes through the scalar tail (:270-286)
            rays = np.zeros((n, 6), np.float32); rays[:, :3] = rs.uniform(-span, span, (n, 3)); d = rs.randn(n, 3); rays[:, 3:] = d / np.linalg.norm(d, axis=1, keepdims=True)
            out.append((scene_name, "tail", rays, rs.uniform(0.5, 2 * span, n).astype(np.float32)))
    return out


def gen_bvh():
    ref = ref_bvh_lib()
    out = {"builds": {}, "half_area": [], "intersect": []}
    built = {}
    for name, geo in bvh_scenes():
        nodes, prims = ref_bvh_build(ref, geo); built[name] = prims
        out["builds"][name] = bvh_digest(nodes, prims)
    rs = np.random.RandomState(3)
    for _ in range(40):  # Q17: Node::half_area() is d.y*d.z only
        lo = rs.uniform(-10, 10, 3).astype(np.float32); hi = (lo + rs.uniform(0, 20, 3)).astype(np.float32)
        out["half_area"].append([[float(v).hex() for v in lo], [float(v).hex() for v in hi], float(ref.ref_node_half_area(lo.ctypes.data, hi.ctypes.data)).hex()])
    for scene_name, kind, rays, tfar in intersect_inputs():
        prims = built[scene_name]; n = len(rays)
        tf = np.zeros(n, np.float32); pid = np.zeros(n, np.int32); occ = np.zeros(n, np.uint8)
        ref.ref_intersect_closest(prims.ctypes.data, len(prims), rays.ctypes.data, n, tf.ctypes.data, pid.ctypes.data)
        ref.ref_intersect_shadow(prims.ctypes.data, len(prims), rays.ctypes.data, tfar.ctypes.data, n, occ.ctypes.data)
        out["intersect"].append({"scene": scene_name, "kind": kind, "n": n, "sha256_tfar": hashlib.sha256(tf.tobytes()).hexdigest(),
                                 "sha256_prim": hashlib.sha256(pid.tobytes()).hexdigest(), "sha256_occluded": hashlib.sha256(occ.tobytes()).hexdigest(),
                                 "hits": int((pid >= 0).sum()), "occluded": int(occ.sum())})
    json.dump(out, open(os.path.join(HERE, "golden", "bvh_kat.json"), "w"), indent=0)


# ---------------------------------------------------------------- the reference's own Renderer<> (oracle/_ref/librefrenderer.so)
def renderer_cases():
    """(name, scene, width, height, max_bounces, samples, first_sample) — small frames the reference renders in well under a second.
    max_bounces is a template argument of Renderer<>: only the instantiated values (oracle_py.REF_MAX_BOUNCES) can be used."""
    # (no zero-light scene: with MIS on the reference divides by the light count and reads prims[0] of an empty list)
    sky = scenes.bvh_test_scene(40, hdri=scenes.synthetic_hdri())
    return [
        ("default_64x48_mb8", scenes.default_scene(), 64, 48, 8, 5, 0),
        ("default_160x96_mb16", scenes.default_scene(), 160, 96, 16, 10, 0),
        ("default_96x64_mb16_from_acc60", scenes.default_scene(), 96, 64, 16, 5, 60),
        ("default_64x64_mb1", scenes.default_scene(), 64, 64, 1, 5, 0),
        ("default_64x64_mb2", scenes.default_scene(), 64, 64, 2, 5, 0),
        ("random300_64x48_mb4", scenes.random_scene(300, light_every=20), 64, 48, 4, 5, 0),
        ("random33_80x48_mb8_third_emissive", scenes.random_scene(33, light_every=3), 80, 48, 8, 5, 0),
        ("random2000_48x32_mb16", scenes.random_scene(2000, light_every=100), 48, 32, 16, 5, 0),
        ("sky_hdri_40_spheres_64x48_mb8", sky, 64, 48, 8, 5, 0),
    ]


def renderer_record(buckets, acted, frame):
    return {"sha256_buckets": hashlib.sha256(np.ascontiguousarray(buckets, np.float32).tobytes()).hexdigest(), "render_acted": bool(acted),
            "sha256_frame": hashlib.sha256(np.ascontiguousarray(frame, np.float32).tobytes()).hexdigest(),
            "sum_rgb": [float(v) for v in np.asarray(buckets).sum(axis=(0, 2), dtype=np.float64)]}


def renderer_ggx_cases():
    """the same for the reference's GGX build (`#define BRDF 1`, oracle/_ref/librefrenderer_ggx.so; gloss_decay_table all zeros)"""
    return [
        ("ggx_default_64x48_mb8", scenes.default_scene(), 64, 48, 8, 5, 0),          # roughness 0.05 .. 1, F0 0.03 .. 0.94
        ("ggx_default_96x64_mb16_from_acc60", scenes.default_scene(), 96, 64, 16, 5, 60),
        ("ggx_random300_64x48_mb4", scenes.ggx_random_scene(300), 64, 48, 4, 5, 0),  # one mirror material (alpha == 0)
        ("ggx_random2000_48x32_mb16", scenes.ggx_random_scene(2000, light_every=100), 48, 32, 16, 5, 0),
        ("ggx_brdf_test_64x48_mb8", scenes.brdf_test_scene(hdri=scenes.synthetic_hdri(24, 12, seed=2)), 64, 48, 8, 5, 0),  # Application.cpp:123-217: the scene made for this closure
    ]


def gen_renderer_ggx():
    oracle_py.build()
    out = {}
    for name, sc, w, h, mb, n, first in renderer_ggx_cases():
        r = oracle_py.ReferenceRenderer(sc, w, h, mb, ggx=True)
        if first:
            r.set_accumulations(first)
        r.accumulate(n)
        acted, frame = r.render()
        out[name] = renderer_record(r.buckets(), acted, frame)
        r.close()
    json.dump(out, open(os.path.join(HERE, "golden", "renderer_ggx_kat.json"), "w"), indent=0)


def gen_renderer():
    oracle_py.build()
    out = {}
    for name, sc, w, h, mb, n, first in renderer_cases():
        r = oracle_py.ReferenceRenderer(sc, w, h, mb)
        if first:
            r.set_accumulations(first)
        r.accumulate(n)
        acted, frame = r.render()
        out[name] = renderer_record(r.buckets(), acted, frame)
        r.close()
    json.dump(out, open(os.path.join(HERE, "golden", "renderer_kat.json"), "w"), indent=0)
    # actual values (not digests) for the tolerance-based GPU comparison: the reference's first sample (bucket 1, linear radiance, tile
    # order) and its converged 200-spp tonemapped frame of the default scene at 160x96, max_bounces 16
    r = oracle_py.ReferenceRenderer(scenes.default_scene(), 160, 96, 16)
    r.accumulate(1); first = r.buckets()[1].copy()
    r.accumulate(199); acted, frame = r.render(); r.close()
    assert acted
    np.savez_compressed(os.path.join(HERE, "golden", "reference_default_160x96_mb16.npz"), first_sample=first, frame_200spp=frame)


def gen_survey():
    H = float.fromhex
    table = {
        "hash_u32": [[0, 0xE6FE3BEB], [1, 0xE02DC198], [2, 0xEB59CF0C], [5, 0xA0787FC7], [0xFFFFFFFF, 0x3F4B5D68]],
        "hash_2d": [[0, 0, 0], [1, 0, 0xEF386249], [0, 1, 0xC2A29A69], [1, 33, 0x56410662], [2, 33, 0xDDEBF9B1], [5, 8447, 0xDC9B143B],
                    [64, (2088959 * 33) & 0xFFFFFFFF, 0x28716627]],
        "pcg_chain": {"state": 0x12345678, "out_next": [[0x28AE66B1, 0xCFF935DD], [0x995312E1, 0x439D1B46], [0x3A39CE3D, 0x18041D83], [0x83FD0318, 0xE9AD0DA4]]},
        "bounded": [[0xDEADBEEF, 1, 0], [0xDEADBEEF, 2, 1], [0xDEADBEEF, 3, 2], [0xDEADBEEF, 100, 96]],
        "unit_float_max": "0x1p+0",
        "bitreverse": [[1, 0x80000000], [6, 0x60000000]],
        # per-pixel streams: x, y, W, max_bounces, acc, branch, seed, state0, u0, u1, u2
        "streams": [
            [0, 0, 1280, 16, 1, 0, 0, 0xEF386249, "0x1.223c1p-3", "0x1.ed82e4p-1", "0x1.4676d8p-2"],
            [17, 3, 1280, 16, 1, 0, 10065, 0xCD0C5AA7, "0x1.0fcc7p-1", "0x1.5fffdp-1", "0x1.9dc21ap-2"],
            [17, 3, 1280, 16, 1, 1, 10065, 0x773C7B1F, "0x1.738daap-1", "0x1.ae810cp-1", "0x1.0d5ce6p-1"],
            [17, 3, 1280, 16, 5, 6, 10065, 0x29579D97, "0x1.858172p-1", "0x1.a0357ep-3", "0x1.5bb406p-1"],
            [1279, 719, 1280, 16, 1, 0, 30412767, 0x8F2CF485, "0x1.010e96p-1", "0x1.c7ae78p-2", "0x1.dc142cp-1"],
            [17, 3, 1280, 8, 1, 0, 5185, 0xA61E83FC, "0x1.0b0022p-6", "0x1.856128p-2", "0x1.483d54p-1"],
            [1919, 1087, 1920, 16, 64, 31, 68935647, 0x8AB9E15C, "0x1.811f82p-1", "0x1.42f1e2p-2", "0x1.c28206p-2"],
            [3839, 2159, 3840, 16, 1024, 0, 273715167, 0x33C0024F, "0x1.7c257cp-4", "0x1.ebef7ap-2", "0x1.7df18p-1"],
        ],
    }
    assert H(table["unit_float_max"]) == 1.0
    json.dump(table, open(os.path.join(HERE, "golden", "survey_kat.json"), "w"), indent=0)


def frame_record(o, sc):
    b = o.buckets()
    return {"sha256_buckets": hashlib.sha256(b.tobytes()).hexdigest(), "sum_rgb": [float(v) for v in b.sum(axis=(0, 2), dtype=np.float64)],
            "counters": {k: int(v) for k, v in o.counters().items()}}


def gen_frames():
    out = {}
    sc = scenes.default_scene()
    o = oracle_py.Oracle(320, 192, max_bounces=8, K=5); o.set_scene(sc); o.accumulate(5, threads=1)
    out["default_320x192_mb8_K5_acc5"] = frame_record(o, sc)
    nodes, prims, ids = o.bvh()
    out["default_bvh"] = {"prim_ids": [int(v) for v in ids], "first_id": [int(v) for v in nodes["first_id"]], "prim_count": [int(v) for v in nodes["prim_count"]],
                          "sha256_nodes": hashlib.sha256(nodes.tobytes()).hexdigest()}
    rc, img = o.render()
    out["default_320x192_render_sha256"] = hashlib.sha256(img.tobytes()).hexdigest()
    o2 = oracle_py.Oracle(160, 96, max_bounces=16, K=1); o2.set_scene(sc); o2.set_accumulations(63); o2.accumulate(1, threads=1)
    out["default_160x96_mb16_K1_acc64"] = frame_record(o2, sc)
    sc3 = scenes.random_scene(2000, light_every=50)
    o3 = oracle_py.Oracle(160, 96, max_bounces=8, K=1); o3.set_scene(sc3); o3.accumulate(1, threads=1)
    out["random2000_160x96_mb8_K1_acc1"] = frame_record(o3, sc3)
    n3, p3, i3 = o3.bvh()
    out["random2000_bvh"] = {"sha256_nodes": hashlib.sha256(n3.tobytes()).hexdigest(), "sha256_prim_ids": hashlib.sha256(i3.tobytes()).hexdigest(),
                             "scene_sha256": hashlib.sha256(sc3["geometry"].tobytes()).hexdigest()}
    json.dump(out, open(os.path.join(HERE, "golden", "oracle_frames.json"), "w"), indent=0)


if __name__ == "__main__":
    gen_rng(); gen_sampling(); gen_camera_move(); gen_bvh(); gen_bvh_heuristics(); gen_renderer(); gen_renderer_ggx(); gen_survey(); gen_frames()
    print("golden vectors written")
