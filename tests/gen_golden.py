"""Regenerates tests/golden/*.json. Run in the build container (needs /root/reference for the reference-derived part):

    python tests/gen_golden.py

* rng_kat.json    — outputs of the reference's OWN Random.hpp/Bitmanip.hpp, compiled verbatim into oracle/_ref/librefrng.so
                    (oracle/Makefile `ref`): hash_u32, hash_2d, pcg streams, rand_bounded_int, make_unit_float, bitreverse.
                    These pin the oracle's RNG layer to the reference itself.
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
    gen_rng(); gen_survey(); gen_frames()
    print("golden vectors written")
