"""Device-built sweep tree (B2R_FLAG_GPU_TREE | B2R_FLAG_GPU_SAH) against its host twin and against the other trees — a numpy-only script
(no torch import: seconds on a GPU box), also run by tests/test_gpu_parity.py::test_gpu_sweep_tree_*. Prints one line per case.

    python tests/gpucheck/sweep_tree_check.py [--perf]     # --perf: upload and frame times of the three trees on the 100k-sphere scene
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cpu-raytracing-experiments_b200")]
import b2r      # noqa: E402
import scenes   # noqa: E402

HC = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
SAH = b2r.FLAG_FORCE_BVH | b2r.FLAG_GPU_TREE | b2r.FLAG_GPU_SAH
SAH3 = b2r.FLAG_FORCE_BVH | b2r.FLAG_GPU_TREE | b2r.FLAG_GPU_SAH3   # the three-axis sweep


def twin_nodes(prims, obox, three=False):
    """node count of the host twin's sweep tree (0: too deep)"""
    nw = C.c_uint32(0); m2 = C.c_uint32(0)
    return nw.value if (HC.hc_sweep3_tree if three else HC.hc_sweep_tree)(vp(prims), len(prims), vp(obox), None, C.byref(nw), C.byref(m2)) == 0 else 0


def check(n, w=160, h=96, spp=2, three=False):
    twin_of = HC.hc_sweep3_tree if three else HC.hc_sweep_tree
    sc = scenes.random_scene(max(n, 2), light_every=20); sc = dict(sc); sc["geometry"] = sc["geometry"][:n]
    r = b2r.Renderer(sc, w, h, max_bounces=6, buckets=2, flags=SAH3 if three else SAH); r.Accumulate(spp)
    wide, ms = r.wide_nodes(); obox = r.origin_box()
    prims = np.ascontiguousarray(r.scene.prims)
    nw = C.c_uint32(0); m2 = C.c_uint32(0)
    rc = twin_of(vp(prims), n, vp(obox), None, C.byref(nw), C.byref(m2))
    twin = np.zeros((nw.value, 4, 8), np.float32)
    if rc == 0:
        twin_of(vp(prims), n, vp(obox), vp(twin), C.byref(nw), C.byref(m2))
    else:  # too deep: the library falls back to the packed tree
        HC.hc_packed_tree(vp(prims), n, vp(obox), None, C.byref(nw), C.byref(m2)); twin = np.zeros((nw.value, 4, 8), np.float32)
        HC.hc_packed_tree(vp(prims), n, vp(obox), vp(twin), C.byref(nw), C.byref(m2))
    same_shape = nw.value == len(wide) and m2.value == ms
    same_links = same_shape and np.array_equal(wide[:, :, 6].view(np.uint32), twin[:, :, 6].view(np.uint32))
    same_tree = same_shape and wide.tobytes() == twin.tobytes()
    s = b2r.Renderer(sc, w, h, max_bounces=6, buckets=2, flags=b2r.FLAG_FORCE_BVH); s.Accumulate(spp)
    same_frame = r.buckets_host().tobytes() == s.buckets_host().tobytes()
    # a scene edit afterwards: the device-built tree is refitted (leaves matched to geometry lazily), the SAH-tree renderer gets a fresh upload
    rs = np.random.RandomState(n); geo2 = np.ascontiguousarray(r.scene.geometry).copy()
    geo2["position"] += (rs.uniform(-0.3, 0.3, (n, 3)) * np.sqrt(geo2["radius_sq"])[:, None]).astype(np.float32)
    r.RefitScene(geo2); r.ResetAccumulator(); r.Accumulate(spp)
    sc2 = dict(sc); sc2["geometry"] = geo2
    s.SetScene(sc2); s.ResetAccumulator(); s.Accumulate(spp)
    same_after_refit = r.buckets_host().tobytes() == s.buckets_host().tobytes()
    print(f"{'three-axis ' if three else ''}n={n}: nodes {len(wide)} (twin {nw.value}) max_stack {ms} (twin {m2.value}) links_equal={same_links} tree_equal={same_tree} frame_equal_sah_tree={same_frame} "
          f"frame_equal_after_refit={same_after_refit}", flush=True)
    r.close(); s.close()
    return same_tree and same_frame and same_after_refit


def perf(n=100000, w=1920, h=1088, spp=4):
    sc = scenes.random_scene(n); ps = b2r.PreparedScene(sc, w, h)
    out = {}
    for name, flags in (("host_sah", b2r.FLAG_FORCE_BVH), ("gpu_packed", b2r.FLAG_FORCE_BVH | b2r.FLAG_GPU_TREE), ("gpu_sweep", SAH), ("gpu_sweep3", SAH3)):
        r = b2r.Renderer(ps, w, h, max_bounces=16, buckets=8, flags=flags); r.Accumulate(spp); r.sync()
        t0 = time.perf_counter(); r.SetScene(ps); r.sync(); t1 = time.perf_counter()
        t2 = time.perf_counter(); r.SetScene(ps); r.sync(); t3 = time.perf_counter()
        r.ResetAccumulator(); r.Accumulate(spp); r.sync()
        t4 = time.perf_counter(); r.ResetAccumulator(); r.Accumulate(spp); r.sync(); t5 = time.perf_counter()
        out[name] = dict(upload_ms=(t3 - t2) * 1e3, first_upload_ms=(t1 - t0) * 1e3, frame_ms=(t5 - t4) * 1e3, nodes=len(r.wide_nodes()[0]), digest=hash(r.buckets_host().tobytes()))
        print(name, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in out[name].items()}, flush=True)
        r.close()
    assert out["host_sah"]["digest"] == out["gpu_packed"]["digest"] == out["gpu_sweep"]["digest"] == out["gpu_sweep3"]["digest"]
    return out


if __name__ == "__main__":
    sizes = (2, 3, 5, 6, 17, 700, 20000)
    ok = all([check(n, three=True) for n in sizes]) if "--three-only" in sys.argv else all([check(n) for n in sizes] + [check(n, three=True) for n in sizes])
    if "--big" in sys.argv:
        ok = check(1000000, 320, 192, 1) and check(1000000, 320, 192, 1, three=True) and ok
    if "--perf" in sys.argv:
        ok = check(100000, 320, 192, 1) and check(100000, 320, 192, 1, three=True) and ok
        perf()
    print("SWEEP_TREE_OK" if ok else "SWEEP_TREE_MISMATCH")
    sys.exit(0 if ok else 1)
