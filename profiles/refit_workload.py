#!/usr/bin/env python
"""Workload for the ncu launch list of the scene-edit path: C3's scene (N=100000) or C4's (N=1000000), one frame, then every sphere
moves by up to a quarter of its radius and b2r_refit_scene re-links and refits the traversal tree on the GPU (k_refit_level per tree
level + k_tree_cost), then one more frame. Prints host wall-clock times (never under ncu for a reported number)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpu-raytracing-experiments_b200"))
import numpy as np
import b2r, scenes

n = int(os.environ.get("N", "100000"))
sc = scenes.random_scene(n)
t0 = time.perf_counter(); ps = b2r.PreparedScene(sc, 640, 368); t1 = time.perf_counter()
r = b2r.Renderer(ps, 640, 368, max_bounces=8, buckets=1, samples_in_flight=1); r.sync(); t2 = time.perf_counter()
r.Accumulate(1); r.sync()
rs = np.random.RandomState(5); geo = ps.geometry.copy()
geo["position"] += (rs.uniform(-1, 1, (n, 3)) * (0.25 * np.sqrt(geo["radius_sq"]))[:, None]).astype(np.float32)
t3 = time.perf_counter(); r.RefitScene(geo, want_quality=False, keep_order=True); r.sync(); t4 = time.perf_counter()
q = r.RefitScene(geo, keep_order=True)
t5 = time.perf_counter(); r.RefitScene(geo, want_quality=False); r.sync(); t6 = time.perf_counter()
r.ResetAccumulator(); r.Accumulate(1); r.sync()
wide, ms = r.wide_nodes()
print(f"n={n}: reference BVH build (host) {1e3 * (t1 - t0):.1f} ms, create + upload_scene (traversal tree build, flatten, H2D) {1e3 * (t2 - t1):.1f} ms, "
      f"refit (order kept) {1e3 * (t4 - t3):.2f} ms, refit after a host rebuild of the reference BVH {1e3 * (t6 - t5):.1f} ms, quality ratio {q:.4f}, "
      f"{len(wide)} wide nodes, worst-case stack {ms}")
r.close()
