#!/usr/bin/env python
"""Workload for the ncu launch list of the device tree builds: C3's scene (N=100000), b2r_upload_scene with B2R_FLAG_GPU_TREE (packed tree)
with B2R_FLAG_GPU_TREE | B2R_FLAG_GPU_SAH (curve sweep tree) and with B2R_FLAG_GPU_TREE | B2R_FLAG_GPU_SAH3 (three-axis sweep tree), b2r_create's upload + three more each, one small frame after each. Prints host
wall-clock times (never under ncu for a reported number). numpy only: no torch import."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpu-raytracing-experiments_b200"))
import b2r, scenes

n = int(os.environ.get("N", "100000"))
ps = b2r.PreparedScene(scenes.random_scene(n), 320, 192)
for name, flags in (("packed", b2r.FLAG_GPU_TREE), ("sweep", b2r.FLAG_GPU_TREE | b2r.FLAG_GPU_SAH), ("sweep3", b2r.FLAG_GPU_TREE | b2r.FLAG_GPU_SAH3)):
    r = b2r.Renderer(ps, 320, 192, max_bounces=4, buckets=1, flags=b2r.FLAG_FORCE_BVH | flags); r.sync()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r.SetScene(ps); r.sync(); ts.append((time.perf_counter() - t0) * 1e3)
    r.Accumulate(1); r.sync()
    wide, ms = r.wide_nodes()
    print(f"n={n} {name}: b2r_upload_scene {ts[0]:.2f} / {ts[1]:.2f} / {ts[2]:.2f} ms (host wall clock incl. stream sync), {len(wide)} wide nodes, worst-case stack {ms}")
    r.close()
