"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel of both pipelines, tiny sizes.
    compute-sanitizer --tool memcheck  python tests/sanitizer_workload.py
    compute-sanitizer --tool racecheck python tests/sanitizer_workload.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cpu-raytracing-experiments_b200")]
import numpy as np  # noqa: E402
import b2r  # noqa: E402
import scenes  # noqa: E402

r = b2r.Renderer(scenes.default_scene(), 64, 48, max_bounces=6, buckets=5, samples_in_flight=3)
r.Accumulate(5); assert r.Render()
r.Resize(48, 32); r.Accumulate(5); assert r.Render(); r.close()
for flags in (b2r.FLAG_FORCE_BVH, b2r.FLAG_FORCE_BRUTE, b2r.FLAG_FORCE_BVH | b2r.FLAG_REFERENCE_TREE | b2r.FLAG_COUNT_TESTS | b2r.FLAG_NO_GRAPH):
    r = b2r.Renderer(scenes.bvh_test_scene(300), 64, 32, max_bounces=5, buckets=2, flags=flags, samples_in_flight=2)
    r.Accumulate(4); assert r.Render()
    rays = np.random.RandomState(0).randn(256, 6).astype(np.float32)
    r.trace_closest(rays); r.trace_shadow(rays, np.full(256, 50.0, np.float32)); r.generate_rays(3)
    print(flags, r.counters())
    r.close()
print("sanitizer workload done")
