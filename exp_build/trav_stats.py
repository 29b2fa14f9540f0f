# scratch: traversal statistics of the BVH pipeline on the C3 scene (counts per ray, kernel times)
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpu-raytracing-experiments_b200"))
import b2r, scenes
n = int(os.environ.get("N", "100000")); spp = int(os.environ.get("SPP", "8"))
sc = scenes.random_scene(n)
for flags, tag in ((b2r.FLAG_COUNT_TESTS | b2r.FLAG_NO_GRAPH, "count"), (b2r.FLAG_NO_GRAPH, "plain")):
    r = b2r.Renderer(sc, 1920, 1088, max_bounces=16, buckets=8, flags=flags)
    r.Accumulate(spp); r.sync() if hasattr(r, "sync") else None
    r.reset_counters(); r.kernel_times()
    r.Accumulate(spp)
    kt = r.kernel_times(); c = r.counters()
    ext, sh = c["extension_rays"], c["shadow_rays"]
    print(tag, os.environ.get("B2R_LIB_PATH", "default").split("/")[-1], {k: round(v[0], 2) for k, v in kt.items() if v[1]},
          "box/ray %.2f sphere/ray %.2f" % (c.get("box_tests", 0) / max(ext + sh, 1), c.get("sphere_tests", 0) / max(ext + sh, 1)), "rays", ext, sh)
    r.close()
