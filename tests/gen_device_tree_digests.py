#!/usr/bin/env python
"""tests/golden/device_tree_digests.json: SHA-256 of the 128-byte node arrays the host twins of the three device tree builds produce
(build_packed_tree, build_sweep_tree, build_sweep3_tree through tests/hostcheck) for fixed scenes. The GPU tests compare the device trees
with the twins bit for bit; this fixture pins the twins themselves, so that a change of the builders' arithmetic, tie rules or sort keys
shows up on a CPU-only box. Regenerate with `python tests/gen_device_tree_digests.py` after a DELIBERATE change of a builder."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cpu-raytracing-experiments_b200")]
import b2r      # noqa: E402
import scenes   # noqa: E402

SIZES = (17, 700, 5000)
BUILDERS = ("packed", "sweep", "sweep3")


def digests(lib):
    out = {}
    for n in SIZES:
        sc = scenes.random_scene(n, light_every=5); _, prims, _ = b2r.build_bvh(sc["geometry"][:n])
        for name in BUILDERS:
            f = getattr(lib, f"hc_{name}_tree"); nw = C.c_uint32(0); ms = C.c_uint32(0)
            assert f(C.c_void_p(prims.ctypes.data), n, None, None, C.byref(nw), C.byref(ms)) == 0
            w = np.zeros((nw.value, 4, 8), np.float32)
            f(C.c_void_p(prims.ctypes.data), n, None, C.c_void_p(w.ctypes.data), C.byref(nw), C.byref(ms))
            out[f"{name}:{n}"] = {"nodes": nw.value, "max_stack": ms.value, "sha256": hashlib.sha256(w.tobytes()).hexdigest()}
    return out


if __name__ == "__main__":
    import __graft_entry__ as g
    g.build_hostcheck()
    d = digests(C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")))
    json.dump(d, open(os.path.join(ROOT, "tests", "golden", "device_tree_digests.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(d, indent=1, sort_keys=True))
