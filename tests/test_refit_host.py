"""Scene edit without a rebuild (b2r_refit_scene, cf. Application.cpp:508-510), CPU tier: the shared routine `refit_slot`
(csrc/b2r_shade.h — the code k_refit_level runs per slot) through its host twin `refit_wide`, compiled by tests/hostcheck.

* identity: refitting with unchanged spheres reproduces flatten_bvh's array bit for bit (flatten_bvh fills its boxes with the same routine);
* containment: after moving spheres every inner slot's box contains every sphere below it, leaves hold the moved spheres;
* semantics: closest hit through the refitted tree == brute force over the moved spheres in the same order (BVH.hpp:265 ties);
* the quality ratio is 1 on identity and grows when spheres are scattered.
The GPU tier (tests/test_gpu_parity.py::test_refit_*) checks the kernel against this twin and against a fresh upload."""
import ctypes as C
import os

import numpy as np
import pytest

import b2r
import scenes

EMPTY = -2 ** 31


def vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def refit(hc, prims_a, prims_b, nodes=None, rays=None, remap=None):
    n = len(prims_a); nw = C.c_uint32(0); cost = (C.c_double * 2)()
    nn = 0 if nodes is None else len(nodes)
    hc.hc_refit(vp(prims_a), vp(nodes), nn, vp(prims_b), vp(remap), n, None, C.byref(nw), cost, None, 0, None, None, None)
    wide = np.zeros((nw.value, 4, 8), np.float32)
    nr = 0 if rays is None else len(rays)
    tfar = np.zeros(nr, np.float32); prim = np.zeros(nr, np.int32)
    hc.hc_refit(vp(prims_a), vp(nodes), nn, vp(prims_b), vp(remap), n, vp(wide), C.byref(nw), cost, vp(rays), nr, vp(tfar), vp(prim), None)
    return wide, (cost[0], cost[1]), tfar, prim


def moved(prims, rs, jitter=0.5, far=0):
    """Every sphere moves by up to `jitter` radii and changes radius by up to 20 %; `far` of them jump anywhere in the scene."""
    out = prims.copy()
    r = np.sqrt(out["radius_sq"])
    out["position"] += (rs.uniform(-1, 1, (len(out), 3)) * (jitter * r)[:, None]).astype(np.float32)
    out["radius_sq"] = ((r * rs.uniform(0.8, 1.2, len(out))) ** 2).astype(np.float32)
    if far:
        idx = rs.choice(len(out), far, replace=False)
        lo, hi = prims["position"].min(0), prims["position"].max(0)
        out["position"][idx] = rs.uniform(lo, hi, (far, 3)).astype(np.float32)
    return out


def camera_rays(n, rs, prims):
    lo, hi = prims["position"].min(0), prims["position"].max(0)
    o = rs.uniform(lo - 5, hi + 5, (n, 3)); t = rs.uniform(lo, hi, (n, 3)); d = t - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    return np.ascontiguousarray(np.concatenate([o, d], 1), np.float32)


def origin_box(prims, points=None):
    """The library's rule (origin_box_rule, csrc/b2r_host.cpp): sphere bounds + points, an eighth of the extent + 1 of slack per side."""
    r = np.sqrt(prims["radius_sq"].astype(np.float64))[:, None]; c = prims["position"].astype(np.float64)
    lo, hi = (c - r).min(0), (c + r).max(0)
    if points is not None and len(points):
        lo = np.minimum(lo, np.asarray(points, np.float64).min(0)); hi = np.maximum(hi, np.asarray(points, np.float64).max(0))
    slack = 0.125 * (hi - lo) + 1.0
    return lo - slack, hi + slack


def check_contains(wide, prims, points=None):
    """Slot layout {c.xyz, r2}{h.x, h.y, link, h.z}. Asserts: every leaf slot holds its sphere and a cube extent H that covers
    sqrt(r^2 + kHitNoise * D^2) for the farthest origin-box corner D (and is no looser than that plus the padding); every inner slot box
    contains the union of its child node's slot boxes and is tight around it; every sphere appears exactly once."""
    seen = []
    olo, ohi = origin_box(prims, points)

    def walk(w):
        lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
        for k in range(4):
            link = int(wide[w, k, 6:7].view(np.int32)[0])
            if link == EMPTY: continue
            c = wide[w, k, 0:3].astype(np.float64); h = wide[w, k, [4, 5, 7]].astype(np.float64)
            if link < 0:
                pr = ~link; seen.append(pr)
                assert np.array_equal(wide[w, k, :3], prims["position"][pr]) and wide[w, k, 3] == prims["radius_sq"][pr]
                assert h[0] == h[1] == h[2]
                far2 = (np.maximum(np.abs(c - olo), np.abs(ohi - c)) ** 2).sum()
                need = np.sqrt(float(prims["radius_sq"][pr]) + 1.5e-6 * far2)
                assert h[0] >= need * (1 - 1e-6) and h[0] <= need * (1 + 2e-4) + (np.abs(c).max() + need + 1) * 1e-6
            else:
                clo, chi = walk(link)
                blo, bhi = c - h, c + h
                assert np.all(blo <= clo) and np.all(bhi >= chi)                         # conservative
                assert np.all(clo - blo <= (np.abs(c) + h + 1) * 3e-6) and np.all(bhi - chi <= (np.abs(c) + h + 1) * 3e-6)  # and tight
            lo = np.minimum(lo, c - h); hi = np.maximum(hi, c + h)
        return lo, hi
    walk(0)
    assert sorted(seen) == list(range(len(prims)))


@pytest.mark.parametrize("n", [1, 2, 9, 40, 777, 5000])
def test_refit_identity_is_bit_exact(hostcheck, n):
    sc = scenes.default_scene() if n == 9 else scenes.random_scene(n, light_every=5)
    nodes, prims, _ = b2r.build_bvh(sc["geometry"])
    base, cost, _, _ = refit(hostcheck, prims, prims)
    # what flatten_bvh itself stores
    if n > 1:
        from test_hostcheck import flatten
        ref_flat = flatten(nodes, prims)
        ref_refit, cost_r, _, _ = refit(hostcheck, prims, prims, nodes=nodes)
        assert ref_flat.tobytes() == ref_refit.tobytes()      # reference topology: refit(identity) == flatten, bit for bit
        assert cost_r[0] == cost_r[1]
    assert cost[0] == cost[1]
    untouched, _, _, _ = refit(hostcheck, prims, None)
    assert base.tobytes() == untouched.tobytes()              # traversal-tree topology (libb2r's default): the same identity
    check_contains(base, prims)


@pytest.mark.parametrize("n,far", [(9, 0), (40, 3), (777, 0), (777, 50), (5000, 100)])
def test_refit_moved_spheres_boxes_and_closest_hit(hostcheck, n, far):
    rs = np.random.RandomState(1000 + n + far)
    sc = scenes.default_scene() if n == 9 else scenes.random_scene(n, light_every=5)
    _, prims, _ = b2r.build_bvh(sc["geometry"])
    new = moved(prims, rs, far=far)
    rays = camera_rays(3000, rs, prims)
    wide, cost, tfar, prim = refit(hostcheck, prims, new, rays=rays)
    check_contains(wide, new, rays[:, :3])
    bt = np.zeros(len(rays), np.float32); bp = np.zeros(len(rays), np.int32)
    hostcheck.hc_closest_brute(vp(new), len(new), vp(rays), len(rays), vp(bt), vp(bp))
    assert (bp >= 0).mean() > 0.2
    assert np.array_equal(bp, prim) and bt.tobytes() == tfar.tobytes()
    if far:
        assert cost[1] > cost[0]          # scattered spheres make the kept topology worse, and the ratio says so
    # topology untouched: same links everywhere
    base, _, _, _ = refit(hostcheck, prims, prims)
    assert np.array_equal(base[:, :, 6].view(np.int32), wide[:, :, 6].view(np.int32))


@pytest.mark.parametrize("n,far", [(9, 0), (777, 50), (5000, 0)])
def test_refit_into_a_rebuilt_bvh_order(hostcheck, n, far):
    """The reference re-sorts its prims on every rebuild (BVH.hpp:201-205). Refit into the NEW order of the moved scene: leaf links are
    remapped old index -> geometry index -> new index (match_prims_to_geometry pairs prims with geometry by value, as b2r_refit_scene
    does), and the closest hit then equals brute force over the new prims array, indices included."""
    rs = np.random.RandomState(2000 + n + far)
    sc = scenes.default_scene() if n == 9 else scenes.random_scene(n, light_every=5)
    geo = np.ascontiguousarray(sc["geometry"])
    _, prims, ids = b2r.build_bvh(geo)
    geo2 = moved(geo, rs, far=far)
    _, prims2, ids2 = b2r.build_bvh(geo2)                 # what the app does after an edit
    # value matching recovers the builder's own index map
    m = np.zeros(n, np.uint32)
    assert hostcheck.hc_match(vp(prims), vp(geo), n, vp(m)) == 1 and np.array_equal(m, ids)
    m2 = np.zeros(n, np.uint32)
    assert hostcheck.hc_match(vp(prims2), vp(geo2), n, vp(m2)) == 1 and np.array_equal(m2, ids2)
    assert hostcheck.hc_match(vp(prims2), vp(geo), n, vp(m2)) == 0      # not a permutation of each other
    prim_of_geom2 = np.zeros(n, np.uint32); prim_of_geom2[ids2] = np.arange(n, dtype=np.uint32)
    remap = np.ascontiguousarray(prim_of_geom2[ids])
    rays = camera_rays(3000, rs, prims)
    wide, cost, tfar, prim = refit(hostcheck, prims, prims2, rays=rays, remap=remap)
    check_contains(wide, prims2, rays[:, :3])
    bt = np.zeros(len(rays), np.float32); bp = np.zeros(len(rays), np.int32)
    hostcheck.hc_closest_brute(vp(prims2), n, vp(rays), len(rays), vp(bt), vp(bp))
    assert np.array_equal(bp, prim) and bt.tobytes() == tfar.tobytes()


def test_match_pairs_equal_spheres(hostcheck):
    sc = scenes.random_scene(64, light_every=5)
    geo = np.ascontiguousarray(sc["geometry"]); geo[10] = geo[3]; geo[40] = geo[3]      # three identical spheres
    perm = np.random.RandomState(3).permutation(64)
    prims = np.ascontiguousarray(geo[perm]); m = np.zeros(64, np.uint32)
    assert hostcheck.hc_match(vp(prims), vp(geo), 64, vp(m)) == 1
    assert sorted(m.tolist()) == list(range(64))
    for f in ("position", "radius_sq", "material_ID"):
        assert np.array_equal(geo[m][f], prims[f])


def test_refit_fuzz_duplicates_and_random_orders(hostcheck):
    """Randomised: scenes with duplicated spheres, a random new leaf order, spheres moved; value matching + remap + refit must still give
    brute-force hits over the new order (indices included, modulo which of two identical twins a tie names) and a valid tree."""
    rs = np.random.RandomState(99)
    for trial in range(12):
        n = int(rs.randint(2, 400))
        geo = np.ascontiguousarray(scenes.random_scene(max(n, 2), light_every=5, seed=1000 + trial)["geometry"])
        for _ in range(int(rs.randint(0, 4))):                       # identical twins
            a, b = rs.randint(0, n, 2); geo[a] = geo[b]
        _, prims, ids = b2r.build_bvh(geo)
        geo2 = moved(geo, rs, far=int(rs.randint(0, 3)))
        for a in range(n):                                            # twins stay twins after the move
            for b in range(a):
                if all(np.array_equal(geo[a][k], geo[b][k]) for k in ("position", "radius_sq", "material_ID")): geo2[a] = geo2[b]
        perm = rs.permutation(n).astype(np.uint32)                    # any order, not just the builder's
        prims2 = np.ascontiguousarray(geo2[perm])
        m_old = np.zeros(n, np.uint32); m_new = np.zeros(n, np.uint32)
        assert hostcheck.hc_match(vp(prims), vp(geo), n, vp(m_old)) == 1 and hostcheck.hc_match(vp(prims2), vp(geo2), n, vp(m_new)) == 1
        assert sorted(m_old.tolist()) == list(range(n)) and sorted(m_new.tolist()) == list(range(n))
        prim_of_geom = np.zeros(n, np.uint32); prim_of_geom[m_new] = np.arange(n, dtype=np.uint32)
        remap = np.ascontiguousarray(prim_of_geom[m_old])
        rays = camera_rays(400, rs, prims)
        wide, _, tfar, prim = refit(hostcheck, prims, prims2, rays=rays, remap=remap)
        check_contains(wide, prims2, rays[:, :3])
        bt = np.zeros(len(rays), np.float32); bp = np.zeros(len(rays), np.int32)
        hostcheck.hc_closest_brute(vp(prims2), n, vp(rays), len(rays), vp(bt), vp(bp))
        assert bt.tobytes() == tfar.tobytes() and np.array_equal(bp, prim)


def test_match_is_linear_in_the_number_of_twins(hostcheck):
    """50 000 copies of one sphere plus 50 000 distinct ones: pairing by value stays linear (twins are chained off one table slot and
    handed out in index order) and rejects an array with one copy too many."""
    import time
    n = 100000
    geo = np.ascontiguousarray(scenes.random_scene(n, light_every=50)["geometry"]); geo[::2] = geo[0]
    perm = np.random.RandomState(1).permutation(n)
    prims = np.ascontiguousarray(geo[perm]); m = np.zeros(n, np.uint32)
    t0 = time.perf_counter()
    assert hostcheck.hc_match(vp(prims), vp(geo), n, vp(m)) == 1
    assert time.perf_counter() - t0 < 2.0
    assert sorted(m.tolist()) == list(range(n))
    for f in ("position", "radius_sq", "material_ID"):
        assert np.array_equal(geo[m][f], prims[f])
    twins = np.flatnonzero(perm % 2 == 0)                                  # positions in prims that hold the repeated sphere
    assert np.array_equal(m[twins], np.arange(0, n, 2))                    # handed out first in, first out
    prims[np.flatnonzero(perm % 2 == 1)[0]] = geo[0]                       # one twin too many, one distinct sphere missing
    assert hostcheck.hc_match(vp(prims), vp(geo), n, vp(m)) == 0


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 17, 300, 5000])
def test_packed_morton_tree_host_twin(hostcheck, n):
    """The tree b2r_upload_scene builds on the GPU with B2R_FLAG_GPU_TREE, through its host twin (build_packed_tree): every sphere in exactly
    one leaf, boxes conservative and tight (check_contains), and the closest hit through it equals brute force, indices included."""
    rs = np.random.RandomState(n)
    sc = scenes.random_scene(max(n, 2), light_every=5)
    _, prims, _ = b2r.build_bvh(sc["geometry"][:n])
    nw = C.c_uint32(0); ms = C.c_uint32(0)
    hostcheck.hc_packed_tree(vp(prims), n, None, None, C.byref(nw), C.byref(ms))
    wide = np.zeros((nw.value, 4, 8), np.float32)
    hostcheck.hc_packed_tree(vp(prims), n, None, vp(wide), C.byref(nw), C.byref(ms))
    assert ms.value + 3 <= 64
    check_contains(wide, prims)
    rays = camera_rays(2000, rs, prims) if n > 1 else np.ascontiguousarray(np.concatenate([prims["position"][[0] * 50] + rs.uniform(3, 9, (50, 3)), -np.ones((50, 3)) / np.sqrt(3)], 1), np.float32)
    m = len(rays); st = np.zeros(m, np.uint32); bx = np.zeros(m, np.uint32); sp = np.zeros(m, np.uint32); pr = np.zeros(m, np.int32)
    hostcheck.hc_trace_stats(None, 0xffffffff, vp(prims), n, vp(rays), m, vp(st), vp(bx), vp(sp), vp(pr))
    bt = np.zeros(m, np.float32); bp = np.zeros(m, np.int32)
    hostcheck.hc_closest_brute(vp(prims), n, vp(rays), m, vp(bt), vp(bp))
    assert np.array_equal(bp, pr)


def test_packed_tree_sort_key_is_a_hilbert_curve(hostcheck):
    """The key build_packed_tree / k_morton_keys sort by (b2r_shade.h: morton_key) is the Hilbert index of the sphere centre's cell on a
    1024^3 grid: a bijection on aligned blocks whose consecutive cells are face neighbours — so a node of the packed tree, a run of
    consecutive keys, is a connected set of cells (a Z curve's runs are not: 58 -> 28 node visits per ray on the 100k-sphere scene)."""
    def keys_of(cells):
        p = np.zeros(len(cells), scenes.SPHERE_DTYPE); p["position"] = cells.astype(np.float32) + 0.5
        lo = np.zeros(3, np.float32); hi = np.full(3, 1024, np.float32); k = np.zeros(len(cells), np.uint32)
        hostcheck.hc_curve_keys(vp(p), len(cells), vp(lo), vp(hi), vp(k)); return k
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3)
    k = keys_of(g << 7) >> 21                                               # the 8^3 coarse cells: the leading 9 key bits
    assert sorted(k.tolist()) == list(range(512))
    assert (np.abs(np.diff(g[np.argsort(k)], axis=0)).sum(1) == 1).all()
    for off in ([0, 0, 0], [512, 96, 992], [224, 608, 960]):                # whole 32^3 blocks anywhere on the grid
        g = np.stack(np.meshgrid(*[np.arange(32)] * 3, indexing="ij"), -1).reshape(-1, 3) + np.array(off)
        k = keys_of(g)
        assert len(set(k.tolist())) == 32 ** 3 and int(k.max()) - int(k.min()) == 32 ** 3 - 1
        assert (np.abs(np.diff(g[np.argsort(k)], axis=0)).sum(1) == 1).all()


@pytest.mark.parametrize("three", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 17, 300, 5000])
def test_sweep_tree_host_twin(hostcheck, n, three):
    """The trees b2r_upload_scene builds on the GPU with B2R_FLAG_GPU_TREE | B2R_FLAG_GPU_SAH (curve sweep) or B2R_FLAG_GPU_SAH3 (three-axis
    sweep), through their host twins (build_sweep_tree / build_sweep3_tree): every sphere in exactly one leaf, boxes conservative and tight,
    the closest hit through it equals brute force (indices included) — and they are the better trees: fewer node visits than the packed one
    on the same rays, the three-axis sweep fewer than the curve sweep."""
    build = hostcheck.hc_sweep3_tree if three else hostcheck.hc_sweep_tree; mode = 0xfffffffd if three else 0xfffffffe
    rs = np.random.RandomState(n)
    sc = scenes.random_scene(max(n, 2), light_every=5)
    _, prims, _ = b2r.build_bvh(sc["geometry"][:n])
    nw = C.c_uint32(0); ms = C.c_uint32(0)
    assert build(vp(prims), n, None, None, C.byref(nw), C.byref(ms)) == 0
    wide = np.zeros((nw.value, 4, 8), np.float32)
    assert build(vp(prims), n, None, vp(wide), C.byref(nw), C.byref(ms)) == 0
    assert ms.value + 3 <= 64
    check_contains(wide, prims)
    rays = camera_rays(2000, rs, prims) if n > 1 else np.ascontiguousarray(np.concatenate([prims["position"][[0] * 50] + rs.uniform(3, 9, (50, 3)), -np.ones((50, 3)) / np.sqrt(3)], 1), np.float32)
    m = len(rays); st = np.zeros(m, np.uint32); bx = np.zeros(m, np.uint32); sp = np.zeros(m, np.uint32); pr = np.zeros(m, np.int32)
    assert hostcheck.hc_trace_stats(None, mode, vp(prims), n, vp(rays), m, vp(st), vp(bx), vp(sp), vp(pr)) == 0
    bt = np.zeros(m, np.float32); bp = np.zeros(m, np.int32)
    hostcheck.hc_closest_brute(vp(prims), n, vp(rays), m, vp(bt), vp(bp))
    assert np.array_equal(bp, pr)
    if n >= 300:
        st2 = np.zeros(m, np.uint32)
        hostcheck.hc_trace_stats(None, 0xfffffffe if three else 0xffffffff, vp(prims), n, vp(rays), m, vp(st2), vp(bx), vp(sp), vp(pr))
        assert st.sum() < st2.sum()


def test_sweep_tree_gives_up_on_a_tree_deeper_than_the_stack(hostcheck):
    """Concentric spheres whose radii grow by half from one to the next: the cheapest cut always peels the largest one off, so the sweep tree
    would be dozens of levels deep; build_sweep_tree says so and the library falls back to the balanced packed tree."""
    n = 100
    geo = np.zeros(n, scenes.SPHERE_DTYPE); geo["radius_sq"] = ((1.5 ** np.arange(n)) ** 2).astype(np.float32); geo["position"][:, 0] = 0.3
    nw = C.c_uint32(0); ms = C.c_uint32(0)
    assert hostcheck.hc_sweep_tree(vp(geo), n, None, None, C.byref(nw), C.byref(ms)) == 1
    hostcheck.hc_packed_tree(vp(geo), n, None, None, C.byref(nw), C.byref(ms))
    assert ms.value + 3 <= 64
    # the same with copies of one sphere: every cut of the run costs the same and the first one is taken
    d = scenes.default_scene()["geometry"]; geo = np.ascontiguousarray(np.concatenate([d, np.repeat(d[3:4], 100)]), dtype=scenes.SPHERE_DTYPE)
    assert hostcheck.hc_sweep_tree(vp(geo), len(geo), None, None, C.byref(nw), C.byref(ms)) == 1
    assert hostcheck.hc_sweep_tree(vp(np.ascontiguousarray(geo[:40])), 40, None, None, C.byref(nw), C.byref(ms)) == 0 and ms.value + 3 <= 64


def _odd_sphere_sets(rs):
    """Sphere sets a scene editor can produce and a curve order dislikes: clusters, a line, coincident centres, a shell around everything,
    points (radius 0), all in one cell, twins."""
    def arr(pos, r):
        g = np.zeros(len(pos), scenes.SPHERE_DTYPE); g["position"] = np.asarray(pos, np.float32); g["radius_sq"] = (np.asarray(r, np.float32) ** 2); return g
    n = 600
    c = rs.uniform(-50, 50, (6, 3)); yield "clusters", arr(c[rs.randint(0, 6, n)] + rs.normal(0, 0.4, (n, 3)), rs.uniform(0.01, 0.3, n))
    t = np.linspace(0, 1, n)[:, None]; yield "line", arr(t * np.array([30.0, 20.0, -10.0]), rs.uniform(0.01, 0.2, n))
    yield "coincident centres", arr(np.zeros((40, 3)) + 0.5, np.linspace(0.1, 3.0, 40))
    yield "shell", arr(np.concatenate([rs.uniform(-5, 5, (n - 1, 3)), [[0, 0, 0]]]), np.concatenate([rs.uniform(0.05, 0.5, n - 1), [40.0]]))
    yield "points", arr(rs.uniform(-5, 5, (n, 3)), np.where(rs.uniform(size=n) < 0.5, 0.0, 0.2))
    yield "one cell", arr(1000.0 + rs.uniform(0, 1e-3, (50, 3)), rs.uniform(1e-4, 2e-4, 50))
    g = arr(rs.uniform(-5, 5, (n // 2, 3)), rs.uniform(0.05, 0.5, n // 2)); yield "twins", np.concatenate([g, g[: n // 4], g])
    yield "plane", arr(np.concatenate([rs.uniform(-20, 20, (n, 2)), np.zeros((n, 1))], 1), rs.uniform(0.05, 0.6, n))


@pytest.mark.parametrize("builder", ["packed", "sweep", "sweep3"])
def test_device_tree_twins_on_awkward_sphere_sets(hostcheck, builder):
    """Both trees the GPU builds by itself, through their host twins, on sphere sets that stress a curve order (see _odd_sphere_sets): the tree
    is valid (every sphere once, boxes conservative and tight), fits the traversal stack, and the closest hit through it equals brute force,
    indices included; where the sweep tree would be too deep the twin says so (the library then links the packed tree)."""
    rs = np.random.RandomState(11)
    for name, prims in _odd_sphere_sets(rs):
        prims = np.ascontiguousarray(prims, dtype=scenes.SPHERE_DTYPE); n = len(prims)
        nw = C.c_uint32(0); ms = C.c_uint32(0)
        build = {"sweep": hostcheck.hc_sweep_tree, "sweep3": hostcheck.hc_sweep3_tree, "packed": hostcheck.hc_packed_tree}[builder]
        rc = build(vp(prims), n, None, None, C.byref(nw), C.byref(ms))
        if rc != 0:
            assert builder != "packed", name      # too deep: only the sweep trees can be
            continue
        wide = np.zeros((nw.value, 4, 8), np.float32)
        build(vp(prims), n, None, vp(wide), C.byref(nw), C.byref(ms))
        assert ms.value + 3 <= 64, name
        check_contains(wide, prims)
        rays = camera_rays(1500, rs, prims)
        m = len(rays); st = np.zeros(m, np.uint32); bx = np.zeros(m, np.uint32); sp = np.zeros(m, np.uint32); pr = np.zeros(m, np.int32)
        assert hostcheck.hc_trace_stats(None, {"sweep": 0xfffffffe, "sweep3": 0xfffffffd, "packed": 0xffffffff}[builder], vp(prims), n, vp(rays), m, vp(st), vp(bx), vp(sp), vp(pr)) == 0
        bt = np.zeros(m, np.float32); bp = np.zeros(m, np.int32)
        hostcheck.hc_closest_brute(vp(prims), n, vp(rays), m, vp(bt), vp(bp))
        assert np.array_equal(bp, pr), name


def test_device_tree_twins_against_the_committed_digests(hostcheck):
    """tests/golden/device_tree_digests.json (tests/gen_device_tree_digests.py): the node arrays of the three host twins for fixed scenes,
    hashed — the builders' arithmetic, tie rules and sort keys are pinned on a CPU-only box; the GPU tests pin the device to the twins."""
    import json
    import gen_device_tree_digests as gen
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "device_tree_digests.json")))
    assert gen.digests(hostcheck) == want
