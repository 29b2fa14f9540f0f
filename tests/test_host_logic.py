"""Host-side product code (libb2r.so, no GPU needed): BVH builder == oracle bit-for-bit, 128-byte flattening round-trips,
light list, camera set-up, argument checking."""
import ctypes as C

import numpy as np
import pytest

import b2r
import oracle_py
import scenes


@pytest.mark.parametrize("n", [1, 2, 9, 100, 5000])
def test_bvh_bit_exact_vs_oracle(n):
    sc = scenes.default_scene() if n == 9 else scenes.random_scene(n, light_every=10)
    nodes, prims, ids = b2r.build_bvh(sc["geometry"])
    o = oracle_py.Oracle(16, 16); o.set_scene(sc)
    on, op, oi = o.bvh()
    assert nodes.tobytes() == on.tobytes()      # node order + boxes, bit-exact (north_star parity gate)
    assert prims.tobytes() == op.tobytes()      # leaf order
    assert np.array_equal(ids, oi)


def test_bvh_ties_are_stable():
    """Equal centroids (default scene: spheres 0/4 share x and z) — ties keep the lower index (Q19 canonical rule)."""
    g = np.zeros(8, scenes.SPHERE_DTYPE); g["radius_sq"] = 1.0; g["position"][:, 1] = np.arange(8) % 2
    nodes, prims, ids = b2r.build_bvh(g)
    o = oracle_py.Oracle(16, 16)
    sc = scenes.Scene(geometry=g, material=scenes.default_scene()["material"], camera=scenes.default_scene()["camera"], ambient=(0, 0, 0), hdri=None)
    o.set_scene(sc)
    assert nodes.tobytes() == o.bvh()[0].tobytes() and np.array_equal(ids, o.bvh()[2])


def test_bvh_builder_on_all_cores_equals_the_serial_algorithm():
    """From 20 000 spheres on b2r_bvh_build fans out: the three centroid sorts and the passes over a large node run side by side, and
    subtrees are built by a thread pool into node ranges fixed by the subtree sizes. The oracle's builder is the serial restatement of
    BVH.hpp:90-206: same nodes, same leaf order — also with heavily tied centroids (quantised positions, equal radii), where the
    merge of the three independent axis sweeps has to reproduce the reference's first-strict-minimum rule."""
    rs = np.random.RandomState(8)
    g = np.ascontiguousarray(scenes.random_scene(30000, light_every=10)["geometry"])
    tied = g.copy(); tied["position"][:, :2] = np.floor(rs.uniform(-100, 100, (len(g), 2)) / 10).astype(np.float32); tied["radius_sq"] = 1.0
    for geo in (g, tied):
        nodes, prims, ids = b2r.build_bvh(geo)
        sc = scenes.Scene(geometry=geo, material=scenes.random_scene(16)["material"], camera=scenes.default_scene()["camera"], ambient=(0, 0, 0), hdri=None)
        o = oracle_py.Oracle(16, 16); o.set_scene(sc)
        on, op, oi = o.bvh()
        assert nodes.tobytes() == on.tobytes() and np.array_equal(ids, oi)
        again = b2r.build_bvh(geo)
        assert again[0].tobytes() == nodes.tobytes() and np.array_equal(again[2], ids)     # schedule-independent


def test_lights_and_camera_match_oracle():
    for sc in (scenes.default_scene(), scenes.random_scene(3000)):
        o = oracle_py.Oracle(1280, 720); o.set_scene(sc)
        assert np.array_equal(b2r.find_lights(sc["geometry"], sc["material"]), o.lights())
        cam = b2r.camera_lookat(sc["camera"]["eye"], sc["camera"]["dir"], 1280, 720, sc["camera"]["focal_length"], 1.0)
        assert cam.tobytes() == o.camera_raw().tobytes()
    rs = np.random.RandomState(5)
    L = oracle_py.lib()
    for d in rs.randn(200, 3).astype(np.float32):  # every quat_cast branch
        q = np.zeros(4, np.float32); L.orc_quat_look_at(d.ctypes.data_as(C.POINTER(C.c_float)), q.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(b2r.camera_lookat((0, 0, 0), d, 16, 16, 50.0)[3:7], q)


def test_abi_argument_errors():
    L = b2r.lib()
    assert L.b2r_bvh_build(None, 3, None, None, None, None) == b2r.ERR_ARG
    h = C.c_void_p(None)
    for w, hgt, k in [(100, 64, 5), (64, 0, 5), (64, 64, 0), (64, 64, 65)]:
        cfg = b2r.Config(w, hgt, 16, k, 0, 0, 0, 0, 0)
        assert L.b2r_create(C.byref(h), C.byref(cfg)) == b2r.ERR_ARG and not h.value
    assert L.b2r_accumulate(None, 1) == b2r.ERR_ARG and b"null" in L.b2r_last_error()


def _read_hdr(path):
    """Minimal Radiance RGBE reader (RLE scanlines) for the test."""
    data = open(path, "rb").read()
    head, _, rest = data.partition(b"\n\n-Y ")
    assert head.startswith(b"#?RADIANCE") and b"FORMAT=32-bit_rle_rgbe" in head
    dims, _, body = rest.partition(b"\n")
    h, _, w = dims.partition(b" +X "); h, w = int(h), int(w)
    out = np.zeros((h, w, 4), np.uint8); pos = 0
    for y in range(h):
        if 8 <= w < 32768:
            assert body[pos] == 2 and body[pos + 1] == 2 and (body[pos + 2] << 8 | body[pos + 3]) == w; pos += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = body[pos]; pos += 1
                    if n > 128:
                        out[y, x:x + n - 128, c] = body[pos]; pos += 1; x += n - 128
                    else:
                        out[y, x:x + n, c] = np.frombuffer(body[pos:pos + n], np.uint8); pos += n; x += n
        else:
            out[y] = np.frombuffer(body[pos:pos + 4 * w], np.uint8).reshape(w, 4); pos += 4 * w
    assert pos == len(body)
    scale = np.ldexp(1.0, out[..., 3].astype(np.int32) - 136)
    return out[..., :3] * scale[..., None] * (out[..., 3:] > 0)


@pytest.mark.parametrize("w,h", [(5, 3), (64, 17), (333, 9)])
def test_hdr_writer_roundtrip(tmp_path, w, h):
    """Image::Store (Image.cpp:71-74): RGBE precision is 8 mantissa bits of the largest channel; rows are flipped."""
    rs = np.random.RandomState(w)
    img = np.zeros((h, w, 4), np.float32); img[..., :3] = rs.rand(h, w, 3) ** 3 * 5; img[..., 3] = 1
    img[0, : w // 2, :3] = 0.25; img[1:, -1, :3] = 0  # runs and exact zeros
    b2r.write_hdr(tmp_path / "f.hdr", img)
    back = _read_hdr(tmp_path / "f.hdr")[::-1]
    tol = img[..., :3].max(axis=2, keepdims=True) / 128 + 1e-30
    assert np.all(np.abs(back - img[..., :3]) <= tol)
    assert np.array_equal(back[0, : w // 2], img[0, : w // 2, :3])


# ------------------------------------------------------------------------------------------------ sky texture input (.hdr reader)
def _hdr_file(path, w, h, body, magic=b"#?RADIANCE", fmt=b"FORMAT=32-bit_rle_rgbe", extra=b"# a comment\nEXPOSURE=1.0\n", res=None):
    res = res if res is not None else b"-Y %d +X %d" % (h, w)
    open(path, "wb").write(magic + b"\n" + extra + fmt + b"\n\n" + res + b"\n" + bytes(body))
    return path


@pytest.mark.parametrize("w,h", [(5, 3), (64, 17), (333, 9)])
def test_hdr_reader_equals_stb_restatement(tmp_path, w, h):
    """b2r_read_hdr = stbi_loadf(path, ..., 4) as the reference loads its sky (Application.cpp:225-231): the file written by the
    library's own Radiance writer (flat below 8 pixels per row, RLE above) decodes to mantissa * 2^(e-136), alpha 1, rows in file
    order — bit for bit the values of the independent Python decoder of this test file."""
    rs = np.random.RandomState(w + 1)
    img = np.zeros((h, w, 4), np.float32); img[..., :3] = rs.rand(h, w, 3) ** 4 * 50; img[..., 3] = 1
    img[0, : w // 2, :3] = 0.25; img[1:, -1, :3] = 0
    b2r.write_hdr(tmp_path / "f.hdr", img)
    got = b2r.read_hdr(tmp_path / "f.hdr")
    assert got.shape == (h, w, 4) and got.dtype == np.float32 and np.all(got[..., 3] == 1.0)
    want = _read_hdr(tmp_path / "f.hdr").astype(np.float32)
    assert got[..., :3].tobytes() == want.tobytes()
    assert np.array_equal(got[::-1][0, : w // 2, :3], img[0, : w // 2, :3])   # the writer flips rows (Image.cpp:72), the reader does not


def test_hdr_reader_known_answers_and_old_format(tmp_path):
    # one RGBE quadruple by hand: (128, 64, 32 | e=129) -> 128 * 2^-7 = 1, 0.5, 0.25; e = 0 -> black whatever the mantissas
    px = [128, 64, 32, 129, 200, 100, 50, 0, 255, 255, 255, 136 + 3, 1, 2, 3, 128]
    got = b2r.read_hdr(_hdr_file(tmp_path / "a.hdr", 4, 1, px, magic=b"#?RGBE"))
    assert got.tolist() == [[[1.0, 0.5, 0.25, 1.0], [0.0, 0.0, 0.0, 1.0], [2040.0, 2040.0, 2040.0, 1.0], [1 / 256, 2 / 256, 3 / 256, 1.0]]]
    # a wide image stored WITHOUT run-length encoding (old Radiance files): the first scanline does not start with {2, 2, <0x80}, stb
    # switches to flat pixels for the whole file ("yes, this makes no sense")
    rs = np.random.RandomState(4); w, h = 12, 3
    raw = rs.randint(0, 256, (h, w, 4)).astype(np.uint8); raw[..., 0] |= 128; raw[0, 0, :2] = (7, 9)
    got = b2r.read_hdr(_hdr_file(tmp_path / "b.hdr", w, h, raw.tobytes()))
    want = (raw[..., :3] * np.ldexp(1.0, raw[..., 3].astype(np.int32) - 136)[..., None] * (raw[..., 3:] > 0)).astype(np.float32)
    assert got[..., :3].tobytes() == want.tobytes()
    # hand-made RLE scanline, w = 8: R = run of 8 x 64; G = literals; B = run 3 + literals 5; E = run of 8 x 128
    line = [2, 2, 0, 8, 128 + 8, 64, 8, 1, 2, 3, 4, 5, 6, 7, 8, 128 + 3, 9, 5, 10, 11, 12, 13, 14, 128 + 8, 128]
    got = b2r.read_hdr(_hdr_file(tmp_path / "c.hdr", 8, 1, line))
    assert got[0, :, 0].tolist() == [0.25] * 8 and got[0, :, 1].tolist() == [k / 256 for k in range(1, 9)]
    assert got[0, :, 2].tolist() == [9 / 256] * 3 + [k / 256 for k in range(10, 15)]


def test_hdr_reader_rejects_what_stb_rejects(tmp_path):
    import ctypes as C
    ok_line = [2, 2, 0, 8] + [128 + 8, 1] * 4
    w, h = C.c_int32(0), C.c_int32(0)
    def rc(path, out=None):
        return b2r.lib().b2r_read_hdr(str(path).encode(), out, C.byref(w), C.byref(h))
    assert rc(_hdr_file(tmp_path / "ok.hdr", 8, 1, ok_line)) == b2r.OK and (w.value, h.value) == (8, 1)
    buf = np.zeros((1, 8, 4), np.float32); p = C.c_void_p(buf.ctypes.data)
    assert rc(tmp_path / "ok.hdr", p) == b2r.OK
    assert rc(tmp_path / "missing.hdr") == b2r.ERR_ARG
    assert rc(_hdr_file(tmp_path / "m.hdr", 8, 1, ok_line, magic=b"#?RADIANC")) == b2r.ERR_ARG                 # "not HDR"
    assert rc(_hdr_file(tmp_path / "f.hdr", 8, 1, ok_line, fmt=b"FORMAT=32-bit_rle_xyze")) == b2r.ERR_ARG      # "unsupported format"
    assert rc(_hdr_file(tmp_path / "r.hdr", 8, 1, ok_line, res=b"+Y 1 +X 8")) == b2r.ERR_ARG                   # "unsupported data layout"
    assert rc(_hdr_file(tmp_path / "x.hdr", 8, 1, ok_line, res=b"-Y 1 -X 8")) == b2r.ERR_ARG
    assert rc(_hdr_file(tmp_path / "l.hdr", 8, 1, [2, 2, 0, 9] + [128 + 8, 1] * 4), p) == b2r.ERR_ARG           # "invalid decoded scanline length"
    assert rc(_hdr_file(tmp_path / "c.hdr", 8, 1, [2, 2, 0, 8, 128 + 9, 1]), p) == b2r.ERR_ARG                  # run longer than the row: "corrupt"
    assert rc(_hdr_file(tmp_path / "z.hdr", 8, 1, [2, 2, 0, 8, 0]), p) == b2r.ERR_ARG                           # zero count
    two = np.zeros((2, 8, 4), np.float32)
    assert rc(_hdr_file(tmp_path / "t.hdr", 8, 2, ok_line), C.c_void_p(two.ctypes.data)) == b2r.ERR_ARG        # second scanline missing
    big = np.zeros((2, 3, 4), np.float32)
    assert rc(_hdr_file(tmp_path / "s.hdr", 3, 2, [1] * 20), C.c_void_p(big.ctypes.data)) == b2r.ERR_ARG        # flat file four bytes short


def test_hdr_reader_fuzz_against_an_independent_encoder(tmp_path):
    """Random RGBE images encoded by a small encoder of this test (random mixture of runs of 1..127 and literal spans of 1..128, per
    component, as the format allows — not only the run/literal choices b2r_write_hdr makes), decoded by b2r_read_hdr."""
    rs = np.random.RandomState(31)
    for trial in range(25):
        w, h = int(rs.randint(8, 300)), int(rs.randint(1, 6))
        rgbe = rs.randint(0, 256, (h, w, 4)).astype(np.uint8)
        for y in range(h):                                            # plant long constant stretches so that runs are worth encoding
            for c in range(4):
                x = 0
                while x < w:
                    n = int(rs.randint(1, 60))
                    if rs.rand() < 0.5: rgbe[y, x:x + n, c] = rgbe[y, x, c]
                    x += n
        rgbe[..., 3][rs.rand(h, w) < 0.05] = 0                        # some black pixels (e = 0)
        body = bytearray()
        for y in range(h):
            body += bytes([2, 2, w >> 8, w & 255])
            for c in range(4):
                row = rgbe[y, :, c]; x = 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and row[x + run] == row[x]: run += 1
                    if run >= 2 and rs.rand() < 0.8:
                        n = int(rs.randint(1, run + 1)); body += bytes([128 + n, int(row[x])]); x += n
                    else:
                        n = int(rs.randint(1, min(128, w - x) + 1)); body += bytes([n]) + row[x:x + n].tobytes(); x += n
        got = b2r.read_hdr(_hdr_file(tmp_path / f"z{trial}.hdr", w, h, body))
        want = (rgbe[..., :3] * np.ldexp(1.0, rgbe[..., 3].astype(np.int32) - 136)[..., None] * (rgbe[..., 3:] > 0)).astype(np.float32)
        assert got[..., :3].tobytes() == want.tobytes() and np.all(got[..., 3] == 1.0)


def test_hdr_writer_bytes_follow_stb_image_write(tmp_path):
    """Image::Store = stbi_write_hdr (Image.cpp:71-74; stb_image_write.h is un-vendored third-party code, restated from its published
    source): apart from the "# Written by" comment line the file is what stb writes, byte for byte — frexp-based RGBE conversion with
    truncation, {2, 2, w_hi, w_lo} scanline headers, literals up to the first triple of equal bytes (<= 128 at a time), runs <= 127 at
    a time with one- and two-byte remainders still written as runs. Checked against a second restatement in this test."""
    rs = np.random.RandomState(12)
    w, h = 300, 4
    img = np.zeros((h, w, 4), np.float32); img[..., :3] = rs.rand(h, w, 3) ** 3 * 9; img[..., 3] = 1
    img[0, 10:150, :] = img[0, 10, :]        # a run of 140: 127 + 13
    img[1, 20:148, 1] = 0.5                   # a run of 128 in one component: 127 + a run of 1
    img[2, :, :3] = 0.0                       # black row: all four components constant
    b2r.write_hdr(tmp_path / "s.hdr", img)
    data = open(tmp_path / "s.hdr", "rb").read()
    head = b"#?RADIANCE\n# Written by libb2r\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y 4 +X 300\n"
    assert data.startswith(head)
    want = bytearray()
    for row in img[::-1]:                     # stbi_flip_vertically_on_write(true)
        rgbe = np.zeros((w, 4), np.uint8)
        for x in range(w):
            r, g, b = (np.float32(v) for v in row[x, :3]); m = max(r, g, b)
            if m >= np.float32(1e-32):
                mant, e = np.frexp(m); n = np.float32(np.float32(mant) * np.float32(256.0) / m)
                rgbe[x] = (int(r * n), int(g * n), int(b * n), e + 128)
        want += bytes([2, 2, w >> 8, w & 255])
        for c in range(4):
            comp = rgbe[:, c].tolist(); x = 0
            while x < w:
                r_ = x
                while r_ + 2 < w and not (comp[r_] == comp[r_ + 1] == comp[r_ + 2]): r_ += 1
                if r_ + 2 >= w: r_ = w
                while x < r_:
                    n = min(128, r_ - x); want += bytes([n] + comp[x:x + n]); x += n
                if r_ + 2 < w:
                    while r_ < w and comp[r_] == comp[x]: r_ += 1
                    while x < r_:
                        n = min(127, r_ - x); want += bytes([128 + n, comp[x]]); x += n
    assert data[len(head):] == bytes(want)
    assert np.array_equal(b2r.read_hdr(tmp_path / "s.hdr")[::-1][2, :, :3], np.zeros((w, 3), np.float32))


def test_node_array_validation_rejects_malformed_trees(hostcheck):
    """b2r_upload_scene flattens a CALLER's node array (B2R_FLAG_REFERENCE_TREE): validate_reference_bvh must refuse anything the
    breadth-first flattening would loop on or render wrongly — cycles, shared children, a sphere in two leaves, a missing sphere."""
    def ok(nodes, n):
        a = np.ascontiguousarray(nodes)
        return hostcheck.hc_validate(C.c_void_p(a.ctypes.data), len(a), n) == 1
    for n in (1, 2, 9, 300):
        sc = scenes.default_scene() if n == 9 else scenes.random_scene(n, light_every=10)
        nodes, prims, _ = b2r.build_bvh(sc["geometry"])
        assert ok(nodes, n)                                             # the reference builder's own output
        if n < 3: continue
        inner = np.flatnonzero(nodes["prim_count"] == 0); leaves = np.flatnonzero(nodes["prim_count"] != 0)
        bad = nodes.copy(); bad["first_id"][inner[-1]] = 0              # child pair before its parent: a cycle through the root
        assert not ok(bad, n)
        bad = nodes.copy(); bad["first_id"][inner[1]] = nodes["first_id"][inner[0]]   # two parents share one child pair
        assert not ok(bad, n)
        bad = nodes.copy(); bad["first_id"][leaves[0]] = nodes["first_id"][leaves[1]]  # one sphere in two leaves (another one in none)
        assert not ok(bad, n)
        bad = nodes.copy(); bad["prim_count"][leaves[0]] = 2            # multi-sphere leaf
        assert not ok(bad, n)
        bad = nodes.copy(); bad["first_id"][inner[0]] = len(nodes) - 1  # second child out of range
        assert not ok(bad, n)
        assert not ok(nodes[:-1], n) and not ok(nodes, n + 1)
