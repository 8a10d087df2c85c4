"""CPU tests of the C++ host (volren_b200/host): the `volpy` module surface of reference src/bindings.cpp, the host's
file formats and math against the Python mirrors / numpy, and the CLI's no-device behaviour. No kernel is launched."""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "volren_b200")
ASSETS = os.path.join(ROOT, "tests", "golden", "assets")


@pytest.fixture(scope="module")
def volpy():
    from volren_b200 import build
    build.build_cuda()
    build.build_host()
    sys.path.insert(0, PKG)
    try:
        import volpy as m
    finally:
        sys.path.remove(PKG)
    return m


def test_module_surface_matches_reference_bindings(volpy):
    # classes of bindings.cpp:64-417
    for name in ["ImageDataFloat", "Volume", "Environment", "TransferFunction", "Renderer", "vec2", "vec3", "vec4", "ivec2", "ivec3", "ivec4",
                 "uvec2", "uvec3", "uvec4", "mat3", "mat4", "quat"]:
        assert hasattr(volpy, name), name
    R = volpy.Renderer
    for name in ["init", "commit", "trace", "reset", "scale_and_move_to_unit_cube", "render", "draw", "resolution", "fbo_data", "save", "save_with_alpha",
                 "volume", "environment", "transferfunc", "sample", "sppx", "bounces", "seed", "tonemap_exposure", "tonemap_gamma", "tonemapping",
                 "show_environment", "albedo", "phase", "density_scale", "emission_scale", "vol_clip_min", "vol_clip_max",
                 "cam_pos", "cam_dir", "cam_up", "cam_fov", "cam_near", "cam_far", "view_matrix", "proj_matrix", "cam_aspect",
                 "colmap_view_trans", "colmap_view_rot", "colmap_focal_length", "shutdown"]:
        assert hasattr(R, name), name
    for name in ["load_grid", "clear", "add_grid_frame", "update_grid_frame", "AABB", "grid_frame_counter", "minorant_majorant"]:
        assert hasattr(volpy.Volume, name), name
    for name in ["randomize", "window_left", "window_width"]:
        assert hasattr(volpy.TransferFunction, name), name
    assert hasattr(volpy.Environment, "strength")


def test_renderer_defaults_match_reference_header(volpy):
    r = volpy.Renderer()    # renderer.h:31-44 (no device needed before init())
    assert (r.sample, r.sppx, r.seed, r.bounces) == (0, 1024, 42, 100)
    assert r.tonemap_exposure == 5.0 and abs(r.tonemap_gamma - 2.2) < 1e-6 and r.tonemapping and r.show_environment
    assert np.allclose(np.array(r.albedo), 0.9) and r.phase == 0.0 and r.density_scale == 1.0 and r.emission_scale == 100.0
    assert np.array_equal(np.array(r.vol_clip_min), [0, 0, 0]) and np.array_equal(np.array(r.vol_clip_max), [1, 1, 1])
    assert r.volume is None and r.environment is None and r.transferfunc is None
    # camera defaults (cppgl camera.cpp:42-46); static properties are shared by all instances (bound by address)
    assert r.cam_fov == 70.0 and abs(r.cam_near - 0.01) < 1e-9 and r.cam_far == 1000.0
    r.cam_pos = volpy.vec3(1, 2, 3)
    assert np.array_equal(np.array(volpy.Renderer().cam_pos), [1, 2, 3])
    r.cam_pos = volpy.vec3(0, 0, 0)


def test_vector_arithmetic_against_numpy(volpy):
    rng = np.random.default_rng(0)
    a, b = rng.random(3).astype(np.float32), rng.random(3).astype(np.float32) + 0.5
    va, vb = volpy.vec3(*a), volpy.vec3(*b)
    s = np.float32(1.7)
    for got, want in [(va + vb, a + b), (va - vb, a - b), (va * vb, a * b), (va / vb, a / b), (va + float(s), a + s), (float(s) - va, s - a),
                      (va * float(s), a * s), (float(s) / vb, s / b), (-va, -a)]:
        assert np.allclose(np.array(got), want, rtol=1e-6)
    assert abs(va.length() - np.linalg.norm(a)) < 1e-6
    assert np.allclose(np.array(vb.normalize()), b / np.linalg.norm(b), rtol=1e-6)
    v = volpy.vec3(1, 2, 3)
    v += volpy.vec3(1, 1, 1)
    v *= 2.0
    assert np.array_equal(np.array(v), [4, 6, 8])
    assert repr(volpy.vec3(1, 2, 3)) == "vec3(1.000000, 2.000000, 3.000000)" and repr(volpy.ivec2(1, -2)) == "ivec2(1, -2)"
    assert repr(volpy.uvec3(7) * volpy.uvec3(2)) == "uvec3(14, 14, 14)"
    assert np.array(volpy.vec4(1, 2, 3, 4)).shape == (4,)
    v2 = volpy.vec2(3, 4)
    assert v2.length() == 5.0


def test_matrix_and_quaternion_layout(volpy):
    m = volpy.mat4(volpy.vec4(1, 2, 3, 4), volpy.vec4(5, 6, 7, 8), volpy.vec4(9, 10, 11, 12), volpy.vec4(13, 14, 15, 16))
    a = np.array(m)                       # numpy rows = glm columns
    assert np.array_equal(a[1], [5, 6, 7, 8]) and m.value(2, 1) == 10.0 and np.array_equal(np.array(m.column(3)), [13, 14, 15, 16])
    i = volpy.mat4(1.0)
    assert np.array_equal(np.array(m * i), a) and np.array_equal(np.array(m + m), 2 * a) and np.array_equal(np.array(m * 2.0), 2 * a)
    # matrix product follows column-major math: (A B) x = A (B x)
    A = np.arange(9, dtype=np.float32).reshape(3, 3) + 1
    B = (np.arange(9, dtype=np.float32).reshape(3, 3) * 0.5 - 1)
    mA = volpy.mat3(*[volpy.vec3(*A[:, c]) for c in range(3)])
    mB = volpy.mat3(*[volpy.vec3(*B[:, c]) for c in range(3)])
    assert np.allclose(np.array(mA * mB).T, A @ B)
    q = volpy.quat(volpy.mat3(1.0))
    assert np.array_equal(np.array(q), [0, 0, 0, 1]) and (q.x, q.y, q.z, q.w) == (0, 0, 0, 1)   # memory order x, y, z, w
    # rotation about y by 90 degrees
    R = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], np.float32)
    qy = np.array(volpy.quat(volpy.mat3(*[volpy.vec3(*R[:, c]) for c in range(3)])))
    assert np.allclose(qy, [0, np.sqrt(.5), 0, np.sqrt(.5)], atol=1e-6)


def test_transfer_function_cdf_matches_python_mirror(volpy, lut_raw):
    from volren_b200 import formats
    tf = volpy.TransferFunction(os.path.join(ASSETS, "lut.txt"))
    got = np.array([np.array(v) for v in tf.lut], np.float32)
    assert np.array_equal(got, lut_raw)            # the raw LUT stays untouched on the host object
    assert tf.window_left == 0.0 and tf.window_width == 1.0
    tf2 = volpy.TransferFunction([volpy.vec4(0, 0, 0, 0), volpy.vec4(1, 0, 0, .5), volpy.vec4(0, 1, 0, .25)])
    assert len(tf2.lut) == 3
    tf2.randomize(5)
    assert len(tf2.lut) == 5 and np.array_equal(np.array(tf2.lut[0]), [0, 0, 0, 0])
    assert formats.lut_for_upload(lut_raw).shape == (8, 4)


def test_hdr_loader_matches_python_loader(volpy, env_rgb):
    got = volpy.load_hdr(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"))
    assert got.shape == env_rgb.shape and np.array_equal(got, env_rgb)
    top_down = volpy.load_hdr(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"), False)
    assert np.array_equal(top_down[::-1], env_rgb)


def test_colmap_helpers_against_numpy(volpy):
    """colmap_view_trans / colmap_view_rot (bindings.cpp:196-203): GL_TO_COLMAP * view with view = lookAt(pos, pos + dir, up)
    (cppgl camera.cpp:51-59), checked against an independent numpy / scipy evaluation for random poses."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(4)
    R = volpy.Renderer
    for _ in range(20):
        pos = rng.standard_normal(3).astype(np.float32) * 2
        d = rng.standard_normal(3).astype(np.float32)
        d /= np.linalg.norm(d)
        if abs(d[1]) > 0.95:
            continue
        R.cam_pos, R.cam_dir, R.cam_up = volpy.vec3(*map(float, pos)), volpy.vec3(*map(float, d)), volpy.vec3(0, 1, 0)
        R.update_camera()
        f = d.astype(np.float64)
        s_ = np.cross(f, [0, 1, 0]); s_ /= np.linalg.norm(s_)
        u = np.cross(s_, f)
        view = np.eye(4)
        view[0, :3], view[1, :3], view[2, :3] = s_, u, -f
        view[:3, 3] = [-s_ @ pos, -u @ pos, f @ pos]
        assert np.allclose(np.array(R.view_matrix).T, view, atol=2e-5)          # the binding exposes glm columns as rows
        M = np.diag([1.0, -1.0, -1.0, 1.0]) @ view
        assert np.allclose(np.array(R.colmap_view_trans()), M[:3, 3], atol=2e-5)
        q = np.array(R.colmap_view_rot())                                        # buffer order x, y, z, w
        want = Rotation.from_matrix(M[:3, :3]).as_quat()
        assert abs(np.linalg.norm(q) - 1) < 1e-5 and min(np.abs(q - want).max(), np.abs(q + want).max()) < 5e-5
    R.cam_pos, R.cam_dir = volpy.vec3(1, 0, 1), (-volpy.vec3(1, 0, 1)).normalize()
    R.update_camera()


def test_ldr_png_environment_maps(volpy, tmp_path):
    """Environment(path) accepts LDR .png files like the reference (cppgl uploads them as GL_R8 / RG8 / RGB8 / RGBA8: the
    shader samples u8 / 255, no gamma; image_load flips to bottom-up rows). Every PNG flavour an encoder produces here:
    gray, gray+alpha, RGB, RGBA, 16-bit, palette, and all five scanline filters (cv2 / PIL choose them adaptively)."""
    import cv2
    from PIL import Image
    rng = np.random.default_rng(9)
    H, W = 37, 53
    smooth = (np.linspace(0, 255, W)[None, :, None] * np.ones((H, 1, 4)) * 0.5 + rng.integers(0, 128, (H, W, 4))).astype(np.uint8)
    cases = {"rgb": smooth[..., :3], "rgba": smooth, "gray": smooth[..., 0], "ga": None, "rgb16": None, "pal": None}
    for name in cases:
        p = str(tmp_path / f"{name}.png")
        if name == "ga":
            Image.fromarray(smooth[..., :2], "LA").save(p)
            want = np.zeros((H, W, 3), np.float32); want[..., 0] = smooth[..., 0] / np.float32(255); want[..., 1] = smooth[..., 1] / np.float32(255)
        elif name == "rgb16":
            img16 = (smooth[..., :3].astype(np.uint16) << 8) | rng.integers(0, 256, (H, W, 3)).astype(np.uint16)
            cv2.imwrite(p, img16[..., ::-1])
            want = (img16 >> 8).astype(np.float32) / np.float32(255)
        elif name == "pal":
            pimg = Image.fromarray(smooth[..., :3], "RGB").quantize(32)
            pimg.save(p)
            want = np.asarray(pimg.convert("RGB"), np.float32) / np.float32(255)
        elif name == "gray":
            cv2.imwrite(p, cases[name])
            want = np.zeros((H, W, 3), np.float32); want[..., 0] = cases[name] / np.float32(255)
        elif name == "rgba":
            cv2.imwrite(p, cases[name][..., [2, 1, 0, 3]])
            want = cases[name][..., :3].astype(np.float32) / np.float32(255)
        else:
            cv2.imwrite(p, cases[name][..., ::-1])
            want = cases[name].astype(np.float32) / np.float32(255)
        got = volpy.load_environment_image(p)
        assert got.shape == (H, W, 3) and np.array_equal(got, want[::-1]), name        # bottom-up rows
    env = volpy.Environment(str(tmp_path / "rgb.png"))                                    # and through the class the scripts use
    assert env.strength == 1.0
    with pytest.raises(RuntimeError):
        volpy.Environment(str(tmp_path / "missing.png"))
    (tmp_path / "x.jpg").write_bytes(b"\xff\xd8\xff")
    with pytest.raises(RuntimeError, match="hdr and .png"):
        volpy.Environment(str(tmp_path / "x.jpg"))


def _read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h, ct = 8, b"", 0, 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body)
        if tag == b"IHDR":
            w, h, depth, ct = struct.unpack(">IIBB", body[:10])
            assert depth == 8
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    ch = {0: 1, 4: 2, 2: 3, 6: 4}[ct]
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
    assert np.all(raw[:, 0] == 0)
    return raw[:, 1:].reshape(h, w, ch)


def test_png_writer_roundtrip_and_flip(volpy, tmp_path):
    rng = np.random.default_rng(1)
    for ch in (3, 4):
        img = rng.integers(0, 256, (5, 7, ch), dtype=np.uint8)
        p = str(tmp_path / f"t{ch}.png")
        volpy.save_ldr(p, img)                   # flipped on write by default (cppgl image_store_ldr)
        assert np.array_equal(_read_png(p), img[::-1])
        volpy.save_ldr(p, img, False)
        assert np.array_equal(_read_png(p), img)
        import cv2                               # an independent decoder accepts the file
        dec = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        assert dec is not None and dec.shape == img.shape


def test_brick_and_dense_files_through_the_host(volpy, smoke_grid, tmp_path):
    vol = volpy.Volume(os.path.join(ASSETS, "smoke.brick"))
    mn, mj = vol.minorant_majorant()
    assert (mn, mj) == smoke_grid.min_maj
    bb_min, bb_max = vol.AABB("density")
    M = smoke_grid.matrix()
    ext = np.array([*smoke_grid.index_extent(), 1], np.float32)
    assert np.allclose(np.array(bb_min), (M @ np.array([0, 0, 0, 1], np.float32))[:3], rtol=1e-6)
    assert np.allclose(np.array(bb_max), (M @ ext)[:3], rtol=1e-6)
    text = repr(vol)
    assert "brick dim: uvec3(16, 32, 16)" in text and "bricks in atlas: 3297" in text and "atlas dim: uvec3(128, 256, 56)" in text
    # CPU decode (BrickGrid::lookup) against the vectorised Python decode
    g = volpy.Volume.load_grid(os.path.join(ASSETS, "smoke.brick"))
    dec = smoke_grid.decode_all()
    rng = np.random.default_rng(2)
    for _ in range(200):
        x, y, z = (int(rng.integers(0, n)) for n in smoke_grid.index_extent())
        assert g.lookup(volpy.uvec3(x, y, z)) == dec[z, y, x]
    # .dense written by the Python mirror is read by the host
    from volren_b200 import formats
    vox = rng.integers(0, 256, (6, 5, 4), dtype=np.uint8)
    formats.save_dense(str(tmp_path / "t.dense"), formats.DenseGridData(vox, -1.0, 3.0))
    d = volpy.Volume.load_grid(str(tmp_path / "t.dense"))
    assert repr(d.index_extent()) == "uvec3(4, 5, 6)" and d.minorant_majorant() == (-1.0, 3.0)
    assert d.lookup(volpy.uvec3(1, 2, 3)) == np.float32(-1.0) + (np.float32(vox[3, 2, 1]) / np.float32(255)) * np.float32(4.0)
    assert d.lookup(volpy.uvec3(4, 0, 0)) == 0.0        # out of bounds -> 0 (grid_dense.cpp:100)
    with pytest.raises(RuntimeError):
        volpy.Volume(str(tmp_path / "missing.brick"))
    with pytest.raises(RuntimeError):
        volpy.Volume.load_grid(str(tmp_path / "x.vdb"))


def test_dat_loader_and_folder_order(volpy, tmp_path):
    vox = (np.arange(2 * 3 * 4) * 9 % 256).astype(np.uint8)
    (tmp_path / "scan.raw").write_bytes(vox.tobytes())
    (tmp_path / "scan.dat").write_text("ObjectFileName: scan.raw\nResolution: 4 3 2\nSliceThickness: 1 1 2\nFormat: UCHAR\nBitsUsed: 8\n")
    g = volpy.Volume.load_grid(str(tmp_path / "scan.dat"))
    assert repr(g.index_extent()) == "uvec3(4, 3, 2)" and g.minorant_majorant() == (0.0, 1.0)
    assert g.lookup(volpy.uvec3(3, 2, 1)) == np.float32(vox[23]) / np.float32(255)
    T = np.array(g.transform).T        # z-up -> y-up rotation by 270 deg about x, then slice thickness
    assert np.allclose(T[:3, :3], np.array([[1, 0, 0], [0, 0, 2], [0, -1, 0]], np.float32), atol=1e-6)
    # animation folders: shorter names first, then lexicographic (volume.cpp:263-270)
    from volren_b200 import formats
    seq = tmp_path / "seq"
    seq.mkdir()
    for i in (10, 2, 1):
        formats.save_dense(str(seq / f"f{i}.dense"), formats.DenseGridData(np.full((2, 2, 2), i, np.uint8), 0.0, float(i)))
    vol = volpy.Volume.load_folder(str(seq), ["density"])
    assert vol.n_grid_frames() == 3
    majs = []
    for i in range(3):
        vol.grid_frame_counter = i
        majs.append(vol.minorant_majorant()[1])
    assert majs == [1.0, 2.0, 10.0]


def test_cli_fails_loudly_without_a_device(volpy, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([os.path.join(PKG, "volren"), os.path.join(ASSETS, "smoke.brick"), "-w", "32", "-h", "32", "--render", "--spp", "1"],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "no usable CUDA device" in r.stderr
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        volpy.Renderer().init()


def test_colormaps_equal_tinycolormap_bit_for_bit(volpy):
    """TransferFunction::colormap (transferfunc.cpp:69-77): the host's tables / closed forms and the Python mirror against LUTs
    produced by the reference's own tinycolormap header (tests/golden/colormap_golden.npz, make_golden.py): all 14 types."""
    from volren_b200 import colormaps
    gold = np.load(os.path.join(ROOT, "tests", "golden", "colormap_golden.npz"))
    for t, name in enumerate(colormaps.TYPES):
        want = gold[f"type{t}_256"]
        got_host = np.array([np.array(v) for v in volpy._colormap_lut(t, 256)], np.float32)
        assert np.array_equal(got_host.view(np.uint32), want.view(np.uint32)), name
        assert np.array_equal(colormaps.colormap_lut(name, 256).view(np.uint32), want.view(np.uint32)), name
    for t in (3, 9):       # Turbo, Viridis: the two CLI flags, at bin counts that do not hit the table entries
        for n in (7, 1000):
            want = gold[f"type{t}_{n}"]
            got_host = np.array([np.array(v) for v in volpy._colormap_lut(t, n)], np.float32)
            assert np.array_equal(got_host.view(np.uint32), want.view(np.uint32))
            assert np.array_equal(colormaps.colormap_lut(colormaps.TYPES[t], n).view(np.uint32), want.view(np.uint32))
    with pytest.raises(Exception):
        volpy._colormap_lut(99, 4)


def test_files_written_by_the_reference_serialisation(volpy, golden):
    """.dense / .brick files written by the reference's own cereal path (serialization.cpp:36-43, 66-80; make_golden.py through
    oracle/_ref) are read by the C++ host and by the Python formats module."""
    from volren_b200 import formats
    d = os.path.join(ROOT, "tests", "golden", "ref_written")
    exp = np.load(os.path.join(d, "expected.npz"))
    vox, (lo, hi) = exp["dense_vox"], exp["dense_minmax"]
    g = volpy.Volume.load_grid(os.path.join(d, "ref_7x5x6.dense"))
    assert repr(g.index_extent()) == "uvec3(7, 5, 6)" and g.minorant_majorant() == (float(lo), float(hi))
    for z in range(6):
        for y in range(5):
            for x in range(7):
                assert g.lookup(volpy.uvec3(x, y, z)) == np.float32(lo) + (np.float32(vox[z, y, x]) / np.float32(255)) * (np.float32(hi) - np.float32(lo))
    pd = formats.load_dense(os.path.join(d, "ref_7x5x6.dense"))
    assert np.array_equal(pd.voxels, vox) and (pd.min_value, pd.max_value) == (float(lo), float(hi))
    name = "ragged_70x33x20"
    pb = formats.load_brick(os.path.join(d, "ref_ragged_70x33x20.brick"))
    assert tuple(pb.n_bricks) == tuple(golden[name + ".n_bricks"]) and np.array_equal(pb.indirection, golden[name + ".indirection"])
    assert np.array_equal(pb.range, golden[name + ".range"])
    for i in range(3):
        assert np.array_equal(pb.mips[i], golden[name + f".mip{i}"])
    hb = volpy.Volume.load_grid(os.path.join(d, "ref_ragged_70x33x20.brick"))
    dec = pb.decode_all()
    rng = np.random.default_rng(4)
    for _ in range(300):
        x, y, z = int(rng.integers(0, 70)), int(rng.integers(0, 33)), int(rng.integers(0, 20))
        assert hb.lookup(volpy.uvec3(x, y, z)) == dec[z, y, x]


def test_range_xy_kernel_on_the_cpu(tmp_path):
    """volren_b200/csrc/vr_brick_range.cuh (x/y part of the brick build's 12^3 window min/max, u16x2 packed) uses no shuffles and
    no shared memory: tests/cpu_harness/range_xy_host.cpp compiles the SAME source as plain C++ and checks every (z, by, bx)
    entry against a brute-force loop -- ragged widths, padding columns / rows of n_bricks, several y-bands, dense and sparse data."""
    import subprocess
    exe = tmp_path / "range_xy_host"
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_harness", "range_xy_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), src], check=True)
    res = subprocess.run([str(exe), "11"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "FAIL" not in res.stdout, res.stdout[-2000:]


def test_fast_encode_expression_on_the_cpu(tmp_path):
    """encode_code_fast (vr_brick.cuh: division by the brick's correctly rounded reciprocal + two fma corrections, trunc-based
    round) is the brick build's table entry; tests/cpu_harness/encode_fast_host.c evaluates the same IEEE operations in C against
    the reference's plain expression (grid_brick.cpp:45-48) on random fp16 ranges, incl. collapsed (span == 0) ones."""
    import subprocess
    exe = tmp_path / "encode_fast_host"
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_harness", "encode_fast_host.c")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), src, "-lm"], check=True)
    res = subprocess.run([str(exe), "20000000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "bad 0" in res.stdout, res.stdout[-2000:]
