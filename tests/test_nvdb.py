"""NanoVDB ingestion: voldata::NanoVDBGrid (grid_nvdb.cpp) + BrickGrid(const Grid&) for non-dense
sources. The fixture tests/golden/nvdb_golden.npz holds a two-grid .nvdb file written by the reference's own NanoVDB
headers and what the UNMODIFIED reference adapter and brick constructor make of it (make_nvdb_golden.py).

CPU: the oracle restatements (oracle/nvdb_np.py, vro_brick_build_values) and the product's host-side reader/accessor
(vrb_nvdb_open / vrb_nvdb_lookup, no device needed) against the fixture. GPU: the device accessor + any-Grid brick build
through the C ABI, bit for bit."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN

NAMES = ("density", "temperature")
# + NanoVDB's own fog-volume sphere (tools/CreatePrimitives.h): interior stored as active constant tiles of the lower nodes
NAMES_ALL = NAMES + ("sphere",)


def _source(g, name):
    """(file bytes, grid name inside the file) of a golden entry."""
    return (g["sphere_file"], "density") if name == "sphere" else (g["nvdb_file"], name)


@pytest.fixture(scope="module")
def nvdb_golden():
    return np.load(os.path.join(GOLDEN, "nvdb_golden.npz"))


def _volpy():
    import sys
    pkg = os.path.join(os.path.dirname(GOLDEN), os.pardir, "volren_b200")
    pkg = os.path.abspath(pkg)
    if not os.path.exists(os.path.join(pkg, "volren")):
        pytest.fail("the C++ host is not built (python -m volren_b200.build --host)")
    sys.path.insert(0, pkg)
    try:
        import volpy as m
    finally:
        sys.path.remove(pkg)
    return m


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _golden_brick_equals(g, name, got):
    assert tuple(g[name + ".n_bricks"]) == got.n_bricks and tuple(g[name + ".atlas_dim"]) == got.atlas_dim
    assert int(g[name + ".brick_count"][0]) == got.brick_count
    assert np.array_equal(g[name + ".range"], got.range), "range"
    assert np.array_equal(g[name + ".indirection"], got.indirection), "indirection"
    assert np.array_equal(g[name + ".atlas"], got.atlas), "atlas"
    for i in range(3):
        assert np.array_equal(g[name + f".mip{i}"], got.mips[i]), f"mip{i}"


# ---- CPU: oracle pins ----------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", NAMES_ALL)
def test_oracle_nvdb_restatement_matches_reference(nvdb_golden, name):
    from oracle.nvdb_np import Grid
    data, gridname = _source(nvdb_golden, name)
    g = Grid(data.tobytes(), gridname)
    d = g.derived()
    assert d["extent"] == tuple(nvdb_golden[name + ".extent"]) and d["ibb_min"] == tuple(nvdb_golden[name + ".ibb_min"])
    assert np.array_equal(_bits(d["min_maj"]), _bits(nvdb_golden[name + ".min_maj"]))
    assert np.array_equal(_bits(d["transform"]), _bits(nvdb_golden[name + ".transform"]))
    assert np.array_equal(_bits(g.padded_lattice(tuple(nvdb_golden[name + ".n_bricks"]))), _bits(nvdb_golden[name + ".padded"]))


@pytest.mark.parametrize("name", NAMES_ALL)
def test_oracle_any_grid_brick_build_matches_reference(oracle, nvdb_golden, name):
    """vro_brick_build_values == the reference's BrickGrid(NanoVDBGrid), incl. the sign of a -0 minimum (first-seen std::min)."""
    got = oracle.brick_build_values(nvdb_golden[name + ".padded"], tuple(int(v) for v in nvdb_golden[name + ".extent"]))
    _golden_brick_equals(nvdb_golden, name, got)


def test_fixture_regenerates_from_the_reference(voldata_ref, nvdb_golden, tmp_path):
    """Where the reference is present: the committed file bytes still load to the committed values through the real thing."""
    p = tmp_path / "fixture.nvdb"
    p.write_bytes(nvdb_golden["nvdb_file"].tobytes())
    for name in NAMES:
        g = voldata_ref.nvdb_load(str(p), name)
        assert g is not None and g["extent"] == tuple(nvdb_golden[name + ".extent"])
        assert np.array_equal(_bits(g["padded"]), _bits(nvdb_golden[name + ".padded"]))
        assert np.array_equal(g["brick"].atlas, nvdb_golden[name + ".atlas"])
    assert voldata_ref.nvdb_load(str(p), "nope") is None
    p2 = tmp_path / "sphere.nvdb"
    p2.write_bytes(nvdb_golden["sphere_file"].tobytes())
    g = voldata_ref.nvdb_load(str(p2), "density")
    assert np.array_equal(_bits(g["padded"]), _bits(nvdb_golden["sphere.padded"])) and np.array_equal(g["brick"].atlas, nvdb_golden["sphere.atlas"])


# ---- CPU: the product's host-side reader and accessor ------------------------------------------------------------------

@pytest.mark.parametrize("name", NAMES_ALL)
def test_reader_matches_reference(nvdb_golden, name):
    import volren_b200 as vr
    n = vr.NanoVDBGridData(*_source(nvdb_golden, name))
    assert n.extent == tuple(nvdb_golden[name + ".extent"]) and n.ibb_min == tuple(nvdb_golden[name + ".ibb_min"])
    assert np.array_equal(_bits(n.min_maj), _bits(nvdb_golden[name + ".min_maj"]))
    assert np.array_equal(_bits(n.transform), _bits(nvdb_golden[name + ".transform"]))
    assert np.array_equal(_bits(n.padded_lattice()), _bits(nvdb_golden[name + ".padded"]))
    if name == "sphere":
        import struct
        tiles = struct.unpack_from("<3I", n.grid.tobytes(), 672 + 44)[0]
        assert tiles > 0 and (nvdb_golden["sphere.padded"] == 1.0).sum() >= 512 * tiles        # the interior tiles are inside the lattice
    else:
        assert n.num_voxels == {"density": 43018, "temperature": 5050}[name]       # active voxels written by make_nvdb_golden.py
    # far outside every root tile -> the root's background (the fog sphere keeps the level set's 3.0 there)
    bg = 3.0 if name == "sphere" else 0.0
    assert n.lookup([[1 << 20, 5, 5], [5, 1 << 22, 5]]).tolist() == [bg, bg]


def test_reader_raw_buffer_files_tiles_and_background(nvdb_golden):
    """Raw grid-buffer files (GridHandle::read by name), constant tiles at every level and a non-zero background:
    the product accessor against the painting oracle on patched copies of the fixture."""
    import volren_b200 as vr
    from oracle.nvdb_np import Grid, find_grid
    data = nvdb_golden["nvdb_file"].tobytes()
    off, size = find_grid(data, "density")
    raw = bytearray(data[off:off + size])
    a = vr.NanoVDBGridData(bytes(raw), "density")                      # a raw buffer is found by GridData::mGridName
    assert a.info.grid_offset == 0 and a.extent == tuple(nvdb_golden["density.extent"])
    with pytest.raises(vr.VrbError, match="No raw grid named"):
        vr.NanoVDBGridData(bytes(raw), "temperature")
    # patch: background 0.25; the first child of the first lower node becomes a constant tile 7.5; the table entry next to
    # the only child of the first upper node (child mask off) gets the tile value 3.25
    tree = 672
    leaf0, lower0, upper0, root = struct.unpack_from("<4q", raw, tree)
    struct.pack_into("<f", raw, tree + root + 28, 0.25)
    for node, mask_off, table_off, words, value, clear in ((tree + lower0, 32 + 512, 1088, 64, 7.5, True), (tree + upper0, 32 + 4096, 8256, 512, 3.25, False)):
        mask = np.frombuffer(bytes(raw[node + mask_off:node + mask_off + words * 8]), np.uint64)
        on = [w * 64 + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1]
        idx = on[0] if clear else on[0] ^ 1
        assert clear or idx not in on
        if clear:
            w = struct.unpack_from("<Q", raw, node + mask_off + (idx >> 6) * 8)[0] & ~(1 << (idx & 63))
            struct.pack_into("<Q", raw, node + mask_off + (idx >> 6) * 8, w)
        struct.pack_into("<q", raw, node + table_off + idx * 8, 0)
        struct.pack_into("<f", raw, node + table_off + idx * 8, value)
    # re-wrap as a one-grid segment file so that both readers take it
    name = b"density\0"
    meta = bytearray(176)
    struct.pack_into("<4Q", meta, 0, len(raw), len(raw), 0, 0)
    struct.pack_into("<I", meta, 136, len(name))
    seg = struct.pack("<QIHH", 0x304244566f6e614e, 32 << 21 | 7 << 10, 1, 0) + bytes(meta) + name + bytes(raw)
    n = vr.NanoVDBGridData(seg, "density")
    want = Grid(seg, "density").padded_lattice(vr._capi.brick_lattice(n.extent)[0])
    got = n.padded_lattice()
    assert np.array_equal(_bits(got), _bits(want))
    vals = set(np.unique(got).tolist())
    assert 7.5 in vals                                                # the lower-node tile lies inside the lattice
    assert n.lookup([[1 << 20, 5, 5]]).tolist() == [0.25]             # outside every root tile -> the root's background


def test_reader_rejects_what_the_reference_rejects(nvdb_golden):
    import volren_b200 as vr
    from oracle.nvdb_np import find_grid
    data = bytearray(nvdb_golden["nvdb_file"].tobytes())
    off, size = find_grid(bytes(data), "density")
    with pytest.raises(vr.VrbError, match="Grid name 'nope' not found in file"):
        vr.NanoVDBGridData(bytes(data), "nope")
    with pytest.raises(vr.VrbError, match="unknown type"):
        vr.NanoVDBGridData(b"\0" * 4096, "density")
    with pytest.raises(vr.VrbError, match="Failed to read Tree"):
        vr.NanoVDBGridData(bytes(data[:off + size // 2]), "density")                  # truncated
    for field, value in ((632, 1), (636, 2)):                                           # level set; double grid
        bad = bytearray(data)
        struct.pack_into("<I", bad, off + field, value)
        with pytest.raises(vr.VrbError, match="Empty or invalid NanoVDB grid!"):
            vr.NanoVDBGridData(bytes(bad), "density")
    bad = bytearray(data)
    struct.pack_into("<H", bad, 14, 1)                                                 # codec ZIP: the reference is built without it
    with pytest.raises(vr.VrbError, match="ZIP compression codec was disabled"):
        vr.NanoVDBGridData(bytes(bad), "density")
    bad = bytearray(data)
    struct.pack_into("<I", bad, 8, 31 << 21)                                           # older ABI
    with pytest.raises(vr.VrbError, match="Incompatible file format"):
        vr.NanoVDBGridData(bytes(bad), "density")
    # a child link that leaves the buffer is caught by the open-time walk, not by a crash in the accessor
    bad = bytearray(data)
    tree = off + 672
    root = struct.unpack_from("<4q", bad, tree)[3]
    struct.pack_into("<q", bad, tree + root + 64 + 8, 1 << 40)
    with pytest.raises(vr.VrbError, match="node offsets leave the buffer"):
        vr.NanoVDBGridData(bytes(bad), "density")


def test_host_volume_loads_nvdb(nvdb_golden, tmp_path):
    """The C++ host (voldata::Volume::load_grid -> NanoVDBGrid, volume.cpp:198-200) through volpy: members, lookup, errors."""
    volpy = _volpy()
    p = tmp_path / "fixture.nvdb"
    p.write_bytes(nvdb_golden["nvdb_file"].tobytes())
    for name in NAMES:
        g = volpy.Volume.load_grid(str(p), name)
        ext = tuple(int(v) for v in nvdb_golden[name + ".extent"])
        assert repr(g.index_extent()) == "uvec3(%d, %d, %d)" % ext
        assert np.array_equal(_bits(g.minorant_majorant()), _bits(nvdb_golden[name + ".min_maj"]))
        assert np.array_equal(_bits(np.array(g.transform)), _bits(nvdb_golden[name + ".transform"]))
        pad = nvdb_golden[name + ".padded"]
        rng = np.random.default_rng(5)
        for _ in range(300):
            x, y, z = (int(rng.integers(-2, pad.shape[2 - i] - 2)) for i in range(3))
            assert np.float32(g.lookup(volpy.uvec3(x & 0xffffffff, y & 0xffffffff, z & 0xffffffff))) == pad[z + 2, y + 2, x + 2]
    with pytest.raises(RuntimeError, match="not found in file"):
        volpy.Volume.load_grid(str(p), "flame")
    # a folder of .nvdb files: every requested name is looked up in every file, absent names are skipped (volume.cpp:286-290)
    seq = tmp_path / "seq"
    seq.mkdir()
    (seq / "a.nvdb").write_bytes(nvdb_golden["nvdb_file"].tobytes())
    vol = volpy.Volume.load_folder(str(seq), ["density", "flame", "temperature"])
    assert vol.n_grid_frames() == 1
    assert np.float32(vol.minorant_majorant("temperature")[1]) == nvdb_golden["temperature.min_maj"][1]
    assert np.float32(vol.minorant_majorant("density")[1]) == nvdb_golden["density.min_maj"][1]


# ---- GPU: device accessor + any-Grid brick build ------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_build_from_nvdb_matches_reference(ctx, nvdb_golden, name):
    import volren_b200 as vr
    n = vr.NanoVDBGridData(nvdb_golden["nvdb_file"], name)
    ctx.grid_clear()
    ctx.grid_build_from_nvdb(n)
    _golden_brick_equals(nvdb_golden, name, ctx.grid_download())
    # the same bricks from the host-tabulated lattice (the path of any other Grid source)
    ctx.grid_build_from_values(nvdb_golden[name + ".padded"], n.extent, frame=1)
    _golden_brick_equals(nvdb_golden, name, ctx.grid_download(frame=1))


@pytest.mark.gpu
@pytest.mark.parametrize("extent", [(1, 1, 1), (9, 8, 8), (30, 17, 41), (64, 64, 64)])
def test_gpu_build_from_values_matches_oracle(ctx, oracle, extent):
    """Arbitrary float lattices: negative values, +-0 ties, constant bricks, huge and denormal ranges, values outside the
    extent (the constructor reads the whole brick lattice), NaN-free."""
    nb, val = _value_lattice(extent)                             # pinned against the reference in the CPU test of the same lattices
    ctx.grid_clear()
    ctx.grid_build_from_values(val, extent)
    got, want = ctx.grid_download(), oracle.brick_build_values(val, extent)
    assert got.n_bricks == want.n_bricks and got.atlas_dim == want.atlas_dim and got.brick_count == want.brick_count
    assert np.array_equal(got.range, want.range) and np.array_equal(got.indirection, want.indirection)
    assert np.array_equal(got.atlas, want.atlas)
    for i in range(3):
        assert np.array_equal(got.mips[i], want.mips[i])


@pytest.mark.gpu
def test_gpu_host_renders_a_nvdb_volume(nvdb_golden, ctx, env_rgb, tmp_path):
    """volpy: Volume('x.nvdb') -> commit (BrickGrid(NanoVDBGrid) on the device) -> render == the C-ABI path on the same grid."""
    import volren_b200 as vr
    from test_gpu_host import _params_of
    volpy = _volpy()
    p = tmp_path / "fixture.nvdb"
    p.write_bytes(nvdb_golden["nvdb_file"].tobytes())
    W = H = 48
    volpy.create_context(W, H)
    r = volpy.Renderer()
    r.init()
    r.volume = volpy.Volume(str(p))
    r.environment = volpy.Environment(os.path.join(GOLDEN, "assets", "table_mountain_2_puresky_1k.hdr"))
    r.scale_and_move_to_unit_cube()
    r.bounces = 8
    r.commit()
    r.render(4)
    data = np.array(r.fbo_data()).reshape(H, W, 3)
    params = _params_of(r)
    n = vr.NanoVDBGridData(nvdb_golden["nvdb_file"], "density")
    ctx.grid_clear()
    ctx.grid_build_from_nvdb(n)
    ctx.env_upload(env_rgb)
    ctx.resize(W, H)
    ctx.trace(params, 1, 4)
    assert np.array_equal(data, ctx.download_color()[..., :3])
    assert data.max() > 0 and np.isfinite(data).all()


@pytest.mark.gpu
def test_gpu_empty_extent_is_an_error(ctx):
    """An empty NanoVDB grid has index extent 0 (grid_nvdb.cpp:15); the reference's brick constructor then divides by
    n_bricks.x * n_bricks.y (grid_brick.cpp:112). Here: a clear error instead of undefined behaviour."""
    import volren_b200 as vr
    with pytest.raises(vr.VrbError, match="empty grid"):
        nb, pd = vr._capi.brick_lattice((0, 5, 5))
        ctx.grid_build_from_values(np.zeros((pd[2], pd[1], pd[0]), np.float32), (0, 5, 5), frame=3)


@pytest.mark.gpu
def test_gpu_nvdb_volume_renders_like_the_oracle(ctx, oracle, nvdb_golden, env_rgb, env_pyramid):
    """End to end for a .nvdb source: device accessor + brick build + tracking kernels against the CPU oracle running on
    the REFERENCE's bricks of the same file (golden). T1 (deterministic, per-pixel rel. error < 1e-3) and T3 (RMSE below
    the oracle's own two-run RMSE) as for the other volume sources."""
    import volren_b200 as vr
    from volren_b200 import formats
    from helpers import default_scene, rel_err, rmse
    g = nvdb_golden
    n = vr.NanoVDBGridData(g["nvdb_file"], "density")
    ref_grid = formats.BrickGridData(g["density.n_bricks"], g["density.atlas_dim"], g["density.brick_count"][0], g["density.indirection"],
                                     g["density.range"], g["density.atlas"], [g[f"density.mip{i}"] for i in range(3)],
                                     min_maj=tuple(g["density.min_maj"]), transform=g["density.transform"])
    W, H = 96, 72
    mk = lambda seed: default_scene(ref_grid, W, H, bounces=32, seed=seed, index_extent=n.extent, density_scale=40.0)
    sc = oracle.make_scene(ref_grid, env_rgb, env_pyramid)
    ctx.grid_clear()
    ctx.grid_build_from_nvdb(n)
    ctx.env_upload(env_rgb)
    ctx.resize(W, H)
    p = mk(42)
    ctx.trace_deterministic(p)
    want = oracle.trace_deterministic(sc, p)
    got = ctx.download_color()
    # One brick of this grid has a NaN majorant IN THE REFERENCE: its window minimum is an active -0.0 voxel, and
    # encode_range ORs the sign-extended int16 half of the minimum over the majorant's bits (grid_brick.cpp:24-26 with
    # glm's `hdata` = short): range word 0xffff8000. Both sides reproduce the word, so rays through that brick are NaN on
    # both sides (the path tracer's sanitize() drops them); everything else must agree to 1e-3.
    assert (nvdb_golden["density.range"] == 0xffff8000).sum() == 1
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan) and 0 < nan.sum() < 0.05 * nan.size
    err = rel_err(np.where(nan, 0, got), np.where(nan, 0, want), eps=1e-3)
    assert err.max() < 1e-3 and np.nanmax(want[..., 3]) > 0.5
    SPP = 128
    ref_a, _ = oracle.trace(sc, p, 1, SPP)
    ref_b, _ = oracle.trace(sc, mk(4242), 1, SPP)
    ctx.clear()
    ctx.trace(p, 1, SPP)
    img = ctx.download_color()
    assert np.isfinite(img).all() and np.isfinite(ref_a).all()          # sanitize() (pathtracer_brick.glsl:36)
    # paths that enter the NaN brick continue from a NaN position, where the float -> int conversions are undefined in
    # GLSL and differ between CPU and GPU: compare the pixels whose (jittered) camera rays stay clear of it
    m = nan[..., 3]
    for _ in range(3):
        m = m | np.roll(m, 1, 0) | np.roll(m, -1, 0) | np.roll(m, 1, 1) | np.roll(m, -1, 1)
    ok = ~m
    assert ok.mean() > 0.8
    assert rmse(img[ok][:, :3], ref_a[ok][:, :3]) < rmse(ref_a[ok][:, :3], ref_b[ok][:, :3])


def _value_lattice(extent, nan=False):
    """The float lattices of test_gpu_build_from_values_matches_oracle (same seeds), optionally with NaN / inf voxels."""
    import volren_b200 as vr
    rng = np.random.default_rng(sum(extent))
    nb, pd = vr._capi.brick_lattice(extent)
    val = rng.standard_normal((pd[2], pd[1], pd[0])).astype(np.float32)
    val[rng.random(val.shape) < 0.5] = 0.0
    val[rng.random(val.shape) < 0.05] = -0.0
    val[:, :, pd[0] // 2:] *= 1e-6
    val[: pd[2] // 3] = np.float32(0.75)
    val[-5:, -5:, -5:] = np.float32(7e4)
    val[2:6, 2:6, 2:6] = np.float32(1e-41)
    if nan:
        val[rng.random(val.shape) < 0.01] = np.nan
        val[rng.random(val.shape) < 0.002] = np.inf
    return nb, val


@pytest.mark.parametrize("extent,nan", [((1, 1, 1), False), ((9, 8, 8), False), ((30, 17, 41), False), ((64, 64, 64), False),
                                        ((70, 33, 20), True), ((70, 33, 20), False)])
def test_oracle_any_grid_build_equals_the_reference_on_arbitrary_floats(oracle, voldata_ref, extent, nan):
    """vro_brick_build_values against the UNMODIFIED BrickGrid(const Grid&) run on a table-backed Grid subclass
    (oracle/ref_harness.cpp TableGrid): negative values (encode_range's sign extension), +-0 ties, denormals, values beyond
    the fp16 range, NaN and inf voxels (std::min / std::max order). These are the lattices the GPU test feeds the device
    builder, so GPU == oracle (test_gpu_build_from_values_matches_oracle) and oracle == reference (here) close the chain."""
    nb, val = _value_lattice(extent, nan)
    got = oracle.brick_build_values(val, extent)
    want = voldata_ref.brick_build_values(val, extent, nb)
    assert want is not None and got.brick_count == want.brick_count and got.atlas_dim == want.atlas_dim
    assert np.array_equal(got.range, want.range) and np.array_equal(got.indirection, want.indirection)
    assert np.array_equal(got.atlas, want.atlas)
    for i in range(3):
        assert np.array_equal(got.mips[i], want.mips[i])


@pytest.mark.gpu
def test_gpu_build_from_nvdb_with_interior_tiles(ctx, nvdb_golden):
    """NanoVDB's own fog-volume sphere: the device accessor takes the constant-tile branch of the lower internal nodes;
    bricks == the reference's BrickGrid(NanoVDBGrid) of the same file. (The host accessor, which shares get_value() with
    the device one, is checked against the same golden on the CPU.)"""
    import volren_b200 as vr
    n = vr.NanoVDBGridData(*_source(nvdb_golden, "sphere"))
    ctx.grid_clear()
    ctx.grid_build_from_nvdb(n)
    _golden_brick_equals(nvdb_golden, "sphere", ctx.grid_download())


def test_reader_survives_corrupted_files(nvdb_golden, tmp_path):
    """Random byte corruption anywhere in the file (headers, masks, child offsets, values): vrb_nvdb_open either rejects the
    file or accepts a grid whose every link stays inside the buffer -- the accessor must then read the whole padded lattice
    without faulting. Runs in a child process so that a fault shows up as a failed test, not as a dead test session."""
    import subprocess
    import sys
    np.save(tmp_path / "file.npy", nvdb_golden["nvdb_file"])
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import volren_b200 as vr
data = np.load(%r)
rng = np.random.default_rng(11)
accepted = rejected = 0
for it in range(600):
    bad = data.copy()
    n = int(rng.integers(1, 6))
    # a quarter each: file header + metadata + GridData + TreeData + root tiles | masks and child offsets of the first
    # upper node | the first lower node | anywhere   (density grid: buffer at byte 200, upper nodes at +256, lower at +1081856)
    lo, hi = ((0, 1400), (200 + 672 + 256, 200 + 672 + 256 + 12352), (200 + 672 + 1081856, 200 + 672 + 1081856 + 33856), (0, bad.size))[it %% 4]
    pos = rng.integers(lo, hi, n)
    bad[pos] = rng.integers(0, 256, n).astype(np.uint8)
    try:
        g = vr.NanoVDBGridData(bad, "density")
    except vr.VrbError:
        rejected += 1
        continue
    accepted += 1
    if all(0 < e <= 512 for e in g.extent):
        g.padded_lattice()
print("accepted", accepted, "rejected", rejected)
assert accepted > 0 and rejected > 0
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "file.npy"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.returncode, res.stdout[-300:], res.stderr[-800:])


def test_reader_rejects_wrapping_size_fields(nvdb_golden):
    """Sizes read from the file must not wrap the 64-bit offset arithmetic: 0, 2^64 - 256 and 2^63 in every
    size field of the container (FileMetaData.fileSize, GridData.mGridSize) -- of the segment file and of the raw grid buffer
    inside it. The reader rejects the file or finds the grid; it never reads out of bounds or spins."""
    import volren_b200 as vr
    data = nvdb_golden["nvdb_file"]
    first = vr.NanoVDBGridData(data, "density")
    off = int(first.info.grid_offset)
    raw = data[off:].copy()                                     # raw grid buffers back to back (GridHandle::read path)
    assert vr.NanoVDBGridData(raw, "density").extent == first.extent
    fields = [(data, 16 + 8), (data, off + 32), (raw, 32)]      # FileMetaData.fileSize of entry 0; mGridSize (GridData + 32)
    for src, pos in fields:
        for val in (0, 0xFFFFFFFFFFFFFF00, 1 << 63, 8):
            bad = src.copy()
            bad[pos:pos + 8] = np.frombuffer(np.uint64(val).tobytes(), np.uint8)
            for name in ("density", "temperature", "nope"):
                try:
                    g = vr.NanoVDBGridData(bad, name)
                except vr.VrbError:
                    continue
                assert g.extent == first.extent or name != "density"
