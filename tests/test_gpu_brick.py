"""GPU (through the C ABI): the brick-grid builder and the dense quantiser are bit-exact against the oracle
(which is pinned against the compiled reference voldata) and against the committed golden vectors."""
import hashlib

import numpy as np
import pytest

from helpers import synth_cases

pytestmark = pytest.mark.gpu
CASES = synth_cases()


def _sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def _assert_same(a, b):
    assert a.n_bricks == b.n_bricks and a.atlas_dim == b.atlas_dim and a.brick_count == b.brick_count
    assert np.array_equal(a.range, b.range), "range"
    assert np.array_equal(a.indirection, b.indirection), "indirection"
    for i in range(3):
        assert np.array_equal(a.mips[i], b.mips[i]), f"mip{i}"
    assert np.array_equal(a.atlas, b.atlas), "atlas"


@pytest.mark.parametrize("name", sorted(CASES))
def test_build_matches_oracle_and_golden(ctx, oracle, golden, name):
    vox, lo, hi = CASES[name]
    ctx.grid_clear()
    ctx.grid_build_from_dense(vox, lo, hi)
    got = ctx.grid_download()
    _assert_same(got, oracle.brick_build(vox, lo, hi))
    assert np.array_equal(golden[name + ".range"], got.range)
    assert np.array_equal(golden[name + ".indirection"], got.indirection)
    assert np.array_equal(golden[name + ".atlas_sha1"], _sha(got.atlas))
    assert np.array_equal(golden[name + ".decode_sha1"], _sha(got.decode_all()))


@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 9, 8), (8, 8, 8), (9, 8, 8), (63, 64, 65), (128, 24, 200), (3, 300, 5)])
def test_build_random_shapes(ctx, oracle, shape):
    rng = np.random.default_rng(sum(shape))
    vox = (rng.random(shape) * 255).astype(np.uint8)
    vox[rng.random(shape) < 0.7] = 0
    ctx.grid_clear()
    ctx.grid_build_from_dense(vox, 0.0, 1.0)
    _assert_same(ctx.grid_download(), oracle.brick_build(vox, 0.0, 1.0))


def test_build_large_grid_properties(ctx):
    """256^3 at full size: size-independent properties (mip = min/max of children, decode error <= half a code,
    allocation is the raster-order prefix sum)."""
    n = 256
    z, y, x = np.mgrid[0:n, 0:n, 0:n].astype(np.float32) / n
    f = np.clip(np.sin(9 * x) * np.sin(7 * y) * np.sin(11 * z), 0, 1)
    vox = (f * 255).astype(np.uint8)
    ctx.grid_clear()
    ctx.grid_build_from_dense(vox, 0.0, 2.0)
    g = ctx.grid_download()
    assert g.n_bricks == (32, 32, 32)
    lo = (g.range & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
    hi = (g.range >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
    dense = (vox.astype(np.float32) / np.float32(255)) * np.float32(2.0)
    # dilated 12^3 range contains the brick's own voxels (up to the fp16 rounding of the bounds)
    blocks = dense.reshape(32, 8, 32, 8, 32, 8)
    assert np.all(blocks.min(axis=(1, 3, 5)) >= lo - 1e-3) and np.all(blocks.max(axis=(1, 3, 5)) <= hi + 2e-3)
    nonempty = hi != lo
    # raster-order allocation: ids are the exclusive prefix sum of the non-empty flags
    px, py, pz = g.decode_ptr()
    ids = (pz.astype(np.int64) * 32 + py) * 32 + px
    want = np.cumsum(nonempty.ravel()) - 1
    assert np.array_equal(ids.ravel()[nonempty.ravel()], want[nonempty.ravel()])
    assert g.brick_count == int(nonempty.sum())
    assert g.atlas_dim[2] == 8 * int(np.ceil(g.brick_count / (32 * 32)))
    # decode error: half a code of the brick's range
    dec = g.decode_all()
    tol = np.repeat(np.repeat(np.repeat((hi - lo) / 255 * 0.51 + 1e-6, 8, 0), 8, 1), 8, 2)
    assert np.all(np.abs(dec - dense) <= tol)
    # mips
    src_lo, src_hi = lo, hi
    for i in range(3):
        m = g.mips[i]
        mlo = (m & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
        mhi = (m >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
        s = src_lo.shape[0]
        assert np.array_equal(mlo, src_lo.reshape(s // 2, 2, s // 2, 2, s // 2, 2).min(axis=(1, 3, 5)))
        assert np.array_equal(mhi, src_hi.reshape(s // 2, 2, s // 2, 2, s // 2, 2).max(axis=(1, 3, 5)))
        src_lo, src_hi = mlo, mhi


def test_upload_brick_roundtrip(ctx, smoke_grid):
    ctx.grid_clear()
    ctx.grid_upload_brick(smoke_grid)
    _assert_same(ctx.grid_download(), smoke_grid)


def test_dense_from_float(ctx, oracle, golden):
    for k in ("dense", "dense_neg"):
        q, mm = ctx.dense_from_float(golden[k + ".input"])
        assert np.array_equal(q, golden[k + ".u8"])
        assert np.array_equal(np.array(mm, np.float32), golden[k + ".minmax"])
    rng = np.random.default_rng(8)
    d = (rng.standard_normal((37, 61, 129)) * 5).astype(np.float32)
    q, mm = ctx.dense_from_float(d)
    q2, mm2 = oracle.dense_from_float(d)
    assert mm == mm2 and np.array_equal(q, q2)


def test_too_many_bricks_is_an_error(ctx):
    from volren_b200 import VrbError, _capi
    vox = np.zeros((1, 1, 8 * 1017), np.uint8)
    with pytest.raises(VrbError) as e:
        ctx.grid_build_from_dense(vox, 0.0, 1.0)
    assert e.value.status == _capi.VRB_ERR_TOO_MANY_BRICKS and "1024" in str(e.value)


@pytest.mark.parametrize("lo,hi", [(-3.0, 7.5), (1e-30, 3e-30), (5.0, 5.0), (2.0, -1.0), (0.0, 70000.0), (-1e20, 1e20)],
                         ids=["negative", "tiny", "collapsed", "reversed", "beyond-fp16", "huge"])
def test_fast_path_value_ranges(ctx, oracle, lo, hi):
    """The table-driven encode (k_brick_encode_lut: division by the brick's rounded reciprocal + fma corrections, with the IEEE
    division as the guarded fallback) against the oracle for DenseGrid value ranges that leave the comfortable regime: negative
    minima (sign-extended range words), denormal-scale spans, min == max, min > max, values beyond fp16 (infinite range halves),
    1e20 (the fallback). Widths that take the fast path, with padding bricks and a ragged last chunk."""
    rng = np.random.default_rng(7)
    vox = (rng.random((21, 30, 136)) * 255).astype(np.uint8)
    vox[rng.random(vox.shape) < 0.5] = 0
    vox[:, :, 100:] = 0                                  # whole empty bricks as well
    ctx.grid_build_from_dense(vox, lo, hi)
    _assert_same(ctx.grid_download(), oracle.brick_build(vox, lo, hi))


@pytest.mark.parametrize("name", ["smoke", "blobs"])
def test_decoded_apron_blocks_equal_the_canonical_fetch(ctx, smoke_grid, name):
    """The production kernel's trilinear fetch reads DECODED 9^3 apron blocks (one slot load + 8 fp32 loads); it must
    return the canonical fetch (records + u8 atlas, lookup_density_trilinear, common.glsl:289-297) bit for bit at any
    point -- inside bricks, across brick faces / edges / corners, in empty bricks, at and beyond the grid border --
    and both must agree with trilinear interpolation of BrickGrid::lookup (grid_brick.cpp:148-154) evaluated on the host."""
    ctx.grid_clear()
    if name == "smoke":
        ctx.grid_upload_brick(smoke_grid)
    else:
        rng = np.random.default_rng(5)
        vox = (rng.random((40, 72, 56)) * 255).astype(np.uint8)
        vox[rng.random(vox.shape) < 0.5] = 0
        vox[:, :24, :] = 0                    # whole empty bricks next to full ones
        vox[8:16, 40:48, 8:16] = 77           # a constant brick (collapsed range)
        ctx.grid_build_from_dense(vox, 0.25, 3.0)      # lo = 0.25: "empty" bricks decode to a non-zero constant
    g = ctx.grid_download()
    dec = g.decode_all()
    ez, ey, ex = dec.shape
    rng = np.random.default_rng(11)
    n = 200_000
    pts = (rng.random((n, 3)) * (np.array([ex, ey, ez]) + 6) - 3).astype(np.float32)        # up to 3 voxels outside
    # brick faces / edges / corners: coordinates within +-1 voxel of multiples of 8
    k = n // 2
    snap = rng.integers(0, 3, size=(k, 3)) > 0
    near = (np.round(pts[:k] / 8) * 8 + (rng.random((k, 3)) * 2 - 1)).astype(np.float32)
    pts[:k] = np.where(snap, near, pts[:k])
    a = ctx.sample_density(pts, mode=0)
    b = ctx.sample_density(pts, mode=1)
    assert np.array_equal(a, b)
    assert np.count_nonzero(a) > n // 20
    # host evaluation: 8 taps of the decoded voxels at ipos - 0.5 (0 outside), mix(x, y, a) = x * (1 - a) + y * a in fp32
    q = pts - np.float32(0.5)
    fl = np.floor(q)
    f = (q - fl).astype(np.float32)
    i0 = fl.astype(np.int64)
    pad = np.zeros((ez + 8, ey + 8, ex + 8), np.float32)
    pad[4:4 + ez, 4:4 + ey, 4:4 + ex] = dec

    def tap(dx, dy, dz):
        x, y, z = i0[:, 0] + dx + 4, i0[:, 1] + dy + 4, i0[:, 2] + dz + 4
        ok = (x >= 0) & (x < ex + 8) & (y >= 0) & (y < ey + 8) & (z >= 0) & (z < ez + 8)
        return np.where(ok, pad[np.clip(z, 0, ez + 7), np.clip(y, 0, ey + 7), np.clip(x, 0, ex + 7)], np.float32(0))

    def mix(x, y, w):
        return (x * (np.float32(1) - w) + y * w).astype(np.float32)
    rows = [mix(tap(0, dy, dz), tap(1, dy, dz), f[:, 0]) for dz in (0, 1) for dy in (0, 1)]
    want = mix(mix(rows[0], rows[1], f[:, 1]), mix(rows[2], rows[3], f[:, 1]), f[:, 2])
    # the voxel decode differs by an ulp between GL's u8 / 255.f and voldata's u8 * (1 / 255.f), the lerps by FMA contraction
    assert np.allclose(a, want, rtol=2e-6, atol=1e-7)
    # nearest voxel (lookup_density_brick)
    c = ctx.sample_density(pts, mode=2)
    ip = np.floor(pts).astype(np.int64)
    inside = np.all((ip >= 0) & (ip < np.array([ex, ey, ez])), axis=1)
    assert np.all(c[~inside] == 0)
    assert np.allclose(c[inside], dec[ip[inside, 2], ip[inside, 1], ip[inside, 0]], rtol=1e-6, atol=0)
