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
