"""CPU: pins the oracle's voldata restatement (oracle/vr_oracle.c) against
 (1) golden vectors generated from the unmodified reference voldata sources (tests/golden/make_golden.py),
 (2) the compiled reference itself where oracle/_ref exists, and
 (3) the invariants of the reference's data/smoke.brick."""
import hashlib

import numpy as np
import pytest

from helpers import synth_cases

CASES = synth_cases()


def _sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("name", sorted(CASES))
def test_brick_build_matches_golden(oracle, golden, name):
    vox, lo, hi = CASES[name]
    g = oracle.brick_build(vox, lo, hi)
    assert tuple(golden[name + ".n_bricks"]) == g.n_bricks
    assert tuple(golden[name + ".atlas_dim"]) == g.atlas_dim
    assert int(golden[name + ".brick_count"][0]) == g.brick_count
    assert np.array_equal(golden[name + ".indirection"], g.indirection)
    assert np.array_equal(golden[name + ".range"], g.range)
    assert np.array_equal(golden[name + ".atlas_sha1"], _sha(g.atlas))
    for i in range(3):
        assert np.array_equal(golden[name + f".mip{i}"], g.mips[i])
    assert np.array_equal(golden[name + ".decode_sha1"], _sha(g.decode_all()))


@pytest.mark.parametrize("name", sorted(CASES))
def test_brick_build_matches_compiled_reference(oracle, voldata_ref, name):
    vox, lo, hi = CASES[name]
    a = oracle.brick_build(vox, lo, hi)
    b, dec = voldata_ref.brick_build(vox, lo, hi, decode=True)
    assert a.n_bricks == b.n_bricks and a.atlas_dim == b.atlas_dim and a.brick_count == b.brick_count
    for f in ("indirection", "range", "atlas"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for i in range(3):
        assert np.array_equal(a.mips[i], b.mips[i])
    assert np.array_equal(a.decode_all().view(np.uint32), dec.view(np.uint32))   # numpy decode == BrickGrid::lookup, bit for bit (NaN-safe)


def test_half_round_half_up_golden(oracle, golden):
    got = np.array([oracle.to_half(f) for f in golden["half.inputs"]], np.uint16)
    assert np.array_equal(got, golden["half.outputs"])
    # it is NOT IEEE ties-to-even: at least one tie must differ from numpy's float16 cast
    ieee = golden["half.inputs"].astype(np.float16).view(np.uint16)
    finite = np.isfinite(golden["half.inputs"]) & (np.abs(golden["half.inputs"]) < 65504)
    assert np.any(got[finite] != ieee[finite])


def test_half_decode_is_exact(oracle):
    hs = np.arange(0, 0x7c00, dtype=np.uint16)
    want = hs.view(np.float16).astype(np.float32)
    got = np.array([oracle.from_half(int(h)) for h in hs[::5]], np.float32)
    assert np.array_equal(got, want[::5])


def test_dense_from_float_golden(oracle, golden):
    for k in ("dense", "dense_neg"):
        q, mm = oracle.dense_from_float(golden[k + ".input"])
        assert np.array_equal(q, golden[k + ".u8"])
        assert np.array_equal(np.array(mm, np.float32), golden[k + ".minmax"])
    # the FLT_MIN quirk (grid_dense.cpp:61): an all-negative grid keeps max = FLT_MIN
    assert golden["dense_neg.minmax"][1] == np.float32(1.1754943508222875e-38)


def test_too_many_bricks(oracle):
    st, nb = oracle.brick_dims((8 * 1017, 8, 8))  # rounds up to 1024 bricks -> rejected (>= MAX_BRICKS)
    assert st != 0
    st, nb = oracle.brick_dims((8 * 1016, 8, 8))
    assert st == 0 and nb == (1016, 8, 8)


def test_n_bricks_rounds_to_multiple_of_8(oracle):
    for dim, want in [((1, 1, 4), (8, 8, 8)), ((65, 64, 63), (16, 8, 8)), ((512, 512, 1800), (64, 64, 232))]:
        st, nb = oracle.brick_dims(dim)
        assert st == 0 and nb == want


# ---- data/smoke.brick ------------------------------------------------------------------------------

def test_smoke_brick_parser_matches_reference_loader(smoke_grid, smoke_golden):
    g = smoke_grid
    assert g.n_bricks == tuple(smoke_golden["n_bricks"]) == (16, 32, 16)
    assert g.atlas_dim == tuple(smoke_golden["atlas_dim"]) == (128, 256, 56)
    assert g.brick_count == int(smoke_golden["brick_count"][0]) == 3297
    assert np.array_equal(np.array(g.min_maj, np.float32), smoke_golden["min_maj"])
    assert g.min_maj[1] == 5.71484375
    assert np.array_equal(g.transform, smoke_golden["transform"])
    assert np.array_equal(_sha(g.indirection), smoke_golden["indirection_sha1"])
    assert np.array_equal(_sha(g.range), smoke_golden["range_sha1"])
    assert np.array_equal(_sha(g.atlas), smoke_golden["atlas_sha1"])
    for i in range(3):
        assert np.array_equal(_sha(g.mips[i]), smoke_golden["mips_sha1"][i])
    assert np.array_equal(_sha(g.decode_all()), smoke_golden["decode_sha1"])


def test_smoke_brick_mip_invariants(oracle, smoke_grid):
    """Every mip texel is encode_range(min of child mins, max of child maxes) (grid_brick.cpp:114-141)."""
    g = smoke_grid

    def halves(words):
        lo = (words & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
        hi = (words >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
        return lo, hi

    src = g.range
    for i in range(3):
        lo, hi = halves(src)
        z, y, x = lo.shape
        lo_c = lo.reshape(z // 2, 2, y // 2, 2, x // 2, 2).min(axis=(1, 3, 5))
        hi_c = hi.reshape(z // 2, 2, y // 2, 2, x // 2, 2).max(axis=(1, 3, 5))
        want = np.array([[[(oracle.to_half(a) | (oracle.to_half(b) << 16)) for a, b in zip(ra, rb)] for ra, rb in zip(pa, pb)]
                         for pa, pb in zip(lo_c, hi_c)], np.uint32)
        assert np.array_equal(want, g.mips[i]), f"mip {i}"
        src = g.mips[i]


def test_smoke_brick_pointer_invariants(smoke_grid):
    g = smoke_grid
    lo = (g.range & 0xFFFF).astype(np.uint16).view(np.float16)
    hi = (g.range >> 16).astype(np.uint16).view(np.float16)
    nonempty = hi != lo
    assert int(nonempty.sum()) == 3297
    px, py, pz = g.decode_ptr()
    slot = (pz.astype(np.int64) * g.n_bricks[1] + py) * g.n_bricks[0] + px
    used = slot[nonempty]
    assert len(np.unique(used)) == len(used) and used.max() == 3296  # a permutation of 0..3296 (TBB order, Q8)
    assert np.all(g.indirection[~nonempty] == 0)
    assert float(hi.astype(np.float32).max()) == 5.71484375


def test_brick_file_roundtrip(tmp_path, smoke_grid):
    from volren_b200 import formats
    p = tmp_path / "rt.brick"
    formats.save_brick(p, smoke_grid)
    import os
    with open(p, "rb") as a, open(os.path.join(os.path.dirname(__file__), "golden", "assets", "smoke.brick"), "rb") as b:
        assert a.read() == b.read()


def test_brick_file_written_by_reference_is_parsed(tmp_path, voldata_ref, oracle):
    from volren_b200 import formats
    vox, lo, hi = CASES["ragged_70x33x20"]
    p = str(tmp_path / "ref.brick")
    voldata_ref.brick_roundtrip_write(vox, lo, hi, p)
    g = formats.load_brick(p)
    want = oracle.brick_build(vox, lo, hi)
    assert g.n_bricks == want.n_bricks and g.brick_count == want.brick_count
    assert np.array_equal(g.atlas, want.atlas) and np.array_equal(g.indirection, want.indirection)
    pd = str(tmp_path / "ref.dense")
    voldata_ref.dense_write(vox, lo, hi, pd)
    d = formats.load_dense(pd)
    assert np.array_equal(d.voxels, vox) and (d.min_value, d.max_value) == (lo, hi)
