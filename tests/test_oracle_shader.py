"""CPU: known-answer and property tests of the shader-layer restatement (RNG, LUT, env pyramid, tracer).
The GLSL reference cannot run here, so these pin the oracle against independently computed values
(pure-Python integer arithmetic, numpy) and against the reference's data fixtures."""
import numpy as np

from helpers import blob_volume, default_scene, readme_scene


def tea_py(v0, v1, n=32):
    s0 = 0
    M = 0xFFFFFFFF
    for _ in range(n):
        s0 = (s0 + 0x9E3779B9) & M
        v0 = (v0 + ((((v1 << 4) & M) + 0xA341316C) & M ^ ((v1 + s0) & M) ^ (((v1 >> 5) + 0xC8013EA4) & M))) & M
        v1 = (v1 + ((((v0 << 4) & M) + 0xAD90777D) & M ^ ((v0 + s0) & M) ^ (((v0 >> 5) + 0x7E95761E) & M))) & M
    return v0


def test_tea_matches_pure_python(oracle):
    rng = np.random.default_rng(5)
    for a, b in [(0, 0), (0, 1), (1, 0), (42 * 1023, 7), (0xFFFFFFFF, 0xFFFFFFFF)] + [tuple(int(x) for x in rng.integers(0, 2**32, 2)) for _ in range(50)]:
        assert oracle.tea(a, b) == tea_py(a, b)
    # pixel 0 ignores the user seed (seed * 0 == 0): common quirk Q1
    assert oracle.tea(0, 1) == tea_py(0 * 12345, 1)


def test_lcg_stream(oracle):
    vals, states = oracle.rng_stream(12345, 1000)
    s = 12345
    for i in range(1000):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        assert int(states[i]) == s
        assert vals[i] == np.float32((s & 0xFFFFFF) / 16777216.0)
    assert vals.min() >= 0.0 and vals.max() < 1.0


def test_lcg_skip_ahead_constants():
    """The product skips the 9 dead emission draws in O(1): A9 = a^9, C9 = c (a^8 + ... + 1) mod 2^32."""
    a, c, M = 1664525, 1013904223, 1 << 32
    A9, C9 = pow(a, 9, M), sum(c * pow(a, k, M) for k in range(9)) % M
    s = 987654321
    t = s
    for _ in range(9):
        t = (t * a + c) % M
    assert (s * A9 + C9) % M == t


def test_lut_cdf_on_reference_lut(oracle, lut_raw):
    """data/lut.txt has non-monotone alpha -> upload_gpu rewrites alpha to a normalised CDF (transferfunc.cpp:33-58)."""
    from volren_b200 import formats
    assert lut_raw.shape == (8, 4)
    up, changed = oracle.lut_upload(lut_raw)
    assert changed
    a = lut_raw[:, 3].astype(np.float32)
    acc = np.float32(0)
    want = []
    for v in a:
        acc = np.float32(acc + v)
        want.append(acc)
    want = np.array(want, np.float32) / want[-1]
    assert np.array_equal(up[:, 3], want)
    assert np.array_equal(up[:, :3], lut_raw[:, :3]) and up[-1, 3] == 1.0 and np.all(np.diff(up[:, 3]) >= 0)
    assert np.array_equal(formats.lut_for_upload(lut_raw), up)      # host logic == oracle
    mono = np.array([[0, 0, 0, 0], [1, 0, 0, .25], [0, 1, 0, .5], [1, 1, 1, 1]], np.float32)
    up2, changed2 = oracle.lut_upload(mono)
    assert not changed2 and np.array_equal(up2, mono)
    zero = np.zeros((4, 4), np.float32)
    zero[1, 3] = -1.0   # non-monotone with integral <= 0 -> uniform ramp (i+1)/n
    up3, changed3 = oracle.lut_upload(np.array([[0, 0, 0, 1], [0, 0, 0, -1], [0, 0, 0, 0], [0, 0, 0, 0]], np.float32))
    assert changed3 and np.allclose(up3[:, 3], [.25, .5, .75, 1.0])


def test_hdr_decode(env_rgb):
    import hashlib
    assert env_rgb.shape == (512, 1024, 3) and env_rgb.dtype == np.float32
    assert float(env_rgb.max()) == 60416.0        # the sun texel: 236 * 2^(144-136)
    # bottom-up after the flip: the bright sky is in the upper half (large row index)
    assert env_rgb[384:].mean() > env_rgb[:128].mean()
    assert hashlib.sha1(env_rgb.tobytes()).hexdigest() == "697389af23e9d83c10bebdd5fac37be53c593ae3"


def test_env_pyramid_properties(oracle, env_rgb, env_pyramid):
    lv = [oracle.pyramid_level(env_pyramid, l) for l in range(10)]
    assert lv[0].shape == (512, 512) and lv[9].shape == (1, 1)
    for l in range(1, 10):
        s = lv[l - 1]
        want = np.float32(0.25) * ((s[0::2, 0::2] + s[0::2, 1::2]) + (s[1::2, 0::2] + s[1::2, 1::2]))
        assert np.array_equal(want, lv[l])
    # level 0 = mean luma of 64 bilinear taps; the taps tile the env map exactly -> global mean matches
    luma = env_rgb @ np.array([0.212671, 0.715160, 0.072169], np.float32)
    assert abs(float(lv[9][0, 0]) - float(luma.mean())) / float(luma.mean()) < 2e-3
    assert lv[0].min() >= 0


def test_trace_determinism_and_tiles(oracle, smoke_grid, env_rgb, env_pyramid):
    p = readme_scene(smoke_grid, 48, 40, bounces=8)
    sc = oracle.make_scene(smoke_grid, env_rgb, env_pyramid)
    a, ca = oracle.trace(sc, p, 1, 2, n_threads=1)
    b, cb = oracle.trace(sc, p, 1, 2, n_threads=4)
    assert np.array_equal(a, b) and ca.as_dict() == cb.as_dict()
    # two dispatches of one sample == one call with two samples (running mean, pathtracer_brick.glsl:36)
    c, _ = oracle.trace(sc, p, 1, 1)
    c, _ = oracle.trace(sc, p, 2, 1, color=c)
    assert np.array_equal(a, c)
    # tiles partition the image exactly
    d = np.zeros_like(a)
    for tile in [(0, 0, 20, 40), (20, 0, 48, 17), (20, 17, 48, 40)]:
        oracle.trace(sc, p, 1, 2, color=d, tile=tile)
    assert np.array_equal(a, d)
    # sum mode / n == mean mode up to fp32 rounding
    s, _ = oracle.trace(sc, p, 1, 2, accum_mode=1)
    assert np.allclose(s / 2, a, rtol=1e-6, atol=1e-7)
    assert ca.n_samples == 48 * 40 * 2 and ca.n_nee == ca.n_real and ca.n_emis == 0
    assert set(np.unique(a[..., 3])) <= {0.0, 0.5, 1.0}     # alpha = mean hit flag


def test_trace_white_furnace_like_energy(oracle, env_rgb, env_pyramid):
    """albedo 1 + constant white environment + no RR bias: radiance stays ~1 everywhere (energy conservation)."""
    vox, lo, hi = blob_volume(32)
    g = oracle.brick_build(vox, lo, hi)
    white = np.ones((4, 8, 3), np.float32)
    pyr = oracle.env_build(white)
    p = default_scene(g, 32, 32, bounces=10000, index_extent=(32, 32, 32), albedo=(1, 1, 1), density_scale=20.0)
    sc = oracle.make_scene(g, white, pyr)
    img, cnt = oracle.trace(sc, p, 1, 64)
    assert cnt.n_real > 0
    assert abs(float(img[..., :3].mean()) - 1.0) < 0.02


def test_trace_tf_variant_runs(oracle, smoke_grid, env_rgb, env_pyramid, lut_raw):
    lut, _ = oracle.lut_upload(lut_raw)
    p = default_scene(smoke_grid, 40, 24, bounces=16, use_tf=True)
    sc = oracle.make_scene(smoke_grid, env_rgb, env_pyramid, lut=lut)
    img, cnt = oracle.trace(sc, p, 1, 4)
    assert np.isfinite(img).all() and cnt.n_dens > 0 and cnt.n_env == 0   # show_environment=false with a TF (main.cpp:76)
    assert img[..., :3].max() > 0


def test_deterministic_mode_oracle(oracle, smoke_grid, env_rgb, env_pyramid):
    p = readme_scene(smoke_grid, 32, 32)
    sc = oracle.make_scene(smoke_grid, env_rgb, env_pyramid)
    img = oracle.trace_deterministic(sc, p)
    assert np.isfinite(img).all()
    a = img[..., 3]
    assert a.min() >= 0 and a.max() <= 1 and a.max() > 0.9 and a[0, 0] == 0.0   # opaque core, empty corner


def test_tonemap(oracle):
    rng = np.random.default_rng(0)
    c = rng.random((8, 8, 4)).astype(np.float32) * 4
    t = oracle.tonemap_inplace(c, 3.0, 2.0)

    def hable(x):
        A, B, C_, D, E, F = .15, .5, .1, .2, .02, .3
        return ((x * (A * x + C_ * B) + D * E) / (x * (A * x + B) + D * F)) - E / F
    want = (hable(3.0 * c[..., :3].astype(np.float64)) / hable(11.2)) ** 0.5
    assert np.allclose(t[..., :3], want, rtol=1e-5) and np.array_equal(t[..., 3], c[..., 3])
    ldr = oracle.draw(c, 3.0, 2.0, True)
    assert np.abs(ldr[..., :3].astype(np.int32) - np.rint(np.clip(want, 0, 1) * 255)).max() <= 1
    assert np.array_equal(oracle.draw(c, 3.0, 2.0, False), oracle.color_to_ldr(c))
