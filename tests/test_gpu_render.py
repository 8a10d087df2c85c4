"""GPU (through the C ABI): environment set-up, the tracking kernels, the deterministic mode, tonemapping and
accumulation against the CPU oracle. Tolerances are written next to each assertion:
  T0 integer / bit-exact layers; T1 deterministic mode: per-pixel rel. error < 1e-3 (north star);
  T2 same-seed replay: diagnostic fraction; T3 statistical: RMSE(gpu, oracle) < RMSE(oracle seed A, oracle seed B)."""
import numpy as np
import pytest

from helpers import blob_volume, default_scene, readme_scene, rel_err, rmse

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smoke_ctx(ctx, smoke_grid, env_rgb):
    ctx.grid_clear()
    ctx.grid_upload_brick(smoke_grid)
    ctx.env_upload(env_rgb)
    return ctx


def test_env_importance_pyramid(smoke_ctx, oracle, env_pyramid):
    for level in range(10):
        got = smoke_ctx.env_download_impmap(level)
        want = oracle.pyramid_level(env_pyramid, level)
        # fp32 sums of 64 taps, FMA contraction on the device: 1e-5 relative
        assert np.allclose(got, want, rtol=1e-5, atol=1e-7), level
    l0 = smoke_ctx.env_download_impmap(0)
    l1 = smoke_ctx.env_download_impmap(1)
    want = np.float32(0.25) * ((l0[0::2, 0::2] + l0[0::2, 1::2]) + (l0[1::2, 0::2] + l0[1::2, 1::2]))
    assert np.array_equal(l1, want)       # the box filter itself is bit-exact


def test_deterministic_mode_T1(smoke_ctx, oracle, smoke_grid, env_rgb, env_pyramid):
    W, H = 160, 120
    p = readme_scene(smoke_grid, W, H)
    smoke_ctx.resize(W, H)
    smoke_ctx.trace_deterministic(p)
    got = smoke_ctx.download_color()
    want = oracle.trace_deterministic(oracle.make_scene(smoke_grid, env_rgb, env_pyramid), p)
    # north star: per-pixel relative error < 1e-3 (relative to max(|ref|, 1e-3) so that fully transparent alpha = 0 pixels count)
    err = rel_err(got, want, eps=1e-3)
    assert err.max() < 1e-3, (err.max(), np.unravel_index(err.argmax(), err.shape))
    assert want[..., 3].max() > 0.9


@pytest.mark.parametrize("use_tf", [False, True])
def test_same_seed_replay_T2(smoke_ctx, oracle, smoke_grid, env_rgb, env_pyramid, lut_raw, use_tf):
    """1 spp, few bounces, same TEA/LCG seeds: the production (fast-math) kernel replays the oracle pixel for pixel to 1e-3 --
    measured 1.0000 of the pixels on B200 for both programs (round 1 asked for 0.97); a last-bit difference of a MUFU function
    can flip a comparison, hence >= 0.999 rather than all."""
    W, H = 128, 96
    lut, _ = oracle.lut_upload(lut_raw)
    smoke_ctx.tf_upload(lut)
    p = default_scene(smoke_grid, W, H, bounces=3, use_tf=True) if use_tf else readme_scene(smoke_grid, W, H, bounces=3)
    smoke_ctx.resize(W, H)
    smoke_ctx.trace(p, 1, 1)
    got = smoke_ctx.download_color()
    want, _ = oracle.trace(oracle.make_scene(smoke_grid, env_rgb, env_pyramid, lut=lut), p, 1, 1)
    ok = np.all(rel_err(got, want, eps=1e-3) < 1e-3, axis=-1)
    frac = ok.mean()
    print(f"T2 same-seed pixel match fraction (tf={use_tf}): {frac:.4f}")
    assert frac >= 0.999
    assert np.array_equal(got[..., 3] > 0, want[..., 3] > 0) or (got[..., 3] != want[..., 3]).mean() < 0.01


@pytest.mark.parametrize("use_tf,W,H,SPP", [(False, 96, 96, 256), (True, 96, 96, 256), (False, 48, 48, 4096), (True, 48, 48, 4096)],
                         ids=["notf-256spp", "tf-256spp", "notf-4096spp-converged", "tf-4096spp-converged"])
def test_statistical_parity_T3_and_counters(smoke_ctx, oracle, smoke_grid, env_rgb, env_pyramid, lut_raw, use_tf, W, H, SPP):
    """The north star's image criterion, literally at 4096 spp (on a 48^2 frame so that the two CPU oracle runs take
    seconds) and at 256 spp on a larger frame."""
    lut, _ = oracle.lut_upload(lut_raw)
    smoke_ctx.tf_upload(lut)
    mk = (lambda seed: default_scene(smoke_grid, W, H, bounces=128, use_tf=True, seed=seed)) if use_tf else (lambda seed: readme_scene(smoke_grid, W, H, seed=seed))
    sc = oracle.make_scene(smoke_grid, env_rgb, env_pyramid, lut=lut)
    ref_a, cnt_a = oracle.trace(sc, mk(42), 1, SPP)
    ref_b, _ = oracle.trace(sc, mk(4242), 1, SPP)
    smoke_ctx.resize(W, H)
    smoke_ctx.set_counting(True)
    smoke_ctx.trace(mk(42), 1, SPP)
    got = smoke_ctx.download_color()
    cnt = smoke_ctx.get_counters().as_dict()
    smoke_ctx.set_counting(False)
    # criterion of the north star at reduced spp: RMSE against the oracle is below the oracle's own 2-run RMSE
    # (same seed -> it is in fact far below)
    two_run = rmse(ref_a[..., :3], ref_b[..., :3])
    assert rmse(got[..., :3], ref_a[..., :3]) < two_run
    # and an independent-seed GPU run is statistically indistinguishable: within 1.25x of the 2-run RMSE
    smoke_ctx.clear()
    smoke_ctx.trace(mk(777), 1, SPP)
    got_c = smoke_ctx.download_color()
    assert rmse(got_c[..., :3], ref_a[..., :3]) < 1.25 * two_run
    assert abs(got_c[..., :3].mean() - ref_a[..., :3].mean()) / ref_a[..., :3].mean() < 0.01
    # event counters define the algorithmic bytes: must agree with the oracle to < 1 %
    want = cnt_a.as_dict()
    assert cnt["n_samples"] == want["n_samples"] == W * H * SPP
    for k in ("n_maj", "n_dens", "n_nee", "n_env", "n_real"):
        if want[k]:
            assert abs(cnt[k] - want[k]) / want[k] < 0.01, (k, cnt[k], want[k])
    assert cnt["n_emis"] == 0


def test_running_mean_equals_successive_dispatches(smoke_ctx, smoke_grid):
    """vrb_trace(first, n) == n calls of one sample (renderer.cpp:138 ++sample; pathtracer_brick.glsl:36 mix)."""
    W, H = 64, 48
    p = readme_scene(smoke_grid, W, H, bounces=8)
    smoke_ctx.resize(W, H)
    smoke_ctx.trace(p, 1, 5)
    a = smoke_ctx.download_color()
    smoke_ctx.clear()
    for s in range(1, 6):
        smoke_ctx.trace(p, s, 1)
    b = smoke_ctx.download_color()
    assert np.array_equal(a, b)
    # tiles partition the image exactly
    smoke_ctx.clear()
    for tile in [(0, 0, 30, 48), (30, 0, 64, 11), (30, 11, 64, 48)]:
        smoke_ctx.trace(p, 1, 5, tile=tile)
    assert np.array_equal(a, smoke_ctx.download_color())
    # sum mode + scale == mean mode up to fp32 rounding
    smoke_ctx.clear()
    smoke_ctx.trace(p, 1, 5, accum_mode=1)
    smoke_ctx.scale(1 / 5)
    assert np.allclose(smoke_ctx.download_color(), a, rtol=2e-6, atol=1e-7)


def test_accumulate_matches_oracle_given_same_L(smoke_ctx, oracle, smoke_grid, env_rgb, env_pyramid):
    """With bounces=0... no volume interaction differences: pixels whose rays miss the volume see only the environment,
    a pure function of the jitter -> the running mean must agree with the oracle to fp32 lookup precision."""
    W, H = 64, 64
    p = readme_scene(smoke_grid, W, H, bounces=4)
    smoke_ctx.resize(W, H)
    smoke_ctx.trace(p, 1, 8)
    got = smoke_ctx.download_color()
    want, _ = oracle.trace(oracle.make_scene(smoke_grid, env_rgb, env_pyramid), p, 1, 8)
    miss = want[..., 3] == 0
    assert miss.sum() > 500
    assert rel_err(got[miss][:, :3], want[miss][:, :3], eps=1e-3).max() < 1e-3


def test_synthetic_dense_volume_end_to_end(ctx, oracle, env_rgb, env_pyramid):
    """DenseGrid -> GPU brick build -> trace, against the oracle built from the oracle's own brick grid."""
    vox, lo, hi = blob_volume(48)
    ctx.grid_build_from_dense(vox, lo, hi, frame=3)     # frame 3: leaves the module's smoke grid (frame 0) alone
    ctx.env_upload(env_rgb)
    g = oracle.brick_build(vox, lo, hi)
    W, H, SPP = 64, 64, 128
    p = default_scene(g, W, H, bounces=64, index_extent=(48, 48, 48), density_scale=8.0, frame=3)
    sc = oracle.make_scene(g, env_rgb, env_pyramid)
    ref_a, _ = oracle.trace(sc, p, 1, SPP)
    p2 = p.copy()
    p2.seed = 9001
    ref_b, _ = oracle.trace(sc, p2, 1, SPP)
    ctx.resize(W, H)
    ctx.trace(p, 1, SPP)
    got = ctx.download_color()
    assert rmse(got[..., :3], ref_a[..., :3]) < rmse(ref_a[..., :3], ref_b[..., :3])
    assert ref_a[..., 3].max() == 1.0


def test_emission_grid_end_to_end(ctx, oracle, env_rgb, env_pyramid):
    """a density grid plus a `temperature` emission grid with its OWN transform (half resolution, so
    emis_from_density is a real matrix), through lookup_emission / lookup_temperature_brick (common.glsl:307-328,
    renderer.cpp:64-74,117-124). Statistical tier against the oracle + exact n_emis accounting."""
    from volren_b200 import scene
    vox, lo, hi = blob_volume(48)
    zz, yy, xx = np.mgrid[0:24, 0:24, 0:24].astype(np.float32) / 24
    temp = np.exp(-(((xx - .45) / .2) ** 2 + ((yy - .5) / .2) ** 2 + ((zz - .5) / .25) ** 2))
    tvox = (np.clip(temp - 0.1, 0, 1) / 0.9 * 255).astype(np.uint8)
    ctx.grid_build_from_dense(vox, lo, hi, frame=5)
    ctx.grid_build_from_dense(tvox, 0.0, 2.0, slot=1, frame=5)
    ctx.env_upload(env_rgb)
    g, ge = oracle.brick_build(vox, lo, hi), oracle.brick_build(tvox, 0.0, 2.0)
    got_e = ctx.grid_download(slot=1, frame=5)
    assert np.array_equal(got_e.range, ge.range) and np.array_equal(got_e.atlas, ge.atlas)
    W, H, SPP = 64, 64, 128
    emat = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32) @ np.asarray(g.matrix(), np.float32)   # 24^3 voxels of size 2 cover the 48^3 grid

    def mk(seed):
        s = scene.RenderSettings(bounces=32, seed=seed, density_scale=8.0, emission_scale=40.0, albedo=(.7, .7, .7), frame=5)
        scene.scale_and_move_to_unit_cube(g.matrix(), (48, 48, 48), s)
        s.density_scale = 8.0
        return scene.make_params(W, H, scene.Camera(), s, g.matrix(), (48, 48, 48), g.min_maj,
                                 emission_matrix=emat, majorant_emission=ge.min_maj[1])
    sc = oracle.make_scene(g, env_rgb, env_pyramid, emission=ge)
    ref_a, cnt_a = oracle.trace(sc, mk(42), 1, SPP)
    ref_b, _ = oracle.trace(sc, mk(777), 1, SPP)
    p0 = mk(42)
    p0.has_emission = 0
    ref_0, _ = oracle.trace(oracle.make_scene(g, env_rgb, env_pyramid), p0, 1, SPP)
    assert ref_a[..., :3].mean() > 1.05 * ref_0[..., :3].mean()      # the emission term is visible in the image
    ctx.resize(W, H)
    ctx.set_counting(True)
    ctx.trace(mk(42), 1, SPP)
    got = ctx.download_color()
    cnt = ctx.get_counters().as_dict()
    ctx.set_counting(False)
    two_run = rmse(ref_a[..., :3], ref_b[..., :3])
    assert rmse(got[..., :3], ref_a[..., :3]) < two_run
    ctx.clear()
    ctx.trace(mk(4711), 1, SPP)
    got_c = ctx.download_color()
    assert rmse(got_c[..., :3], ref_a[..., :3]) < 1.25 * two_run
    want = cnt_a.as_dict()
    assert want["n_emis"] > 0
    for k in ("n_maj", "n_dens", "n_emis", "n_nee", "n_real"):
        assert abs(cnt[k] - want[k]) / want[k] < 0.01, (k, cnt[k], want[k])
    # the cross-check kernels agree with the production kernel's event counts path for path
    for kind in (1, 2):
        ctx.set_kernel(kind)
        ctx.clear()
        ctx.set_counting(True)
        ctx.trace(mk(42), 1, 4)
        c = ctx.get_counters().as_dict()
        ctx.set_counting(False)
        if kind == 1:
            c1 = c
        else:
            assert c == c1
    ctx.set_kernel(0)
    ctx.grid_free(1, 5)
    ctx.grid_free(0, 5)


def test_tonemap_and_readback(smoke_ctx, oracle, smoke_grid):
    W, H = 80, 56
    p = readme_scene(smoke_grid, W, H, bounces=8)
    smoke_ctx.resize(W, H)
    smoke_ctx.trace(p, 1, 4)
    lin = smoke_ctx.download_color()
    assert np.array_equal(smoke_ctx.download_color(3), lin[..., :3])
    # draw(): framebuffer with / without tonemapping (renderer.cpp:147-153)
    smoke_ctx.tonemap(3.0, 2.0, in_place=False, tonemapping=True)
    fb = smoke_ctx.download_framebuffer()
    assert np.abs(fb.astype(np.int32) - oracle.draw(lin, 3.0, 2.0, True).astype(np.int32)).max() <= 1   # powf ulp -> at most 1 code
    smoke_ctx.tonemap(3.0, 2.0, in_place=False, tonemapping=False)
    assert np.array_equal(smoke_ctx.download_framebuffer(), oracle.color_to_ldr(lin))
    assert np.array_equal(smoke_ctx.download_color_ldr(), oracle.color_to_ldr(lin))
    # offline CLI path: tonemap.glsl in place (main.cpp:540-550)
    smoke_ctx.tonemap(3.0, 2.0, in_place=True)
    got = smoke_ctx.download_color()
    want = oracle.tonemap_inplace(lin, 3.0, 2.0)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_loose_anchor_on_reference_example_image(smoke_ctx, smoke_grid):
    """imgs/example.jpg (README command, 4096 spp) box-filtered to 64x64 is the only rendered golden the
    reference ships (lossy JPEG): a loose end-to-end anchor, RMSE < 8/255 on the tonemapped image."""
    import os
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "example_64x64_rgb8.npy")).astype(np.float32)
    W = H = 256
    p = readme_scene(smoke_grid, W, H)
    smoke_ctx.resize(W, H)
    smoke_ctx.trace(p, 1, 256)
    smoke_ctx.tonemap(3.0, 2.0, in_place=False, tonemapping=True)
    fb = smoke_ctx.download_framebuffer()[::-1, :, :3].astype(np.float32)     # PNG order: top row first
    small = fb.reshape(64, 4, 64, 4, 3).mean(axis=(1, 3))
    err = rmse(small, want)
    print("RMSE vs imgs/example.jpg (8-bit codes):", err)
    assert err < 8.0


def test_error_paths(ctx, smoke_grid):
    from volren_b200 import Context, VrbError, _capi
    c = Context(0)
    try:
        p = readme_scene(smoke_grid, 16, 16)
        with pytest.raises(VrbError) as e:
            c.trace(p)
        assert e.value.status == _capi.VRB_ERR_STATE
        c.resize(16, 16)
        with pytest.raises(VrbError):
            c.trace(p)          # no grid
        c.grid_upload_brick(smoke_grid)
        with pytest.raises(VrbError) as e:
            c.trace(p)          # no environment
        assert "environment" in str(e.value)
        c.env_upload(np.ones((2, 4, 3), np.float32))
        p.use_transferfunc = 1
        with pytest.raises(VrbError):
            c.trace(p)          # TF requested, none uploaded
        p.use_transferfunc = 0
        with pytest.raises(VrbError):
            c.trace(p, 0, 1)    # current_sample is 1-based
        c.trace(p, 1, 1)
        c.sync()
    finally:
        c.close()


@pytest.mark.parametrize("use_tf", [False, True])
def test_persistent_kernel_equals_simple_kernel(smoke_ctx, oracle, smoke_grid, lut_raw, use_tf):
    """The production persistent-thread kernel and the straightforward one-thread-per-pixel kernel evaluate the
    same paths with the same random numbers: images agree to fp32 rounding of the (rare) fractional shadow
    transmittance, event counters agree exactly."""
    W, H, SPP = 150, 90, 6       # deliberately not multiples of the 8x4 ticket tiles
    lut, _ = oracle.lut_upload(lut_raw)
    smoke_ctx.tf_upload(lut)
    p = default_scene(smoke_grid, W, H, bounces=128, use_tf=True) if use_tf else readme_scene(smoke_grid, W, H)
    smoke_ctx.resize(W, H)
    out, cnt = [], []
    for kind in (1, 2):      # 1 = simple (IEEE math), 2 = lane-resident persistent (IEEE math); 0 = ray pool with fast math is the default elsewhere
        smoke_ctx.set_kernel(kind)
        smoke_ctx.clear()
        smoke_ctx.set_counting(True)
        smoke_ctx.trace(p, 3, SPP)
        out.append(smoke_ctx.download_color())
        cnt.append(smoke_ctx.get_counters().as_dict())
        smoke_ctx.set_counting(False)
    smoke_ctx.set_kernel(0)
    assert cnt[0] == cnt[1]
    same = np.all(out[0] == out[1], axis=-1).mean()
    print(f"bit-identical pixels persistent vs simple (tf={use_tf}): {same:.5f}")
    assert same > (0.99 if use_tf else 0.999)     # TF: the in-brick trilinear fast path contracts its FMAs differently
    assert np.allclose(out[0], out[1], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("use_tf", [False, True])
def test_ray_pool_kernel_is_bit_identical_to_the_lane_resident_kernel(smoke_ctx, oracle, smoke_grid, lut_raw, use_tf):
    """k_trace_pool (vr_trace_pool.cuh: 64 path states per warp in shared memory, one stage per scheduler iteration) is a
    different SCHEDULE of the same per-path arithmetic as the lane-resident persistent kernel (kind 3, round 1's production
    kernel): images and event counters are equal bit for bit, counting and non-counting builds, odd image sizes, a sample
    range that is not a power of two."""
    W, H, SPP = 150, 90, 6
    lut, _ = oracle.lut_upload(lut_raw)
    smoke_ctx.tf_upload(lut)
    p = default_scene(smoke_grid, W, H, bounces=128, use_tf=True) if use_tf else readme_scene(smoke_grid, W, H)
    smoke_ctx.resize(W, H)
    try:
        out, cnt = [], []
        for kind in (3, 0):
            smoke_ctx.set_kernel(kind)
            smoke_ctx.clear()
            smoke_ctx.set_counting(True)
            smoke_ctx.trace(p, 3, SPP)
            cnt.append(smoke_ctx.get_counters().as_dict())
            smoke_ctx.set_counting(False)
            smoke_ctx.trace(p, 3 + SPP, SPP)          # and the non-counting build on top
            out.append(smoke_ctx.download_color())
        assert cnt[0] == cnt[1]
        assert cnt[0]["n_samples"] == W * H * SPP
        assert np.array_equal(out[0], out[1])
    finally:
        smoke_ctx.set_kernel(0)


IEEE_REPLAY_MIN = 0.999     # fraction of pixels within 1e-3 of the oracle, same seeds, 1 spp


@pytest.mark.parametrize("use_tf,bounces", [(False, 3), (True, 3), (False, 128), (True, 128)], ids=["notf-b3", "tf-b3", "notf-b128", "tf-b128"])
def test_ieee_kernels_replay_the_oracle(smoke_ctx, oracle, smoke_grid, env_rgb, env_pyramid, lut_raw, use_tf, bounces):
    """Kernels 1 (one thread per pixel) and 2 (lane-resident persistent), IEEE math without FMA contraction
    (csrc/vrb200_strict.cu), against the oracle -- which equals the reference's own GLSL compiled as C++ bit for bit
    (tests/test_glsl_ref.py). Same TEA/LCG seeds, 1 spp: the device follows the oracle path for path except where the
    last bit of a libm function (CUDA logf / sincosf / atan2f / acosf vs glibc) flips a comparison. Pass mark: >= 0.999 of
    the pixels within 1e-3 (relative to max(|ref|, 1e-3)), for 3 and for 128 bounces."""
    W, H = 128, 96
    lut, _ = oracle.lut_upload(lut_raw)
    smoke_ctx.tf_upload(lut)
    p = default_scene(smoke_grid, W, H, bounces=bounces, use_tf=True) if use_tf else readme_scene(smoke_grid, W, H, bounces=bounces)
    want, cnt_want = oracle.trace(oracle.make_scene(smoke_grid, env_rgb, env_pyramid, lut=lut), p, 1, 1)
    smoke_ctx.resize(W, H)
    try:
        for kind in (1, 2):
            smoke_ctx.set_kernel(kind)
            smoke_ctx.clear()
            smoke_ctx.set_counting(True)
            smoke_ctx.trace(p, 1, 1)
            got = smoke_ctx.download_color()
            cnt = smoke_ctx.get_counters().as_dict()
            smoke_ctx.set_counting(False)
            ok = np.all(rel_err(got, want, eps=1e-3) < 1e-3, axis=-1)
            exact = np.all(got == want, axis=-1).mean()
            print(f"IEEE kernel {kind} vs oracle (tf={use_tf}, bounces={bounces}): within 1e-3 {ok.mean():.5f}, bit-identical {exact:.5f}")
            assert ok.mean() >= IEEE_REPLAY_MIN, (kind, ok.mean())
            w = cnt_want.as_dict()
            for k in ("n_maj", "n_dens", "n_nee", "n_env", "n_real"):      # the event counts follow: a flipped path changes them by a few events
                assert abs(cnt[k] - w[k]) <= 0.002 * max(w[k], 1) + 2, (kind, k, cnt[k], w[k])
    finally:
        smoke_ctx.set_kernel(0)


def test_scheduling_options_do_not_change_the_image(smoke_ctx, oracle, smoke_grid, lut_raw):
    """Heaviest-tiles-first order, screen-space box culling and the pass size are pure scheduling: bit-identical images."""
    W, H = 200, 120                               # not a multiple of the 8x4 tile: border tiles
    lut, _ = oracle.lut_upload(lut_raw)
    smoke_ctx.tf_upload(lut)
    smoke_ctx.resize(W, H)
    for p in (default_scene(smoke_grid, W, H, bounces=8, use_tf=True),                  # environment hidden: culling is active
              readme_scene(smoke_grid, W, H, bounces=8)):
        images = []
        for lpt, cull, npass in ((1, 1, 16), (0, 0, 16), (1, 1, 3), (0, 1, 1)):
            smoke_ctx.set_option("lpt", lpt); smoke_ctx.set_option("cull", cull); smoke_ctx.set_option("pass", npass)
            smoke_ctx.clear()
            smoke_ctx.trace(p, 1, 5)
            smoke_ctx.trace(p, 6, 5)              # the second launch of the same view uses the measured tile costs
            images.append(smoke_ctx.download_color())
        for img in images[1:]:
            assert np.array_equal(images[0], img)
        # a pixel row far from the volume is exactly zero with the environment hidden
        if not p.show_environment:
            assert np.all(images[0][0] == 0)
    smoke_ctx.set_option("lpt", 1); smoke_ctx.set_option("cull", 1); smoke_ctx.set_option("pass", 32)
    with pytest.raises(Exception):
        smoke_ctx.set_option("nonsense", 1)


def test_screen_space_brick_mask_is_exact(smoke_ctx, oracle, smoke_grid, lut_raw):
    """Hidden environment: tiles onto which no brick with a positive majorant projects get no tickets (k_tile_mask) and
    pixels outside the box's screen rectangle are never traced. Both are exact: the image equals the unculled one bit for
    bit -- off-centre views, a camera INSIDE the volume (mask unusable), a non-monotone LUT (TF majorant does not bound the
    density: mask off), the non-TF kernel with a hidden environment, sum mode."""
    from volren_b200 import scene
    W, H = 232, 136
    lut, _ = oracle.lut_upload(lut_raw)
    bad_lut = lut.copy()
    bad_lut[:, 3] = bad_lut[::-1, 3]                     # decreasing alpha
    smoke_ctx.resize(W, H)

    def params(cam, use_tf, seed=7):
        st = scene.RenderSettings(bounces=6, seed=seed, use_transferfunc=use_tf, show_environment=False)
        scene.scale_and_move_to_unit_cube(smoke_grid.matrix(), smoke_grid.index_extent(), st)
        return scene.make_params(W, H, cam, st, smoke_grid.matrix(), smoke_grid.index_extent(), smoke_grid.min_maj)

    cams = [scene.Camera(),                                                                        # default pose
            scene.Camera(pos=np.array([.9, .5, .3], np.float32), dir=scene.normalize([-1, -.2, -.6]), fov_degree=55.0),   # volume off-centre, partly outside
            scene.Camera(pos=np.array([.05, .1, .02], np.float32), dir=scene.normalize([.3, 1, .2])),                    # inside the volume
            scene.Camera(pos=np.array([1, 0, 1], np.float32), dir=scene.normalize([1, 0, 1]))]                           # looking away: nothing visible
    for table in (lut, bad_lut):
        smoke_ctx.tf_upload(table)
        for cam in cams:
            for use_tf in (True, False):
                for accum in (0, 1):
                    p = params(cam, use_tf)
                    images = []
                    for cull in (1, 0):
                        smoke_ctx.set_option("cull", cull)
                        smoke_ctx.clear()
                        smoke_ctx.trace(p, 1, 3, accum_mode=accum)
                        smoke_ctx.trace(p, 4, 2, accum_mode=accum)
                        images.append(smoke_ctx.download_color())
                    assert np.array_equal(images[0], images[1]), (cam, use_tf, accum)
    smoke_ctx.set_option("cull", 1)
    smoke_ctx.tf_upload(lut)
    assert images[0].max() == 0            # the last view looks away from the volume
    # the default TF view is not trivially empty
    smoke_ctx.clear()
    smoke_ctx.trace(params(cams[0], True), 1, 4)
    assert smoke_ctx.download_color()[..., 3].max() > 0


def test_cached_tile_order_follows_the_grid_contents(ctx, oracle, lut_raw):
    """The brick mask and the tile order are cached per (view, grid contents) and re-used for several passes
    (vrb200.cu ORDER_REUSE). Re-uploading DIFFERENT voxels into the same slot, same shape, same view, must rebuild
    them: with the environment hidden the culled image still equals the unculled one bit for bit, for more passes than
    the re-use window."""
    from volren_b200 import scene
    W, H, N = 160, 96, 40
    lut, _ = oracle.lut_upload(lut_raw)
    ctx.tf_upload(lut)
    ctx.resize(W, H)
    z, y, x = np.mgrid[0:N, 0:N, 0:N].astype(np.float32) / N

    def blob(cx, cy):
        f = np.exp(-(((x - cx) / .12) ** 2 + ((y - cy) / .12) ** 2 + ((z - .5) / .2) ** 2))
        return (np.clip(f - 0.2, 0, 1) / 0.8 * 255).astype(np.uint8)

    class G:            # what helpers.default_scene needs of a grid
        min_maj = (0.0, 1.0)
        def matrix(self): return np.eye(4, dtype=np.float32)
        def index_extent(self): return (N, N, N)

    from helpers import default_scene
    p = default_scene(G(), W, H, bounces=6, use_tf=True)
    assert not p.show_environment
    ctx.set_option("pass", 2)
    try:
        results = []
        for cx, cy in ((.25, .3), (.75, .7)):            # the two volumes light up different tiles
            vox = blob(cx, cy)
            per_cull = []
            for cull in (1, 0):
                ctx.set_option("cull", cull)
                ctx.grid_clear()
                ctx.grid_build_from_dense(vox, 0.0, 1.0)
                ctx.clear()
                ctx.trace(p, 1, 24)                       # 12 passes of 2 samples: beyond the re-use window
                per_cull.append(ctx.download_color())
            assert np.array_equal(per_cull[0], per_cull[1])
            results.append(per_cull[0])
        a, b = results[0][..., 3] > 0, results[1][..., 3] > 0
        assert a.any() and b.any() and (a & ~b).any() and (b & ~a).any()
    finally:
        ctx.set_option("cull", 1); ctx.set_option("pass", 32)
