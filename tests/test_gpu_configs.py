"""GPU parity (through the C ABI) on grids SHAPED like the benchmarked BASELINE.json configs -- not only smoke.brick:

  C3  a 512^3 fBm cloud built on the GPU (64^3 bricks, non-TF kernel, environment visible, density 100)
  C4  the full-size 512x512x1800 CT phantom (64x64x232 bricks after padding) with the 256-entry Turbo LUT (TF kernel)
  C5  a 4-frame animated volume, every frame with its own grid, rendered at params.frame = 0..3
  CLI `--vol_crop_min/max` + `--vol_rot_y` (clip planes and a rotated, translation-free volume transform, main.cpp:417-429)
  a 1024x1024x2048 grid whose brick-linear atlas is 2 GiB (byte offsets beyond 2^31) and whose decoded blocks are 12 GB

Tiers as in test_gpu_render.py: T1 deterministic mode per-pixel rel. error < 1e-3 vs the fp64 oracle; T3 RMSE(gpu, oracle) <
RMSE(oracle seed A, oracle seed B) at equal spp on a small frame, event counters within 1 %. The oracle renders the brick
grid downloaded from the device; that the device build equals the oracle's own build bit for bit is asserted as well.
"""
import os
import sys

import numpy as np
import pytest

from helpers import rel_err, rmse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def _t1(ctx, oracle, sc, p):
    ctx.trace_deterministic(p)
    got = ctx.download_color()
    want = oracle.trace_deterministic(sc, p)
    err = rel_err(got, want, eps=1e-3)
    assert err.max() < 1e-3, (err.max(), np.unravel_index(err.argmax(), err.shape))
    return want


def _t3(ctx, oracle, sc, mk, spp, keys=("n_maj", "n_dens", "n_nee", "n_env", "n_real")):
    ref_a, cnt_a = oracle.trace(sc, mk(42), 1, spp)
    ref_b, _ = oracle.trace(sc, mk(4242), 1, spp)
    ctx.clear()
    ctx.set_counting(True)
    ctx.trace(mk(42), 1, spp)
    got = ctx.download_color()
    cnt = ctx.get_counters().as_dict()
    ctx.set_counting(False)
    two_run = rmse(ref_a[..., :3], ref_b[..., :3])
    assert two_run > 0
    assert rmse(got[..., :3], ref_a[..., :3]) < two_run
    ctx.clear()
    ctx.trace(mk(777), 1, spp)                  # the production (non-counting) build, independent seed
    got_c = ctx.download_color()
    assert rmse(got_c[..., :3], ref_a[..., :3]) < 1.25 * two_run
    want = cnt_a.as_dict()
    assert cnt["n_samples"] == want["n_samples"]
    for k in keys:
        if want[k] > 1000:
            assert abs(cnt[k] - want[k]) / want[k] < 0.01, (k, cnt[k], want[k])
    return ref_a


def test_c3_shaped_fbm_cloud(ctx, oracle, env_rgb, env_pyramid):
    import torch
    import workloads as wl
    n = 512
    vox = wl.fbm_cloud(n)
    torch.cuda.synchronize()              # the voxels are written on torch's stream, the build reads them on the context's own
    ctx.grid_clear()
    ctx.grid_build_from_dense_device(vox.data_ptr(), (n, n, n), 0.0, 1.0)
    ctx.env_upload(env_rgb)
    g = ctx.grid_download()
    g.min_maj = (0.0, 1.0)
    assert tuple(g.n_bricks) == (64, 64, 64) and g.brick_count > 10000
    host = oracle.brick_build(vox.cpu().numpy(), 0.0, 1.0)             # the device build == the oracle's, on the C3-shaped grid too
    assert np.array_equal(g.range, host.range) and np.array_equal(g.indirection, host.indirection) and np.array_equal(g.atlas, host.atlas)
    for i in range(3):
        assert np.array_equal(g.mips[i], host.mips[i])
    del vox
    torch.cuda.empty_cache()
    sc = oracle.make_scene(g, env_rgb, env_pyramid)
    W, H = 96, 54
    ctx.resize(W, H)
    want = _t1(ctx, oracle, sc, wl.synthetic_params((n, n, n), W, H, False))
    assert want[..., 3].max() > 0.9 and (want[..., 3] == 0).mean() > 0.05          # opaque core and empty corners in one frame
    W, H = 64, 36
    ctx.resize(W, H)
    _t3(ctx, oracle, sc, lambda seed: wl.synthetic_params((n, n, n), W, H, False, seed=seed), 256)
    ctx.grid_clear()


def test_c4_shaped_ct_phantom_with_turbo_lut(ctx, oracle, env_rgb, env_pyramid):
    import torch
    import workloads as wl
    dims = (512, 512, 1800)
    vox = wl.ct_phantom(*dims)
    torch.cuda.synchronize()              # the voxels are written on torch's stream, the build reads them on the context's own
    ctx.grid_clear()
    ctx.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0)
    del vox
    torch.cuda.empty_cache()
    ctx.env_upload(env_rgb)
    lut = wl.turbo_lut()
    assert lut.shape == (256, 4) and np.all(np.diff(lut[:, 3]) > 0)       # monotone alpha: uploaded as is (transferfunc.cpp:46-53)
    ctx.tf_upload(lut)
    g = ctx.grid_download()
    g.min_maj = (0.0, 1.0)
    assert tuple(g.n_bricks) == (64, 64, 232) and tuple(g.index_extent()) == (512, 512, 1856)
    sc = oracle.make_scene(g, env_rgb, env_pyramid, lut=lut)
    W, H = 96, 54
    ctx.resize(W, H)
    _t1(ctx, oracle, sc, wl.synthetic_params(dims, W, H, True))
    W, H = 64, 36
    ctx.resize(W, H)
    ref = _t3(ctx, oracle, sc, lambda seed: wl.synthetic_params(dims, W, H, True, seed=seed), 256)
    assert ref[..., :3].max() > 0 and np.ptp(ref[..., 0] - ref[..., 2]) > 0       # the LUT colours the image
    ctx.grid_clear()


def test_c5_shaped_animated_frames(ctx, oracle, env_rgb, env_pyramid):
    """Four frames with different grids in one context; params.frame selects the grid (renderer.cpp:110)."""
    import workloads as wl
    n, frames = 64, 4
    vols = wl.fbm_frames(n, frames, threshold=0.25)
    import torch
    torch.cuda.synchronize()              # the voxels are written on torch's stream, the build reads them on the context's own
    ctx.grid_clear()
    for i, v in enumerate(vols):
        ctx.grid_build_from_dense_device(v.data_ptr(), (n, n, n), 0.0, 1.0, frame=i)
    ctx.env_upload(env_rgb)
    grids = []
    for i in range(frames):
        g = ctx.grid_download(frame=i)
        g.min_maj = (0.0, 1.0)
        grids.append(g)
        host = oracle.brick_build(vols[i].cpu().numpy(), 0.0, 1.0)
        assert np.array_equal(g.range, host.range) and np.array_equal(g.atlas, host.atlas)
    assert not np.array_equal(grids[0].range, grids[3].range)
    q = wl.c5_parameters(frames)
    W = H = 48
    ctx.resize(W, H)
    images = []
    for i in range(frames):
        sc = oracle.make_scene(grids[i], env_rgb, env_pyramid)

        def mk(seed, i=i):
            p = wl.c5_frame_params(q[i], n, W, H, seed)
            p.frame = i
            return p
        _t1(ctx, oracle, sc, mk(1))
        images.append(_t3(ctx, oracle, sc, mk, 64))
    assert rmse(images[0], images[3]) > 0
    ctx.grid_clear()


def test_clip_planes_and_rotated_volume(ctx, oracle, smoke_grid, env_rgb, env_pyramid):
    """`--vol_crop_min .1 .2 0 --vol_crop_max .9 .7 .8 --vol_rot_y 30` (main.cpp:417-429): the scene of the bit-exact CPU case
    `crop_rot` in test_glsl_ref.py, on the device."""
    import test_glsl_ref as T
    ctx.grid_clear()
    ctx.grid_upload_brick(smoke_grid)
    ctx.env_upload(env_rgb)
    a = dict(grid=smoke_grid, env=env_rgb, pyr=env_pyramid, lut=None)
    sc, p, _, _ = T.build_case("crop_rot", oracle, a)
    W, H = p.resolution[0], p.resolution[1]
    ctx.resize(W, H)
    _t1(ctx, oracle, sc, p)

    def mk(seed):
        q = p.copy()
        q.seed = seed
        return q
    _t3(ctx, oracle, sc, mk, 128)
    # and the IEEE kernel replays the clipped / rotated scene same-seed
    want, _ = oracle.trace(sc, p, 1, 1)
    ctx.set_kernel(1)
    try:
        ctx.clear()
        ctx.trace(p, 1, 1)
        ok = np.all(rel_err(ctx.download_color(), want, eps=1e-3) < 1e-3, axis=-1)
    finally:
        ctx.set_kernel(0)
    assert ok.mean() >= 0.999, ok.mean()


def test_two_gib_atlas_addressing(ctx, env_rgb):
    """1024 x 1024 x 2048 voxels with every brick allocated: the brick-linear atlas is 2 GiB and the decoded apron blocks 12 GB, so
    byte offsets pass 2^31 / 2^32. The tracer's own fetch (vrb_debug_sample_density) at points spread over the whole grid --
    nearest decode, 8-tap trilinear through the u8 atlas and through the decoded blocks -- against the values computed on the
    host from the source voxels."""
    import torch
    dims = (1024, 1024, 2048)
    w, h, d = dims
    g = torch.Generator(device="cuda").manual_seed(5)
    vox = torch.empty((d, h, w), device="cuda", dtype=torch.uint8)
    for z0 in range(0, d, 256):           # filled in slabs of 2^28 voxels
        vox[z0:z0 + 256] = torch.randint(1, 256, (256, h, w), device="cuda", dtype=torch.uint8, generator=g)   # no zero voxel: no brick stays empty
    assert int(vox[-8:].min()) >= 1 and int(vox[1024:1032].min()) >= 1
    torch.cuda.synchronize()              # the voxels are written on torch's stream, the build reads them on the context's own
    ctx.grid_clear()
    ctx.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0)
    nb, ad, count = ctx.grid_info()
    assert tuple(nb) == (128, 128, 256) and count == 128 * 128 * 256 and ad[0] * ad[1] * ad[2] == 2 ** 31
    rng = np.random.default_rng(9)
    n = 20000
    pts = (rng.random((n, 3)) * np.array([w - 2, h - 2, d - 2]) + 1).astype(np.float32)
    pts[: n // 4, 2] = (d - 300) + rng.random(n // 4).astype(np.float32) * 290            # a quarter in the last slices: the highest offsets
    near = ctx.sample_density(pts, mode=2)
    tri = ctx.sample_density(pts, mode=0)
    tri_dec = ctx.sample_density(pts, mode=1)
    assert np.array_equal(tri, tri_dec)
    ip = np.floor(pts).astype(np.int64)
    idx = torch.from_numpy((ip[:, 2] * h + ip[:, 1]) * w + ip[:, 0]).cuda()
    u8 = vox.view(-1)[idx].cpu().numpy().astype(np.float32)
    # a random brick spans (almost surely) the full code range of its 12^3 window, whose fp16 range is wide: the decode is within
    # half a range quantum + fp16 rounding of the source value
    assert np.abs(near - u8 / 255.0).max() < 1.5 / 255.0
    # trilinear: the same 8 nearest decodes, lerped (common.glsl:289-297)
    q = pts - 0.5
    f = q - np.floor(q)
    base = np.floor(q).astype(np.int64)
    acc = np.zeros(n, np.float64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                c = base + np.array([dx, dy, dz])
                tap = ctx.sample_density((c + 0.5).astype(np.float32), mode=2).astype(np.float64)
                wgt = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                acc += wgt * tap
    assert np.abs(tri - acc).max() < 1e-5
    del vox
    ctx.grid_clear()
    torch.cuda.empty_cache()


def test_float_field_to_bricks_on_the_device(ctx, oracle):
    """The C3 pipeline as SURVEY 8(d) writes it: an fp32 field in device memory -> DenseGrid(float*) (grid_dense.cpp:57-95: global
    min / max with the reference's FLT_MAX / FLT_MIN start values, 8-bit quantisation) -> BrickGrid(const Grid&), without leaving
    the GPU (vrb_grid_build_from_float_device), bit for bit against the oracle's two steps; a width that takes the float4 kernels
    and one that does not, negative values, and an all-negative field (the FLT_MIN quirk: max stays 1.17e-38)."""
    import torch
    for shape, shift in (((48, 40, 64), 0.0), ((21, 30, 50), -0.3), ((16, 16, 32), -5.0)):
        g = torch.Generator(device="cuda").manual_seed(sum(shape))
        f = torch.rand(shape, device="cuda", generator=g, dtype=torch.float32)
        f = torch.where(f < 0.6, torch.zeros_like(f), f) * 3.0 + shift
        torch.cuda.synchronize()
        d, h, w = shape
        ctx.grid_clear()
        mm = ctx.grid_build_from_float_device(f.data_ptr(), (w, h, d))
        got = ctx.grid_download()
        q, mm_want = oracle.dense_from_float(f.cpu().numpy())
        assert mm == mm_want
        want = oracle.brick_build(q, mm_want[0], mm_want[1])
        assert np.array_equal(got.range, want.range) and np.array_equal(got.indirection, want.indirection) and np.array_equal(got.atlas, want.atlas)
        for i in range(3):
            assert np.array_equal(got.mips[i], want.mips[i])
    ctx.grid_clear()
