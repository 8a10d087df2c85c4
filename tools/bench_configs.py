"""The other BASELINE.json configs on one GPU, as measurements next to bench.py's headline (configs[1]):

  C1  smoke.brick + hdr, 1024x1024, README command (non-TF kernel, environment visible)
  C3  synthetic 1024^3 fBm cloud -> GPU brick build -> 1920x1080, density 100, albedo .8 (atlas ~ 1 GiB: HBM-resident)
  C4  synthetic 512x512x1800 CT-like grid + 256-entry RGBA LUT (TF kernel), 1920x1080
  C5  animated 256^3 fBm frames, datagen_denoise-style noisy (1..33 spp) / clean pairs at 1024x1024, parameters drawn like
      scripts/datagen_denoise.py:60-80 with random.seed(42), fp16 (N, 3, H, W) output; per frame: GPU brick build from the
      dense grid, two renders, two read-backs (--frames, --clean-spp; the named config is 64 frames x 4096 clean spp)

    python tools/bench_configs.py [--configs C1,C3,C4,C5] [--scale 1.0] [--spp 16] [--launches 4]

Per config one JSON line: brick-build time and GB/s (algorithmic bytes: 1 B/voxel read + 1 B/allocated voxel written +
8 B/brick), samples/s of the tracking kernel, event counters per sample and the algorithmic GB/s they imply.
Synthetic volumes are generated on the device with torch (no host copy of a 1 GiB grid).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import volren_b200 as vr  # noqa: E402
from volren_b200 import formats, scene  # noqa: E402
from helpers import readme_scene  # noqa: E402

A = os.path.join(ROOT, "tests", "golden", "assets")
RANK, WORLD = 0, 1


def fbm_cloud(n, seed=42, octaves=5, base=4, threshold=0.30):
    """5-octave value-noise fBm (lacunarity 2, gain .5, base frequency 4) x smoothstep radial falloff, max(0, f - 0.45) -> u8 [z][y][x]."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    acc = torch.zeros((1, 1, n, n, n), device="cuda", dtype=torch.float16)
    amp, norm = 0.5, 0.0
    for o in range(octaves):
        f = base * 2 ** o
        lattice = torch.rand((1, 1, f + 1, f + 1, f + 1), device="cuda", generator=g, dtype=torch.float32).half()
        acc += amp * F.interpolate(lattice, size=(n, n, n), mode="trilinear", align_corners=True)
        norm += amp
        amp *= 0.5
    acc /= norm
    ax = torch.linspace(-1, 1, n, device="cuda", dtype=torch.float16)
    r = torch.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2)
    t = ((1.0 - r) / 0.6).clamp(0, 1)
    fall = t * t * (3 - 2 * t)
    d = (acc[0, 0] * fall - threshold).clamp_(min=0)
    d /= d.max()
    return (d * 255).round().to(torch.uint8).contiguous()


def fbm_frames(n, frames, seed=42, octaves=5, base=4, threshold=0.30):
    """The C5 animation: one fBm field taller in y than the grid; frame t is the window shifted by 0.05 * t of the grid edge."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    ny = n + int(round(0.05 * (frames - 1) * n)) + 1
    acc = torch.zeros((1, 1, n, ny, n), device="cuda", dtype=torch.float16)
    amp, norm = 0.5, 0.0
    for o in range(octaves):
        f = base * 2 ** o
        fy = max(2, int(round(f * ny / n)))
        lattice = torch.rand((1, 1, f + 1, fy + 1, f + 1), device="cuda", generator=g, dtype=torch.float32).half()
        acc += amp * F.interpolate(lattice, size=(n, ny, n), mode="trilinear", align_corners=True)
        norm += amp
        amp *= 0.5
    acc /= norm
    ax = torch.linspace(-1, 1, n, device="cuda", dtype=torch.float16)
    r = torch.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2)
    t = ((1.0 - r) / 0.6).clamp(0, 1)
    fall = t * t * (3 - 2 * t)
    out = []
    for k in range(frames):
        y0 = int(round(0.05 * k * n))
        d = (acc[0, 0, :, y0:y0 + n, :] * fall - threshold).clamp_(min=0)
        d = d / d.max().clamp(min=1e-3)
        out.append((d * 255).round().to(torch.uint8).contiguous())
    return out


def run_c5(ctx, frames, clean_spp, n=256, W=1024, H=1024):
    """datagen_denoise.py:60-130 on synthetic frames: parameters drawn in the script's order with random.seed(42)."""
    import math
    import random
    random.seed(42)

    def sphere():
        z = 1.0 - 2.0 * random.random()
        r = math.sqrt(max(0.0, 1.0 - z * z))
        phi = 2.0 * math.pi * random.random()
        return np.array([r * math.cos(phi), r * math.sin(phi), z], np.float32)

    def draw():
        p = {}
        p["samples"] = random.randint(1, 32 + 1); p["max_bounces"] = random.randint(1, 128 + 1)
        p["seed_input"] = random.randint(0, 2 ** 31); p["seed_target"] = random.randint(0, 2 ** 31)
        p["env_strength"] = 0.5 + random.random() * 10; p["env_show"] = random.random() < 0.1
        p["lut_n_bins"] = random.randint(2, 32 + 1); p["lut_window_left"] = random.random() * 0.25; p["lut_window_width"] = random.random()
        p["vol_albedo"] = (random.random(), random.random(), random.random()); p["vol_phase"] = -0.9 + random.random() * 1.8
        p["vol_density_scale"] = 0.01 + random.random() * 5
        p["cam_pos_sample"] = sphere(); p["cam_dir_sample"] = sphere(); p["cam_fov"] = 25 + random.random() * 70
        return p

    vols = fbm_frames(n, frames)
    plist = [draw() for _ in range(frames)]
    ctx.resize(W, H)
    inputs = np.zeros((frames, 3, H, W), np.float16)
    targets = np.zeros((frames, 3, H, W), np.float16)
    g = DenseInfo((n, n, n))
    samples = 0
    build_ms = []
    import time
    torch.cuda.synchronize()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    for i, q in enumerate(plist):
        if i % WORLD != RANK:
            continue
        build_ms.append(timed(lambda: ctx.grid_build_from_dense_device(vols[i].data_ptr(), (n, n, n), 0.0, 1.0)))
        s = scene.RenderSettings(bounces=q["max_bounces"], albedo=q["vol_albedo"], phase=q["vol_phase"], env_strength=q["env_strength"],
                                 show_environment=q["env_show"], use_transferfunc=False)
        scene.scale_and_move_to_unit_cube(g.matrix(), g.index_extent(), s)      # Renderer::commit
        s.density_scale = q["vol_density_scale"]
        M = np.asarray(s.volume_transform, np.float32) @ g.matrix()
        bb_min, bb_max = (M @ np.array([0, 0, 0, 1], np.float32))[:3], (M @ np.array([n, n, n, 1], np.float32))[:3]
        center = bb_min + (bb_max - bb_min) * 0.5
        radius = float(np.linalg.norm(bb_max - center))
        pos = center + q["cam_pos_sample"] * radius
        cam = scene.Camera(pos=pos.astype(np.float32), dir=scene.normalize(center + q["cam_dir_sample"] * radius * 0.1 - pos), fov_degree=q["cam_fov"])
        for seed, spp, dst in ((q["seed_input"], q["samples"], inputs), (q["seed_target"], clean_spp, targets)):
            s.seed = seed if seed < 2 ** 31 else seed - 2 ** 32
            p = scene.make_params(W, H, cam, s, g.matrix(), g.index_extent(), g.min_maj)
            ctx.clear()
            ctx.trace(p, 1, spp)
            rgb = ctx.download_color(3)                                      # fbo_data(): linear RGB fp32
            dst[i] = np.transpose(np.flip(rgb, axis=0).astype(np.float16), [2, 1, 0])     # datagen_denoise.py:113-114 (square images)
            samples += W * H * spp
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    part = "single GPU (tile partition: volren_b200.multigpu, tests/test_multigpu_gloo.py)"
    if WORLD > 1:
        import torch.distributed as dist
        t = torch.tensor([dt, float(samples)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, samples = float(tmax[0]), int(t[1])
        part = f"{WORLD} GPUs, frames dealt round-robin (every frame is an independent job: own volume, own parameters), max wall time over ranks"
        if RANK != 0:
            return
    print(json.dumps(dict(config=f"C5 {frames} animated {n}^3 fBm frames, datagen_denoise-style pairs (noisy 1..33 spp, clean {clean_spp} spp; named: 64 frames, 4096 spp)",
                          resolution=[W, H], frames=frames, clean_spp=clean_spp, wall_s=dt, frames_per_s=frames / dt, samples=samples,
                          samples_per_s=samples / dt, brick_build_ms_median=float(np.median(build_ms)), output="fp16 (N,3,H,W) noisy + clean",
                          finite=bool(np.isfinite(inputs.astype(np.float32)).all() and np.isfinite(targets.astype(np.float32)).all()),
                          mean_clean=float(targets.astype(np.float32).mean()), n_gpus=WORLD, partition=part)), flush=True)


def ct_phantom(w, h, d, seed=42):
    """Nested ellipsoid body (.25), cylinder bones (.8), sphere organs (.45), + U(-.02, .02) noise, air 0 -> u8 [z][y][x]."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.linspace(-1, 1, d, device="cuda", dtype=torch.float16)[:, None, None]
    y = torch.linspace(-1, 1, h, device="cuda", dtype=torch.float16)[None, :, None]
    x = torch.linspace(-1, 1, w, device="cuda", dtype=torch.float16)[None, None, :]
    v = torch.zeros((d, h, w), device="cuda", dtype=torch.float16)
    body = (x / 0.8) ** 2 + (y / 0.6) ** 2 + (z / 0.95) ** 2 < 1
    v[body] = 0.25
    for cx, cy in ((-0.3, 0.0), (0.3, 0.0), (0.0, 0.35)):
        v[((x - cx) ** 2 + (y - cy) ** 2 < 0.006).expand_as(v) & body] = 0.8
    for cx, cy, cz, rr in ((0.25, -0.2, 0.3, 0.2), (-0.3, 0.15, -0.2, 0.25), (0.0, -0.1, -0.6, 0.18)):
        v[((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 < rr * rr)] = 0.45
    noise = (torch.rand((d, h, w), device="cuda", generator=g, dtype=torch.float32).half() - 0.5) * 0.04
    v = torch.where(v > 0, (v + noise).clamp(0, 1), v)
    return (v * 255).round().to(torch.uint8).contiguous()


def turbo_like_lut(n=256):
    """RGBA LUT with alpha = i / n (transferfunc.cpp:69-77 shape); colours from a simple blue-green-red ramp."""
    f = np.arange(n, dtype=np.float32) / n
    rgb = np.stack([np.clip(1.5 - np.abs(4 * f - 3), 0, 1), np.clip(1.5 - np.abs(4 * f - 2), 0, 1), np.clip(1.5 - np.abs(4 * f - 1), 0, 1)], -1)
    return np.concatenate([rgb, f[:, None]], -1).astype(np.float32)


class DenseInfo:
    def __init__(self, dims_whd):
        self.dims, self.min_maj = dims_whd, (0.0, 1.0)

    def matrix(self):
        return np.eye(4, dtype=np.float32)

    def index_extent(self):
        return self.dims


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def alg_bytes(c, tf):
    per_maj, per_dens = (36, 104) if tf else (4, 9)
    return per_maj * c["n_maj"] + per_dens * c["n_dens"] + 9 * c["n_emis"] + 200 * c["n_nee"] + 100 * c["n_env"] + 32 * c["n_samples"]


def run(name, ctx, params, W, H, tf, spp, launches, extra):
    ctx.resize(W, H)
    ctx.set_counting(True)
    ctx.trace(params, 1, 2)
    c = ctx.get_counters().as_dict()
    ctx.set_counting(False)
    n = c["n_samples"]
    per = {k: v / n for k, v in c.items()}
    ctx.clear()
    ms = [timed(lambda i=i: ctx.trace(params, 1 + i * spp, spp)) for i in range(launches)]
    best = min(ms[1:]) if launches > 1 else ms[0]
    sps = W * H * spp / (best * 1e-3)
    out = dict(config=name, resolution=[W, H], spp_per_launch=spp, ms_per_launch=best, samples_per_s=sps, counters_per_sample=per,
               algorithmic_bytes_per_sample=alg_bytes(c, tf) / n, algorithmic_GBps=alg_bytes(c, tf) / n * sps / 1e9, **extra)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C3,C4")
    ap.add_argument("--scale", type=float, default=1.0, help="scales the synthetic grid edge (1.0 = the named sizes)")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--launches", type=int, default=4)
    ap.add_argument("--frames", type=int, default=8, help="C5: number of animation frames (named config: 64)")
    ap.add_argument("--clean-spp", type=int, default=256, help="C5: samples of the clean image (named config: 4096)")
    a = ap.parse_args()
    # under torchrun (C5 only): one process per GPU, the animation frames are dealt round-robin to the ranks
    global RANK, WORLD
    RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        assert a.configs == "C5", "multi-rank runs are for C5 (frame-parallel); bench.py covers the spp-sliced scaling"
    ctx = vr.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    env = formats.load_hdr(os.path.join(A, "table_mountain_2_puresky_1k.hdr"))
    ctx.env_upload(env)
    for name in a.configs.split(","):
        ctx.grid_clear()
        if name == "C1":
            grid = formats.load_brick(os.path.join(A, "smoke.brick"))
            ctx.grid_upload_brick(grid)
            run("C1 smoke.brick + hdr, README command", ctx, readme_scene(grid, 1024, 1024), 1024, 1024, False, a.spp, a.launches, {})
            continue
        if name == "C5":
            run_c5(ctx, a.frames, a.clean_spp)
            continue
        if name == "C3":
            n = int(1024 * a.scale) // 8 * 8
            vox = fbm_cloud(n)
            dims, tf, label = (n, n, n), False, f"C3 synthetic {n}^3 fBm cloud"
        else:
            w, d = int(512 * a.scale) // 8 * 8, int(1800 * a.scale) // 8 * 8
            vox = ct_phantom(w, w, d)
            dims, tf, label = (w, w, d), True, f"C4 synthetic {w}x{w}x{d} CT phantom + 256-entry LUT"
        torch.cuda.synchronize()
        torch.cuda.empty_cache()          # the generator's temporaries go back to the driver before the pool grows
        build_first_ms = timed(lambda: ctx.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0))
        build_ms = min(timed(lambda: ctx.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0)) for _ in range(3))
        nb, _, count = ctx.grid_info()
        n_vox = dims[0] * dims[1] * dims[2]
        n_bricks = nb[0] * nb[1] * nb[2]
        build_bytes = n_vox + count * 512 + 8 * n_bricks
        extra = dict(grid=list(dims), n_bricks=list(nb), bricks_allocated=count, atlas_MiB=count * 512 / 2 ** 20,
                     build_first_ms=build_first_ms, build_ms=build_ms, build_algorithmic_GBps=build_bytes / (build_ms * 1e-3) / 1e9)
        del vox
        torch.cuda.empty_cache()
        s = scene.RenderSettings(bounces=128, albedo=(.8, .8, .8), use_transferfunc=tf, show_environment=not tf)
        g = DenseInfo(dims)
        scene.scale_and_move_to_unit_cube(g.matrix(), g.index_extent(), s)
        if not tf:
            s.density_scale = 100.0
        else:
            ctx.tf_upload(turbo_like_lut())
        p = scene.make_params(1920, 1080, scene.Camera(), s, g.matrix(), g.index_extent(), g.min_maj)
        run(label, ctx, p, 1920, 1080, tf, a.spp, a.launches, extra)
    ctx.close()


if __name__ == "__main__":
    main()
