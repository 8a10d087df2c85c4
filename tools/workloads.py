"""The BASELINE.json configs as reproducible workloads, shared by bench.py, tools/ and the GPU tests.

  C1  data/smoke.brick + hdr, 1024x1024, README command (non-TF kernel, environment visible)
  C2  data/smoke.brick + hdr + data/lut.txt (TF kernel, environment hidden by load_transferfunc), 1920x1080
  C3  synthetic N^3 fBm cloud (named size 1024^3) -> DenseGrid -> GPU brick build, density 100, albedo .8, non-TF, 1920x1080
  C4  synthetic 512x512x1800 CT phantom + 256-entry Turbo LUT with alpha = i/256 (TF kernel), 1920x1080
  C5  animated 256^3 fBm frames, datagen_denoise-style noisy/clean pairs at 1024x1024

Synthetic volumes are generated ON THE DEVICE with torch (no host copy of a 1 GiB grid), deterministically (seed 42); the
generators return u8 tensors [z][y][x]. `scale` shrinks the grid edge for parity tests (the oracle needs the voxels on the host).
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ASSETS = os.path.join(ROOT, "tests", "golden", "assets")

from volren_b200 import formats, scene  # noqa: E402


class DenseInfo:
    """What the host knows about a DenseGrid source (voldata/grid_dense.cpp): identity transform, value range [0, 1]."""

    def __init__(self, dims_whd, min_maj=(0.0, 1.0)):
        self.dims, self.min_maj = tuple(int(d) for d in dims_whd), min_maj

    def matrix(self):
        return np.eye(4, dtype=np.float32)

    def index_extent(self):
        return self.dims


def load_assets():
    grid = formats.load_brick(os.path.join(ASSETS, "smoke.brick"))
    env = formats.load_hdr(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"))
    lut = formats.lut_for_upload(formats.load_lut_txt(os.path.join(ASSETS, "lut.txt")))
    return grid, env, lut


def readme_params(grid, w, h, bounces=128, seed=42):
    """README offline command (README.md:72-73): albedo .8, phase .3, density 100, env_strength 3, env_rot 270, cam_fov 40."""
    s = scene.RenderSettings(bounces=bounces, seed=seed, albedo=(.8, .8, .8), phase=.3, env_strength=3.0,
                             env_transform=scene.rotate_y(270), show_environment=True, use_transferfunc=False)
    scene.scale_and_move_to_unit_cube(grid.matrix(), grid.index_extent(), s)
    s.density_scale = 100.0
    return scene.make_params(w, h, scene.Camera(fov_degree=40.0), s, grid.matrix(), grid.index_extent(), grid.min_maj)


def default_params(grid, w, h, bounces=128, seed=42, use_tf=False, **kw):
    """`./volren vol env [lut]` with defaults: unit-cube scale, fov 70, albedo .9, g 0; a LUT hides the environment (main.cpp:76)."""
    s = scene.RenderSettings(bounces=bounces, seed=seed, use_transferfunc=use_tf, show_environment=not use_tf, **kw)
    scene.scale_and_move_to_unit_cube(grid.matrix(), grid.index_extent(), s)
    return scene.make_params(w, h, scene.Camera(), s, grid.matrix(), grid.index_extent(), grid.min_maj)


def synthetic_params(dims, w, h, use_tf, bounces=128, seed=42):
    """C3 / C4: default camera, unit cube; C3 density 100 + albedo .8, C4 unit-cube density + LUT."""
    s = scene.RenderSettings(bounces=bounces, seed=seed, albedo=(.8, .8, .8), use_transferfunc=use_tf, show_environment=not use_tf)
    g = DenseInfo(dims)
    scene.scale_and_move_to_unit_cube(g.matrix(), g.index_extent(), s)
    if not use_tf:
        s.density_scale = 100.0
    return scene.make_params(w, h, scene.Camera(), s, g.matrix(), g.index_extent(), g.min_maj)


# ---- device generators (torch) ---------------------------------------------------------------------------------

def fbm_cloud(n, seed=42, octaves=5, base=4, threshold=0.30):
    """5-octave value-noise fBm (lacunarity 2, gain .5, base frequency 4) x smoothstep radial falloff, max(0, f - threshold) -> u8 [z][y][x]."""
    import torch
    import torch.nn.functional as F
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
    import torch
    import torch.nn.functional as F
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


def ct_phantom(w, h, d, seed=42):
    """Nested ellipsoid body (.25), cylinder bones (.8), sphere organs (.45), + U(-.02, .02) noise, air 0 -> u8 [z][y][x]."""
    import torch
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


def turbo_lut(n=256):
    """C4's LUT: TransferFunction::colormap(tinycolormap Turbo) with alpha = i / n (transferfunc.cpp:69-77), as uploaded
    (upload_gpu leaves a monotone alpha alone)."""
    from volren_b200 import colormaps
    rgb = np.array([colormaps.get_color(i / float(n), "Turbo") for i in range(n)], np.float32)
    alpha = (np.arange(n, dtype=np.float32) / np.float32(n)).astype(np.float32)
    return np.concatenate([rgb, alpha[:, None]], -1).astype(np.float32)


def c5_parameters(frames, seed=42):
    """datagen_denoise.py:60-80: the per-image parameter draws, in the script's order, with random.seed(42)."""
    import random
    random.seed(seed)

    def sphere():
        z = 1.0 - 2.0 * random.random()
        r = math.sqrt(max(0.0, 1.0 - z * z))
        phi = 2.0 * math.pi * random.random()
        return np.array([r * math.cos(phi), r * math.sin(phi), z], np.float32)

    out = []
    for _ in range(frames):
        p = {}
        p["samples"] = random.randint(1, 32 + 1); p["max_bounces"] = random.randint(1, 128 + 1)
        p["seed_input"] = random.randint(0, 2 ** 31); p["seed_target"] = random.randint(0, 2 ** 31)
        p["env_strength"] = 0.5 + random.random() * 10; p["env_show"] = random.random() < 0.1
        p["lut_n_bins"] = random.randint(2, 32 + 1); p["lut_window_left"] = random.random() * 0.25; p["lut_window_width"] = random.random()
        p["vol_albedo"] = (random.random(), random.random(), random.random()); p["vol_phase"] = -0.9 + random.random() * 1.8
        p["vol_density_scale"] = 0.01 + random.random() * 5
        p["cam_pos_sample"] = sphere(); p["cam_dir_sample"] = sphere(); p["cam_fov"] = 25 + random.random() * 70
        out.append(p)
    return out


def c5_frame_params(q, n, w, h, seed):
    """The uniform block datagen_denoise.py:85-108 sets up for one image of an n^3 frame."""
    g = DenseInfo((n, n, n))
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
    s.seed = seed if seed < 2 ** 31 else seed - 2 ** 32
    return scene.make_params(w, h, cam, s, g.matrix(), g.index_extent(), g.min_maj)


def algorithmic_bytes(c, use_tf):
    """bytes the algorithm must touch, from the event counters."""
    per_maj, per_dens = (4 + 32, 8 * 9 + 32) if use_tf else (4, 9)
    return (per_maj * c["n_maj"] + per_dens * c["n_dens"] + 9 * c["n_emis"] + 200 * c["n_nee"] + 100 * c["n_env"] + 32 * c["n_samples"])
