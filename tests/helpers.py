"""Shared scene builders for the tests (host-side only; kernels are reached through the C ABI)."""
import numpy as np

from volren_b200 import scene


def synth_cases():
    import importlib.util
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
    spec = importlib.util.spec_from_file_location("make_golden", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.synth_cases()


def readme_scene(grid, w, h, bounces=128, seed=42, show_environment=True, use_tf=False, fov=40.0):
    """README offline command (reference README.md:72-73): albedo .8, phase .3, density 100, env_strength 3,
    env_rot 270, cam_fov 40, default camera (main.cpp:458-459), volume scaled to the unit cube."""
    s = scene.RenderSettings(bounces=bounces, seed=seed, albedo=(.8, .8, .8), phase=.3, env_strength=3.0,
                             env_transform=scene.rotate_y(270), show_environment=show_environment, use_transferfunc=use_tf)
    scene.scale_and_move_to_unit_cube(grid.matrix(), grid.index_extent(), s)
    s.density_scale = 100.0
    cam = scene.Camera(fov_degree=fov)
    return scene.make_params(w, h, cam, s, grid.matrix(), grid.index_extent(), grid.min_maj)


def default_scene(grid, w, h, bounces=100, seed=42, use_tf=False, show_environment=None, index_extent=None, **kw):
    """`./volren vol env [lut]` with defaults: unit-cube scale, fov 70, albedo .9, g 0."""
    s = scene.RenderSettings(bounces=bounces, seed=seed, use_transferfunc=use_tf,
                             show_environment=(not use_tf) if show_environment is None else show_environment, **kw)
    ext = index_extent if index_extent is not None else grid.index_extent()
    scene.scale_and_move_to_unit_cube(grid.matrix(), ext, s)
    return scene.make_params(w, h, scene.Camera(), s, grid.matrix(), ext, grid.min_maj)


def blob_volume(n=48, seed=3):
    """Small synthetic u8 density grid (two soft blobs + noise) -> (voxels[z][y][x], vmin, vmax)."""
    rng = np.random.default_rng(seed)
    z, y, x = np.mgrid[0:n, 0:n, 0:n].astype(np.float32) / n
    f = np.exp(-(((x - .4) / .18) ** 2 + ((y - .5) / .22) ** 2 + ((z - .5) / .2) ** 2))
    f += 0.7 * np.exp(-(((x - .7) / .1) ** 2 + ((y - .3) / .1) ** 2 + ((z - .6) / .12) ** 2))
    f += 0.05 * rng.random(f.shape).astype(np.float32)
    f = np.clip(f - 0.15, 0, 1)
    return (f / f.max() * 255).astype(np.uint8), 0.0, 4.0


def rel_err(a, b, eps=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), eps)


def rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))
