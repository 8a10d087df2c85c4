"""Host-side scene assembly: turns renderer-level settings into the `vrb_params` uniform block, following
RendererOpenGL::trace (reference src/renderer.cpp:88-139), CameraImpl::update (cppgl camera.cpp:51-59),
Volume::AABB (voldata volume.cpp:102-107) and scale_and_move_to_unit_cube (renderer.cpp:227-242).

All matrices here are conventional math matrices (row-major numpy, column vectors); they are transposed into
glm's column-major layout when written into Params.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from ._capi import Params

f32 = np.float32


def normalize(v):
    v = np.asarray(v, f32)
    return (v / f32(np.sqrt(np.dot(v, v)))).astype(f32)


def look_at(pos, target, up):
    """glm::lookAt (right-handed)."""
    f = normalize(np.asarray(target, f32) - np.asarray(pos, f32))
    s = normalize(np.cross(f, np.asarray(up, f32)))
    u = np.cross(s, f).astype(f32)
    m = np.eye(4, dtype=f32)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(s, pos), -np.dot(u, pos), np.dot(f, pos)
    return m


def rotate_y(deg):
    """glm::rotate(mat4(1), radians(deg), (0,1,0)) as mat3 (main.cpp:379-380)."""
    a = math.radians(deg)
    c, s = f32(math.cos(a)), f32(math.sin(a))
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], f32)


def translate_scale(scale, translate):
    """glm::translate(glm::scale(mat4(1), vec3(scale)), translate)."""
    m = np.eye(4, dtype=f32)
    m[:3, :3] *= f32(scale)
    m[:3, 3] = f32(scale) * np.asarray(translate, f32)
    return m


@dataclass
class Camera:
    """cppgl CameraImpl defaults (camera.cpp:42-46) with the volren default pose (main.cpp:458-459)."""
    pos: np.ndarray = field(default_factory=lambda: np.array([1, 0, 1], f32))
    dir: np.ndarray = field(default_factory=lambda: normalize([-1, 0, -1]))
    up: np.ndarray = field(default_factory=lambda: np.array([0, 1, 0], f32))
    fov_degree: float = 70.0

    def view(self):
        d = normalize(self.dir)
        return look_at(self.pos, np.asarray(self.pos, f32) + d, normalize(self.up))


@dataclass
class RenderSettings:
    """The public data members of RendererOpenGL (renderer.h:31-62) that feed the uniform block."""
    bounces: int = 100
    seed: int = 42
    show_environment: bool = True
    albedo: tuple = (0.9, 0.9, 0.9)
    phase: float = 0.0
    density_scale: float = 1.0
    emission_scale: float = 100.0
    vol_clip_min: tuple = (0.0, 0.0, 0.0)
    vol_clip_max: tuple = (1.0, 1.0, 1.0)
    env_strength: float = 1.0
    env_transform: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=f32))
    tf_window_left: float = 0.0
    tf_window_width: float = 1.0
    use_transferfunc: bool = False
    volume_transform: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=f32))
    frame: int = 0


def scale_and_move_to_unit_cube(grid_matrix, index_extent, settings: RenderSettings):
    """renderer.cpp:227-242 for a single-frame volume; mutates settings (volume_transform, density_scale)."""
    g = np.asarray(grid_matrix, f32)
    bb_min = (g @ np.array([0, 0, 0, 1], f32))[:3]
    bb_max = (g @ np.array([*index_extent, 1], f32))[:3]
    lo = np.minimum(np.full(3, np.finfo(f32).max, f32), bb_min)
    hi = np.maximum(np.full(3, np.finfo(f32).tiny, f32), bb_max)
    extent = hi - lo
    size = f32(max(extent))
    if size != 1.0:
        settings.volume_transform = translate_scale(f32(1) / size, -lo - f32(0.5) * extent)
        settings.density_scale = float(f32(settings.density_scale) * size)
    return settings


def _set(arr, values):
    v = np.asarray(values, f32).ravel()
    for i, x in enumerate(v):
        arr[i] = float(x)


def make_params(width, height, camera: Camera, settings: RenderSettings, grid_matrix, index_extent, minorant_majorant,
                emission_matrix=None, majorant_emission=0.0) -> Params:
    s = settings
    p = Params()
    p.bounces, p.seed, p.show_environment, p.frame = int(s.bounces), int(np.int32(s.seed)), int(bool(s.show_environment)), int(s.frame)
    view = camera.view()
    _set(p.cam_pos, camera.pos)
    p.cam_fov = float(camera.fov_degree)
    _set(p.cam_transform, np.linalg.inv(view[:3, :3].astype(np.float64)).astype(f32).T)  # column-major
    model = (np.asarray(s.volume_transform, f32) @ np.asarray(grid_matrix, f32)).astype(f32)
    bb_min = (model @ np.array([0, 0, 0, 1], f32))[:3]
    bb_max = (model @ np.array([*index_extent, 1], f32))[:3]
    cmin, cmax = np.asarray(s.vol_clip_min, f32), np.asarray(s.vol_clip_max, f32)
    _set(p.vol_bb_min, bb_min + cmin * (bb_max - bb_min))
    _set(p.vol_bb_max, bb_min + cmax * (bb_max - bb_min))
    mn, mj = f32(minorant_majorant[0]), f32(minorant_majorant[1])
    ds = f32(s.density_scale)
    p.vol_minorant, p.vol_majorant = float(mn * ds), float(mj * ds)
    p.vol_inv_majorant = float(f32(1) / (mj * ds))
    _set(p.vol_albedo, s.albedo)
    p.vol_phase_g, p.vol_density_scale, p.vol_emission_scale = float(s.phase), float(ds), float(s.emission_scale)
    p.vol_emission_norm = float(f32(1) / f32(max(majorant_emission, 1e-4))) if majorant_emission > 0 else 1.0
    _set(p.vol_density_transform, model.T)
    _set(p.vol_density_inv_transform, np.linalg.inv(model.astype(np.float64)).astype(f32).T)
    if emission_matrix is not None:
        em = (np.asarray(s.volume_transform, f32) @ np.asarray(emission_matrix, f32)).astype(f32)
        p.has_emission = 1
        _set(p.vol_emission_transform, em.T)
        _set(p.vol_emission_inv_transform, np.linalg.inv(em.astype(np.float64)).astype(f32).T)
    p.use_transferfunc = int(bool(s.use_transferfunc))
    p.tf_window_left, p.tf_window_width = float(s.tf_window_left), float(s.tf_window_width)
    et = np.asarray(s.env_transform, f32)
    _set(p.env_transform, et.T)
    _set(p.env_inv_transform, np.linalg.inv(et.astype(np.float64)).astype(f32).T)
    p.env_strength = float(s.env_strength)
    p.resolution[0], p.resolution[1] = int(width), int(height)
    return p
