"""CPU: pins the shader layer of the oracle (oracle/vr_oracle.c: tea ... trace_path, env_setup, tonemap) against the
reference's OWN shader text.

oracle/_ref/libglsl_ref.so is the UNMODIFIED /root/reference/shader/{common,pathtracer_brick,pathtracer_brick_tf,env_setup,
tonemap}.glsl compiled as C++ (oracle/glsl_ref/glsl2cpp.py: ten lexical rewrites, listed there; glsl_shim.h: the reference's
own glm + the GL fixed-function decisions). The restatement must equal it BIT FOR BIT -- every pixel, every case:
TF and non-TF programs, bounces 1 / 3 / 128, with and without an emission grid, clipped + rotated volumes, hidden
environment, several dispatches into the running mean. The same images are committed as golden vectors
(tests/golden/glsl_ref_golden.npz, make_glsl_golden.py) so the check also runs where the reference tree is absent.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import ASSETS, GOLDEN
from helpers import blob_volume, default_scene, readme_scene

# sha256 of the shader files the library was generated from (nihofm/volren @ e8aea40)
SHADER_SHA256 = {
    "shader/common.glsl": "7748f0df0b3776b2ab7a7f58d9271b53f70a107d96237dba809e7c47bb1abef8",
    "shader/pathtracer_brick.glsl": "32135e5f166de724c4cf37399f13b03527994ff90cead579000bfee31d949231",
    "shader/pathtracer_brick_tf.glsl": "3f21fff3c4f0b8059f18b81ce8163776a44434874a219f317974bcb74e849c80",
    "shader/env_setup.glsl": "a5ba43187af290d4a619695fe849b3dce6a1fefdb2b5672f7a81e3461a0acd8f",
    "shader/tonemap.glsl": "dd89accb59d477aadfd874f99bddd3061da16d8e5be769719f0eccdb36665021",
}

# name -> (program, bounces, full-size frame); the golden vectors use a 48x36 frame of the same scene
CASES = {
    "notf_b1": ("notf", 1), "notf_b3": ("notf", 3), "notf_b128": ("notf", 128),
    "tf_b1": ("tf", 1), "tf_b3": ("tf", 3), "tf_b128": ("tf", 128),
    "notf_hidden_env": ("notf_hidden", 16), "tf_shown_env": ("tf_shown", 16),
    "emission": ("emission", 8), "emission_tf": ("emission_tf", 8),
    "crop_rot": ("crop_rot", 32), "running_mean": ("running_mean", 8),
}


def load_assets(oracle):
    from volren_b200 import formats
    env = formats.load_hdr(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"))
    lut, _ = oracle.lut_upload(formats.load_lut_txt(os.path.join(ASSETS, "lut.txt")))
    return dict(grid=formats.load_brick(os.path.join(ASSETS, "smoke.brick")), env=env, pyr=oracle.env_build(env), lut=lut)


def build_case(name, oracle, a, golden_size=False):
    """-> (scene, params, first_sample, n_samples)"""
    from volren_b200 import scene
    kind, bounces = CASES[name]
    W, H = (48, 36) if golden_size else (128, 96)
    grid, env, pyr, lut = a["grid"], a["env"], a["pyr"], a["lut"]
    if kind == "notf":
        return oracle.make_scene(grid, env, pyr), readme_scene(grid, W, H, bounces=bounces), 1, 1
    if kind == "tf":
        return oracle.make_scene(grid, env, pyr, lut), default_scene(grid, W, H, bounces=bounces, use_tf=True), 1, 1
    if kind == "notf_hidden":
        return oracle.make_scene(grid, env, pyr), readme_scene(grid, W, H, bounces=bounces, show_environment=False, seed=7), 3, 1
    if kind == "tf_shown":
        return (oracle.make_scene(grid, env, pyr, lut),
                default_scene(grid, W, H, bounces=bounces, use_tf=True, show_environment=True, phase=-0.4, seed=-5), 2, 1)
    if kind in ("emission", "emission_tf"):
        # the scene of test_emission_grid_end_to_end: a 48^3 density blob + a 24^3 `temperature` grid with its own transform
        vox, lo, hi = blob_volume(48)
        zz, yy, xx = np.mgrid[0:24, 0:24, 0:24].astype(np.float32) / 24
        temp = np.exp(-(((xx - .45) / .2) ** 2 + ((yy - .5) / .2) ** 2 + ((zz - .5) / .25) ** 2))
        tvox = (np.clip(temp - 0.1, 0, 1) / 0.9 * 255).astype(np.uint8)
        g, ge = oracle.brick_build(vox, lo, hi), oracle.brick_build(tvox, 0.0, 2.0)
        emat = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32) @ np.asarray(g.matrix(), np.float32)
        tf = kind == "emission_tf"
        s = scene.RenderSettings(bounces=bounces, seed=42, density_scale=8.0, emission_scale=40.0, albedo=(.7, .6, .5),
                                 phase=0.2, use_transferfunc=tf, show_environment=True)
        scene.scale_and_move_to_unit_cube(g.matrix(), (48, 48, 48), s)
        s.density_scale = 8.0
        p = scene.make_params(W, H, scene.Camera(), s, g.matrix(), (48, 48, 48), g.min_maj,
                              emission_matrix=emat, majorant_emission=ge.min_maj[1])
        return oracle.make_scene(g, env, pyr, lut if tf else None, emission=ge), p, 1, 2
    if kind == "crop_rot":
        # --vol_crop_min/max + --vol_rot_y (main.cpp:417-429): clip planes and a rotated (translation-free) volume transform
        s = scene.RenderSettings(bounces=bounces, seed=11, albedo=(.8, .8, .8), phase=.3, env_strength=2.0,
                                 env_transform=scene.rotate_y(90), vol_clip_min=(.1, .2, 0.), vol_clip_max=(.9, .7, .8))
        scene.scale_and_move_to_unit_cube(grid.matrix(), grid.index_extent(), s)
        rot = np.eye(4, dtype=np.float32)
        rot[:3, :3] = scene.rotate_y(30) @ np.asarray(s.volume_transform, np.float32)[:3, :3]   # mat3-truncated (Q11)
        s.volume_transform = rot
        s.density_scale = 60.0
        cam = scene.Camera(pos=np.array([.3, .2, 1.2], np.float32), dir=scene.normalize([-.3, -.2, -1.2]), fov_degree=55.0)
        return oracle.make_scene(grid, env, pyr), scene.make_params(W, H, cam, s, grid.matrix(), grid.index_extent(), grid.min_maj), 1, 1
    if kind == "running_mean":
        return oracle.make_scene(grid, env, pyr), readme_scene(grid, W, H, bounces=bounces, seed=1234567), 1, 5
    raise KeyError(name)


def tonemap_input():
    rng = np.random.default_rng(17)
    img = (rng.random((24, 32, 4), np.float32) * np.float32(8.0)).astype(np.float32)
    img[0, 0] = (np.nan, 1.0, np.inf, 0.5)
    img[0, 1] = (-np.inf, 0.0, -1.0, np.nan)
    img[0, 2] = (0.0, 1e-30, 1e30, 1.0)
    return img


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def assets(oracle):
    return load_assets(oracle)


@pytest.fixture(scope="module")
def glsl():
    from oracle.binding import GlslRef
    if not GlslRef.available():
        pytest.skip("oracle/_ref/libglsl_ref.so not built (no /root/reference here); the golden vectors still pin the oracle")
    return GlslRef()


@pytest.fixture(scope="module")
def glsl_golden():
    return np.load(os.path.join(GOLDEN, "glsl_ref_golden.npz"))


def test_library_was_generated_from_the_unmodified_shaders(glsl, glsl_golden):
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "gen", "SHA256SUMS")
    if os.path.exists(here):
        got = dict(reversed(line.split("  ")) for line in open(here).read().strip().split("\n"))
        assert got == SHADER_SHA256
    ref = "/root/reference"
    if os.path.isdir(ref):
        for rel, want in SHADER_SHA256.items():
            assert hashlib.sha256(open(os.path.join(ref, rel), "rb").read()).hexdigest() == want
    gold = dict(reversed(line.split("  ")) for line in bytes(glsl_golden["sha256sums"]).decode().strip().split("\n"))
    assert gold == SHADER_SHA256


def test_tea_and_lcg_equal_the_shader_functions(glsl, oracle):
    rng = np.random.default_rng(3)
    for a, b in [(0, 1), (0, 0), (0xFFFFFFFF, 0xFFFFFFFF)] + [tuple(int(x) for x in rng.integers(0, 2**32, 2)) for _ in range(200)]:
        assert glsl.tea(a, b) == oracle.tea(a, b)
    vals, states = oracle.rng_stream(987654321, 500)
    s = 987654321
    for i in range(500):
        v, s = glsl.rng(s)
        assert s == int(states[i]) and np.float32(v) == vals[i]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_equals_compiled_glsl_bit_for_bit(glsl, oracle, assets, name):
    """128x96, every pixel, bitwise (NaN patterns included): vr_oracle.c == the reference's shader text."""
    sc, p, first, n = build_case(name, oracle, assets)
    want = glsl.trace(sc, p, first, n)
    got, _ = oracle.trace(sc, p, first, n)
    assert np.isfinite(want).all()
    assert want[..., :3].max() > 0 and (want[..., 3] > 0).mean() > 0.02        # the case really hits the volume
    neq = (bits(got) != bits(want)).any(axis=2)
    assert not neq.any(), (name, int(neq.sum()), np.argwhere(neq)[:4])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_equals_committed_glsl_golden(oracle, assets, glsl_golden, name):
    """The same comparison against vectors generated by make_glsl_golden.py (runs without the reference tree)."""
    sc, p, first, n = build_case(name, oracle, assets, golden_size=True)
    got, _ = oracle.trace(sc, p, first, n)
    assert np.array_equal(bits(got), bits(glsl_golden[name]))


def test_env_setup_equals_compiled_glsl(glsl, oracle, assets, glsl_golden):
    lvl0 = oracle.pyramid_level(assets["pyr"], 0)
    assert np.array_equal(bits(glsl.env_setup(assets["env"])), bits(lvl0))
    assert np.array_equal(bits(glsl_golden["env_impmap_level0"]), bits(lvl0))


def test_env_setup_on_a_small_odd_sized_map(glsl, oracle):
    rng = np.random.default_rng(1)
    env = (rng.random((7, 13, 3), np.float32) * np.float32(20.0)).astype(np.float32)
    assert np.array_equal(bits(glsl.env_setup(env)), bits(oracle.pyramid_level(oracle.env_build(env), 0)))


def test_tonemap_equals_compiled_glsl(glsl, oracle, glsl_golden):
    img = tonemap_input()
    for exposure, gamma in [(3.0, 2.0), (10.0, 2.2), (0.5, 1.0)]:
        assert np.array_equal(bits(glsl.tonemap(img, exposure, gamma)), bits(oracle.tonemap_inplace(img, exposure, gamma)))
    assert np.array_equal(bits(glsl_golden["tonemap_out"]), bits(oracle.tonemap_inplace(glsl_golden["tonemap_in"], 3.0, 2.0)))


def test_nan_path_rng_zero_under_a_zero_majorant(glsl, oracle, assets):
    """rng() == 0 makes the free-flight draw tau = -log(1 - 0) = 0; under a zero majorant the shader then computes t = 0 / 0
    (common.glsl:434/481) and goes on to `tf_lut[int(floor(NaN))]` (common.glsl:209). Sample 1436 of the headline frame at
    480x270 (bench.py's CPU leg) contains such a path: a plain C++ cast there indexed the LUT at INT_MIN and crashed the
    compiled shaders (found by the bench run on the GPU box); with the GPU conversion pinned for `int(x)` (R10 of glsl2cpp.py,
    the same decision as f2i() in vr_oracle.c and cvt.rzi on the device) both sides agree bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import workloads as wl
    grid = assets["grid"]
    p = wl.default_params(grid, 480, 270, bounces=128, use_tf=True)
    sc = oracle.make_scene(grid, assets["env"], assets["pyr"], assets["lut"])
    want = glsl.trace(sc, p, 1436, 1)
    got, _ = oracle.trace(sc, p, 1436, 1)
    assert np.isfinite(want).all() and want[..., 3].max() > 0
    assert np.array_equal(bits(got), bits(want))
