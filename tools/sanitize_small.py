"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every build path and every tracking kernel once.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import volren_b200 as vr
from volren_b200 import formats
from helpers import default_scene, readme_scene
A = os.path.join(ROOT, "tests", "golden", "assets")
grid = formats.load_brick(os.path.join(A, "smoke.brick"))
env = formats.load_hdr(os.path.join(A, "table_mountain_2_puresky_1k.hdr"))
lut = formats.lut_for_upload(formats.load_lut_txt(os.path.join(A, "lut.txt")))
g = np.load(os.path.join(ROOT, "tests", "golden", "nvdb_golden.npz"))
ctx = vr.Context(0)
W, H = 72, 44
ctx.resize(W, H); ctx.env_upload(env); ctx.tf_upload(lut)
rng = np.random.default_rng(1)
vox = (rng.random((20, 24, 40)) * 255).astype(np.uint8); vox[rng.random(vox.shape) < 0.6] = 0
ctx.grid_build_from_dense(vox, 0.0, 1.0, frame=1)
n = vr.NanoVDBGridData(g["nvdb_file"], "temperature")
ctx.grid_build_from_nvdb(n, frame=2)
ctx.grid_build_from_values(g["density.padded"], tuple(int(v) for v in g["density.extent"]), frame=3)
ctx.grid_download(frame=2)
ctx.grid_upload_brick(grid)
for kind in (0, 1, 2, 3, 4):
    ctx.set_kernel(kind)
    for tf in (1, 0):
        p = default_scene(grid, W, H, bounces=16, use_tf=True) if tf else readme_scene(grid, W, H, bounces=16)
        ctx.clear(); ctx.trace(p, 1, 3); ctx.trace(p, 4, 2)
        img = ctx.download_color()
        assert np.isfinite(img).all()
ctx.set_kernel(0)
ctx.trace_deterministic(readme_scene(grid, W, H))
ctx.tonemap(3.0, 2.2); ctx.download_color_ldr()
ctx.close()
print("sanitize_small ok")
