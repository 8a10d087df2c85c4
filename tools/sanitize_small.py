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
ctx.grid_build_from_dense(vox, 0.0, 1.0, frame=1)                      # fast path: k_range_xy, k_range_z_count, scan, k_brick_encode_lut, k_range_mips3
vox2 = (rng.random((21, 30, 136)) * 255).astype(np.uint8); vox2[rng.random(vox2.shape) < 0.5] = 0
ctx.grid_build_from_dense(vox2, -3.0, 7.5, frame=4)                    # padding bricks, ragged last chunk, several y-groups
vox3 = (rng.random((9, 11, 13)) * 255).astype(np.uint8)
ctx.grid_build_from_dense(vox3, 0.0, 1.0, frame=5)                     # width not a multiple of 8: warp-per-brick kernels
try:
    import torch
    f = torch.rand((24, 20, 32), device="cuda") * 3 - 1
    torch.cuda.synchronize()
    ctx.grid_build_from_float_device(f.data_ptr(), (32, 20, 24), frame=6)   # DenseGrid(float*) + bricks on the device (float4 kernels)
    f2 = torch.rand((7, 9, 10), device="cuda")
    torch.cuda.synchronize()
    ctx.grid_build_from_float_device(f2.data_ptr(), (10, 9, 7), frame=7)    # scalar kernels
except ImportError:
    pass
n = vr.NanoVDBGridData(g["nvdb_file"], "temperature")
ctx.grid_build_from_nvdb(n, frame=2)
ctx.grid_build_from_values(g["density.padded"], tuple(int(v) for v in g["density.extent"]), frame=3)
ctx.grid_download(frame=2)
ctx.grid_upload_brick(grid)
for kind in (0, 1, 2, 3):
    ctx.set_kernel(kind)
    for tf in (1, 0):
        p = default_scene(grid, W, H, bounces=16, use_tf=True) if tf else readme_scene(grid, W, H, bounces=16)
        ctx.set_option("pass", 2)
        ctx.clear(); ctx.trace(p, 1, 3); ctx.trace(p, 4, 5)                 # several passes per call: both pass lanes, cached mask / order re-use
        img = ctx.download_color()
        assert np.isfinite(img).all()
ctx.set_kernel(0)
ctx.trace_deterministic(readme_scene(grid, W, H))
ctx.tonemap(3.0, 2.2); ctx.download_color_ldr()
ctx.close()
print("sanitize_small ok")
