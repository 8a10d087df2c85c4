"""Diagnostic for test_scheduling_options_do_not_change_the_image / test_cached_tile_order: which option sets differ, where."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import volren_b200 as vr
from volren_b200 import formats
from helpers import default_scene, readme_scene
from oracle.binding import Oracle
A = os.path.join(ROOT, "tests", "golden", "assets")
grid = formats.load_brick(os.path.join(A, "smoke.brick"))
env = formats.load_hdr(os.path.join(A, "table_mountain_2_puresky_1k.hdr"))
o = Oracle()
lut, _ = o.lut_upload(formats.load_lut_txt(os.path.join(A, "lut.txt")))
ctx = vr.Context(0)
ctx.grid_upload_brick(grid); ctx.env_upload(env); ctx.tf_upload(lut)
W, H = 200, 120
ctx.resize(W, H)
for kern in (0, 3):
    ctx.set_kernel(kern)
    for name, p in (("tf", default_scene(grid, W, H, bounces=8, use_tf=True)), ("notf", readme_scene(grid, W, H, bounces=8))):
        images = []
        for lpt, cull, npass in ((1, 1, 16), (0, 0, 16), (1, 1, 3), (0, 1, 1)):
            ctx.set_option("lpt", lpt); ctx.set_option("cull", cull); ctx.set_option("pass", npass)
            ctx.clear()
            ctx.trace(p, 1, 5)
            ctx.trace(p, 6, 5)
            images.append(ctx.download_color())
        for i, img in enumerate(images[1:], 1):
            d = np.argwhere(np.any(images[0] != img, axis=-1))
            print(f"kernel {kern} {name} option set {i}: {len(d)} pixels differ", (d.min(0), d.max(0)) if len(d) else "", flush=True)
            if len(d):
                y, x = d[0]
                print("   first", (y, x), images[0][y, x], img[y, x])
ctx.set_kernel(0)
