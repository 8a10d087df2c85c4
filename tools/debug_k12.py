import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import volren_b200 as vr
import workloads as wl
from helpers import default_scene
grid, env, lut = wl.load_assets()
ctx = vr.Context(0)
ctx.grid_upload_brick(grid); ctx.env_upload(env); ctx.tf_upload(lut)
W, H = 150, 90
for bounces in (1, 2, 128):
    for cull in (1, 0):
        p = default_scene(grid, W, H, bounces=bounces, use_tf=True)
        ctx.resize(W, H)
        ctx.set_option("cull", cull)
        out = []
        for kind in (1, 2):
            ctx.set_kernel(kind); ctx.clear(); ctx.trace(p, 3, 6); out.append(ctx.download_color())
        d = np.abs(out[0] - out[1]).max(axis=-1)
        bad = np.argwhere(d > 0)
        print(f"bounces {bounces} cull {cull}: differing pixels {len(bad)} max abs {d.max():.3e}")
        for y, x in bad[:6]:
            print("   ", y, x, out[0][y, x], out[1][y, x])
ctx.close()
