"""Short driver for ncu: renders a few launches of the tracking kernel on the bench workload.
    python tools/profile_trace.py [--tf 0|1] [--w 1920 --h 1080] [--spp 4] [--launches 3] [--scene smoke|cloud]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import volren_b200 as vr  # noqa: E402
from volren_b200 import formats  # noqa: E402
from helpers import default_scene, readme_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tf", type=int, default=1)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--count", type=int, default=0)
ap.add_argument("--kernel", type=int, default=0)
a = ap.parse_args()
A = os.path.join(ROOT, "tests", "golden", "assets")
grid = formats.load_brick(os.path.join(A, "smoke.brick"))
env = formats.load_hdr(os.path.join(A, "table_mountain_2_puresky_1k.hdr"))
lut = formats.lut_for_upload(formats.load_lut_txt(os.path.join(A, "lut.txt")))
ctx = vr.Context(0)
ctx.resize(a.w, a.h)
ctx.grid_upload_brick(grid)
ctx.env_upload(env)
ctx.tf_upload(lut)
p = default_scene(grid, a.w, a.h, bounces=128, use_tf=True) if a.tf else readme_scene(grid, a.w, a.h)
import time
ctx.set_kernel(a.kernel)
if a.count:
    ctx.set_counting(True)
for i in range(a.launches):
    ctx.sync(); t = time.perf_counter()
    ctx.trace(p, 1 + i * a.spp, a.spp)
    ctx.sync(); dt = time.perf_counter() - t
    print(f"launch {i}: {dt*1e3:.2f} ms  {a.w*a.h*a.spp/dt/1e6:.1f} Msamples/s")
if a.count:
    c = ctx.get_counters().as_dict(); n = c["n_samples"]
    print({k: v / n for k, v in c.items()})
ctx.close()
