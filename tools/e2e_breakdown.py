"""Host-side time of each public-API call of one e2e bench step (single context, stream synced after every call):
    python tools/e2e_breakdown.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import volren_b200 as vr
from volren_b200 import formats
from helpers import default_scene
A = os.path.join(ROOT, "tests", "golden", "assets")
grid = formats.load_brick(os.path.join(A, "smoke.brick"))
env = formats.load_hdr(os.path.join(A, "table_mountain_2_puresky_1k.hdr"))
lut = formats.lut_for_upload(formats.load_lut_txt(os.path.join(A, "lut.txt")))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
grid.indirection, grid.range, grid.atlas, grid.mips = pin(grid.indirection), pin(grid.range), pin(grid.atlas), [pin(m) for m in grid.mips]
env, lut = pin(env), pin(lut)
W, H, S = 1920, 1080, 16
ctx = vr.Context(0); ctx.resize(W, H)
p = default_scene(grid, W, H, bounces=128, use_tf=True)
host = pin(np.empty((H, W, 4), np.float32))
acc = {}
def T(name, f):
    ctx.sync(); t = time.perf_counter(); f(); ctx.sync(); acc.setdefault(name, []).append((time.perf_counter() - t) * 1e3)
for it in range(12):
    T("grid_upload_brick", lambda: ctx.grid_upload_brick(grid))
    T("env_upload", lambda: ctx.env_upload(env))
    T("tf_upload", lambda: ctx.tf_upload(lut))
    T("trace", lambda: ctx.trace(p, 1 + it * S, S))
    T("download_color", lambda: ctx.lib.vrb_download_color(ctx.handle, host.ctypes.data, 4))
tot = 0
for k, v in acc.items():
    m = float(np.median(v[2:])); tot += m
    print(f"{k:20s} {m:7.3f} ms")
print(f"{'sum':20s} {tot:7.3f} ms  -> {W*H*S/tot/1e6:.2f} Gsamples/s unpipelined")
