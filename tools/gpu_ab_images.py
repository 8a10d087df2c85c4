"""A/B of two builds of libvrb200 (tools/_sweep/*.so): images must be bit-identical, timings printed.
    python tools/gpu_ab_images.py"""
import glob, os, subprocess, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, hashlib, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, volren_b200 as vr
from volren_b200 import formats
from helpers import default_scene, readme_scene
A = os.path.join(%r, "tests", "golden", "assets")
grid = formats.load_brick(os.path.join(A, "smoke.brick")); env = formats.load_hdr(os.path.join(A, "table_mountain_2_puresky_1k.hdr"))
lut = formats.lut_for_upload(formats.load_lut_txt(os.path.join(A, "lut.txt")))
ctx = vr.Context(0); W, H = 640, 360; ctx.resize(W, H); ctx.grid_upload_brick(grid); ctx.env_upload(env); ctx.tf_upload(lut)
for tf in (1, 0):
    p = default_scene(grid, W, H, bounces=128, use_tf=True) if tf else readme_scene(grid, W, H)
    ctx.clear(); ctx.trace(p, 1, 16); img = ctx.download_color()
    print("tf", tf, hashlib.sha1(img.tobytes()).hexdigest(), float(img[..., :3].mean()))
''' % (ROOT, ROOT, ROOT)
for so in sorted(glob.glob(os.path.join(ROOT, "tools", "_sweep", "*.so"))):
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, VRB200_LIB=so), capture_output=True, text=True)
    print(os.path.basename(so)); print(out.stdout.strip() or out.stderr[-500:])
