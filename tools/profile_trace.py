"""Timing / ncu driver for the tracking kernel on the BASELINE.json configs (tools/workloads.py).

    python tools/profile_trace.py --scene c1|c2|c3|c4 [--scale 1.0] [--spp 32] [--launches 4] [--kernel 0|1|2|3] [--count 1] [--w W --h H]

Prints per-launch wall time (host timer around vrb_trace + sync) and, with --count 1, the event counters per sample.
`--lib path.so` (or VRB200_LIB) selects a sweep build of the library (tools/sweep.py).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="c2")
ap.add_argument("--tf", type=int, default=None, help="legacy: 1 = c2, 0 = c1 at 1920x1080")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--w", type=int, default=0)
ap.add_argument("--h", type=int, default=0)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--count", type=int, default=0)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--lib", default=None)
ap.add_argument("--json", type=int, default=0)
ap.add_argument("--l2-persist", type=int, default=0, help="MiB of persisting L2 with a window over records + majorant tables (vrb_set_option l2_persist)")
a = ap.parse_args()
if a.lib:
    os.environ["VRB200_LIB"] = a.lib

import volren_b200 as vr  # noqa: E402
import workloads as wl  # noqa: E402

if a.tf is not None:
    a.scene = "c2" if a.tf else "c1"
    a.w, a.h = a.w or 1920, a.h or 1080
ctx = vr.Context(0)
grid, env, lut = wl.load_assets()
ctx.env_upload(env)
name = a.scene.lower()
if name in ("c1", "c2"):
    W, H = (a.w or 1024, a.h or 1024) if name == "c1" else (a.w or 1920, a.h or 1080)
    ctx.grid_upload_brick(grid)
    if name == "c2":
        ctx.tf_upload(lut)
        p = wl.default_params(grid, W, H, use_tf=True)
    else:
        p = wl.readme_params(grid, W, H)
    tf = name == "c2"
else:
    import torch
    W, H = a.w or 1920, a.h or 1080
    if name == "c3":
        n = int(1024 * a.scale) // 8 * 8
        vox, dims, tf = wl.fbm_cloud(n), (n, n, n), False
    else:
        w, d = int(512 * a.scale) // 8 * 8, int(1800 * a.scale) // 8 * 8
        vox, dims, tf = wl.ct_phantom(w, w, d), (w, w, d), True
        ctx.tf_upload(wl.turbo_lut())
    torch.cuda.synchronize()
    ctx.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0)
    ctx.sync()
    del vox
    torch.cuda.empty_cache()
    p = wl.synthetic_params(dims, W, H, tf)
ctx.resize(W, H)
ctx.set_kernel(a.kernel)
if a.l2_persist:
    ctx.set_option("l2_persist", a.l2_persist)
if a.count:
    ctx.set_counting(True)
ms = []
for i in range(a.launches):
    ctx.sync(); t = time.perf_counter()
    ctx.trace(p, 1 + i * a.spp, a.spp)
    ctx.sync(); dt = time.perf_counter() - t
    ms.append(dt * 1e3)
    print(f"launch {i}: {dt*1e3:.2f} ms  {W*H*a.spp/dt/1e6:.1f} Msamples/s", flush=True)
out = {"scene": name, "kernel": a.kernel, "l2_persist_MiB": a.l2_persist, "lib": os.path.basename(a.lib) if a.lib else "default", "res": [W, H], "spp": a.spp,
       "best_ms": min(ms[1:]) if len(ms) > 1 else ms[0]}
out["gsamples_per_s"] = W * H * a.spp / (out["best_ms"] * 1e-3) / 1e9
if a.count:
    c = ctx.get_counters().as_dict(); n = c["n_samples"]
    out["counters_per_sample"] = {k: v / n for k, v in c.items()}
    out["alg_bytes_per_sample"] = wl.algorithmic_bytes(c, tf) / n
    print(out["counters_per_sample"])
if a.json:
    print("JSON " + json.dumps(out), flush=True)
ctx.close()
