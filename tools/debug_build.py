"""Diagnostic: GPU brick build of the C3-shaped fBm grid vs the oracle -- which buffer differs, where, how."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import volren_b200 as vr
import workloads as wl
from oracle.binding import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
o = Oracle()
ctx = vr.Context(0)
vox = wl.fbm_cloud(n)
print("vox", vox.shape, vox.dtype, vox.is_contiguous(), hex(vox.data_ptr()), int(vox.max()))
torch.cuda.synchronize()
ctx.grid_build_from_dense_device(vox.data_ptr(), (n, n, n), 0.0, 1.0)
g = ctx.grid_download()
host = o.brick_build(vox.cpu().numpy(), 0.0, 1.0)
for name in ("range", "indirection", "atlas"):
    a, b = getattr(g, name), getattr(host, name)
    print(name, a.shape, b.shape, a.dtype, b.dtype)
    if a.shape != b.shape:
        continue
    d = np.argwhere(a != b)
    print("  mismatches:", len(d))
    for idx in d[:8]:
        t = tuple(idx)
        print("   at", t, "gpu", hex(int(a[t])), "oracle", hex(int(b[t])))
    if len(d):
        print("   bbox min", d.min(0), "max", d.max(0))
for i in range(3):
    print("mip", i, np.array_equal(g.mips[i], host.mips[i]))
print("brick_count", g.brick_count, host.brick_count)
