"""Times successive brick builds of one synthetic grid (device-resident u8 voxels) through the C ABI."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import volren_b200 as vr
from bench_configs import fbm_cloud, timed
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = vr.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
vox = fbm_cloud(n)
torch.cuda.synchronize()
for i in range(5):
    t0 = time.perf_counter()
    ms = timed(lambda: ctx.grid_build_from_dense_device(vox.data_ptr(), (n, n, n), 0.0, 1.0))
    print(f"build {i}: events {ms:.2f} ms  wall {(time.perf_counter()-t0)*1e3:.2f} ms  info {ctx.grid_info()}", flush=True)
ctx.close()
