"""Brick-build baseline: the UNMODIFIED reference voldata::BrickGrid(DenseGrid) compiled from its own
sources (oracle/_ref, serial: TBB is absent) against the GPU builder through the C ABI, same voxels, results compared bit
for bit.   python tools/brick_build_baseline.py [edge=384]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import volren_b200 as vr
from oracle.binding import VoldataRef

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
z, y, x = np.meshgrid(*(np.linspace(-1, 1, n, dtype=np.float32),) * 3, indexing="ij")
f = np.clip(np.sin(7 * x) * np.sin(5 * y + 1) * np.sin(9 * z + 2) - 0.2, 0, 1) * np.clip(1.2 - np.sqrt(x * x + y * y + z * z), 0, 1)
vox = (f / f.max() * 255).astype(np.uint8)
del x, y, z, f
ctx = vr.Context(0)
ctx.grid_build_from_dense(vox, 0.0, 1.0)           # warm-up (pool growth, first launches)
t0 = time.perf_counter(); ctx.grid_build_from_dense(vox, 0.0, 1.0); ctx.sync(); gpu_s = time.perf_counter() - t0
got = ctx.grid_download()
out = dict(grid=[n, n, n], bricks_allocated=got.brick_count, gpu_build_from_host_ms=gpu_s * 1e3)
if VoldataRef.available():
    ref = VoldataRef()
    t0 = time.perf_counter(); want = ref.brick_build(vox, 0.0, 1.0); cpu_s = time.perf_counter() - t0
    same = (np.array_equal(got.range, want.range) and np.array_equal(got.indirection, want.indirection) and np.array_equal(got.atlas, want.atlas)
            and all(np.array_equal(a, b) for a, b in zip(got.mips, want.mips)))
    out.update(reference_cpu_build_ms=cpu_s * 1e3, reference_threads=1, speedup=cpu_s / gpu_s, bit_identical=bool(same))
    assert same
else:
    out["reference"] = "oracle/_ref not built"
print(json.dumps(out))
ctx.close()
