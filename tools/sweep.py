"""Build several variants of libvrb200 with different -D tunables (here, on CPU) and time them (on the GPU box).
    python tools/sweep.py build  name:"-DVR_POOL_SLOTS=96 -DVR_POOL_MIN_BLOCKS=4" ...
    python tools/sweep.py run [--scenes c1,c2,c3] [--spp 32] [--scale 0.5]
Variants live in tools/_sweep/*.so (they travel with the snapshot since *.so is not gpurun-ignored).
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "_sweep")
sys.path.insert(0, ROOT)


def build(specs):
    from volren_b200 import build as vb
    os.makedirs(OUT, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, "*.so")):
        os.remove(f)
    for spec in specs:          # sequential: each build already runs its translation units in parallel
        name, _, flags = spec.partition(":")
        os.environ["VRB200_NVCC_FLAGS"] = flags
        vb.OBJ_DIR = os.path.join(OUT, "obj_" + name)
        vb.build_cuda(force=True, lib=os.path.join(OUT, name + ".so"))
        print("built", name, flags, flush=True)


def run(argv):
    scenes, rest = "c1,c2", []
    it = iter(argv)
    for x in it:
        if x == "--scenes":
            scenes = next(it)
        else:
            rest.append(x)
    for so in sorted(glob.glob(os.path.join(OUT, "*.so"))):
        for sc in scenes.split(","):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_trace.py"), "--scene", sc, "--lib", so, "--json", "1"] + rest,
                                 capture_output=True, text=True)
            line = [l for l in out.stdout.splitlines() if l.startswith("JSON ")]
            print(f"{os.path.basename(so):40s} {sc}  {line[-1][5:] if line else 'FAILED ' + out.stderr[-300:]}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
