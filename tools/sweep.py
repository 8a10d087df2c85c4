"""Build several variants of libvrb200 with different -D tunables (here, on CPU) and time them (on the GPU box).
    python tools/sweep.py build  name:"-DVR_K_NEE=8 -DVR_K_FINISH=16" ...
    python tools/sweep.py run [--tf 0|1] [--spp 16]
Variants live in gpurun_out/../build/sweep/*.so (they travel with the snapshot since *.so is not gpurun-ignored).
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
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        cmd = [vb._nvcc()] + vb.NVCC_FLAGS + flags.split() + ["-o", os.path.join(OUT, name + ".so"), os.path.join(vb.CSRC, "vrb200.cu")]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        assert p.wait() == 0, name


def run(argv):
    for so in sorted(glob.glob(os.path.join(OUT, "*.so"))):
        for tf in (1, 0):
            env = dict(os.environ, VRB200_LIB=so)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_trace.py"), "--tf", str(tf), "--spp", "32", "--launches", "3"] + argv,
                                 env=env, capture_output=True, text=True).stdout.strip().splitlines()
            print(f"{os.path.basename(so):40s} tf={tf}  {out[-1] if out else 'FAILED'}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
