"""Build script for the native parts of volren_b200 (in-tree, sm_100a only).

    python -m volren_b200.build [--host] [--force]

libvrb200.so   : CUDA kernels + C ABI (include/vrb200.h)          <- volren_b200/csrc/*.cu
volpy*.so, volren (with --host): C++ host mirroring the reference Renderer/CLI/pybind surface
"""
from __future__ import annotations

import os
import shlex
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libvrb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
# translation units: (source, extra flags). The IEEE cross-check kernels are compiled without FMA contraction.
UNITS = [("vrb200.cu", []), ("vrb200_strict.cu", ["-fmad=false"])]
OBJ_DIR = os.path.join(PKG, "csrc", "_build")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_cuda(force=False, verbose=False, lib=None):
    sources = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [os.path.join(ROOT, "include", "vrb200.h")]
    if not force and lib is None and not _newer(LIB, sources):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = shlex.split(os.environ.get("VRB200_NVCC_FLAGS", ""))          # e.g. -DVR_POOL_SLOTS=96 for sweep builds
    jobs = []
    for src, flags in UNITS:
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        jobs.append((obj, subprocess.Popen(cmd)))
    for obj, proc in jobs:
        if proc.wait() != 0:
            raise subprocess.CalledProcessError(proc.returncode, "nvcc -c " + obj)
    out = lib or LIB
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + [o for o, _ in jobs], check=True)
    return out


def host_targets():
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(PKG, "volpy" + ext), os.path.join(PKG, "volren")


def build_host(force=False):
    if not os.path.isdir(HOST) or not os.path.exists(os.path.join(HOST, "Makefile")):
        return None
    subprocess.run(["make", "-s", "-C", HOST] + (["-B"] if force else []), check=True)
    return host_targets()


def build_all(force=False):
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    force = "--force" in sys.argv
    build_cuda(force, verbose="-v" in sys.argv)
    if "--host" in sys.argv:
        build_host(force)
    print("built", LIB)
