#!/usr/bin/env python3
"""Generates tests/golden/glsl_ref_golden.npz: images rendered by the reference's OWN GLSL compiled as C++
(oracle/_ref/libglsl_ref.so, built by oracle/Makefile from /root/reference/shader/*.glsl through
oracle/glsl_ref/glsl2cpp.py). Run in a container that has /root/reference:

    make -C oracle all && python tests/golden/make_glsl_golden.py

The vectors travel with the repository (the .so and the reference tree may not), so that `vr_oracle.c == the reference's
shader text` is checked wherever the tests run. tests/test_glsl_ref.py::CASES defines the scenes; the same function
builds them for the generator and for the test.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_glsl_ref as T
    from oracle.binding import GlslRef, Oracle
    o, g = Oracle(), GlslRef()
    out = {}
    assets = T.load_assets(o)
    for name in T.CASES:
        sc, p, first, n = T.build_case(name, o, assets, golden_size=True)
        out[name] = g.trace(sc, p, first, n)
    out["env_impmap_level0"] = g.env_setup(assets["env"])
    out["tonemap_in"] = T.tonemap_input()
    out["tonemap_out"] = g.tonemap(out["tonemap_in"], 3.0, 2.0)
    with open(os.path.join(ROOT, "oracle", "_ref", "gen", "SHA256SUMS")) as f:
        out["sha256sums"] = np.frombuffer(f.read().encode(), np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "glsl_ref_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
