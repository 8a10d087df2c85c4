#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- the committed recipe that turns the reference's UNMODIFIED GLSL compute shaders
into C++ translation-unit fragments, so that `oracle/_ref/libglsl_ref.so` is "the reference compiled here" for
the shader layer (the GLSL cannot run without a GL stack).

    glsl2cpp.py <reference root> <output dir>

reads   shader/common.glsl, shader/pathtracer_brick.glsl, shader/pathtracer_brick_tf.glsl,
        shader/env_setup.glsl, shader/tonemap.glsl            (never copied into this repository)
writes  <output dir>/{pathtracer_brick,pathtracer_brick_tf,env_setup,tonemap}.inc + SHA256SUMS of the inputs

The output goes to oracle/_ref/gen/ (git-ignored, derived from reference sources). The fragments are included
by glsl_ref.cpp inside one namespace each, after glsl_shim.h (glm of the reference's own submodule + the GL
fixed-function behaviour that no source states).

Every rewrite is purely lexical and listed here; none touches an arithmetic expression:

 R1  `#include "common.glsl"` is spliced in place          (what cppgl::Shader does, cppgl/shader.cpp:57-92)
 R2  `#version ...` and `layout (local_size_x ...) in;` lines are dropped
 R3  `layout (...)` qualifiers and the `uniform` keyword are dropped: uniforms, samplers and images become
     namespace-scope variables that the harness assigns before "dispatching"
 R4  `layout(std430, binding = 4) buffer LUTBuffer { vec4 tf_lut[]; };`  ->  `const vec4* tf_lut;`
 R5  parameter qualifiers: `inout T x` / `out T x` -> `T& x`, `in T x` -> `T x`
 R6  floating literals without suffix get an `f` (a GLSL literal is a 32-bit float, a C++ one a double)
 R7  the entry point's `uint seed = tea(seed * (...` reads the UNIFORM `seed` in GLSL (a name's scope starts
     after its initializer, GLSL 4.50 spec 4.2.2) but the new local in C++: the uniform is renamed `u_seed`
     in its declaration and in that one initializer
 R8  `#define M_PI` / `#define FLT_MAX` get an `#undef` in front (libc macros of the same name)
 R9  `vecN(rng(previous), rng(previous), ...)` -> `vecN{rng(previous), rng(previous), ...}`: GLSL evaluates constructor
     arguments left to right (4.50 spec 5.9 / 6.1: "in order, from left to right"); C++ leaves call arguments
     unsequenced (g++ goes right to left) but orders a braced list. These are the only expressions in the five files
     with two side effects in one argument list.
 R10 the scalar conversion `int(x)` -> `glsl_int(x)` (glsl_shim.h): for a float argument the GPU conversion -- saturating,
     NaN -> 0 -- instead of the C++ cast, which is undefined there (x86 yields INT_MIN). GLSL leaves the case undefined too and
     the reference reaches it: rng() == 0 under a zero majorant makes t = 0 / 0 (common.glsl:434/481), the next lookups run
     on NaN positions and `tf_lut[int(floor(NaN))]` (common.glsl:209) must not index wildly. Same decision as f2i() of
     oracle/vr_oracle.c and as cvt.rzi.s32.f32 on the device; identical results wherever the argument is a finite in-range value.
"""
import hashlib
import os
import re
import sys

FILES = ["pathtracer_brick", "pathtracer_brick_tf", "env_setup", "tonemap"]
TYPES = r"(?:float|int|uint|bool|vec2|vec3|vec4|ivec2|ivec3|ivec4|uvec2|uvec3|uvec4)"

FLOAT_LIT = re.compile(
    r"(?<![\w.])("                       # not glued to an identifier / another number
    r"(?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?"   # 1.  1.5  .5  1.5e3
    r"|\d+[eE][-+]?\d+"                  # 1e-6
    r")(?![\w.])")                       # no suffix yet (f, F, lf) and not part of a longer token


def lexical_rewrites(src: str, entry: bool) -> str:
    out = []
    for line in src.split("\n"):
        code, sep, comment = line.partition("//")
        if re.match(r"\s*#version\b", code):
            continue                                                              # R2
        if re.match(r"\s*layout\s*\(\s*local_size_x", code):
            continue                                                              # R2
        code = re.sub(r"\blayout\s*\([^)]*\)\s*", "", code)                         # R3
        code = re.sub(r"^\s*uniform\s+", "", code)                                # R3
        code = re.sub(r"\b(?:inout|out)\s+(" + TYPES + r")\s+(\w+)", r"\1& \2", code)  # R5
        code = re.sub(r"([(,]\s*)in\s+(" + TYPES + r")\s+(\w+)", r"\1\2 \3", code)     # R5
        if not re.match(r"\s*#", code):
            code = FLOAT_LIT.sub(lambda m: m.group(1) + "f", code)                # R6
        else:
            m = re.match(r"(\s*#define\s+\w+(?:\([^)]*\))?)(.*)", code)
            if m:                                                                 # R6 inside macro bodies
                code = m.group(1) + FLOAT_LIT.sub(lambda k: k.group(1) + "f", m.group(2))
            m = re.match(r"\s*#define\s+(M_PI|FLT_MAX)\b", code)
            if m:
                out.append("#undef " + m.group(1))                               # R8
        code = re.sub(r"\b(vec[234])\((rng\(previous\)(?:,\s*rng\(previous\))+)\)", r"\1{\2}", code)   # R9
        code = re.sub(r"(?<![\w.])int\(", "glsl_int(", code)                        # R10
        if entry:                                                                 # R7
            code = re.sub(r"^int seed;", "int u_seed;", code)
            code = code.replace("uint seed = tea(seed * (", "uint seed = tea(u_seed * (")
        out.append(code + sep + comment)
    text = "\n".join(out)
    text = re.sub(r"buffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", r"const \1* \2;", text)  # R4
    return text


def main(ref_root: str, out_dir: str) -> None:
    shader_dir = os.path.join(ref_root, "shader")
    os.makedirs(out_dir, exist_ok=True)
    sums = []

    def read(name):
        with open(os.path.join(shader_dir, name), "rb") as f:
            raw = f.read()
        sums.append(f"{hashlib.sha256(raw).hexdigest()}  shader/{name}")
        return raw.decode("utf-8")

    common = lexical_rewrites(read("common.glsl"), entry=False)
    for stem in FILES:
        body = lexical_rewrites(read(stem + ".glsl"), entry=True)
        body = re.sub(r'^\s*#include\s+"common\.glsl"\s*$', lambda m: common, body, flags=re.M)   # R1
        with open(os.path.join(out_dir, stem + ".inc"), "w") as f:
            f.write(f"// GENERATED by oracle/glsl_ref/glsl2cpp.py from the reference's shader/{stem}.glsl -- do not commit\n")
            f.write(body)
    with open(os.path.join(out_dir, "SHA256SUMS"), "w") as f:
        f.write("\n".join(sorted(set(sums))) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
