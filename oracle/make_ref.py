#!/usr/bin/env python3
"""Generate the translation units of oracle/_ref from the reference sources WHERE THEY LIE.

TEST INFRASTRUCTURE ONLY (see oracle/barbu_hair_oracle.h).  Nothing from /root/reference is copied
into the repository: this script reads the reference at build time and writes generated files
into oracle/_ref/ (git-ignored), which oracle/Makefile then compiles with g++.

What it produces
  _ref/cs_simulation.gen.inc   src/shaders/hair/01_simulation/cs_simulation.glsl with the purely
                               lexical changes needed to parse GLSL 4.30 as C++17 on top of GLM:
      * `#version`, `#include "hair/interop.h"` and the `layout(local_size_x=..) in;` line removed
        (interop.h is included by the shim from its reference path instead);
      * the four `layout(std430, binding=..) buffer X { vec4 name[]; };` blocks become
        `static thread_local vec4* name;` pointers the harness binds (== glBindBufferRange);
      * `inout T x` -> `T& x`; the `in` qualifier is dropped (by-value); `uniform`/`shared`
        become `static thread_local` (via macros in the shim);
      * unsuffixed floating literals get an `f` suffix (GLSL literals are 32-bit floats);
      * the `.xyz` swizzle becomes GLM's function swizzle `.xyz()` (GLM only offers operator
        swizzles together with its SIMD code paths, which would change the arithmetic);
      * `main` -> `shader_main`.
    The two documented semantic patches of SURVEY.md §8c are applied by the SHIM, not the text:
      (1) HAIR_MAX_PARTICLE_PER_STRAND is re-#defined to REF_N after interop.h;
      (2) `memoryBarrierShared()` is a real workgroup barrier (fiber yield) — the evident intent,
          and what a <=32-wide workgroup computes on lock-step hardware.
  _ref/hair_init_simulation.gen.inc   body of Hair::init_simulation, src/fx/hair.cc:236-328
                                      (up to, not including, the GL buffer creation).
  _ref/hair_init_mesh.gen.inc         element loop of Hair::init_mesh, src/fx/hair.cc:397-409.
"""
import os
import re
import sys

REF = os.environ.get("BARBU_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def gen_shader():
    src = open(os.path.join(REF, "src/shaders/hair/01_simulation/cs_simulation.glsl")).read()
    # buffers: layout(std430, binding = X) \n qualifier buffer Name {\n  vec4 name[];\n};
    src, n = re.subn(
        r"layout\(std430,\s*binding\s*=\s*\w+\)\s*\n\s*(?:readonly|writeonly)\s+buffer\s+\w+\s*\{\s*vec4\s+(\w+)\[\];\s*\};",
        r"static thread_local vec4* \1;", src)
    assert n == 4, f"expected 4 SSBO blocks, found {n}"
    src, n = re.subn(r"^#version.*$", "", src, flags=re.M); assert n == 1
    src, n = re.subn(r'^#include\s+"hair/interop.h".*$', "", src, flags=re.M); assert n == 1
    src, n = re.subn(r"^layout\(local_size_x\s*=\s*\w+\)\s*in;\s*$", "", src, flags=re.M); assert n == 1
    src, n = re.subn(r"\binout\s+(\w+)\s+", r"\1& ", src); assert n >= 4
    src = re.sub(r"\bin\s+(vec3|float|Particle_t)\b", r"\1", src)
    src, n = re.subn(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src); assert n == 1
    src, n = re.subn(r"\.xyz\b(?!\()", ".xyz()", src); assert n >= 3
    # float literals: digits '.' digits [exponent], not already suffixed, not part of an identifier
    src = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)(?![\w.])", r"\1f", src)
    open(os.path.join(OUT, "cs_simulation.gen.inc"), "w").write(src)


def slice_lines(path, first, last):
    lines = open(os.path.join(REF, path)).read().split("\n")
    return "\n".join(lines[first - 1:last]) + "\n"


def gen_host():
    body = slice_lines("src/fx/hair.cc", 237, 328)
    assert "using AttribBuffer_t" in body.split("\n")[1] + body.split("\n")[2], "hair.cc moved; re-pin line numbers"
    assert body.rstrip().endswith("}"), "init_simulation tangent block should end at line 328"
    open(os.path.join(OUT, "hair_init_simulation.gen.inc"), "w").write(body)
    mesh = slice_lines("src/fx/hair.cc", 397, 409)
    assert "mesh_.patchsize = 6;" in mesh and "elements[idx++] = e + 1;" in mesh, "hair.cc moved; re-pin line numbers"
    open(os.path.join(OUT, "hair_init_mesh.gen.inc"), "w").write(mesh)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference tree not found at {REF}")
    os.makedirs(OUT, exist_ok=True)
    gen_shader()
    gen_host()
    print("generated into", OUT)
