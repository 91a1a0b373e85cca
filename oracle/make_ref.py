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
  _ref/marschner.gen.inc       the Marschner LUT shaders: shared/inc_constants.glsl, the two
                               `gaussian` overloads of shared/inc_maths.glsl:270-276, shared/inc_fresnel.glsl,
                               shared/inc_solver.glsl, hair/marschner/inc_marschner_n.glsl and the `main` of
                               cs_marschner_m.glsl / cs_marschner_n.glsl (-> marschner_m_main / marschner_n_main), with
                               the same lexical edits plus: `#include`/include guards, `precision`, `layout(...) in;`
                               and the image declaration removed (the harness provides imageStore), and the one
                               swizzle assignment `roots.xyz = e;` spelled component-wise.
  _ref/tess_skin.gen.inc       the stages either side of the simulation, same lexical edits:
                               shared/inc_constants.glsl; from shared/inc_maths.glsl the functions sample_triangle2
                               (l.107-115), hermite_mix (210-228), maprange (239-241), smoothstep2 (263-266); from
                               shared/inc_skinning.glsl apply_skinning (22-31), get_dual_quaternions_matrices (37-52) and
                               skinning_DQBS (54-82) — `out`/`inout` parameters become references, the subroutine uniform
                               `uSkinning` is bound to skinning_DQBS by the harness, the one swizzle assignment
                               `_weights.xyz *= e;` is spelled component-wise; and the `main` bodies of the four
                               02_tess_stream stages (vs/tcs/tes/gs -> vs_main/tcs_main/tes_main/gs_main, one namespace
                               each) with their interface declarations (layout(...) lines) removed — the harness
                               declares the stage inputs and outputs.
  _ref/hair_init_simulation.gen.inc   body of Hair::init_simulation, src/fx/hair.cc:236-328
                                      (up to, not including, the GL buffer creation).
  _ref/hair_init_mesh.gen.inc         element loop of Hair::init_mesh, src/fx/hair.cc:397-409.
  _ref/raw_recalc_normals.gen.inc     RawMeshData::recalculateNormals, src/utils/raw_mesh_file.cc:11-50 (scalps without normals).
  _ref/generate_skinning_datas.gen.inc  SkeletonController::generate_skinning_datas, src/fx/animation/skeleton_controller.cc:248-265
                                      (skinning matrices -> dual-quaternion palette).
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


FLOAT_LITERAL = r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)(?![\w.])"


def glsl_to_cpp(src):
    """The lexical edits shared by the Marschner sources (see the module docstring)."""
    src = re.sub(r"^#version.*$", "", src, flags=re.M)
    src = re.sub(r'^#include\s+".*$', "", src, flags=re.M)
    src = re.sub(r"^#(ifndef|define|endif)\s*(//\s*)?SHADERS?_\w*GLSL_\s*$", "", src, flags=re.M)
    src = re.sub(r"^precision\s+highp\s+float;\s*$", "", src, flags=re.M)
    src = re.sub(r"^layout\(local_size_x[^)]*\)\s*in;\s*$", "", src, flags=re.M)
    src = re.sub(r"^\s*writeonly\s*(\n\s*)?uniform\s+layout\(rgba16f\)\s+image2D\s+uDstImg;\s*$", "", src, flags=re.M)
    src = re.sub(r"\bin\s+(vec[234]|float|int)\b", r"\1", src)
    src = re.sub(r"(\w+)\.xyz\s*=\s*([^;]*);", r"{ const vec3 t_ = \2; \1.x = t_.x; \1.y = t_.y; \1.z = t_.z; }", src)
    src = re.sub(r"\.xyz\b(?!\()", ".xyz()", src)
    src = re.sub(r"\.xy\b(?!\()", ".xy()", src)
    return re.sub(FLOAT_LITERAL, r"\1f", src)


def gen_marschner():
    sh = os.path.join(REF, "src/shaders")
    parts = []
    parts.append(glsl_to_cpp(open(os.path.join(sh, "shared/inc_constants.glsl")).read()))
    maths = open(os.path.join(sh, "shared/inc_maths.glsl")).read()
    g = re.findall(r"^(?:float|vec3) gaussian\([^)]*\) \{\n.*?\n\}\n", maths, flags=re.M | re.S)
    assert len(g) == 2, f"expected two gaussian overloads in inc_maths.glsl, found {len(g)}"
    parts.append(glsl_to_cpp("".join(g)))
    for f in ("shared/inc_fresnel.glsl", "shared/inc_solver.glsl", "hair/marschner/inc_marschner_n.glsl"):
        parts.append(glsl_to_cpp(open(os.path.join(sh, f)).read()))
    for f, name in (("hair/marschner/cs_marschner_m.glsl", "marschner_m_main"), ("hair/marschner/cs_marschner_n.glsl", "marschner_n_main")):
        src = glsl_to_cpp(open(os.path.join(sh, f)).read())
        src, n = re.subn(r"\bvoid\s+main\s*\(\s*\)", f"void {name}()", src); assert n == 1
        src = re.sub(r"^#ifndef BLOCK_DIM\s*\n\s*#define BLOCK_DIM\s+16\s*\n#endif\s*$", "", src, flags=re.M)
        if name == "marschner_n_main":                     # both programs declare it; one C++ translation unit holds both
            src, n = re.subn(r"^uniform float uInvResolution;\s*$", "", src, flags=re.M); assert n == 1
        parts.append(src)
    out = "\n".join(parts)
    assert "imageStore" in out and "uDstImg;" not in out and "#include" not in out
    open(os.path.join(OUT, "marschner.gen.inc"), "w").write(out)


def extract_function(text, name):
    """The definition of GLSL function `name` (return type line to the closing brace in column 0)."""
    m = re.findall(r"^(?:vec[234]|float|void) " + name + r"\([^)]*\) \{\n.*?\n\}\n", text, flags=re.M | re.S)
    assert len(m) == 1, f"expected one definition of {name}, found {len(m)}"
    return m[0]


def glsl_fn_to_cpp(src):
    """glsl_to_cpp plus what function parameters and the skinning source need."""
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+", r"\1& ", src)
    src = re.sub(r"\bin\s+(uvec4|mat4)\b", r"\1", src)
    src = re.sub(r"(\w+)\.xyz\s*\*=\s*([^;]*);", r"{ const vec3 t_ = \2; \1.x *= t_.x; \1.y *= t_.y; \1.z *= t_.z; }", src)
    src = re.sub(r"\.xyzw\b(?!\()", ".xyzw()", src)
    return glsl_to_cpp(src)


def stage_main(path, name):
    """`void main() {...}` of one tess-stream stage, renamed; everything else in those files is interface declaration."""
    src = open(path).read()
    m = re.findall(r"^void main\(\) \{\n.*?\n\}\n", src, flags=re.M | re.S)
    assert len(m) == 1, f"{path}: expected one main()"
    body = re.sub(r"^#if 1\s*$", "#if 1", m[0], flags=re.M)
    return glsl_fn_to_cpp(body).replace("void main()", f"void {name}()")


def gen_tess_skin():
    sh = os.path.join(REF, "src/shaders")
    maths = open(os.path.join(sh, "shared/inc_maths.glsl")).read()
    skin = open(os.path.join(sh, "shared/inc_skinning.glsl")).read()
    parts = ["namespace shared_inc {"]
    parts.append(glsl_to_cpp(open(os.path.join(sh, "shared/inc_constants.glsl")).read()))
    for fn in ("sample_triangle2", "hermite_mix", "maprange", "smoothstep2"):
        parts.append(glsl_fn_to_cpp(extract_function(maths, fn)))
    parts.append("void skinning_DQBS(uvec4 _indices, vec4 _weights, vec3& v, vec3& n);")
    for fn in ("apply_skinning", "get_dual_quaternions_matrices", "skinning_DQBS"):
        parts.append(glsl_fn_to_cpp(extract_function(skin, fn)))
    parts.append("}  // namespace shared_inc")
    ts = os.path.join(sh, "hair/02_tess_stream")
    for stage in ("vs", "tcs", "tes", "gs"):
        parts.append(f"namespace stage_{stage} {{\nusing namespace shared_inc;\nSTAGE_{stage.upper()}_INTERFACE")
        parts.append(stage_main(os.path.join(ts, f"{stage}_stream_hair.glsl"), f"{stage}_main"))
        parts.append(f"}}  // namespace stage_{stage}")
    out = "\n".join(parts)
    assert "layout(" not in out and "#include" not in out and "subroutine" not in out
    open(os.path.join(OUT, "tess_skin.gen.inc"), "w").write(out)


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
    # RawMeshData::recalculateNormals, src/utils/raw_mesh_file.cc:11-50 (the whole member function definition)
    rn = slice_lines("src/utils/raw_mesh_file.cc", 11, 50)
    assert rn.startswith("void RawMeshData::recalculateNormals() {") and rn.rstrip().endswith("}") and "face.z = static_cast<uint32_t>(normals.size());" in rn, \
        "raw_mesh_file.cc moved; re-pin line numbers"
    open(os.path.join(OUT, "raw_recalc_normals.gen.inc"), "w").write(rn)
    # SkeletonController::generate_skinning_datas, src/fx/animation/skeleton_controller.cc:248-265
    gs = slice_lines("src/fx/animation/skeleton_controller.cc", 248, 265)
    assert gs.startswith("void SkeletonController::generate_skinning_datas(") and gs.rstrip().endswith("}") and \
        "glm::dualquat(skinning_matrices_[i])" in gs, "skeleton_controller.cc moved; re-pin line numbers"
    open(os.path.join(OUT, "generate_skinning_datas.gen.inc"), "w").write(gs)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference tree not found at {REF}")
    os.makedirs(OUT, exist_ok=True)
    gen_shader()
    gen_host()
    gen_marschner()
    gen_tess_skin()
    print("generated into", OUT)
