"""ctypes access to the CPU oracle (oracle/libbarbu_hair_oracle.so) and, when built, oracle/_ref.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under barbu_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libbarbu_hair_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
BHO_MAX_COLLIDERS = 8


class BhoCapsule(C.Structure):
    _fields_ = [("a", C.c_float * 3), ("b", C.c_float * 3), ("radius", C.c_float)]


class BhoParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("scale", C.c_float), ("sphere", C.c_float * 4), ("iterations", C.c_int),
                ("gravity", C.c_float * 3), ("force_coeff", C.c_float), ("damp", C.c_float),
                ("wind", C.c_float * 3), ("drag", C.c_float), ("ncapsules", C.c_int),
                ("capsules", BhoCapsule * BHO_MAX_COLLIDERS)]


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("barbu_hair_oracle.c", "barbu_marschner_oracle.c", "barbu_hair_oracle.h")]
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    return ORACLE_SO


def build_ref() -> bool:
    """Build oracle/_ref from /root/reference when that tree exists (this container only)."""
    if not os.path.isdir(os.environ.get("BARBU_REFERENCE", "/root/reference")):
        return False
    subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
    return True


_oracle = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        _oracle = C.CDLL(build_oracle())
        _oracle.bho_simplex2.restype = C.c_float
        _oracle.bho_simplex2.argtypes = [C.c_float, C.c_float]
        _oracle.bho_fnv1a64.restype = C.c_uint64
        _oracle.bho_patch_indices.restype = C.c_int
    return _oracle


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**kw) -> BhoParams:
    p = BhoParams()
    oracle().bho_default_params(C.byref(p))
    for k, v in kw.items():
        cur = getattr(p, k)
        if isinstance(cur, C.Array):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)
    return p


def step(pos4: np.ndarray, vel4: np.ndarray, nstrands: int, nverts: int, params: BhoParams, nthreads: int = 1):
    """In-place reference step on (V,4) float32 arrays."""
    assert pos4.dtype == np.float32 and vel4.dtype == np.float32 and pos4.flags.c_contiguous and vel4.flags.c_contiguous
    oracle().bho_step_mt(_p(pos4), _p(vel4), C.c_int64(nstrands), C.c_int(nverts), C.byref(params), C.c_int(nthreads))


def random_values(seed: int, nstrands: int) -> np.ndarray:
    out = np.empty(nstrands, np.float32)
    oracle().bho_random_values(C.c_uint(seed), C.c_int64(nstrands), _p(out))
    return out


def sphere_scalp(rows: int, cols: int, column_major: bool = False):
    """Synthetic sphere scalp. column_major: the same mesh with its vertices numbered meridian by meridian
    (vertex c * rows + r instead of r * cols + c); faces keep their order."""
    S = rows * cols
    pos, nrm = np.empty((S, 3), np.float32), np.empty((S, 3), np.float32)
    tri = np.empty((2 * (rows - 1) * cols, 3), np.int32)
    oracle().bho_sphere_scalp(C.c_int(rows), C.c_int(cols), _p(pos), _p(nrm), _p(tri))
    if column_major:
        old_of_new = np.arange(S).reshape(rows, cols).T.reshape(-1)          # new vertex c * rows + r  <-  old r * cols + c
        new_of_old = np.empty(S, np.int64); new_of_old[old_of_new] = np.arange(S)
        pos, nrm = np.ascontiguousarray(pos[old_of_new]), np.ascontiguousarray(nrm[old_of_new])
        tri = new_of_old[tri].astype(np.int32)
    return pos, nrm, tri


def init_strands(root_pos3, root_nrm3, random_value, nverts: int, maxlength: float = 0.5):
    S = root_pos3.shape[0]
    pos, vel = np.empty((S * nverts, 4), np.float32), np.empty((S * nverts, 4), np.float32)
    oracle().bho_init_strands(_p(np.ascontiguousarray(root_pos3, np.float32)), _p(np.ascontiguousarray(root_nrm3, np.float32)),
                              _p(np.ascontiguousarray(random_value, np.float32)), C.c_int64(S), C.c_int(nverts),
                              C.c_float(maxlength), _p(pos), _p(vel))
    return pos, vel


def init_tangents(root_nrm3, nverts: int, maxlength: float = 0.5) -> np.ndarray:
    S = root_nrm3.shape[0]
    tan = np.empty((S * nverts, 4), np.float32)
    oracle().bho_init_tangents(_p(np.ascontiguousarray(root_nrm3, np.float32)), C.c_int64(S), C.c_int(nverts),
                               C.c_float(maxlength), _p(tan))
    return tan


def patch_indices(tri: np.ndarray, nverts: int) -> np.ndarray:
    tri = np.ascontiguousarray(tri, np.int32).reshape(-1, 3)
    out = np.empty(6 * tri.shape[0] * max(nverts - 1, 0), np.int32)
    rc = oracle().bho_patch_indices(_p(tri), C.c_int64(tri.shape[0]), C.c_int(nverts), _p(out))
    if rc != 0:
        raise OverflowError("patch indices exceed int32")
    return out


def skin_roots_dq(rest_pos3, rest_nrm3, joints4, weights3, dq):
    S = rest_pos3.shape[0]
    op, on = np.empty((S, 3), np.float32), np.empty((S, 3), np.float32)
    oracle().bho_skin_roots_dq(_p(np.ascontiguousarray(rest_pos3, np.float32)), _p(np.ascontiguousarray(rest_nrm3, np.float32)),
                               _p(np.ascontiguousarray(joints4, np.int32)), _p(np.ascontiguousarray(weights3, np.float32)),
                               _p(np.ascontiguousarray(dq, np.float32)), C.c_int64(S), _p(op), _p(on))
    return op, on


def fnv1a64(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a)
    return int(oracle().bho_fnv1a64(_p(a), C.c_uint64(a.nbytes)))


# ---- oracle/_ref: the reference sources compiled here --------------------------------------------

def ref_available(nverts: int) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libbarbu_ref_glsl_N{nverts}.so"))


_ref_cache = {}


def _ref(kind: str, nverts: int) -> C.CDLL:
    key = (kind, nverts)
    if key not in _ref_cache:
        lib = C.CDLL(os.path.join(REF_DIR, f"libbarbu_ref_{kind}_N{nverts}.so"))
        if kind == "host":
            lib.ref_host_patch_indices.restype = C.c_int64
            lib.ref_host_simplex2.restype = C.c_float
            lib.ref_host_simplex2.argtypes = [C.c_float, C.c_float]
        _ref_cache[key] = lib
    return _ref_cache[key]


def ref_update(pos4, vel4, nstrands: int, nverts: int, dt: float, scale: float, sphere):
    """One Hair::update of the reference shader source (READ -> dispatch -> swap), in place."""
    _ref("glsl", nverts).ref_glsl_update(_p(pos4), _p(vel4), C.c_int64(nstrands), C.c_float(dt), C.c_float(scale),
                                         (C.c_float * 4)(*sphere))


def ref_init_simulation(root_pos3, root_nrm3, seed: int, nverts: int, maxlength: float = 0.5):
    S = root_pos3.shape[0]
    planes = [np.empty((S * nverts, 4), np.float32) for _ in range(3)]
    _ref("host", nverts).ref_host_init_simulation(_p(np.ascontiguousarray(root_pos3, np.float32)),
                                                  _p(np.ascontiguousarray(root_nrm3, np.float32)), C.c_int64(S),
                                                  C.c_uint(seed), C.c_float(maxlength), *[_p(a) for a in planes])
    return tuple(planes)


def ref_patch_indices(tri, nstrands: int, nverts: int) -> np.ndarray:
    tri = np.ascontiguousarray(tri, np.int32).reshape(-1, 3)
    out = np.empty(6 * tri.shape[0] * max(nverts - 1, 0), np.int32)
    n = _ref("host", nverts).ref_host_patch_indices(_p(tri), C.c_int64(tri.shape[0]), C.c_int64(nstrands), _p(out))
    assert n == out.size
    return out


def recalc_normals(pos3, corner_v) -> np.ndarray:
    """RawMeshData::recalculateNormals (raw_mesh_file.cc:11-50): one normal per face corner. corner_v: (F, 3) zero-based."""
    pos3 = np.ascontiguousarray(pos3, np.float32).reshape(-1, 3); corner_v = np.ascontiguousarray(corner_v, np.int32).reshape(-1, 3)
    out = np.empty((corner_v.size, 3), np.float32)
    oracle().bho_recalc_normals(_p(pos3), C.c_int64(pos3.shape[0]), _p(corner_v), C.c_int64(corner_v.shape[0]), _p(out))
    return out


def ref_recalc_normals(pos3, corner_v):
    """The reference's own function (oracle/_ref host harness). Returns (normals per corner (3F, 3), normal index per corner)."""
    pos3 = np.ascontiguousarray(pos3, np.float32).reshape(-1, 3); corner_v = np.ascontiguousarray(corner_v, np.int32).reshape(-1, 3)
    out = np.empty((corner_v.size, 3), np.float32); idx = np.empty(corner_v.size, np.int32)
    _ref("host", 4).ref_host_recalc_normals(_p(pos3), C.c_int64(pos3.shape[0]), _p(corner_v), C.c_int64(corner_v.shape[0]), _p(out), _p(idx))
    return out, idx


def dq_palette_from_matrices(global_pose, inverse_bind) -> np.ndarray:
    """generate_skinning_datas (skeleton_controller.cc:248-265). Matrices (J, 16) in GLM's column-major layout -> (J, 8)."""
    A = np.ascontiguousarray(global_pose, np.float32).reshape(-1, 16); B = np.ascontiguousarray(inverse_bind, np.float32).reshape(-1, 16)
    out = np.empty((A.shape[0], 8), np.float32)
    oracle().bho_dq_palette_from_matrices(_p(A), _p(B), C.c_int(A.shape[0]), _p(out))
    return out


def ref_dq_palette_from_matrices(global_pose, inverse_bind) -> np.ndarray:
    A = np.ascontiguousarray(global_pose, np.float32).reshape(-1, 16); B = np.ascontiguousarray(inverse_bind, np.float32).reshape(-1, 16)
    out = np.empty((A.shape[0], 8), np.float32)
    _ref("host", 4).ref_host_dq_palette(_p(A), _p(B), C.c_int(A.shape[0]), _p(out))
    return out


def ref_simplex2(x: float, y: float) -> float:
    return float(_ref("host", 4).ref_host_simplex2(x, y))


def tess_stream(pos4, tan4, patch, nverts: int, scale: float, ninstances: int, nlines: int, nsubsegments: int, seed: int):
    """Interpolated render strands of the tess-stream stage: (count, 4) float32 GL_LINES vertices (xyz, relPos)."""
    patch = np.ascontiguousarray(patch, np.int32)
    npatches = patch.size // 6
    lib = oracle()
    lib.bho_tess_stream_count.restype = C.c_int64
    n = lib.bho_tess_stream_count(C.c_int64(npatches), C.c_int(ninstances), C.c_int(nlines), C.c_int(nsubsegments))
    out = np.empty((n, 4), np.float32)
    lib.bho_tess_stream(_p(np.ascontiguousarray(pos4, np.float32)), _p(np.ascontiguousarray(tan4, np.float32)), _p(patch),
                        C.c_int64(npatches), C.c_int(nverts), C.c_float(scale), C.c_int(ninstances), C.c_int(nlines),
                        C.c_int(nsubsegments), C.c_uint32(seed), _p(out))
    return out


def tess_random_table(seed: int, size: int = 4096) -> np.ndarray:
    """(size, 2) float32: the oracle's seeded pair for every index of the reference's random table (HAIR_TF_RANDOMBUFFER_SIZE)."""
    out = np.empty((size, 2), np.float32)
    st = np.zeros(2, np.float32)
    for i in range(size):
        oracle().bho_tess_random_pair(C.c_uint32(seed), C.c_int(i), _p(st))
        out[i] = st
    return out


def ref_tess_skin_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libbarbu_ref_tess_skin.so"))


def _ref_tess_skin() -> C.CDLL:
    if "tess_skin" not in _ref_cache:
        _ref_cache["tess_skin"] = C.CDLL(os.path.join(REF_DIR, "libbarbu_ref_tess_skin.so"))
    return _ref_cache["tess_skin"]


def ref_tess_stream(pos4, tan4, patch, nverts: int, scale: float, ninstances: int, nlines: int, nsubsegments: int, randtable):
    """The reference's four tess-stream shader stages (SOURCE compiled over GLM, oracle/_ref) on the CPU; `randtable` is the
    content of the random buffer, (4096, 2) float32."""
    patch = np.ascontiguousarray(patch, np.int32)
    npatches = patch.size // 6
    out = np.empty((npatches * ninstances * nlines * nsubsegments * 2, 4), np.float32)
    randtable = np.ascontiguousarray(randtable, np.float32)
    assert randtable.shape == (4096, 2)
    _ref_tess_skin().ref_tess_stream(_p(np.ascontiguousarray(pos4, np.float32)), _p(np.ascontiguousarray(tan4, np.float32)), _p(patch),
                                     C.c_int64(npatches), C.c_int(nverts), C.c_float(scale), C.c_int(ninstances), C.c_int(nlines),
                                     C.c_int(nsubsegments), _p(randtable), _p(out))
    return out


def ref_skin_dq(rest_pos3, rest_nrm3, joints4, weights3, dq):
    """apply_skinning + skinning_DQBS of the reference's inc_skinning.glsl (SOURCE compiled over GLM, oracle/_ref)."""
    S = rest_pos3.shape[0]
    op, on = np.empty((S, 3), np.float32), np.empty((S, 3), np.float32)
    _ref_tess_skin().ref_skin_dq(_p(np.ascontiguousarray(rest_pos3, np.float32)), _p(np.ascontiguousarray(rest_nrm3, np.float32)),
                                 _p(np.ascontiguousarray(joints4, np.int32)), _p(np.ascontiguousarray(weights3, np.float32)),
                                 _p(np.ascontiguousarray(dq, np.float32)), C.c_int64(S), _p(op), _p(on))
    return op, on


# ---- scalp input ----------------------------------------------------------------------------------

def obj_scalp(path: str):
    """Plain-Python restatement of how the reference reads an OBJ scalp (small files only):
    ParseOBJ, src/memory/resources/mesh_data_manager.cc:69-223 — '\\n'-terminated lines; lines whose first byte is odd
    ('#', 'o', 'g', 's', 'u', 'm') carry no geometry; `v`, `vn`; `f` corners v/vt/vn, quads split (x, y, z), (z, w, x);
    MeshData::setup, src/memory/resources/mesh_data.cc:384-406 — vertices = unique (v, vt, vn) triples in
    first-appearance order. Returns (positions (S,3) f32, normals (S,3) f32, triangles (F,3) i32)."""
    raw = open(path, "rb").read().decode("latin-1")
    lines = raw.split("\n")[:-1]                       # a last line without '\n' is never reached by the strchr loop
    pos, nrm, ntex, corners = [], [], 0, []
    for s in lines:
        if not s or (ord(s[0]) & 1):
            continue
        if s[0] == "v":
            if s[1] == " ":
                pos.append([np.float32(x) for x in s[2:].split()[:3]])
            elif s[1] == "t":
                ntex += 1
            else:
                nrm.append([np.float32(x) for x in s[3:].split()[:3]])
        elif s[0] == "f":
            cs = []
            for tok in s[2:].split()[:4]:
                parts = tok.split("/")
                v = int(parts[0])
                t = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                n = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                cs.append((v, t, n))
            corners += cs[:3]
            if len(cs) == 4 and cs[3][0] > 0:
                corners += [cs[2], cs[3], cs[0]]
    if not nrm:
        # MeshData::setup -> RawMeshData::recalculateNormals (mesh_data.cc:366-372, raw_mesh_file.cc:11-50): every corner gets
        # a normal entry of its own, so all 3 * F corner triples are unique
        cv = np.array([c[0] - 1 for c in corners], np.int32).reshape(-1, 3)
        nrm = [list(x) for x in recalc_normals(np.array(pos, np.float32), cv)]
        corners = [(v, t, q + 1) for q, (v, t, _) in enumerate(corners)]
    seen, uniq, idx = {}, [], []
    for v, t, n in corners:
        key = (v - 1, t - 1, n - 1)
        if key not in seen:
            seen[key] = len(uniq)
            uniq.append(key)
        idx.append(seen[key])
    P = np.array([pos[k[0]] for k in uniq], np.float32).reshape(-1, 3)
    Nn = np.array([nrm[k[2]] for k in uniq], np.float32).reshape(-1, 3)
    return P, Nn, np.array(idx, np.int32).reshape(-1, 3)


# ---- Marschner lookup tables -------------------------------------------------------------------
MARSCHNER_DEFAULTS = dict(eta=1.55, absorption=0.20, eccentricity=0.85, ar=-5.0, br=5.0, glintScale=0.5, azimuthalWidth=10.0,
                          deltaCaustic=0.2, deltaHm=0.5)                       # marschner.h:38-52


def marschner_params(**kw) -> np.ndarray:
    d = dict(MARSCHNER_DEFAULTS); d.update(kw)
    return np.array([d[k] for k in MARSCHNER_DEFAULTS], np.float32)


def marschner_luts(params, resolution: int = 128):
    """(M, N) lookup tables as (res, res, 4) float32, texel (x, y) at [y, x]."""
    m = np.empty((resolution, resolution, 4), np.float32); n = np.empty_like(m)
    oracle().bho_marschner_luts(_p(np.ascontiguousarray(params, np.float32)), C.c_int(resolution), _p(m), _p(n))
    return m, n


def float_to_half(a) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty(a.shape, np.uint16)
    oracle().bho_float_to_half(_p(a), C.c_int64(a.size), _p(out))
    return out


def ref_marschner_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libbarbu_ref_marschner.so"))


def ref_marschner_luts(params, resolution: int = 128):
    """The reference shader SOURCES run on the CPU (oracle/_ref/libbarbu_ref_marschner.so)."""
    lib = C.CDLL(os.path.join(REF_DIR, "libbarbu_ref_marschner.so"))
    m = np.empty((resolution, resolution, 4), np.float32); n = np.empty_like(m)
    lib.ref_marschner_luts(_p(np.ascontiguousarray(params, np.float32)), C.c_int(resolution), _p(m), _p(n))
    return m, n
