"""Scalp input: the OBJ reading rules of the reference loader and its vertex re-indexing, as
bh_load_obj_scalp implements them, against the oracle's Python restatement — on synthetic files everywhere, and on the
reference's own Head_scalp.obj (counts + checksums in tests/golden/head_scalp.json) where the reference tree exists."""
import json
import os

import numpy as np
import pytest

import barbu_b200 as bb
from barbu_b200 import hair
from oracle import pyoracle as po
from tests.util import DT, GOLDEN, assert_bit_equal

ASSET = "/root/reference/assets/models/InfiniteScan/Head_scalp.obj"

OBJ_QUADS = """# comment
o patch
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 0 0.5
v 2 1 0.5
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
vn 0.6 0 0.8
s off
usemtl skin
f 1/1/1 2/2/1 3/3/1 4/4/1
f 2/2/1 5/1/2 6/4/2 3/3/1
f 2/2/2 5/1/2 6/4/2
f 1/1/1 2/2/1 3/3/1
v 9 9 9
f 1/1/1 2/2/1 7/1/2
f 4/4/1 3/3/1 2/2/1 without-newline-is-ignored"""


def write(tmp_path, text, name="scalp.obj"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_obj_rules_quads_reindexing_and_last_line(tmp_path):
    path = write(tmp_path, OBJ_QUADS)
    m = bb.load_obj_scalp(path)
    P, Nn, T = po.obj_scalp(path)
    assert_bit_equal(m.positions, P, "positions")
    assert_bit_equal(m.normals, Nn, "normals")
    assert_bit_equal(m.indices, T, "triangles")
    # quads -> (x, y, z), (z, w, x); repeated corners share a vertex, a corner with another normal does not
    assert m.nfaces == 2 + 2 + 1 + 1 + 1 and T[:2].tolist() == [[0, 1, 2], [2, 3, 0]]
    assert m.nvertices == 8     # 1/1/1 2/2/1 3/3/1 4/4/1, 5/1/2 6/4/2, 2/2/2, 7/1/2 in first-appearance order
    assert T[5].tolist() == [0, 1, 2], "a repeated face reuses the vertices"


def test_obj_only_normals_and_errors(tmp_path):
    path = write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n")
    m = bb.load_obj_scalp(path)
    assert (m.nvertices, m.nfaces) == (3, 1)
    assert_bit_equal(m.positions, po.obj_scalp(path)[0])
    with pytest.raises(bb.BarbuHairError) as e:
        bb.load_obj_scalp(write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n", "nonormals_badindex.obj"))
    assert e.value.code == hair.BH_ERR_INVALID
    with pytest.raises(bb.BarbuHairError) as e:
        bb.load_obj_scalp(str(tmp_path / "missing.obj"))
    assert e.value.code == hair.BH_ERR_INVALID
    with pytest.raises(bb.BarbuHairError):
        bb.load_obj_scalp(write(tmp_path, "v 0 0 0\nvn 0 0 1\nf 1//1 2//1 3//1\n", "badindex.obj"))
    h = bb.Hair()
    h.setup(str(tmp_path / "missing.obj"))           # Hair::setup with an unknown resource id: log + uninitialised
    assert not h.initialized() and "not found" in h.log[-1]


def _grid_obj(n, with_vt, quads, seed=0):
    """An n x n height-field patch without `vn` lines: quads or triangles, optionally with texture coordinates."""
    rng = np.random.default_rng(seed)
    lines = ["# scalp without normals", "o patch"]
    for r in range(n):
        for c in range(n):
            lines.append(f"v {c / (n - 1):.6f} {0.2 * np.sin(3.0 * r / n) + 0.01 * rng.standard_normal():.6f} {r / (n - 1):.6f}")
    if with_vt:
        for r in range(n):
            for c in range(n):
                lines.append(f"vt {c / (n - 1):.4f} {r / (n - 1):.4f}")
    vid = lambda r, c: r * n + c + 1
    fmt = (lambda v: f"{v}/{v}") if with_vt else (lambda v: f"{v}")
    for r in range(n - 1):
        for c in range(n - 1):
            a, b, d, e = vid(r, c), vid(r + 1, c), vid(r + 1, c + 1), vid(r, c + 1)
            if quads: lines.append("f " + " ".join(fmt(v) for v in (a, b, d, e)))
            else: lines += ["f " + " ".join(fmt(v) for v in (a, b, e)), "f " + " ".join(fmt(v) for v in (e, b, d))]
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("with_vt,quads", [(False, False), (False, True), (True, False), (True, True)])
def test_obj_without_normals_gets_the_reference_recalculated_normals(tmp_path, with_vt, quads):
    """A scalp file without `vn` lines: the normals RawMeshData::recalculateNormals makes (raw_mesh_file.cc:11-50; the oracle's
    restatement is pinned to that source in test_oracle_vs_reference_live.py), one per face corner — so 3 roots per triangle
    (mesh_data.cc:384-406 finds every corner triple unique), in corner order."""
    n = 9
    path = write(tmp_path, _grid_obj(n, with_vt, quads), "nonormals.obj")
    m = bb.load_obj_scalp(path)
    P, Nn, T = po.obj_scalp(path)
    nf = 2 * (n - 1) * (n - 1)
    assert (m.nvertices, m.nfaces) == (3 * nf, nf) and P.shape == (3 * nf, 3)
    assert_bit_equal(m.positions, P, "root positions"); assert_bit_equal(m.normals, Nn, "recalculated normals"); assert_bit_equal(np.asarray(m.indices, np.int32).reshape(-1, 3), T)
    assert np.asarray(m.indices).reshape(-1).tolist() == list(range(3 * nf))
    ln = np.linalg.norm(m.normals.astype(np.float64), axis=1)
    assert np.all(np.abs(ln - 1.0) < 1e-6), "unit normals"
    # corners of one vertex share the vertex normal: an interior vertex of the triangle grid belongs to six faces
    first = {}
    for q, v in enumerate(np.round(m.positions, 6).tolist()):
        first.setdefault(tuple(v), []).append(q)
    assert max(len(v) for v in first.values()) == (6 if not quads else 6)
    for qs in first.values():
        for q in qs[1:]:
            assert np.array_equal(m.normals[q].view(np.uint32), m.normals[qs[0]].view(np.uint32))


@pytest.mark.skipif(not os.path.exists(ASSET), reason="reference tree not present (GPU box)")
def test_reference_head_scalp_asset_matches_golden():
    g = json.load(open(os.path.join(GOLDEN, "head_scalp.json")))
    m = bb.load_obj_scalp(ASSET)
    assert (m.nvertices, m.nfaces) == (g["nvertices"], g["nfaces"]) == (448, 760)
    assert po.fnv1a64(m.positions) == g["fnv_positions"]
    assert po.fnv1a64(m.normals) == g["fnv_normals"]
    assert po.fnv1a64(m.indices) == g["fnv_triangles"]
    assert m.indices[:3].tolist() == g["first_triangles"]
    # host-side generation from that scalp (tangents; positions need the device)
    assert po.fnv1a64(bb.init_tangents_host(m.normals, m.nvertices, 0, 4)) == g["fnv_tan0"]


@pytest.mark.gpu
@pytest.mark.parametrize("policy", ["adaptor default", "throughput"])
def test_hair_setup_from_obj_reference_default_shape_bit_exact(tmp_path, policy):
    """Hair::setup(resource) + 10 updates with the reference's defaults (N = 4 control points, uScaleFactor 1.45),
    on an OBJ scalp written here (a bumpy quad patch), against the oracle."""
    n = 12
    lines = ["o scalp"]
    rng = np.random.default_rng(5)
    for i in range(n):
        for j in range(n):
            lines.append(f"v {i / n:.6f} {1.0 + 0.05 * rng.standard_normal():.6f} {j / n:.6f}")
    for i in range(n):
        for j in range(n):
            v = rng.standard_normal(3) * 0.2 + (0, 1, 0)
            v /= np.linalg.norm(v)
            lines.append(f"vn {v[0]:.4f} {v[1]:.4f} {v[2]:.4f}")
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j + 1, i * n + j + 2, (i + 1) * n + j + 2, (i + 1) * n + j + 1
            lines.append(f"f {a}//{a} {b}//{b} {c}//{c} {d}//{d}")
    path = write(tmp_path, "\n".join(lines) + "\n")
    P, Nn, T = po.obj_scalp(path)
    S, N = P.shape[0], 4
    sphere = (0.5, 0.9, 0.5, 0.3)
    h = bb.Hair()
    h.init()
    h.set_bounding_sphere(sphere)
    h.setup(path)
    assert h.initialized() and h.nroots == S == n * n
    if policy == "throughput":
        h.sim.set_step_policy(bb.BH_POLICY_THROUGHPUT)
        assert h.sim.kernel_kind == 0                  # N = 4: the streaming kernel, two strands per 128-byte tensor row
    else:
        assert h.sim.kernel_kind == 3                  # the adaptor asks for BH_POLICY_AUTO: a scalp this small takes the latency kernel
    for _ in range(10):
        h.update(float(DT))
    gp, gv, gt = h.sim.download(tan=True)
    pos, vel = po.init_strands(P, Nn, po.random_values(1234, S), N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere)
    for _ in range(10):
        po.step(pos, vel, S, N, par)
    assert_bit_equal(gp, pos, "positions")
    assert_bit_equal(gv, vel, "velocities")
    assert_bit_equal(gt, po.init_tangents(Nn, N), "tangents")
    assert_bit_equal(h.patch_indices, po.patch_indices(T, N), "patch indices")
    h.deinit()
