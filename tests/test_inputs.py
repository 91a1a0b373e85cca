"""Scalp INPUT rows either side of the hot path (SURVEY.md §8 a1 / a-ext, VERDICT r1 "Inputs"): scalps without normals
(RawMeshData::recalculateNormals), the skinning palette from joint matrices (SkeletonController::generate_skinning_datas ->
glm::dualquat) and a skinned glTF scalp — oracle and product against known answers the REFERENCE's host code produced
(tests/golden/inputs.npz, made by tests/golden/make_inputs_golden.py through oracle/_ref), then the whole chain on the GPU."""
import os

import numpy as np
import pytest

import barbu_b200 as bb
from oracle import pyoracle as po
from tests.util import DT, GOLDEN, assert_bit_equal

GLTF = os.path.join(GOLDEN, "skinned_scalp.gltf")


def golden():
    return np.load(os.path.join(GOLDEN, "inputs.npz"))


def test_recalculated_normals_match_reference_golden():
    g = golden()
    got = po.recalc_normals(g["recalc_pos"], g["recalc_tri"])
    assert_bit_equal(got, g["recalc_nrm"], "per-corner normals (raw_mesh_file.cc:11-50)")
    assert np.isnan(g["recalc_nrm"]).any(), "the fixture holds zero-area faces"


def test_dq_palette_matches_reference_golden_oracle_and_product():
    g = golden()
    want = g["dq_palette"]
    assert_bit_equal(po.dq_palette_from_matrices(g["dq_global"], g["dq_inverse_bind"]), want, "oracle palette")
    assert_bit_equal(bb.dq_palette_from_matrices(g["dq_global"], g["dq_inverse_bind"]), want, "bh_dq_palette_from_matrices")
    # all four branches of dualquat_cast are in the fixture: which component of the rotation is 0.5 * r decides
    G = g["dq_global"][:6].reshape(6, 4, 4)
    tr = G[:, 0, 0] + G[:, 1, 1] + G[:, 2, 2]
    assert (tr > 0).any() and (tr <= 0).sum() >= 3
    # unit rotations, and the translation comes back out of the dual part: t = 2 * dual * conj(real)
    real, dual = want[:, :4].astype(np.float64), want[:, 4:].astype(np.float64)
    assert np.allclose(np.linalg.norm(real, axis=1), 1.0, atol=1e-5)
    with pytest.raises(ValueError):
        bb.dq_palette_from_matrices(g["dq_global"], g["dq_inverse_bind"][:-1])


def test_dq_palette_product_equals_oracle_on_random_matrices():
    """bh_dq_palette_from_matrices (host code of the product library) == the oracle (pinned to the reference source in
    tests/test_oracle_vs_reference_live.py) on arbitrary matrices, special values included; njoints = 0 is fine."""
    rng = np.random.default_rng(5)
    G = rng.standard_normal((500, 16)).astype(np.float32); B = rng.standard_normal((500, 16)).astype(np.float32)
    G[:100, [0, 5, 10]] = np.array([0.5, -0.25, -0.25], np.float32); B[:100] = np.eye(4, dtype=np.float32).reshape(16)
    G[100:120] = 0.0; G[120, 0] = np.inf; B[121, 7] = np.nan
    assert_bit_equal(bb.dq_palette_from_matrices(G, B), po.dq_palette_from_matrices(G, B), "palette")
    assert bb.dq_palette_from_matrices(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32)).shape == (0, 8)


def test_gltf_scalp_loads_in_first_appearance_order():
    """mesh_data_manager.cc:755-880 + mesh_data.cc:384-436: world matrix applied, corners (i, i, i) re-indexed in
    first-appearance order over the index list, joints and weights following their vertex."""
    g = golden()
    m = bb.load_gltf_scalp(GLTF)
    assert (m.nvertices, m.nfaces) == (84, 144)
    assert_bit_equal(m.positions, g["gltf_pos"], "positions"); assert_bit_equal(m.normals, g["gltf_nrm"], "normals")
    assert np.array_equal(m.joints, g["gltf_joints"]) and m.joints.dtype == np.int32
    assert_bit_equal(m.weights, g["gltf_weights"], "weights")
    assert np.array_equal(np.asarray(m.indices), g["gltf_indices"])
    first_use = []
    for i in np.asarray(m.indices).reshape(-1).tolist():
        if i not in first_use: first_use.append(i)
    assert first_use == list(range(84)), "vertex k is the k-th distinct index of the index list"
    assert m.joint_nodes == [1, 2, 3, 4]
    assert_bit_equal(m.inverse_bind, g["gltf_inverse_bind"], "inverse bind matrices")
    # rest pose: global * inverse bind = identity -> identity palette
    dq = bb.dq_palette_from_matrices(m.joint_rest_global, m.inverse_bind)
    assert np.array_equal(dq, np.tile(np.array([0, 0, 0, 1, 0, 0, 0, 0], np.float32), (4, 1)))


def _pose(m, frame):
    """Global joint matrices of a swaying chain: joint j rotates about z by an angle growing with j, about its own rest
    position (column-major 16-float rows)."""
    out = []
    for j, rest in enumerate(m.joint_rest_global.reshape(-1, 4, 4)):
        a = 0.12 * (j + 1) * np.sin(0.4 * frame + 0.3 * j)
        c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
        R = np.array([[c, s, 0, 0], [-s, c, 0, 0], [0, 0, 1, 0], [0.02 * frame, 0, 0, 1]], np.float32)   # [column][row]
        M = np.empty((4, 4), np.float32)
        for col in range(4):
            M[col] = ((rest[0] * R[col, 0] + rest[1] * R[col, 1]) + rest[2] * R[col, 2]) + rest[3] * R[col, 3]
        out.append(M.reshape(16))
    return np.stack(out)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [4, 16])
def test_skinned_gltf_scalp_end_to_end_bit_exact(N):
    """configs[2]'s input side on a real file: glTF scalp -> strands -> every frame: joint matrices -> palette
    (bh_dq_palette_from_matrices) -> roots skinned on the device -> step; against the oracle fed the reference-made palette."""
    m = bb.load_gltf_scalp(GLTF)
    S = m.nvertices
    sphere = (0.125, -0.25, 0.5, 0.85)
    rv = po.random_values(77, S)
    pos, vel = po.init_strands(m.positions, m.normals, rv, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere)
    w3 = np.ascontiguousarray(m.weights[:, :3])
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=sphere)
        sim.upload(pos, vel)
        sim.set_skin(m.positions, m.joints, w3)
        for frame in range(6):
            G = _pose(m, frame)
            dq = bb.dq_palette_from_matrices(G, m.inverse_bind)
            assert_bit_equal(dq, po.dq_palette_from_matrices(G, m.inverse_bind), "palette")
            sim.skin_roots(dq)
            sim.step(float(DT), 1)
            sp, _ = po.skin_roots_dq(m.positions, m.normals, m.joints, w3, dq)
            pos.reshape(S, N, 4)[:, 0, :3] = sp
            po.step(pos, vel, S, N, par)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, pos, "positions"); assert_bit_equal(gv, vel, "velocities")
    assert np.abs(gp.reshape(S, N, 4)[:, 0, :3] - m.positions).max() > 1e-3, "the roots moved with the skeleton"


@pytest.mark.gpu
def test_scalp_without_normals_grows_three_strands_per_face(tmp_path):
    """OBJ without `vn` -> recalculated per-corner normals -> Hair::setup: 3 * F roots, state bit-equal to the oracle's."""
    n = 6
    lines = [f"v {c / (n - 1):.5f} {0.1 * ((r * 7 + c * 3) % 5):.5f} {r / (n - 1):.5f}" for r in range(n) for c in range(n)]
    vid = lambda r, c: r * n + c + 1
    for r in range(n - 1):
        for c in range(n - 1):
            lines.append(f"f {vid(r, c)} {vid(r + 1, c)} {vid(r + 1, c + 1)} {vid(r, c + 1)}")
    path = tmp_path / "bare.obj"; path.write_text("\n".join(lines) + "\n")
    P, Nn, T = po.obj_scalp(str(path))
    h = bb.Hair()
    h.set_bounding_sphere((0.5, -2.0, 0.5, 0.4))
    h.setup(str(path))
    assert h.initialized() and h.nroots == 3 * 2 * (n - 1) ** 2 == P.shape[0]
    gp, gv, gt = h.sim.download(tan=True)
    pos, vel = po.init_strands(P, Nn, po.random_values(h.params.seed, h.nroots), h.params.ncontrol_points)
    assert_bit_equal(gp, pos, "strands grown along the recalculated normals")
    assert_bit_equal(gt, po.init_tangents(Nn, h.params.ncontrol_points), "tangents")
    assert_bit_equal(h.patch_indices, po.patch_indices(T, h.params.ncontrol_points))
    h.deinit()
