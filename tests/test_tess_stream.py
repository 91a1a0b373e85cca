"""Tess-stream stage: interpolated render strands as the GL_LINES vertex stream the reference's
transform-feedback pass produces (shaders/hair/02_tess_stream, hair.cc:141-173). The CUDA kernel against the oracle
(bit-exact), and the oracle against properties that follow from the shader text."""
import numpy as np
import pytest

import barbu_b200 as bb
from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, sphere_state


def scene(rows, cols, N, steps=0):
    root_pos, root_nrm, tri, rv, pos, vel = sphere_state(rows, cols, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(steps):
        po.step(pos, vel, rows * cols, N, par)
    return root_nrm, tri, pos, vel, po.init_tangents(root_nrm, N), po.patch_indices(tri, N)


def test_oracle_stream_properties():
    N, ninst, nlines, nsub = 4, 3, 2, 16                                  # the reference defaults (hair.h:33-35, interop.h:8)
    _, tri, pos, _, tan, patch = scene(6, 8, N, steps=5)
    out = po.tess_stream(pos, tan, patch, N, 1.45, ninst, nlines, nsub, 11)
    npatches = patch.size // 6
    assert npatches == tri.shape[0] * (N - 1) and out.shape == (npatches * ninst * nlines * nsub * 2, 4)
    seg = out.reshape(ninst, npatches, nlines, nsub, 2, 4)
    assert_bit_equal(seg[:, :, :, :-1, 1], seg[:, :, :, 1:, 0], "a segment starts where the previous one ended")
    # the curve through segment j of a triangle ends where segment j+1 starts (Hermite end points are the control points)
    p = seg.reshape(ninst, tri.shape[0], N - 1, nlines, nsub, 2, 4)
    assert np.abs(p[:, :, :-1, :, -1, 1, :3] - p[:, :, 1:, :, 0, 0, :3]).max() < 1e-6
    # x = 0 of the first segment is a barycentric mix of the three master roots, with the pair of (instance, line)
    st = np.zeros(2, np.float32)
    for inst in range(ninst):
        for line in range(nlines):
            y = np.float32(line) / np.float32(nlines)
            po.oracle().bho_tess_random_pair(11, int(np.float32(y * np.float32(40)) + np.float32(inst)) % 4096, st.ctypes.data_as(po.C.c_void_p))
            a, b = float(st[0]), float(st[1])
            if a + b > 1:
                a = 1 - max(a, b)
            w = np.array([a, b, 1 - (a + b)])
            roots = pos[:, :3].reshape(-1, N, 3)[tri[:, :], 0]                # (F, 3 corners, xyz)
            want = (roots * w[None, :, None]).sum(axis=1)
            assert np.abs(p[inst, :, 0, line, 0, 0, :3] - want).max() < 1e-5
    rel = seg[..., 3]
    assert rel.min() >= 0 and rel.max() <= 1 and (np.diff(rel.reshape(-1, nsub * 2)[:, ::2], axis=1) >= 0).all()
    assert 0 <= st.min() and st.max() < 1


@pytest.mark.gpu
@pytest.mark.parametrize("N,ninst,nlines,nsub,rows,cols", [(4, 3, 2, 16, 6, 8), (16, 1, 1, 1, 5, 7), (32, 2, 5, 7, 8, 16), (8, 4, 3, 64, 4, 8),
                                                          # > 7 (instance, isoline) pairs: the thread-per-segment kernel; 31 / 33 points per patch: warp seams
                                                          (8, 5, 3, 8, 4, 8), (16, 7, 4, 30, 6, 8), (4, 2, 2, 32, 16, 16), (16, 13, 1, 3, 5, 7)])
def test_tess_stream_bit_exact_vs_oracle(N, ninst, nlines, nsub, rows, cols):
    S = rows * cols
    _, tri, pos, vel, tan, patch = scene(rows, cols, N, steps=3)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE)
        sim.upload(pos, vel, tan)
        sim.tess_set_patches(patch)
        got = sim.tess_stream(ninst, nlines, nsub, seed=99)
        # and after a device step, from the stepped planes
        sim.step(float(DT), 1)
        got2 = sim.tess_stream(ninst, nlines, nsub, seed=99)
        p1, _, _ = sim.download()
    assert_bit_equal(got, po.tess_stream(pos, tan, patch, N, 1.45, ninst, nlines, nsub, 99), "stream from the uploaded state")
    assert_bit_equal(got2, po.tess_stream(p1, tan, patch, N, 1.45, ninst, nlines, nsub, 99), "stream after a step")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hair_N4_s145", "hair_N16_s100"])
def test_tess_stream_bit_exact_vs_reference_fixture(name):
    """The CUDA kernel against tests/golden/tess_skin.npz — vertices the REFERENCE's four tess-stream shader stages and
    inc_maths.glsl produced when run over GLM (tests/golden/make_tess_skin_golden.py), no oracle in between."""
    from tests.util import golden
    g, t = golden(name), golden("tess_skin")
    ninst, nlines, nsub, seed = (int(x) for x in t[f"tess_{name}_args"])
    N, S = int(g["nverts"]), g["root_pos"].shape[0]
    npatch = t[f"tess_{name}"].shape[0] // (ninst * nlines * nsub * 2)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=float(t[f"tess_{name}_scale"]))
        sim.upload(g["pos10"], g["vel10"], g["tan0"])
        sim.tess_set_patches(g["patch"][:6 * npatch])
        got = sim.tess_stream(ninst, nlines, nsub, seed=seed)
    assert_bit_equal(got, t[f"tess_{name}"], "tess-stream vertices vs the reference shader stages")


@pytest.mark.gpu
def test_tess_stream_api_errors():
    with bb.HairSim(8, 4) as sim:
        with pytest.raises(bb.BarbuHairError):
            sim.tess_stream()                                              # no strand state
        sim.upload(np.zeros((32, 4), np.float32), np.zeros((32, 4), np.float32), np.zeros((32, 4), np.float32))
        with pytest.raises(bb.BarbuHairError):
            sim.tess_stream()                                              # no patches
        with pytest.raises(bb.BarbuHairError):
            sim.tess_set_patches(np.arange(5, dtype=np.int32))             # not 6 per patch
        with pytest.raises(bb.BarbuHairError):
            sim.tess_set_patches(np.full(6, 99, np.int32))                 # outside the vertex range
        sim.tess_set_patches(np.array([0, 1, 4, 5, 8, 9], np.int32))
        with pytest.raises(ValueError):
            sim.tess_stream(0, 1, 1)
        assert sim.tess_stream(2, 3, 4).shape == (2 * 3 * 4 * 2, 4)
