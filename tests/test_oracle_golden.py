"""The oracle against what the REFERENCE SOURCES compute (fixtures made by tests/golden/make_golden.py
from oracle/_ref). Bit-exact: this is what pins the oracle."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_bit_equal, golden

CASES = ["hair_N2_s100", "hair_N3_s145", "hair_N4_s145", "hair_N8_s145", "hair_N16_s100", "hair_N32_s100",
         "hair_N32_s145", "hair_N64_s100"]


@pytest.mark.parametrize("name", CASES)
def test_init_simulation_matches_reference(name):
    g = golden(name)
    N, S = int(g["nverts"]), g["root_pos"].shape[0]
    rv = po.random_values(int(g["seed"]), S)
    pos, vel = po.init_strands(g["root_pos"], g["root_nrm"], rv, N)
    assert_bit_equal(pos, g["pos0"], "positions")
    assert_bit_equal(vel, g["vel0"], "velocities")
    assert_bit_equal(po.init_tangents(g["root_nrm"], N), g["tan0"], "tangents")


@pytest.mark.parametrize("name", CASES)
def test_patch_indices_match_reference(name):
    g = golden(name)
    assert_bit_equal(po.patch_indices(g["tri"], int(g["nverts"])), g["patch"], "patch indices")


@pytest.mark.parametrize("name", CASES)
def test_step_matches_reference_shader(name):
    g = golden(name)
    N, S = int(g["nverts"]), g["root_pos"].shape[0]
    p = po.default_params(dt=float(g["dt"]), scale=float(g["scale"]), sphere=tuple(g["sphere"]))
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    for it in range(1, 12):
        po.step(pos, vel, S, N, p)
        if it in (1, 10):
            assert_bit_equal(pos, g[f"pos{it}"], f"pos after {it} updates")
            assert_bit_equal(vel, g[f"vel{it}"], f"vel after {it} updates")
    assert_bit_equal(pos, g["posw1"], "pos after 11 updates")
    assert_bit_equal(vel, g["velw1"], "vel after 11 updates")


def test_step_multithreaded_equals_serial():
    g = golden("hair_N16_s100")
    N, S = 16, g["root_pos"].shape[0]
    p = po.default_params(dt=float(g["dt"]), scale=1.0, sphere=tuple(g["sphere"]))
    a, av, b, bv = g["pos0"].copy(), g["vel0"].copy(), g["pos0"].copy(), g["vel0"].copy()
    for _ in range(3):
        po.step(a, av, S, N, p, nthreads=1)
        po.step(b, bv, S, N, p, nthreads=4)
    assert_bit_equal(a, b)
    assert_bit_equal(av, bv)


def test_simplex_matches_glm():
    g = golden("simplex")
    o = po.oracle()
    val = np.array([o.bho_simplex2(float(x), float(y)) for x, y in g["pts"]], np.float32)
    assert_bit_equal(val, g["val"], "glm::simplex(vec2)")


@pytest.mark.parametrize("name", ["hair_N4_s145", "hair_N16_s100"])
def test_tess_stream_matches_reference_shader_stages(name):
    """tests/golden/tess_skin.npz: the reference's vs/tcs/tes/gs_stream_hair.glsl + inc_maths.glsl run over GLM (oracle/_ref)."""
    g, t = golden(name), golden("tess_skin")
    ninst, nlines, nsub, seed = (int(x) for x in t[f"tess_{name}_args"])
    npatch = t[f"tess_{name}"].shape[0] // (ninst * nlines * nsub * 2)
    got = po.tess_stream(g["pos10"], g["tan0"], g["patch"][:6 * npatch], int(g["nverts"]), float(t[f"tess_{name}_scale"]), ninst, nlines, nsub, seed)
    assert_bit_equal(got, t[f"tess_{name}"], "tess-stream vertices")


def test_dq_skinning_matches_reference_shader():
    """tests/golden/tess_skin.npz: apply_skinning + skinning_DQBS of inc_skinning.glsl run over GLM (oracle/_ref)."""
    t = golden("tess_skin")
    p, n = po.skin_roots_dq(t["skin_pos"], t["skin_nrm"], t["skin_joints"], t["skin_weights"], t["skin_dq"])
    assert_bit_equal(p, t["skin_out_pos"], "skinned positions")
    assert_bit_equal(n, t["skin_out_nrm"], "skinned normals")
