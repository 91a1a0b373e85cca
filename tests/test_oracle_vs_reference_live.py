"""The oracle against the reference shader SOURCE executed here (oracle/_ref: cs_simulation.glsl compiled over the reference's
GLM) on RANDOMISED states — beyond the committed golden shapes: random roots and velocities, spheres on and
off the origin, contacts from the first step, degenerate segments (NaN by the reference's own arithmetic), special values.
Bit-exact. Skipped where oracle/_ref was not built (it needs the reference checkout; the committed fixtures still pin the
oracle there)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_bit_equal

NVERTS = [2, 3, 4, 8, 16, 32, 64, 128]


def random_state(rng, S, N, spread):
    root = rng.standard_normal((S, 3)).astype(np.float32)
    root /= np.linalg.norm(root, axis=1, keepdims=True).astype(np.float32)
    nrm = root + (rng.standard_normal((S, 3)) * spread).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    rv = (1.0 + 0.1 * (1.0 - 2.0 * rng.random(S))).astype(np.float32)
    pos, vel = po.init_strands(root, nrm.astype(np.float32), rv, N)
    vel[:, :3] = (rng.standard_normal((S * N, 3)) * spread).astype(np.float32)
    pos[:, :3] += (rng.standard_normal((S * N, 3)) * spread * 0.05).astype(np.float32)
    return pos, vel


@pytest.mark.parametrize("N", NVERTS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_states_bit_exact_vs_reference_shader(N, seed):
    if not po.ref_available(N):
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(1000 * N + seed)
    S = 48
    spread = [0.02, 0.3, 2.0][seed]                                    # gentle, rough, violent
    pos, vel = random_state(rng, S, N, spread)
    sphere = [(0.0, 0.0, 0.0, 0.98), (0.1, -0.2, 0.05, 1.1), (0.0, 0.5, 0.0, 0.6)][seed]
    scale = [1.0, 1.45, 0.7][seed]
    dt = float(np.float32(1.0) / np.float32(90.0)) * [1.0, 0.25, 2.0][seed]
    par = po.default_params(dt=dt, scale=scale, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for it in range(6):
        po.step(pos, vel, S, N, par)
        po.ref_update(rp, rv, S, N, dt, scale, sphere)
        assert_bit_equal(pos, rp, f"positions after {it + 1} updates")
        assert_bit_equal(vel, rv, f"velocities after {it + 1} updates")


@pytest.mark.parametrize("N", [4, 8, 32])
def test_degenerate_and_special_values_bit_exact_vs_reference_shader(N):
    """Zero-length segments (normalize(0) -> NaN in the reference's arithmetic), a vertex at the sphere centre, -0.0,
    huge and denormal coordinates, infinite velocity: the oracle must produce what the shader source produces."""
    if not po.ref_available(N):
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(77 + N)
    S = 16
    pos, vel = random_state(rng, S, N, 0.05)
    P = pos.reshape(S, N, 4); V = vel.reshape(S, N, 4)
    P[0, 1, :3] = P[0, 0, :3]                                          # coincident with the root: vdiff = 0
    P[1, N - 1, :3] = 0.0                                              # at the sphere centre: inversesqrt(0)
    P[2, 1, :3] = -0.0
    P[3, 1, :3] = 1.0e19; P[3, N - 1, :3] = -3.0e38                    # dot overflows to inf
    P[4, 1, :3] = P[4, 0, :3] + np.float32(1e-30)                      # denormal-range squared length
    V[5, 1, :3] = np.inf; V[6, N - 1, 0] = np.nan
    P[7, :, 3] = 0.0                                                   # zero rest lengths
    P[8, 1:, :3] = P[8, 0, :3]                                         # a whole strand collapsed onto its root
    par = po.default_params(dt=float(np.float32(1.0) / np.float32(90.0)), scale=1.45, sphere=(0.0, 0.0, 0.0, 0.98))
    rp, rv = pos.copy(), vel.copy()
    for it in range(3):
        po.step(pos, vel, S, N, par)
        po.ref_update(rp, rv, S, N, par.dt, 1.45, (0.0, 0.0, 0.0, 0.98))
        assert_bit_equal(pos, rp, f"positions after {it + 1} updates")
        assert_bit_equal(vel, rv, f"velocities after {it + 1} updates")


@pytest.mark.parametrize("N", [2, 4, 16, 128])
@pytest.mark.parametrize("seed", [3, 99, 123456])
def test_host_generation_bit_exact_vs_reference_host_code(N, seed):
    """Hair::init_simulation (hair.cc:236-361: jitter by glibc rand(), rest lengths, tangents with glm::simplex) and
    Hair::init_mesh's element loop (hair.cc:397-409), sliced from the reference's hair.cc and run live, on random scalps."""
    if not po.ref_available(N):
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(seed)
    S = 40
    root = (rng.standard_normal((S, 3)) * 3.0).astype(np.float32)
    nrm = rng.standard_normal((S, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    maxlength = float(np.float32(rng.uniform(0.1, 2.0)))
    rpos, rvel, rtan = po.ref_init_simulation(root, nrm, seed, N, maxlength)
    pos, vel = po.init_strands(root, nrm, po.random_values(seed, S), N, maxlength)
    assert_bit_equal(pos, rpos, "positions + rest lengths")
    assert_bit_equal(vel, rvel, "velocities")
    assert_bit_equal(po.init_tangents(nrm, N, maxlength), rtan, "tangents")
    tri = rng.integers(0, S, (25, 3)).astype(np.int32)
    assert_bit_equal(po.patch_indices(tri, N), po.ref_patch_indices(tri, S, N), "patch indices")


def test_simplex_noise_bit_exact_vs_glm_live():
    if not po.ref_available(4):
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(5)
    o = po.oracle()
    for x, y in (rng.standard_normal((500, 2)) * 37.0).astype(np.float32):
        a, b = np.float32(o.bho_simplex2(float(x), float(y))), np.float32(po.ref_simplex2(float(x), float(y)))
        assert a.view(np.uint32) == b.view(np.uint32), (x, y, a, b)


# ---- the stages either side of the simulation, pinned the same way ---------------------

@pytest.mark.parametrize("N,ninst,nlines,nsub", [(4, 3, 2, 16), (2, 1, 1, 1), (16, 2, 5, 7), (32, 4, 3, 9), (8, 45, 2, 3)])
@pytest.mark.parametrize("seed", [0, 7])
def test_tess_stream_bit_exact_vs_reference_shader_stages(N, ninst, nlines, nsub, seed):
    """bho_tess_stream against the reference's vs/tcs/tes/gs_stream_hair.glsl + inc_maths.glsl (hermite_mix,
    sample_triangle2, maprange, smoothstep2) run live over GLM, on random control points, tangents and patch lists. The
    random table handed to the reference stages holds the oracle's seeded pairs; the index expression is the reference's."""
    if not po.ref_tess_skin_available():
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(31 * N + seed)
    S = 24
    pos = np.zeros((S * N, 4), np.float32); tan = np.zeros((S * N, 4), np.float32)
    pos[:, :3] = (rng.standard_normal((S * N, 3)) * [1.0, 30.0][seed % 2]).astype(np.float32)
    pos[:, 3] = rng.random(S * N).astype(np.float32)
    tan[:, :3] = (rng.standard_normal((S * N, 3)) * 0.3).astype(np.float32)
    tri = rng.integers(0, S, (17, 3)).astype(np.int32)
    patch = po.patch_indices(tri, N)
    scale = [1.45, 0.37][seed % 2]
    table = po.tess_random_table(1234 + seed)
    got = po.tess_stream(pos, tan, patch, N, scale, ninst, nlines, nsub, 1234 + seed)
    want = po.ref_tess_stream(pos, tan, patch, N, scale, ninst, nlines, nsub, table)
    assert_bit_equal(got, want, "tess-stream vertices (xyz, relPos)")


def test_tess_stream_special_values_bit_exact_vs_reference_shader_stages():
    """Pairs on and across the fold of sample_triangle2 (s + t == 1, > 1 with s < t and s > t: the reference's overwrite of
    st.x before st.y reads it), -0.0 / huge / NaN control points."""
    if not po.ref_tess_skin_available():
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    N, S = 4, 8
    rng = np.random.default_rng(3)
    pos = np.zeros((S * N, 4), np.float32); tan = np.zeros((S * N, 4), np.float32)
    pos[:, :3] = rng.standard_normal((S * N, 3)).astype(np.float32)
    tan[:, :3] = rng.standard_normal((S * N, 3)).astype(np.float32)
    pos[1, :3] = -0.0; pos[6, :3] = 3.0e37; pos[9, 0] = np.nan; tan[13, :3] = np.inf
    tri = np.array([[0, 1, 2], [3, 4, 5], [5, 6, 7], [2, 2, 2]], np.int32)
    patch = po.patch_indices(tri, N)
    for seed in range(40):                                               # many tables: both sides of the fold for every (line, instance)
        table = po.tess_random_table(seed)
        want = po.ref_tess_stream(pos, tan, patch, N, 1.45, 3, 4, 5, table)
        assert_bit_equal(po.tess_stream(pos, tan, patch, N, 1.45, 3, 4, 5, seed), want, f"seed {seed}")
    s = po.tess_random_table(0)
    assert ((s.sum(axis=1) > 1) & (s[:, 0] < s[:, 1])).any() and ((s.sum(axis=1) > 1) & (s[:, 0] > s[:, 1])).any()


def random_dq_palette(rng, njoints, flip_some=True):
    """Unit dual quaternions (real xyzw, dual xyzw) of random rigid transforms; some negated (q and -q are one rotation):
    the antipodality fix of skinning_DQBS must see both signs."""
    q = rng.standard_normal((njoints, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.standard_normal((njoints, 3)) * 0.5
    x, y, z, w = q.T
    dual = 0.5 * np.stack([t[:, 0] * w + t[:, 1] * z - t[:, 2] * y, -t[:, 0] * z + t[:, 1] * w + t[:, 2] * x,
                           t[:, 0] * y - t[:, 1] * x + t[:, 2] * w, -t[:, 0] * x - t[:, 1] * y - t[:, 2] * z], axis=1)
    dq = np.concatenate([q, dual], axis=1)
    if flip_some:
        dq[rng.random(njoints) < 0.5] *= -1.0
    return dq.astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_dq_skinning_bit_exact_vs_reference_shader(seed):
    """bho_skin_roots_dq against apply_skinning + skinning_DQBS of the reference's inc_skinning.glsl run live over GLM:
    random palettes with antipodal joints, weights at and below the Epsilon() early-out, zero and negative weights,
    repeated joints, orthogonal real parts (sign(0) = 0)."""
    if not po.ref_tess_skin_available():
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(500 + seed)
    S, J = 400, 13
    dq = random_dq_palette(rng, J)
    if seed == 3:                                                        # exactly orthogonal real parts: dot == 0 -> sign 0
        dq[0, :4] = (1, 0, 0, 0); dq[1, :4] = (0, 1, 0, 0); dq[2, :4] = (0, 0, 0, 1)
    pos = (rng.standard_normal((S, 3)) * 2.0).astype(np.float32)
    nrm = rng.standard_normal((S, 3)).astype(np.float32)
    joints = rng.integers(0, J, (S, 4)).astype(np.int32)
    if seed == 3:
        joints[:50] = rng.integers(0, 3, (50, 4))
    w = rng.dirichlet([1.0, 1.0, 1.0, 1.0], S)[:, :3].astype(np.float32)
    w[0] = (0.0, 0.5, 0.5); w[1] = (1e-6, 0.5, 0.4); w[2] = (np.float32(1e-6) + np.float32(1e-12), 0.3, 0.3)
    w[3] = (1.0000001e-6, 0.0, 0.0); w[4] = (1.0, 0.0, 0.0); w[5] = (0.7, 0.7, 0.7); w[6] = (0.5, -0.25, 0.5)
    w[7] = (-0.5, 0.5, 0.5); w[8] = (np.nan, 0.5, 0.5)
    joints[9] = (4, 4, 4, 4)
    gp, gn = po.skin_roots_dq(pos, nrm, joints, w, dq)
    rp, rn = po.ref_skin_dq(pos, nrm, joints, w, dq)
    assert_bit_equal(gp, rp, "skinned positions")
    assert_bit_equal(gn, rn, "skinned normals")
    assert np.array_equal(gp[0], pos[0]) and np.array_equal(gp[1], pos[1])      # weights.x <= Epsilon(): untouched


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_recalculated_normals_bit_exact_vs_reference_host_code(seed):
    """Scalps without normals: bho_recalc_normals == RawMeshData::recalculateNormals (raw_mesh_file.cc:11-50, compiled here
    from the reference source) on random triangle soups, including zero-area faces (normalize(0): NaN by the reference's own
    arithmetic, which then poisons the three vertex sums), faces that repeat a vertex, and vertices no face uses."""
    if not po.ref_available(4):
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(500 + seed)
    nv, nf = 60 + 40 * seed, 150 + 100 * seed
    pos = (rng.standard_normal((nv, 3)) * [1.0, 1e-3, 1e4][seed]).astype(np.float32)
    tri = rng.integers(0, nv - 3, (nf, 3)).astype(np.int32)                 # the last three vertices stay unused: 0 / 0
    tri[5] = [7, 7, 9]                                                       # a repeated vertex: zero-area face
    pos[11] = pos[12]; tri[6] = [11, 12, 13]                                 # coincident positions: zero-area face
    want, idx = po.ref_recalc_normals(pos, tri)
    assert idx.tolist() == list(range(3 * nf)), "every corner gets a normal entry of its own, in corner order"
    got = po.recalc_normals(pos, tri)
    assert_bit_equal(got, want, "per-corner normals")
    assert np.isnan(want).any() and np.isfinite(want).any()


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_dq_palette_bit_exact_vs_reference_host_code(seed):
    """bho_dq_palette_from_matrices == SkeletonController::generate_skinning_datas + glm::dualquat(mat3x4)
    (skeleton_controller.cc:248-265, gtx/dual_quaternion.inl:303-351; compiled here from the reference sources) on random
    joints: rigid, scaled / sheared (not a rotation: the formula still has to match), near the branch boundaries of the
    rotation extraction (trace ~ 0, equal diagonal entries), and with special values."""
    if not po.ref_available(4):
        pytest.skip("oracle/_ref not built (no reference checkout here)")
    rng = np.random.default_rng(900 + seed)
    J = 256
    G = rng.standard_normal((J, 16)).astype(np.float32); B = rng.standard_normal((J, 16)).astype(np.float32)
    if seed == 0:                                                            # rigid transforms
        for M in (G, B):
            for j in range(J):
                q = rng.standard_normal(4); x, y, z, w = q / np.linalg.norm(q)
                R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w), 0], [2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w), 0],
                              [2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y), 0], [*rng.standard_normal(3), 1]], np.float32)
                M[j] = R.reshape(16)
    if seed == 2:                                                            # branch boundaries: inverse bind = identity, crafted diagonals
        B[:] = np.eye(4, dtype=np.float32).reshape(16)
        G[:, 0], G[:, 5], G[:, 10] = 0.5, -0.25, -0.25                       # trace == 0 exactly -> second branch
        G[64:128, 5] = 0.5                                                   # m00 == m11: neither of the first two diagonal branches
        G[128:192, 10] = 0.5
        G[192:, 0] = -1.0; G[192:, 5] = -1.0; G[192:, 10] = -1.0             # sqrt(1 + m22 - m00 - m11) = sqrt(2)
    if seed == 3:
        G[5, 0] = np.inf; G[6, 3] = np.nan; B[7] = 0.0; G[8] = 0.0           # r = 0 -> 0.5 / 0
    assert_bit_equal(po.dq_palette_from_matrices(G, B), po.ref_dq_palette_from_matrices(G, B), "dual-quaternion palette")
