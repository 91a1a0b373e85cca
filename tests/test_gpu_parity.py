"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the golden
fixtures. Exact profile: BIT-EXACT. Fast profile: <= 1e-5 relative per vertex after one step on
well-conditioned states (BASELINE.json north_star; SURVEY.md App. B for what "well-conditioned" means).
"""
import os
import ctypes as C

import numpy as np
import pytest

import barbu_b200 as bb
from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, golden, ragged_state, rel_err, sphere_state

pytestmark = pytest.mark.gpu

GOLDEN_CASES = ["hair_N2_s100", "hair_N3_s145", "hair_N4_s145", "hair_N8_s145", "hair_N16_s100", "hair_N32_s100",
                "hair_N32_s145", "hair_N64_s100"]


def gpu_steps(pos, vel, S, N, nsteps, dt=DT, substeps=1, **cfg):
    with bb.HairSim(S, N) as sim:
        sim.configure(**cfg)
        sim.upload(pos, vel)
        for _ in range(nsteps):
            sim.step(float(dt), substeps)
        p, v, _ = sim.download()
    return p, v


# ---- against what the reference sources compute (golden fixtures) ---------------------------------

@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_step_bit_exact_vs_reference_golden(name):
    g = golden(name)
    N, S = int(g["nverts"]), g["root_pos"].shape[0]
    cfg = dict(scale=float(g["scale"]), sphere=tuple(g["sphere"]))
    p1, v1 = gpu_steps(g["pos0"], g["vel0"], S, N, 1, g["dt"], **cfg)
    assert_bit_equal(p1, g["pos1"], "pos after 1 update")
    assert_bit_equal(v1, g["vel1"], "vel after 1 update")
    p10, v10 = gpu_steps(g["pos0"], g["vel0"], S, N, 10, g["dt"], **cfg)
    assert_bit_equal(p10, g["pos10"], "pos after 10 updates")
    assert_bit_equal(v10, g["vel10"], "vel after 10 updates")
    pw, vw = gpu_steps(g["pos10"], g["vel10"], S, N, 1, g["dt"], **cfg)
    assert_bit_equal(pw, g["posw1"], "pos, warm state + 1 update")
    assert_bit_equal(vw, g["velw1"], "vel, warm state + 1 update")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_strand_generation_bit_exact_vs_reference_golden(name):
    g = golden(name)
    N, S = int(g["nverts"]), g["root_pos"].shape[0]
    with bb.HairSim(S, N) as sim:
        sim.init_strands(g["root_pos"], g["root_nrm"], bb.random_values(int(g["seed"]), 0, S), 0.5)
        pos, vel, _ = sim.download()
    assert_bit_equal(pos, g["pos0"], "generated positions")
    assert_bit_equal(vel, g["vel0"], "generated velocities")
    assert_bit_equal(bb.build_patch_indices(g["tri"], N), g["patch"], "patch indices")


# ---- against the oracle on seeded inputs -----------------------------------------------------------

@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 48, 64, 100, 128])
@pytest.mark.parametrize("S", [1, 33, 1000])
def test_step_bit_exact_ragged_shapes(S, N):
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.2, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(3):
        po.step(rp, rv, S, N, par)
    gp, gv = gpu_steps(pos, vel, S, N, 3, scale=1.2, sphere=SPHERE)
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


@pytest.mark.parametrize("N", [2, 3, 5, 6, 7, 9, 12, 15, 17, 20, 31, 33, 47, 100, 127])
@pytest.mark.parametrize("S,sphere", [(1000, SPHERE), (4099, (0.1, -0.05, 0.2, 1.05))])
def test_stream_kernel_any_vertex_count_bit_exact(S, N, sphere):
    """Vertices per strand that are not a multiple of 8 stay on the streaming kernel: the last chunk of a strand is ragged,
    the tensor map fills its out-of-bounds slots with NaN and drops them on the way out, and the tip leaves the pipeline at
    step (N % 8) - 1 of the next root chunk. Many steps, contacts from the first one (the second sphere reaches past the
    roots), a ragged last tile."""
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(10):
        po.step(rp, rv, S, N, par, nthreads=16)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=sphere)
        assert sim.kernel_kind == 0
        sim.upload(pos, vel)
        for _ in range(10):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


@pytest.mark.parametrize("N,scale", [(16, 1.45), (32, 1.45), (32, 1.0)])
def test_config1_size_many_steps_bit_exact(N, scale):
    """BASELINE config 1 shape (4,096 strands) over 60 steps from the cold state, contacts included."""
    _, _, _, _, pos, vel = sphere_state(64, 64, N)
    S = 4096
    par = po.default_params(dt=float(DT), scale=scale, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(60):
        po.step(rp, rv, S, N, par, nthreads=8)
    gp, gv = gpu_steps(pos, vel, S, N, 60, scale=scale, sphere=SPHERE)
    assert_bit_equal(gp, rp, "positions after 60 steps")
    assert_bit_equal(gv, rv, "velocities after 60 steps")


@pytest.mark.parametrize("iters", [0, 1, 2, 5, 12])
def test_generic_kernel_other_iteration_counts(iters):
    S, N = 500, 12
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.1, sphere=SPHERE, iterations=iters)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(2):
        po.step(rp, rv, S, N, par)
    gp, gv = gpu_steps(pos, vel, S, N, 2, scale=1.1, sphere=SPHERE, iterations=iters)
    assert_bit_equal(gp, rp)
    assert_bit_equal(gv, rv)


def test_substeps_equal_repeated_steps_with_dt_over_k():
    S, N = 777, 32
    pos, vel = ragged_state(S, N)
    h = np.float32(DT) / np.float32(4)
    par = po.default_params(dt=float(h), scale=1.0, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(8):
        po.step(rp, rv, S, N, par)
    gp, gv = gpu_steps(pos, vel, S, N, 2, DT, substeps=4, scale=1.0, sphere=SPHERE)
    assert_bit_equal(gp, rp)
    assert_bit_equal(gv, rv)


def test_special_values_follow_reference_arithmetic():
    """-0.0 roots, a vertex exactly on its predecessor (normalize(0) -> NaN) and a vertex at the sphere
    centre must come out exactly as the reference arithmetic produces them (App. A note 2: not "fixed")."""
    S, N = 64, 8
    pos, vel = ragged_state(S, N)
    pos[0 * N, :3] = (-0.0, 1.0, -0.0)
    pos[1 * N + 3, :3] = pos[1 * N + 2, :3]
    vel[1 * N + 3, :3] = vel[1 * N + 2, :3]
    pos[2 * N + 5, :3] = (0.0, 0.0, 0.0)
    par = po.default_params(dt=float(DT), scale=1.0, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    po.step(rp, rv, S, N, par)
    gp, gv = gpu_steps(pos, vel, S, N, 1, scale=1.0, sphere=SPHERE)
    assert_bit_equal(gp, rp)
    assert_bit_equal(gv, rv)


# ---- extensions (no reference implementation; parity is against the oracle's definition) -----------

def test_extensions_wind_drag_capsules_bit_exact():
    S, N = 600, 16
    pos, vel = ragged_state(S, N)
    caps = [((0.3, 0.2, 0.1), (-0.4, 0.5, 0.0), 0.35), ((0.0, -0.6, 0.2), (0.0, -0.6, 0.2), 0.5)]
    par = po.default_params(dt=float(DT), scale=1.0, sphere=SPHERE, wind=(3.0, 0.5, -2.0), drag=0.05, ncapsules=len(caps))
    gcfg = bb.default_params()
    gcfg.scale, gcfg.drag, gcfg.ncapsules = 1.0, 0.05, len(caps)
    for i, x in enumerate(SPHERE):
        gcfg.sphere[i] = x
    for i, x in enumerate((3.0, 0.5, -2.0)):
        gcfg.wind[i] = x
    for q, (a, b, r) in enumerate(caps):
        for i in range(3):
            par.capsules[q].a[i], par.capsules[q].b[i] = a[i], b[i]
            gcfg.capsules[q].a[i], gcfg.capsules[q].b[i] = a[i], b[i]
        par.capsules[q].radius = gcfg.capsules[q].radius = r
    rp, rv = pos.copy(), vel.copy()
    for _ in range(5):
        po.step(rp, rv, S, N, par)
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg)
        sim.upload(pos, vel)
        for _ in range(5):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp)
    assert_bit_equal(gv, rv)


def test_dq_skinned_roots_bit_exact():
    rng = np.random.default_rng(5)
    S, N, J = 900, 8, 6
    root_pos, root_nrm, _, rv, pos, vel = sphere_state(30, 30, N)
    q = rng.standard_normal((J, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    dual = (rng.standard_normal((J, 4)) * 0.1).astype(np.float32)
    dq = np.concatenate([q, dual], axis=1).astype(np.float32)
    joints = rng.integers(0, J, (S, 4)).astype(np.int32)
    w = rng.dirichlet(np.ones(4), S).astype(np.float32)[:, :3]
    w[::7, 0] = 0.0                                   # exercises the weights.x <= eps early-out
    exp_pos, _ = po.skin_roots_dq(root_pos, root_nrm, joints, w, dq)
    with bb.HairSim(S, N) as sim:
        sim.upload(pos, vel)
        sim.set_skin(root_pos, joints, w)
        sim.skin_roots(dq)
        gp, _, _ = sim.download()
    expect = pos.copy()
    expect[::N, :3] = exp_pos
    assert_bit_equal(gp, expect, "skinned roots")


def test_dq_skinned_roots_bit_exact_vs_reference_fixture():
    """skin_roots_dq_kernel against tests/golden/tess_skin.npz: positions the REFERENCE's apply_skinning + skinning_DQBS
    (inc_skinning.glsl over GLM, tests/golden/make_tess_skin_golden.py) produced — antipodal joints, Epsilon() early-out."""
    t = golden("tess_skin")
    S, N = t["skin_pos"].shape[0], 4
    pos = np.zeros((S * N, 4), np.float32); pos[:, 3] = 0.125
    pos[::N, :3] = t["skin_pos"]
    with bb.HairSim(S, N) as sim:
        sim.upload(pos, np.zeros_like(pos))
        sim.set_skin(t["skin_pos"], t["skin_joints"], t["skin_weights"])
        sim.skin_roots(t["skin_dq"])
        gp, _, _ = sim.download()
    expect = pos.copy()
    expect[::N, :3] = t["skin_out_pos"]
    assert_bit_equal(gp, expect, "skinned roots vs the reference shader")


def test_skin_roots_rejects_joint_indices_outside_the_palette():
    """A joint index the palette does not hold must be refused (BH_ERR_INVALID), not read out of bounds on the device."""
    S, N = 64, 4
    pos = np.zeros((S * N, 4), np.float32)
    root = np.zeros((S, 3), np.float32)
    w = np.full((S, 3), 0.25, np.float32)
    dq = np.zeros((6, 8), np.float32); dq[:, 3] = 1.0
    with bb.HairSim(S, N) as sim:
        sim.upload(pos, pos.copy())
        j = np.zeros((S, 4), np.int32); j[17, 2] = 6
        sim.set_skin(root, j, w)
        with pytest.raises(bb.BarbuHairError):
            sim.skin_roots(dq)                                             # 6 joints: index 6 is out of range
        sim.skin_roots(np.concatenate([dq, dq[:1]]))                       # 7 joints: fine
        j[17, 2] = -1
        with pytest.raises(bb.BarbuHairError):
            sim.set_skin(root, j, w)                                       # negative index
        sim.step(float(DT), 1)                                             # the context is still healthy


# ---- fast profile: tolerance ----------------------------------------------------------------------

@pytest.mark.parametrize("N", [16, 32])
def test_fast_profile_within_1e5_after_one_step(N):
    """Tolerance of BASELINE.json north_star: <= 1e-5 relative per vertex after one step.
    States: cold with scale 1.0, and warm (after 60 exact steps) with the reference's 1.45."""
    TOL = 1e-5
    S = 4096
    _, _, _, _, pos, vel = sphere_state(64, 64, N)
    par = po.default_params(dt=float(DT), scale=1.0, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    po.step(rp, rv, S, N, par, nthreads=8)
    gp, gv = gpu_steps(pos, vel, S, N, 1, scale=1.0, sphere=SPHERE, math=bb.BH_MATH_FAST)
    assert rel_err(gp, rp).max() <= TOL
    # warm state
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    wp, wv = pos.copy(), vel.copy()
    for _ in range(60):
        po.step(wp, wv, S, N, par, nthreads=8)
    rp, rv = wp.copy(), wv.copy()
    po.step(rp, rv, S, N, par, nthreads=8)
    gp, gv = gpu_steps(wp, wv, S, N, 1, scale=1.45, sphere=SPHERE, math=bb.BH_MATH_FAST)
    err = rel_err(gp, rp)
    # the MAXIMUM over all vertices, not a quantile. The only vertices that may be set aside are those of a strand in
    # grazing contact — some vertex at or before them lies within 4 ulp of the collider surface in the exact result, where
    # a different rounding legitimately flips `dp < r*r` — and they are listed, and there may be at most two such strands.
    excluded = contact_bifurcations(err > TOL, rp, S, N, SPHERE)
    assert err[~excluded].max() <= TOL and excluded.reshape(S, N).any(axis=1).sum() <= 2, \
        f"max {err.max():.3e}; set aside (grazing contact): {np.nonzero(excluded)[0].tolist()}"
    # velocities are cancellation differences: absolute tolerance scaled by segment length (App. B iv)
    seg = 1.45 * np.maximum(rp[:, 3], 1e-3)
    verr = np.abs(gv[:, :3] - rv[:, :3]).max(axis=1) / seg
    assert verr[~excluded].max() <= 1e-4, f"velocity max {verr.max():.3e}"


def contact_bifurcations(over, exact_pos, S, N, sphere, ulps=4):
    """Mask of the vertices flagged in `over` that sit on a strand with a grazing contact at or before them: a vertex of the
    exact result within `ulps` ulp of the sphere surface. Flagged vertices without such a contact stay un-excused."""
    c, r = np.asarray(sphere[:3], np.float64), float(sphere[3])
    dist = np.linalg.norm(exact_pos[:, :3].astype(np.float64) - c, axis=1).reshape(S, N)
    grazing = np.abs(dist - r) <= ulps * float(np.spacing(np.float32(r)))
    grazing[:, 0] = False                                                 # roots do not collide (cs:149-151)
    upstream = np.cumsum(grazing, axis=1) > 0
    return (over.reshape(S, N) & upstream).reshape(-1)


def test_full_size_fast_profile_one_step_max_error():
    """north_star's tolerance at configs[1]'s full size (2^20 strands x 32): from a settled state (30 frames of the exact
    profile, which the tests above hold bit-identical to the oracle), ONE step of the fast profile against ONE step of the
    exact profile over ALL 33,554,432 vertices: max relative error <= 1e-5 (grazing-contact strands listed, at most 0.002 %),
    and 2,048 sampled strands of the exact side replayed by the CPU oracle, bit for bit."""
    rows, cols, N = 1024, 1024, 32
    S = rows * cols
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE, math=bb.BH_MATH_EXACT)
        sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
        for _ in range(30):
            sim.step(float(DT), 1)
        wp, wv, _ = sim.download()
        sim.step(float(DT), 1)
        ep, ev, _ = sim.download()
        sim.upload(wp, wv)
        sim.configure(math=bb.BH_MATH_FAST)
        sim.step(float(DT), 1)
        fp, fv, _ = sim.download()
    finite = np.isfinite(ep[:, :3]).all(axis=1) & np.isfinite(wp[:, :3]).all(axis=1)   # strands the reference arithmetic itself turned into NaN (normalize(0), App. A note 2)
    assert (~finite).reshape(S, N).any(axis=1).sum() <= S * 1e-4
    assert np.array_equal(np.isfinite(fp[:, :3]).all(axis=1), np.isfinite(ep[:, :3]).all(axis=1)), "fast and exact disagree on which vertices are finite"
    err = np.where(finite, rel_err(fp, ep), 0.0)
    excluded = contact_bifurcations(err > 1e-5, ep, S, N, SPHERE)
    nstr = int(excluded.reshape(S, N).any(axis=1).sum())
    assert err[~excluded].max() <= 1e-5 and nstr <= 20, f"max {err.max():.3e}; {nstr} grazing-contact strands set aside: {np.nonzero(excluded.reshape(S, N).any(axis=1))[0][:40].tolist()}"
    assert np.median(err) <= 1e-7 and np.quantile(err, 0.999) <= 2e-6, f"median {np.median(err):.3e} p99.9 {np.quantile(err, 0.999):.3e}"
    idx = np.unique(np.concatenate([np.arange(0, S, 521), np.nonzero(excluded.reshape(S, N).any(axis=1))[0][:64]]))[:2048 + 64]
    p, v = _sample_strands(wp, wv, S, N, idx)
    po.step(p, v, idx.size, N, po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE), nthreads=16)
    gp, gv = _sample_strands(ep, ev, S, N, idx)
    assert_bit_equal(gp, p, "exact profile, sampled strands vs the oracle")
    assert_bit_equal(gv, v, "exact profile, sampled velocities vs the oracle")


# ---- generators, host path, API behaviour ---------------------------------------------------------

@pytest.mark.parametrize("rows,cols,N", [(8, 16, 4), (64, 64, 16), (33, 17, 5)])
def test_sphere_scalp_generation_bit_exact_and_sharded(rows, cols, N):
    S = rows * cols
    _, _, tri, rv, pos, vel = sphere_state(rows, cols, N)
    assert_bit_equal(bb.random_values(1234, 0, S), rv, "rand() jitter")
    assert_bit_equal(bb.sphere_scalp_triangles(rows, cols), tri, "scalp triangles")
    with bb.HairSim(S, N) as sim:
        sim.init_sphere_scalp(rows, cols, 0, rv)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, pos)
    assert_bit_equal(gv, vel)
    # two shards, as two ranks would hold them
    first = S // 2 + 3
    for lo, hi in ((0, first), (first, S)):
        with bb.HairSim(hi - lo, N) as sim:
            sim.init_sphere_scalp(rows, cols, lo, bb.random_values(1234, lo, hi - lo))
            gp, _, _ = sim.download()
        assert_bit_equal(gp, pos[lo * N:hi * N], f"shard [{lo},{hi})")
    assert_bit_equal(bb.build_patch_indices(tri, N), po.patch_indices(tri, N), "patch indices")


@pytest.mark.parametrize("S,N", [(20000, 16), (70001, 20), (33333, 4)])     # several slices; ragged chunks; odd strand count at N = 4
def test_step_host_equals_device_resident_step(S, N):
    pos, vel = ragged_state(S, N)
    gp, gv = gpu_steps(pos, vel, S, N, 1, DT, substeps=4, scale=1.0, sphere=SPHERE)
    hp, hv = bb.PinnedBuffer(4 * S * N), bb.PinnedBuffer(4 * S * N)
    hp.array[:] = pos.reshape(-1)
    hv.array[:] = vel.reshape(-1)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.0, sphere=SPHERE)
        sim.step_host(float(DT), 4, hp.array, hv.array)
        assert sim.launch_count >= 4
    assert_bit_equal(hp.array.reshape(-1, 4), gp)
    assert_bit_equal(hv.array.reshape(-1, 4), gv)
    hp.free(); hv.free()


def test_api_error_behaviour():
    lib = bb.load_library()
    with bb.HairSim(10, 4) as sim:
        with pytest.raises(bb.BarbuHairError) as e:
            sim.step(0.01, 1)                       # no state yet: Hair::update before setup
        assert e.value.code == 3
        with pytest.raises(bb.BarbuHairError):
            sim.configure(iterations=-1)
        with pytest.raises(bb.BarbuHairError):
            sim.configure(math=7)
        with pytest.raises(bb.BarbuHairError):
            sim.register_gl_buffer(1)               # no GL context in this process
        assert b"GL" in lib.bh_last_error() or len(lib.bh_last_error()) > 0
    h = C.c_void_p()
    assert lib.bh_create(C.byref(h), 0, 4, 0) == 1
    assert lib.bh_create(C.byref(h), 4, 4, 99) == 1
    with pytest.raises(bb.BarbuHairError) as e:
        bb.build_patch_indices(np.array([[0, 1, 2 ** 30]], np.int32), 4)
    assert e.value.code == 5


def test_hair_module_mirror():
    """The Hair-shaped adaptor: setup / set_bounding_sphere / update as Renderer::update drives them
    (core/renderer.cc:69-81), against the oracle."""
    rows, cols, N = 16, 28, 4                        # 448 strands x 4 CPs: the reference's default workload size
    root_pos, root_nrm, tri, rv, pos, vel = sphere_state(rows, cols, N)
    hair = bb.Hair(params=bb.Hair.Parameters(ncontrol_points=N))
    hair.init()
    hair.update(float(DT))
    assert not hair.initialized() and "without initialization" in hair.log[-1]
    hair.setup(None)
    assert not hair.initialized() and "not found" in hair.log[-1]
    hair.set_bounding_sphere(SPHERE)
    hair.setup(bb.ScalpMesh(root_pos, root_nrm, tri))
    assert hair.initialized() and hair.nroots == 448
    gp, gv, gt = hair.sim.download(tan=True)
    assert_bit_equal(gp, pos)
    assert_bit_equal(gt, po.init_tangents(root_nrm, N), "tangent plane")
    assert_bit_equal(hair.patch_indices, po.patch_indices(tri, N))
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(4):
        hair.update(float(DT))
        po.step(pos, vel, 448, N, par)
    gp, gv, gt2 = hair.sim.download(tan=True)
    assert_bit_equal(gp, pos)
    assert_bit_equal(gv, vel)
    assert_bit_equal(gt2, gt, "tangent plane untouched by the simulation")
    # the tess-stream half of render(), default tessellation (hair.h:33-35)
    assert_bit_equal(hair.stream(), po.tess_stream(pos, gt, hair.patch_indices, N, 1.45, 3, 2, 16, 1234), "tess-stream")
    # a second setup starts over: new sim, patches uploaded again, same results as a fresh module
    hair.setup(bb.ScalpMesh(root_pos, root_nrm, tri))
    assert hair.initialized() and hair.nroots == 448
    gp2, _, gt3 = hair.sim.download(tan=True)
    pos0 = sphere_state(rows, cols, N)[4]
    assert_bit_equal(gp2, pos0, "state after a second setup")
    assert_bit_equal(hair.stream(), po.tess_stream(pos0, gt3, hair.patch_indices, N, 1.45, 3, 2, 16, 1234), "tess-stream after a second setup")
    hair.deinit()
    assert not hair.initialized() and hair.stream() is None


def test_switching_streams_keeps_steps_in_order():
    """bh_set_stream / bh_reset_stream order the new stream after the work queued on the old one: steps issued back to back
    on alternating streams (they share one tile scheduler and update the state in place) equal the same steps on one stream."""
    import torch
    S, N = 1 << 16, 32
    pos, vel = ragged_state(S, N)
    streams = [torch.cuda.Stream() for _ in range(3)]
    with bb.HairSim(S, N) as a, bb.HairSim(S, N) as b:
        for sim in (a, b):
            sim.configure(scale=1.45, sphere=SPHERE)
            sim.upload(pos, vel)
        for k in range(12):
            a.step(float(DT), 2)
        for k in range(12):
            if k % 4 == 3: b.reset_stream()
            else: b.set_stream(streams[k % 3].cuda_stream)
            b.step(float(DT), 2)                                             # no host synchronisation in between
        wp, wv, _ = a.download()
        gp, gv, _ = b.download()
    assert_bit_equal(gp, wp, "positions"); assert_bit_equal(gv, wv, "velocities")


# ---- streaming kernel (TMA tiles, persistent warps, packed fp32x2) ---------------------------------

def test_exact_inversesqrt_exhaustive():
    """Branch-free 1/sqrt(x) of the exact profile == __frcp_rn(__fsqrt_rn(x)) for every finite float >= 2^-102."""
    assert bb.selftest_math(0) == 0


@pytest.mark.parametrize("N", [8, 16, 24, 32, 64, 128])
@pytest.mark.parametrize("S", [1, 31, 32, 33, 1000, 20000])
def test_stream_kernel_bit_exact(S, N):
    """Shapes that select the streaming kernel: ragged last tile, one chunk per strand (N=8), many tiles per warp."""
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.2, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(2):
        po.step(rp, rv, S, N, par)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.2, sphere=SPHERE)
        assert sim.kernel_kind == 0
        sim.upload(pos, vel)
        for _ in range(2):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


@pytest.mark.parametrize("S", [2, 3, 63, 64, 65, 1000, 20001, 131072])
@pytest.mark.parametrize("sphere", [(0.0, 0.0, 0.0, 1.05), (0.1, -0.05, 0.2, 1.0)])   # reaches past the roots: push-outs from step 1
def test_stream_kernel_reference_nverts_4_bit_exact(S, sphere):
    """The reference's own N = 4 (interop.h:8): two strands per 128-byte tensor row, roots in slots 0 and 4 of every
    chunk; an odd strand count leaves the last strand to the per-strand kernel."""
    N = 4
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(12):
        po.step(rp, rv, S, N, par, nthreads=16)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=sphere)
        assert sim.kernel_kind == 0
        sim.upload(pos, vel)
        for _ in range(12):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")
    touching = np.linalg.norm(rp[:, :3] - np.float32(sphere[:3]), axis=1) < sphere[3] * (1 + 1e-6)
    if S >= 1000:
        assert touching.reshape(S, N)[:, 1:].any(), "the case must exercise the push-out"


def test_stream_kernel_off_origin_sphere_bit_exact():
    S, N = 3000, 32
    sphere = (0.1, -0.05, 0.2, 0.9)
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.3, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(3):
        po.step(rp, rv, S, N, par)
    gp, gv = gpu_steps(pos, vel, S, N, 3, scale=1.3, sphere=sphere)
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


def test_stream_and_per_strand_kernels_agree(monkeypatch):
    """Same state through bh_step (streaming kernel) and through a 9-iteration-free path: compare with the oracle at a
    size where every resident warp owns several tiles (persistent scheduler, pipeline carried across tiles)."""
    rows, cols, N = 256, 512, 16                      # 131,072 strands = 4,096 tiles > 148 SMs x 12 warps
    S = rows * cols
    _, _, _, _, pos, vel = sphere_state(rows, cols, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(4):
        po.step(rp, rv, S, N, par, nthreads=16)
    gp, gv = gpu_steps(pos, vel, S, N, 4, scale=1.45, sphere=SPHERE)
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


def _capsule_params(caps, **kw):
    """(oracle params, device params) carrying the same capsules."""
    par = po.default_params(ncapsules=len(caps), **kw)
    gcfg = bb.default_params()
    gcfg.scale, gcfg.ncapsules = kw["scale"], len(caps)
    for i, x in enumerate(kw["sphere"]):
        gcfg.sphere[i] = x
    for q, (a, b, r) in enumerate(caps):
        for i in range(3):
            par.capsules[q].a[i], par.capsules[q].b[i] = a[i], b[i]
            gcfg.capsules[q].a[i], gcfg.capsules[q].b[i] = a[i], b[i]
        par.capsules[q].radius = gcfg.capsules[q].radius = r
    return par, gcfg


CAPSULE_SETS = {
    # two "arms" leaving the scalp: roots inside the capsules, long contact along the axis
    "arms": [((0.55, 0.25, 0.0), (1.25, -0.35, 0.0), 0.30), ((-0.55, 0.25, 0.0), (-1.25, -0.35, 0.0), 0.30)],
    # overlapping capsules + a degenerate one (a == b: a sphere) + one far away that is never touched
    "overlap": [((0.0, 1.0, 0.0), (0.3, 1.4, 0.2), 0.25), ((0.1, 1.1, 0.1), (0.1, 1.1, 0.1), 0.3),
                ((0.2, 1.2, 0.0), (-0.4, 1.3, 0.1), 0.2), ((50.0, 0.0, 0.0), (51.0, 0.0, 0.0), 0.5)],
    # parameters that defeat the conservative bounds (they must then step aside, not decide): an axis whose squared length
    # overflows, a capsule far from the origin, a zero radius, a NaN end point
    "extreme": [((-1.0e20, 0.3, 0.0), (1.0e20, 0.3, 0.0), 0.4), ((1.0e6, 0.0, 0.0), (1.0e6, 1.0, 0.0), 0.5),
                ((0.2, 0.9, 0.1), (0.4, 1.2, 0.1), 0.0), ((float("nan"), 0.0, 0.0), (0.0, 1.0, 0.0), 0.3)],
    # the shell test in front of the per-capsule bounds (squared distance to the sphere's centre, already known, against the
    # range of distances at which capsules exist): one capsule well beyond the sphere where the strands hang (lower bound
    # active), one entirely inside the sphere (never reached once the sphere has pushed a vertex out), one across its surface
    "shell": [((-0.5, -1.45, 0.0), (0.5, -1.45, 0.1), 0.15), ((0.1, 0.2, 0.0), (0.3, 0.3, 0.1), 0.25),
              ((0.0, 0.9, 0.0), (0.05, 1.3, 0.0), 0.2)],
    # eight capsules (the maximum) in a ring below the scalp, where the strands hang
    "ring8": [((float(np.cos(k * np.pi / 4)), -1.1, float(np.sin(k * np.pi / 4))),
               (float(np.cos((k + 1) * np.pi / 4)), -1.25, float(np.sin((k + 1) * np.pi / 4))), 0.12) for k in range(8)],
}


@pytest.mark.parametrize("caps", sorted(CAPSULE_SETS))
@pytest.mark.parametrize("S,N,sphere", [(600, 16, SPHERE), (4100, 32, SPHERE), (1001, 8, (0.05, 0.1, -0.02, 0.9)),
                                        (2050, 4, (0.0, 0.0, 0.0, 1.02)), (333, 64, SPHERE),
                                        (1500, 12, SPHERE), (777, 21, (0.05, 0.1, -0.02, 0.9)), (900, 5, SPHERE)])   # ragged last chunk
def test_stream_kernel_capsules_bit_exact(caps, S, N, sphere):
    """Capsule colliders (extension, oracle-defined) through the streaming kernel: conservative bounding test, exact
    capsule chain for the warps that may touch, collision chain of the leaving vertex recomputed with its velocity."""
    capsules = CAPSULE_SETS[caps]
    pos, vel = ragged_state(S, N)
    par, gcfg = _capsule_params(capsules, dt=float(DT), scale=1.45, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    nsteps = 14
    for _ in range(nsteps):
        po.step(rp, rv, S, N, par, nthreads=16)
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg)
        assert sim.kernel_kind == 0, "capsules must not leave the streaming kernel"
        sim.upload(pos, vel)
        for _ in range(nsteps):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")
    if caps not in ("ring8", "extreme") or (caps == "ring8" and N >= 16):
        # the case must exercise the capsule push-out: some non-root vertex sits on a capsule surface
        x = rp[:, :3].astype(np.float64).reshape(S, N, 3)[:, 1:].reshape(-1, 3)
        on = np.zeros(len(x), bool)
        for a, b, r in capsules:
            a, b = np.array(a), np.array(b)
            ab = b - a
            t = np.clip((x - a) @ ab / max(ab @ ab, 1e-30), 0, 1) if ab @ ab > 0 else np.zeros(len(x))
            d = np.linalg.norm(x - (a + t[:, None] * ab), axis=1)
            on |= np.abs(d - r) < 1e-5
        assert on.any(), "no vertex rests on a capsule: the test does not cover the push-out"


@pytest.mark.parametrize("S,N", [(2100, 32), (1300, 24), (3000, 4), (999, 13)])
def test_stream_kernel_capsules_temporal_bound_bit_exact(S, N):
    """The temporal bound in front of the capsule tests (a lane re-tests only when its last measured distance has been used
    up by the rest lengths stepped since): strands blown by wind into capsules that only their outer vertices reach, with
    rest lengths that vary wildly along a strand and from strand to strand (a few segments ten times the others, some
    almost zero), over enough steps for whole strands to arrive from far away. Any skipped test that mattered shows as a
    vertex left inside a capsule, i.e. as a mismatch."""
    pos, vel = ragged_state(S, N, seed=7)
    rng = np.random.default_rng(11)
    f = rng.choice(np.array([0.02, 0.5, 1.0, 1.0, 1.0, 2.0, 10.0], np.float32), size=S * N)
    pos[:, 3] = (pos[:, 3] * f).astype(np.float32)
    caps = [((-0.4, -1.9, 0.3), (0.9, -2.3, -0.2), 0.35), ((1.4, -1.0, 0.2), (1.6, -0.1, 0.5), 0.3), ((1.6, -1.9, 0.9), (1.6, -1.9, 0.9), 0.45)]
    wind = (2.5, -1.0, 0.7)
    par, gcfg = _capsule_params(caps, dt=float(DT), scale=1.45, sphere=SPHERE)
    for i, x in enumerate(wind):
        par.wind[i] = x
        gcfg.wind[i] = x
    rp, rv = pos.copy(), vel.copy()
    nsteps = 60
    for _ in range(nsteps):
        po.step(rp, rv, S, N, par, nthreads=16)
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg)
        assert sim.kernel_kind == 0
        sim.upload(pos, vel)
        for _ in range(nsteps):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")
    x = rp[:, :3].astype(np.float64).reshape(S, N, 3)[:, 1:].reshape(-1, 3)
    on = np.zeros(len(x), bool)
    for a, b, r in caps:
        a, b = np.array(a), np.array(b)
        ab = b - a
        t = np.clip((x - a) @ ab / max(ab @ ab, 1e-30), 0, 1) if ab @ ab > 0 else np.zeros(len(x))
        on |= np.abs(np.linalg.norm(x - (a + t[:, None] * ab), axis=1) - r) < 1e-5
    assert on.any(), "no vertex rests on a capsule: the test does not cover the push-out"


def _random_capsule_scene(seed, capsules=True):
    """A random scene for the capsule variant: 1-8 capsules of random size scattered through the volume the hair sweeps (some
    degenerate, some huge, some tiny), random wind and drag, random sphere, rest lengths scaled per vertex, random shape."""
    rng = np.random.default_rng(seed)
    S = int(rng.integers(200, 1500)); N = int(rng.choice([4, 5, 8, 13, 16, 24, 32, 40]) if capsules else rng.integers(1, 71))
    pos, vel = ragged_state(S, N, seed=seed)
    f = rng.choice(np.array([0.05, 0.5, 1.0, 1.0, 1.0, 1.5, 4.0], np.float32), size=S * N)
    pos[:, 3] = (pos[:, 3] * f).astype(np.float32)
    ncaps = int(rng.integers(1, 9)) if capsules else 0
    caps = []
    for _ in range(ncaps):
        a = rng.uniform(-2.2, 2.2, 3)
        kind = rng.integers(0, 5)
        b = a if kind == 0 else a + rng.uniform(-1.0, 1.0, 3) * (3.0 if kind == 1 else 0.8)
        r = float(rng.choice([0.02, 0.1, 0.25, 0.4, 0.9]))
        caps.append((tuple(float(x) for x in a), tuple(float(x) for x in b), r))
    sphere = (float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1)), float(rng.choice([0.0, 0.9, 0.98, 1.05])))
    if rng.integers(0, 3) == 0:
        sphere = (0.0, 0.0, 0.0, sphere[3])                                   # the origin-centred variant of the kernel
    wind = tuple(float(x) for x in rng.uniform(-3.0, 3.0, 3))
    drag = float(rng.choice([0.0, 0.0, 0.03]))
    scale = float(rng.choice([1.0, 1.45, 2.0]))
    return S, N, pos, vel, caps, sphere, wind, drag, scale


def run_capsule_scene(seed, nsteps=40, fused_substeps=0, capsules=True):
    """Steps a random capsule scene on the device and on the oracle; returns the number of differing words (0 = bit-equal).
    fused_substeps = k > 0: the device runs frames of k substeps as ONE launch each (fusion forced), nsteps // k frames."""
    S, N, pos, vel, caps, sphere, wind, drag, scale = _random_capsule_scene(seed, capsules)
    k = max(fused_substeps, 1)
    h = float(np.float32(DT) / np.float32(k)) if k > 1 else float(DT)
    par, gcfg = _capsule_params(caps, dt=h, scale=scale, sphere=sphere)
    par.drag = gcfg.drag = drag
    for i, x in enumerate(wind):
        par.wind[i] = x
        gcfg.wind[i] = x
    rp, rv = pos.copy(), vel.copy()
    for _ in range(nsteps // k * k):
        po.step(rp, rv, S, N, par, nthreads=16)
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg)
        if k > 1:
            sim.set_substep_fusion(True, always=True)
        sim.upload(pos, vel)
        kind = sim.kernel_kind
        for _ in range(nsteps // k):
            sim.step(float(DT), k)
        gp, gv, _ = sim.download()
    # the repo's rule for "bit-equal" (tests/util.py::assert_bit_equal): every word identical, except that a NaN may meet a NaN
    # of another payload (x86 makes 0xFFC00000, sm_100 0x7FFFFFFF) — random scenes do produce NaN strands (normalize(0))
    diff = int(((gp.view(np.uint32) != rp.view(np.uint32)) & ~(np.isnan(gp) & np.isnan(rp))).sum()
               + ((gv.view(np.uint32) != rv.view(np.uint32)) & ~(np.isnan(gv) & np.isnan(rv))).sum())
    return diff, kind, (S, N, len(caps))


@pytest.mark.parametrize("seed", range(3000, 3016))
def test_random_scenes_without_capsules_bit_exact(seed):
    """The same fuzz without capsules (the reference's own collider set: one sphere), 1-70 vertices per strand, whatever
    kernel the shape selects; odd seeds run frames of 3 substeps (fused where the shape allows it)."""
    diff, kind, shape = run_capsule_scene(seed, nsteps=39, fused_substeps=3 if seed % 2 else 0, capsules=False)
    assert diff == 0, f"seed {seed} {shape}: kernel kind {kind}, {diff} words differ"


@pytest.mark.parametrize("seed", range(2000, 2008))
def test_stream_kernel_capsules_random_scenes_fused_bit_exact(seed):
    """The same fuzz with frames of 4 substeps as ONE launch each (the temporal bound across the passes of a fused launch)."""
    diff, kind, shape = run_capsule_scene(seed, nsteps=40, fused_substeps=4)
    assert kind == 0 and diff == 0, f"seed {seed} {shape}: kernel kind {kind}, {diff} words differ"


@pytest.mark.parametrize("seed", range(1000, 1024))
def test_stream_kernel_capsules_random_scenes_bit_exact(seed):
    """Fuzz of the capsule variant's conservative machinery (shell, temporal bound, capsule-shaped bound, packed chain): random
    scenes, 40 steps each, bit-equal to the oracle. tests/reports/capsule_fuzz.py runs the same generator over more seeds."""
    diff, kind, shape = run_capsule_scene(seed)
    assert kind == 0, "capsules must not leave the streaming kernel"
    assert diff == 0, f"seed {seed} {shape}: {diff} words differ"


@pytest.mark.parametrize("sphere", [(0.0, 0.0, 0.0, 5.0), (0.0, 0.0, 0.0, 0.0), (0.3, 2.0, 0.1, 1.2), (0.0, 0.0, 0.0, 1.0e-3), (1.0e3, 0.0, 0.0, 1.0)])
@pytest.mark.parametrize("caps", ["arms", "shell", "overlap"])
def test_stream_kernel_capsules_any_sphere_bit_exact(caps, sphere):
    """The capsule bounds measure distances from the SPHERE's centre: spheres that swallow every capsule and every strand
    (all vertices end on its surface), of radius zero, off to the side, tiny, or a thousand units away stay bit-exact."""
    S, N = 1300, 16
    pos, vel = ragged_state(S, N)
    par, gcfg = _capsule_params(CAPSULE_SETS[caps], dt=float(DT), scale=1.45, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(8):
        po.step(rp, rv, S, N, par, nthreads=16)
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg)
        assert sim.kernel_kind == 0
        sim.upload(pos, vel)
        for _ in range(8):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


def test_stream_kernel_capsules_fast_profile_close_to_exact():
    """Fast profile with capsules: same contacts, positions within the one-step tolerance of the exact profile."""
    S, N = 2048, 32
    pos, vel = ragged_state(S, N)
    par, gcfg = _capsule_params(CAPSULE_SETS["arms"], dt=float(DT), scale=1.45, sphere=SPHERE)
    out = {}
    for math in (bb.BH_MATH_EXACT, bb.BH_MATH_FAST):
        gcfg.math = math
        with bb.HairSim(S, N) as sim:
            sim.set_params(gcfg)
            sim.upload(pos, vel)
            for _ in range(30):                       # settle on the exact profile, then one step in the profile under test
                gcfg.math = bb.BH_MATH_EXACT; sim.set_params(gcfg); sim.step(float(DT), 1)
            gcfg.math = math; sim.set_params(gcfg)
            sim.step(float(DT), 1)
            out[math], _, _ = sim.download()
    err = rel_err(out[bb.BH_MATH_FAST], out[bb.BH_MATH_EXACT])
    assert np.percentile(err, 99.9) <= 1e-5, f"p99.9 {np.percentile(err, 99.9):.3e}"
    # ... and the MAXIMUM over all vertices, no vertex set aside (measured: 1.1e-6)
    assert err.max() <= 1e-5, f"max {err.max():.3e}"


# ---- BASELINE.json sizes: size-independent properties + sampled strands against the oracle -------------

def _sample_strands(pos, vel, S, N, idx):
    p = pos.reshape(S, N, 4)[idx].reshape(-1, 4).copy()
    v = vel.reshape(S, N, 4)[idx].reshape(-1, 4).copy()
    return p, v


@pytest.mark.parametrize("rows,cols,N,math", [(1024, 1024, 32, "exact"), (1024, 1024, 32, "fast"), (2048, 1024, 16, "exact"),
                                              (512, 512, 128, "exact")])
def test_full_size_frame_properties_and_sampled_parity(rows, cols, N, math):
    """configs[1] (2^20 strands x 32, 4 substeps/frame) and two sweep shapes of configs[4]: after two frames
    I1 roots pinned, I3 nothing inside the collider, I4 w planes preserved, everything finite; and because strands are
    independent, 8,192 sampled strands must equal the oracle stepping just those strands — bit-exact (exact profile) or
    within the stated tolerances (fast profile)."""
    S = rows * cols
    substeps, frames = 4, 2
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE, math=bb.BH_MATH_FAST if math == "fast" else bb.BH_MATH_EXACT)
        assert sim.kernel_kind == 0
        sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
        pos0, vel0, _ = sim.download()
        for _ in range(frames):
            sim.step(float(DT), substeps)
        pos1, vel1, _ = sim.download()
    # The cold straight state stretched by 1.45 in one pass is ill-conditioned: a handful of strands
    # fold onto themselves and the REFERENCE arithmetic yields normalize(0) = NaN for them (App. A note 2). They are
    # few, and they are added to the sampled set below, where the oracle must produce the same NaNs.
    bad = np.nonzero(~(np.isfinite(pos1).reshape(S, -1).all(axis=1) & np.isfinite(vel1).reshape(S, -1).all(axis=1)))[0]
    assert bad.size <= S * 1e-4, f"{bad.size} non-finite strands"
    assert_bit_equal(pos1[::N], pos0[::N], "I1 roots")
    assert not vel1[::N].any(), "I1 root velocities"
    assert_bit_equal(pos1[:, 3], pos0[:, 3], "I4 rest lengths")
    assert not vel1[:, 3].any(), "I4 velocity w"
    r = np.linalg.norm(pos1[:, :3].astype(np.float64), axis=1).reshape(S, N)[:, 1:]
    assert np.nanmin(r) >= SPHERE[3] * (1 - 1e-6), "I3 vertex inside the sphere"
    # sampled strands: whole tiles spread over the launch (first, last, and every ~127th tile) + a ragged sprinkle
    tiles = np.unique(np.concatenate([[0, S // 32 - 1], np.arange(0, S // 32, 127)]))[:240]
    idx = np.unique(np.concatenate([(tiles[:, None] * 32 + np.arange(32)).ravel(), np.arange(5, S, 100003), bad[:64]]))
    p, v = _sample_strands(pos0, vel0, S, N, idx)
    h = float(np.float32(DT) / np.float32(substeps))
    par = po.default_params(dt=h, scale=1.45, sphere=SPHERE)
    for _ in range(frames * substeps):
        po.step(p, v, idx.size, N, par, nthreads=16)
    gp, gv = _sample_strands(pos1, vel1, S, N, idx)
    if math == "exact":
        assert_bit_equal(gp, p, "sampled positions")
        assert_bit_equal(gv, v, "sampled velocities")
    else:
        # 8 dependent steps of a contact problem with different rounding: report-style bound (App. B): the bulk agrees
        # to 1e-5, a few strands in sliding contact may bifurcate
        e = rel_err(gp, p)
        e = e[np.isfinite(e)]
        assert np.median(e) < 1e-6 and np.quantile(e, 0.99) < 1e-4


def test_full_size_step_host_round_trip_equals_device_resident():
    """bh_step_host (H2D + substeps + D2H, sliced over internal streams) == upload, bh_step, download at 2^18 x 32."""
    rows, cols, N = 512, 512, 32
    S = rows * cols
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE)
        sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
        pos0, vel0, _ = sim.download()
        sim.step(float(DT), 4)
        want_p, want_v, _ = sim.download()
        hp, hv = pos0.reshape(-1).copy(), vel0.reshape(-1).copy()
        sim.step_host(float(DT), 4, hp, hv)
    assert_bit_equal(hp.reshape(-1, 4), want_p, "positions through host buffers")
    assert_bit_equal(hv.reshape(-1, 4), want_v, "velocities through host buffers")


def test_config3_shape_skinned_roots_and_capsules_sampled_parity():
    """configs[2] of BASELINE.json in shape (extension: oracle-defined, no reference parity): sphere scalp skinned by an
    8-joint dual-quaternion palette re-posed every frame, sphere + 2 capsule colliders, 4 substeps per frame — here at
    2^20 strands x 32 (tests/reports/config3.py runs the 4M-strand size). Whole tiles of strands are replayed on the CPU oracle
    and must match bit for bit; no free vertex may end inside a collider."""
    rows, cols, N, J, frames, substeps = 1024, 1024, 32, 8, 3, 4
    S = rows * cols
    caps = CAPSULE_SETS["arms"]
    root_pos, root_nrm, _ = po.sphere_scalp(rows, cols)
    jy = np.linspace(-1.0, 1.0, J, dtype=np.float32)
    d = np.abs(root_pos[:, 1:2] - jy[None, :])
    joints = np.argsort(d, axis=1)[:, :4].astype(np.int32)
    w = 1.0 / (np.take_along_axis(d, joints, 1) + 0.05)
    weights = np.ascontiguousarray((w / w.sum(1, keepdims=True)).astype(np.float32)[:, :3])

    def palette(frame):
        out = np.zeros((J, 8), np.float32)
        for j in range(J):
            a = 0.05 * np.sin(0.35 * frame + 0.7 * j)
            tx = 0.02 * np.sin(0.2 * frame + j)
            qz, qw = np.sin(a / 2), np.cos(a / 2)
            out[j] = (0.0, 0.0, qz, qw, 0.5 * tx * qw, -0.5 * tx * qz, 0.0, 0.0)   # real xyzw, dual = 0.5 * t * q
        return out

    par, gcfg = _capsule_params(caps, dt=float(np.float32(DT) / np.float32(substeps)), scale=1.45, sphere=SPHERE)
    tiles = np.unique(np.concatenate([[0, S // 32 - 1], np.arange(0, S // 32, 257)]))
    idx = (tiles[:, None] * 32 + np.arange(32)).ravel()
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg)
        assert sim.kernel_kind == 0
        sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S))
        sim.set_skin(root_pos, joints, weights)
        pos0, vel0, _ = sim.download()
        p, v = _sample_strands(pos0, vel0, S, N, idx)
        for f in range(frames):
            dq = palette(f)
            sim.skin_roots(dq)
            sim.step(float(DT), substeps)
            sp, _ = po.skin_roots_dq(root_pos[idx], root_nrm[idx], joints[idx], weights[idx], dq)
            p.reshape(-1, N, 4)[:, 0, :3] = sp
            for _ in range(substeps):
                po.step(p, v, idx.size, N, par, nthreads=16)
        pos1, vel1, _ = sim.download()
    gp, gv = _sample_strands(pos1, vel1, S, N, idx)
    assert_bit_equal(gp, p, "sampled positions")
    assert_bit_equal(gv, v, "sampled velocities")
    assert_bit_equal(pos1[:, 3], pos0[:, 3], "I4 rest lengths")
    x = pos1[:, :3].astype(np.float64).reshape(S, N, 3)[:, 1:].reshape(-1, 3)
    x = x[np.isfinite(x).all(axis=1)]
    # the LAST collider always leaves its vertices on its surface; earlier ones may be re-entered by a later push-out
    a, b, r = (np.array(t, np.float64) for t in caps[-1])
    ab = b - a
    t = np.clip((x - a) @ ab / (ab @ ab), 0.0, 1.0)
    dist = np.linalg.norm(x - (a + t[:, None] * ab), axis=1)
    assert dist.min() >= float(r) * (1 - 1e-5), "vertex inside the last capsule"
    assert (np.abs(dist - float(r)) < 1e-5).any(), "nothing rests on the capsule: the case does not cover it"


@pytest.mark.parametrize("N", [2, 3, 4, 8, 16, 32, 64, 128])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_states_bit_exact(N, seed):
    """The randomised states of tests/test_oracle_vs_reference_live.py (gentle / rough / violent, three spheres, three
    time steps and scales) on the device: every kernel choice (streaming, per-strand) against the oracle, bit for bit."""
    from tests.test_oracle_vs_reference_live import random_state
    rng = np.random.default_rng(1000 * N + seed)
    S = 1500
    pos, vel = random_state(rng, S, N, [0.02, 0.3, 2.0][seed])
    sphere = [(0.0, 0.0, 0.0, 0.98), (0.1, -0.2, 0.05, 1.1), (0.0, 0.5, 0.0, 0.6)][seed]
    scale = [1.0, 1.45, 0.7][seed]
    dt = float(np.float32(1.0) / np.float32(90.0)) * [1.0, 0.25, 2.0][seed]
    par = po.default_params(dt=dt, scale=scale, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(6):
        po.step(rp, rv, S, N, par, nthreads=16)
    gp, gv = gpu_steps(pos, vel, S, N, 6, dt=dt, scale=scale, sphere=sphere)
    assert_bit_equal(gp, rp, "positions")
    assert_bit_equal(gv, rv, "velocities")


# ---- frame-level substep fusion: k substeps as k passes of one launch ------------------------------------------------------

@pytest.mark.parametrize("S,N,k", [(4096, 32, 4), (4000, 32, 2), (1000, 8, 4), (2048 + 77, 16, 3), (33333, 4, 4), (7001, 20, 4), (300, 128, 2),
                                   (96, 8, 4), (40, 8, 4), (65, 4, 2)])
@pytest.mark.parametrize("math", [bb.BH_MATH_EXACT, bb.BH_MATH_FAST])
def test_fused_substeps_bit_identical_to_separate_launches(S, N, k, math):
    """bh_set_substep_fusion: bh_step(dt, k) as k passes of ONE launch must equal k launches bit for bit in BOTH arithmetic
    profiles (same operations per strand, only the schedule differs) — whole tiles and ragged ones, every chunk count
    (group sizes 1, 2 and 4 tiles), the two-strands-per-row shape (N = 4), shapes too small for a group (silent fallback) —
    and the exact profile must still equal the oracle."""
    pos, vel = ragged_state(S, N)
    outs = []
    for fuse in (False, True):
        with bb.HairSim(S, N) as sim:
            sim.configure(scale=1.45, sphere=SPHERE, math=math)
            sim.set_substep_fusion(fuse, always=True)
            sim.upload(pos, vel)
            l0 = sim.launch_count
            for _ in range(3):
                sim.step(float(DT), k)
            launches = sim.launch_count - l0
            outs.append(sim.download()[:2] + (launches,))
    assert_bit_equal(outs[1][0], outs[0][0], "positions, fused vs separate launches")
    assert_bit_equal(outs[1][1], outs[0][1], "velocities, fused vs separate launches")
    ntiles = ((S // 2 if N == 4 else S) + 31) // 32
    chunks = 1 if N == 4 else (N + 7) // 8
    group = 1 if chunks >= 4 else -(-4 // chunks)
    fusable = ntiles >= group and not (N == 4 and S % 2)
    assert outs[0][2] >= 3 * k and outs[1][2] == (3 if fusable else outs[0][2]), (outs[0][2], outs[1][2])
    if math == bb.BH_MATH_EXACT:
        h = float(np.float32(DT) / np.float32(k))
        par = po.default_params(dt=h, scale=1.45, sphere=SPHERE)
        for _ in range(3 * k):
            po.step(pos, vel, S, N, par, nthreads=8)
        assert_bit_equal(outs[1][0], pos, "fused positions vs the oracle")
        assert_bit_equal(outs[1][1], vel, "fused velocities vs the oracle")


def test_fused_substeps_full_size_bit_identical_and_capsules():
    """configs[1] at full size (2^20 x 32, 4 substeps): fused frames == unfused frames over all 33.5M vertices, settled
    state with contacts; and a capsule scene at 2^16 x 32."""
    rows, cols, N = 1024, 1024, 32
    S = rows * cols
    res = []
    for fuse in (False, True):
        with bb.HairSim(S, N) as sim:
            sim.configure(scale=1.45, sphere=SPHERE, math=bb.BH_MATH_EXACT)
            sim.set_substep_fusion(fuse, always=True)
            sim.init_sphere_scalp(rows, cols, 0, bb.random_values(1234, 0, S), order=bb.BH_SCALP_COLUMN_MAJOR)
            for _ in range(12):
                sim.step(float(DT), 4)
            res.append(sim.checksum(3))
            if fuse:
                p1, v1, _ = sim.download()
            else:
                p0, v0, _ = sim.download()
    assert_bit_equal(p1, p0, "positions"); assert_bit_equal(v1, v0, "velocities")
    assert res[0] == res[1]
    S, N = 1 << 16, 32
    pos, vel = ragged_state(S, N)
    cfg = bb.default_params(); cfg.scale = 1.45; cfg.math = bb.BH_MATH_EXACT
    for i, x in enumerate(SPHERE): cfg.sphere[i] = x
    caps = CAPSULE_SETS["arms"]
    cfg.ncapsules = len(caps)
    for q, (a, b, r) in enumerate(caps):
        for i in range(3): cfg.capsules[q].a[i], cfg.capsules[q].b[i] = a[i], b[i]
        cfg.capsules[q].radius = r
    outs = []
    for fuse in (False, True):
        with bb.HairSim(S, N) as sim:
            sim.set_params(cfg); sim.set_substep_fusion(fuse, always=True); sim.upload(pos, vel)
            for _ in range(5):
                sim.step(float(DT), 4)
            outs.append(sim.download()[:2])
    assert_bit_equal(outs[1][0], outs[0][0], "capsule scene positions"); assert_bit_equal(outs[1][1], outs[0][1], "capsule scene velocities")


@pytest.mark.parametrize("S,N,k,fuse", [(20000, 16, 2, False), (70001, 20, 4, True), (33333, 4, 3, False), (5000, 32, 4, True), (1 << 17, 32, 4, True)])
def test_step_readback_equals_step_then_download(S, N, k, fuse):
    """bh_step_readback (the reference's frame: state resident, positions out; sliced so the copy overlaps the steps) ==
    bh_step + bh_download, bit for bit, and the velocities left on the device agree too; also through bh_step_host."""
    pos, vel = ragged_state(S, N)
    with bb.HairSim(S, N) as a, bb.HairSim(S, N) as b:
        for sim in (a, b):
            sim.configure(scale=1.45, sphere=SPHERE)
            sim.set_substep_fusion(fuse, always=True)
            sim.upload(pos, vel)
        out = bb.PinnedBuffer(4 * S * N)
        for _ in range(3):
            a.step(float(DT), k)
            b.step_readback(float(DT), k, out.array)
        wp, wv, _ = a.download()
        assert_bit_equal(out.array.reshape(-1, 4), wp, "positions read back slice by slice")
        gp, gv, _ = b.download()
        assert_bit_equal(gp, wp); assert_bit_equal(gv, wv, "velocities left on the device")
        hp, hv = pos.reshape(-1).copy(), vel.reshape(-1).copy()
        for _ in range(3):
            b.step_host(float(DT), k, hp, hv)
        assert_bit_equal(hp.reshape(-1, 4), wp, "bh_step_host with the same fusion setting")
        out.free()
    with bb.HairSim(64, 8) as sim:
        with pytest.raises(bb.BarbuHairError):
            sim.step_readback(float(DT), 1, np.zeros(4 * 64 * 8, np.float32))     # no strand state


def test_big_blocks_on_every_shape_bit_exact():
    """The launcher gives shards that fill the GPU ONE block of 12 warps per SM instead of three of 4 (same kernel binary,
    warps per block read from blockDim). BH_STREAM_BIG_BLOCKS=2 forces that shape for every launch — shards of a few tiles,
    where most warps of a block find no tile, included — and the fuzz and the capsule cases must stay bit-exact.
    The knob is read once per process, hence the subprocess."""
    import subprocess, sys
    env = dict(os.environ, BH_STREAM_BIG_BLOCKS="2")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "random_scenes or temporal_bound or substeps_equal or (stream_kernel_capsules_bit_exact and arms)"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]
