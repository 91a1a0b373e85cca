"""The latency-oriented step kernel (hair_wave.cu; bh_set_step_policy): eight lanes per strand, one constraint iteration
each. Same arithmetic as the throughput kernels, so the exact profile must equal the CPU oracle bit for bit on every shape,
and the substeps of a frame run as passes of one launch."""
import numpy as np
import pytest

import barbu_b200 as bb
from oracle import pyoracle as po
from tests.util import DT, assert_bit_equal
from tests.test_gpu_parity import SPHERE, CAPSULE_SETS, _capsule_params, ragged_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,N", [(1, 4), (3, 1), (5, 2), (448, 4), (1001, 3), (130, 5), (77, 8), (600, 16), (333, 21), (200, 32), (64, 64), (19, 128)])
@pytest.mark.parametrize("sphere", [SPHERE, (0.05, 0.1, -0.02, 0.9)])
def test_wave_kernel_bit_exact_vs_oracle(S, N, sphere):
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(10):
        po.step(rp, rv, S, N, par)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=sphere)
        sim.set_step_policy(bb.BH_POLICY_LATENCY)
        assert sim.kernel_kind == 3
        sim.upload(pos, vel)
        for _ in range(10):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions"); assert_bit_equal(gv, rv, "velocities")


@pytest.mark.parametrize("caps", sorted(CAPSULE_SETS))
def test_wave_kernel_capsules_wind_drag_bit_exact(caps):
    S, N = 700, 16
    pos, vel = ragged_state(S, N)
    par, gcfg = _capsule_params(CAPSULE_SETS[caps], dt=float(DT), scale=1.45, sphere=SPHERE)
    for q in (par, gcfg):
        q.drag = 0.05
        for i, x in enumerate((3.0, 0.5, -2.0)): q.wind[i] = x
    rp, rv = pos.copy(), vel.copy()
    for _ in range(12):
        po.step(rp, rv, S, N, par, nthreads=8)
    with bb.HairSim(S, N) as sim:
        sim.set_params(gcfg); sim.set_step_policy(bb.BH_POLICY_LATENCY)
        assert sim.kernel_kind == 3
        sim.upload(pos, vel)
        for _ in range(12):
            sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions"); assert_bit_equal(gv, rv, "velocities")


@pytest.mark.parametrize("S,N,k", [(448, 4, 4), (513, 16, 3), (90, 32, 2)])
def test_wave_kernel_substeps_are_passes_of_one_launch(S, N, k):
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(np.float32(DT) / np.float32(k)), scale=1.45, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(3 * k):
        po.step(rp, rv, S, N, par)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE); sim.set_step_policy(bb.BH_POLICY_LATENCY); sim.upload(pos, vel)
        l0 = sim.launch_count
        for _ in range(3):
            sim.step(float(DT), k)
        assert sim.launch_count - l0 == 3, "one launch per frame"
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp, "positions"); assert_bit_equal(gv, rv, "velocities")


def test_step_policy_auto_by_size_and_fallbacks():
    with bb.HairSim(448, 4) as small, bb.HairSim(1 << 15, 16) as large:
        for sim in (small, large):
            sim.configure(scale=1.45, sphere=SPHERE)
        assert small.kernel_kind == 0 and large.kernel_kind == 0              # default: throughput kernels
        small.set_step_policy(bb.BH_POLICY_AUTO); large.set_step_policy(bb.BH_POLICY_AUTO)
        assert small.kernel_kind == 3 and large.kernel_kind == 0              # 2^18 vertices is the line
        small.configure(iterations=5)
        assert small.kernel_kind == 2, "other iteration counts keep the generic kernel"
        with pytest.raises(bb.BarbuHairError):
            small.set_step_policy(7)
    # the Hair adaptor picks AUTO: the reference's default-sized scalp runs on the latency kernel, bit-exact
    rows, cols, N = 16, 28, 4
    root_pos, root_nrm, tri = po.sphere_scalp(rows, cols)
    hair = bb.Hair(params=bb.Hair.Parameters(ncontrol_points=N))
    hair.set_bounding_sphere(SPHERE)
    hair.setup(bb.ScalpMesh(root_pos, root_nrm, tri))
    assert hair.sim.kernel_kind == 3
    pos, vel = po.init_strands(root_pos, root_nrm, po.random_values(hair.params.seed, rows * cols), N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(5):
        hair.update(float(DT)); po.step(pos, vel, rows * cols, N, par)
    gp, gv, _ = hair.sim.download()
    assert_bit_equal(gp, pos); assert_bit_equal(gv, vel)
    hair.deinit()


def test_wave_kernel_fast_profile_close_to_exact():
    S, N = 2000, 16
    pos, vel = ragged_state(S, N)
    outs = []
    for math in (bb.BH_MATH_EXACT, bb.BH_MATH_FAST):
        with bb.HairSim(S, N) as sim:
            sim.configure(scale=1.45, sphere=SPHERE, math=math); sim.set_step_policy(bb.BH_POLICY_LATENCY); sim.upload(pos, vel)
            sim.step(float(DT), 1)
            outs.append(sim.download()[0])
    a, b = outs[0][:, :3].astype(np.float64), outs[1][:, :3].astype(np.float64)
    ok = np.isfinite(a).all(axis=1) & np.isfinite(b).all(axis=1)
    rel = np.abs(a[ok] - b[ok]).max(axis=1) / np.maximum(np.abs(a[ok]).max(axis=1), 1e-3)
    assert np.median(rel) < 1e-6 and np.quantile(rel, 0.99) < 1e-5            # a rough state with contacts: the bulk, not the worst vertex


@pytest.mark.parametrize("fuse", [False, True])
def test_wave_kernel_behind_the_host_buffer_entry_points(fuse):
    """bh_step_host and bh_step_readback slice the shard and launch per slice: under the latency policy (decided by the
    sim's size, not the slice's) they run the wavefront kernel too and still equal step + download bit for bit."""
    S, N, k = 9001, 16, 3
    pos, vel = ragged_state(S, N)
    with bb.HairSim(S, N) as a, bb.HairSim(S, N) as b:
        for sim in (a, b):
            sim.configure(scale=1.45, sphere=SPHERE); sim.set_step_policy(bb.BH_POLICY_AUTO); sim.set_substep_fusion(fuse, always=True)
            assert sim.kernel_kind == 3
            sim.upload(pos, vel)
        out = bb.PinnedBuffer(4 * S * N)
        hp, hv = pos.reshape(-1).copy(), vel.reshape(-1).copy()
        for _ in range(3):
            a.step(float(DT), k)
            b.step_readback(float(DT), k, out.array)
        wp, wv, _ = a.download()
        assert_bit_equal(out.array.reshape(-1, 4), wp, "read-back positions")
        assert_bit_equal(b.download()[1], wv, "velocities left on the device")
        for _ in range(3):
            b.step_host(float(DT), k, hp, hv)
        b.upload(pos, vel)
        assert_bit_equal(hp.reshape(-1, 4), wp, "bh_step_host"); assert_bit_equal(hv.reshape(-1, 4), wv)
        out.free()
