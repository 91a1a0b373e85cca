"""Known-answer tests of the oracle derived from the shader text: the reference has no tests
of its own, so these — together with the bit-exact agreement with the compiled reference sources
(test_oracle_golden.py) — are what pins the checker."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, bits, sphere_state


def step_copy(pos, vel, S, N, **kw):
    p, v = pos.copy(), vel.copy()
    po.step(p, v, S, N, po.default_params(**kw))
    return p, v


@pytest.mark.parametrize("N", [2, 4, 16, 32])
def test_I1_I4_roots_rest_lengths_and_w_planes(N):
    _, _, _, _, pos, vel = sphere_state(8, 16, N)
    S = 128
    p, v = pos.copy(), vel.copy()
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(5):
        po.step(p, v, S, N, par)
    assert_bit_equal(p[::N], pos[::N], "I1 roots (xyz and w) never change")
    assert not v[::N].any(), "I1 root velocity is zero"
    assert_bit_equal(p[:, 3], pos[:, 3], "I4 rest lengths are bit-preserved")
    assert not v[:, 3].any(), "I4 velocity w is zero"


@pytest.mark.parametrize("N,scale", [(4, 1.45), (16, 1.0), (32, 1.2)])
def test_I2_segments_have_scaled_rest_length_without_collider(N, scale):
    _, _, _, _, pos, vel = sphere_state(8, 16, N)
    p, _ = step_copy(pos, vel, 128, N, dt=float(DT), scale=scale, sphere=(0, 0, 0, 0))
    x = p[:, :3].astype(np.float64).reshape(128, N, 3)
    seg = np.linalg.norm(x[:, 1:] - x[:, :-1], axis=2)
    want = scale * p[:, 3].astype(np.float64).reshape(128, N)[:, 1:]
    # positions ~1.5 carry an fp32 ulp of 1.2e-7; a 0.02-long segment therefore closes to a few 1e-6 relative
    assert np.abs(seg / want - 1).max() < 2e-5


def test_I3_no_vertex_inside_the_sphere():
    N = 16
    _, _, _, _, pos, vel = sphere_state(16, 16, N)
    p, v = pos.copy(), vel.copy()
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(30):
        po.step(p, v, 256, N, par)
    r = np.linalg.norm(p[:, :3].astype(np.float64), axis=1).reshape(256, N)[:, 1:]
    assert r.min() >= SPHERE[3] * (1 - 1e-6)
    assert (np.abs(r - SPHERE[3]) < 1e-6).any(), "the test state must actually touch the collider"


def test_I5_hanging_strand_is_a_fixed_point():
    N, scale = 8, 1.0
    pos = np.zeros((N, 4), np.float32)
    rest = np.float32(0.0625)
    for i in range(N):
        pos[i] = (0.25, 2.0 - i * float(rest), -0.5, rest if i else 0.0)
    vel = np.zeros_like(pos)
    p, v = step_copy(pos, vel, 1, N, dt=float(DT), scale=scale, sphere=(0, 0, 0, 0))
    assert np.abs(p[:, :3] - pos[:, :3]).max() < 1e-6
    assert np.abs(v[:, :3]).max() < 1e-6


def test_I6_two_vertex_hand_case():
    L, dt = np.float32(0.5), np.float32(DT)
    pos = np.array([[0, 0, 0, 0], [L, 0, 0, L]], np.float32)
    vel = np.zeros_like(pos)
    p, v = step_copy(pos, vel, 1, 2, dt=float(dt), scale=1.0, sphere=(0, 0, 0, 0))
    f = np.float32(20.0) * np.float32(-9.81)
    x1 = np.array([L, np.float32(np.float32(dt * dt) * f), 0], np.float64)
    want = float(L) * x1 / np.linalg.norm(x1)
    assert np.abs(p[1, :3] - want).max() < 1e-6
    # with 8 iterations the projection is idempotent after the first: d = D(1,8) - C(1,7) ~ 0 => tip velocity ~ 0
    assert np.abs(v[1, :3]).max() < 1e-6


def test_I7_patch_indices_of_one_triangle():
    got = po.patch_indices(np.array([[0, 1, 2]], np.int32), 4)
    assert got.tolist() == [0, 1, 4, 5, 8, 9, 1, 2, 5, 6, 9, 10, 2, 3, 6, 7, 10, 11]
    assert po.patch_indices(np.zeros((0, 3), np.int32), 4).size == 0
    with pytest.raises(OverflowError):
        po.patch_indices(np.array([[0, 1, 2 ** 30]], np.int32), 4)


def test_I8_velocity_is_damped_displacement_of_the_next_vertex():
    """vel_i = 0.8 * d_{i+1} (1 <= i <= N-2), vel_{N-1} = d_{N-1}: check through the 7- and 8-iteration runs,
    d_i = D(i,8) - C(i,7), on a state that does not touch the collider in the last iteration."""
    N, S = 8, 64
    _, _, _, _, pos, vel = sphere_state(4, 16, N)
    rng = np.random.default_rng(3)
    vel[:, :3] = (rng.standard_normal((S * N, 3)) * 1e-2).astype(np.float32)
    vel[::N] = 0
    kw = dict(dt=float(DT), scale=1.0, sphere=(0, 0, 0, 0))
    p8, v8 = step_copy(pos, vel, S, N, iterations=8, **kw)
    p7, _ = step_copy(pos, vel, S, N, iterations=7, **kw)
    d = (p8[:, :3] - p7[:, :3]).reshape(S, N, 3)            # no collider: C(i,7) = D(i,7) = positions of the 7-iteration run
    v = v8[:, :3].reshape(S, N, 3)
    assert_bit_equal(v[:, 1:N - 1], (d[:, 2:] * np.float32(0.8)).astype(np.float32), "inner vertices")
    assert_bit_equal(v[:, N - 1], d[:, N - 1], "tip")


def test_step_is_strand_local_and_order_independent():
    """Strands never interact: stepping a permutation of the strands equals permuting the stepped strands — the
    property multi-GPU sharding relies on."""
    N, S = 16, 96
    _, _, _, _, pos, vel = sphere_state(6, 16, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    p, v = pos.copy(), vel.copy()
    po.step(p, v, S, N, par)
    perm = np.random.default_rng(0).permutation(S)
    pp = pos.reshape(S, N, 4)[perm].reshape(-1, 4).copy()
    vp = vel.reshape(S, N, 4)[perm].reshape(-1, 4).copy()
    po.step(pp, vp, S, N, par)
    assert_bit_equal(pp, p.reshape(S, N, 4)[perm].reshape(-1, 4))
    assert_bit_equal(vp, v.reshape(S, N, 4)[perm].reshape(-1, 4))
    assert bits(p).shape == bits(pos).shape
