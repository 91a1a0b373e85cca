"""Buffer 0 in a buffer somebody else owns (SURVEY.md §8 row g: the renderer's GL buffer, hair.cc:371-389). No GL context
exists in this image (NVIDIA_DRIVER_CAPABILITIES=compute,utility: no libEGL / libGL), so the GL registration itself can
only fail here; what CAN run is everything around the three cudaGraphics* calls — the state moving into the shared buffer,
every entry point bracketing its work with map / unmap, error paths leaving the buffer unmapped, the state moving back —
through bh_register_device_buffer, which puts a plain device allocation in the GL buffer's role on the same code path."""
import numpy as np
import pytest

import barbu_b200 as bb
from tests.util import DT, assert_bit_equal
from tests.test_gpu_parity import SPHERE, ragged_state

pytestmark = pytest.mark.gpu


def test_caller_owned_buffer_follows_the_gl_protocol_bit_exact():
    import torch
    S, N = 3000, 16
    V = S * N
    pos, vel = ragged_state(S, N)
    shared = torch.full((3 * V * 4 + 64,), float("nan"), dtype=torch.float32, device="cuda")   # 3 planes + slack, 256-byte aligned
    with bb.HairSim(S, N) as a, bb.HairSim(S, N) as b:
        for sim in (a, b):
            sim.configure(scale=1.45, sphere=SPHERE)
            sim.upload(pos, vel)
        assert b.buffer_map_stats() == (0, 0, False)
        b.register_device_buffer(shared.data_ptr(), shared.numel() * 4)
        torch.cuda.synchronize()
        assert_bit_equal(shared[:4 * V].cpu().numpy().reshape(-1, 4), pos, "the state moved into the shared buffer")
        for _ in range(5):
            a.step(float(DT), 2)
            b.step(float(DT), 2)
            m, u, now = b.buffer_map_stats()
            assert m == u and not now, "every step maps once and unmaps once"
        b.synchronize()
        wp, wv, _ = a.download()
        # the "renderer" reads plane 0 in place, without any call into the library
        assert_bit_equal(shared[:4 * V].cpu().numpy().reshape(-1, 4), wp, "positions, read in place from the shared buffer")
        assert_bit_equal(shared[4 * V:8 * V].cpu().numpy().reshape(-1, 4), wv, "velocities")
        gp, gv, _ = b.download()
        assert_bit_equal(gp, wp); assert_bit_equal(gv, wv)
        # entry points that need a plane pointer of their own refuse while the buffer is shared, and leave it unmapped
        for call in (lambda: b.device_plane(0), lambda: b.step_host(float(DT), 1, pos.reshape(-1).copy(), vel.reshape(-1).copy()),
                     lambda: b.step_readback(float(DT), 1, np.zeros(4 * V, np.float32))):
            with pytest.raises(bb.BarbuHairError):
                call()
        # a failing call between map and unmap (skinning with a palette that is too small) still unmaps
        root = pos.reshape(S, N, 4)[:, 0, :3].copy()
        b.set_skin(root, np.full((S, 4), 3, np.int32), np.full((S, 3), 0.25, np.float32))
        with pytest.raises(bb.BarbuHairError):
            b.skin_roots(np.tile(np.array([0, 0, 0, 1, 0, 0, 0, 0], np.float32), (2, 1)))      # joint 3 of a 2-joint palette
        m, u, now = b.buffer_map_stats()
        assert m == u and not now
        with pytest.raises(bb.BarbuHairError):
            b.register_device_buffer(shared.data_ptr(), shared.numel() * 4)                  # one buffer at a time
        # state moves back; the shared buffer is no longer touched
        b.unregister_device_buffer()
        shared.fill_(float("nan"))
        for _ in range(3):
            a.step(float(DT), 1); b.step(float(DT), 1)
        wp, wv, _ = a.download(); gp, gv, _ = b.download()
        assert_bit_equal(gp, wp); assert_bit_equal(gv, wv)
        assert bool(torch.isnan(shared).all()), "nothing written after unregistration"
        assert b.buffer_map_stats()[2] is False
        b.unregister_device_buffer()                                                            # idempotent


def test_shared_buffer_registration_errors_leave_the_sim_untouched():
    import torch
    S, N = 512, 8
    pos, vel = ragged_state(S, N)
    small = torch.zeros(S * N * 4, dtype=torch.float32, device="cuda")                          # one plane only
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE)
        sim.upload(pos, vel)
        for ptr, nbytes in ((small.data_ptr(), small.numel() * 4), (small.data_ptr() + 4, 1 << 30), (0, 1 << 30)):
            with pytest.raises(bb.BarbuHairError):
                sim.register_device_buffer(ptr, nbytes)
        host = np.zeros(3 * S * N * 4, np.float32)
        with pytest.raises(bb.BarbuHairError):
            sim.register_device_buffer(host.ctypes.data, host.nbytes)                            # not device memory
        with pytest.raises(bb.BarbuHairError):
            sim.register_gl_buffer(1)                                                            # no GL context in this image
        assert sim.buffer_map_stats() == (0, 0, False)
        sim.step(float(DT), 1)
        gp, _, _ = sim.download()
        assert np.isfinite(gp[:, :3]).any()
