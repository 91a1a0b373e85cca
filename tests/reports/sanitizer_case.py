"""compute-sanitizer workload: every streaming-kernel variant (N = 32 / 8 / 4, off-origin sphere, capsules, both profiles) checked
against the oracle, plus generation, tess-stream and checksum kernels, at sizes the sanitizer finishes in seconds.
Usage on the GPU box: compute-sanitizer --tool {memcheck,racecheck,initcheck,synccheck} python tests/reports/sanitizer_case.py"""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import barbu_b200 as bb
from oracle import pyoracle as po
from tests.util import DT, SPHERE, ragged_state, assert_bit_equal
# small shapes through every streaming variant: N=32 (root + inner chunks), N=8, N=4 (two strands per row), capsules, off-origin sphere
for (S, N, sphere, caps) in [(100, 32, SPHERE, 0), (70, 8, (0.1, 0.0, 0.05, 0.95), 0), (130, 4, (0, 0, 0, 1.02), 0), (90, 16, SPHERE, 2)]:
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere, ncapsules=caps)
    g = bb.default_params(); g.scale = 1.45; g.ncapsules = caps
    for i, x in enumerate(sphere): g.sphere[i] = x
    for q in range(caps):
        a, b, r = ((0.5, 0.3, 0.0), (1.2, -0.3, 0.0), 0.3) if q == 0 else ((0.0, 1.0, 0.0), (0.3, 1.4, 0.2), 0.25)
        for i in range(3): par.capsules[q].a[i] = g.capsules[q].a[i] = a[i]; par.capsules[q].b[i] = g.capsules[q].b[i] = b[i]
        par.capsules[q].radius = g.capsules[q].radius = r
    rp, rv = pos.copy(), vel.copy()
    for _ in range(6): po.step(rp, rv, S, N, par)
    for math in (bb.BH_MATH_EXACT, bb.BH_MATH_FAST):
        g.math = math
        with bb.HairSim(S, N) as sim:
            sim.set_params(g); sim.upload(pos, vel)
            for _ in range(6): sim.step(float(DT), 1)
            gp, gv, _ = sim.download()
        if math == bb.BH_MATH_EXACT:
            assert_bit_equal(gp, rp); assert_bit_equal(gv, rv)
    print("ok", S, N, caps, flush=True)
# tess-stream + checksum + generation kernels
sim = bb.HairSim(64 * 64, 16); sim.configure(scale=1.45, sphere=SPHERE)
sim.init_sphere_scalp(64, 64, 0, bb.random_values(1, 0, 64 * 64))
tri = bb.sphere_scalp_triangles(64, 64); sim.tess_set_patches(bb.build_patch_indices(tri, 16))
out = sim.tess_stream(3, 2, 16, 7); print("tess", out.shape, sim.checksum(3)); sim.close()
# ---- round 2 paths: fused substeps (forced), capsules behind the shell test (capsule inside / across / beyond the sphere), the
# sliced read-back, the shared-buffer protocol on a caller-owned device buffer, skinned roots from a matrix palette
import torch
for (S, N) in [(260, 32), (300, 8), (140, 4), (97, 20)]:
    pos, vel = ragged_state(S, N)
    with bb.HairSim(S, N) as a, bb.HairSim(S, N) as b:
        for sim, fuse in ((a, False), (b, True)):
            sim.configure(scale=1.45, sphere=SPHERE); sim.set_substep_fusion(fuse, always=True); sim.upload(pos, vel)
            for _ in range(3): sim.step(float(DT), 4)
        assert_bit_equal(b.download()[0], a.download()[0]); assert_bit_equal(b.download()[1], a.download()[1])
    print("ok fused", S, N, flush=True)
shell = [((-0.5, -1.45, 0.0), (0.5, -1.45, 0.1), 0.15), ((0.1, 0.2, 0.0), (0.3, 0.3, 0.1), 0.25), ((0.0, 0.9, 0.0), (0.05, 1.3, 0.0), 0.2)]
for (S, N, sphere) in [(150, 16, SPHERE), (90, 32, (0.3, 2.0, 0.1, 1.2)), (64, 8, (0.0, 0.0, 0.0, 5.0))]:
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=sphere, ncapsules=len(shell))
    g = bb.default_params(); g.scale = 1.45; g.ncapsules = len(shell)
    for i, x in enumerate(sphere): g.sphere[i] = x
    for q, (ca, cb, r) in enumerate(shell):
        for i in range(3): par.capsules[q].a[i] = g.capsules[q].a[i] = ca[i]; par.capsules[q].b[i] = g.capsules[q].b[i] = cb[i]
        par.capsules[q].radius = g.capsules[q].radius = r
    rp, rv = pos.copy(), vel.copy()
    for _ in range(6): po.step(rp, rv, S, N, par)
    with bb.HairSim(S, N) as sim:
        sim.set_params(g); sim.upload(pos, vel)
        for _ in range(6): sim.step(float(DT), 1)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp); assert_bit_equal(gv, rv)
    print("ok shell", S, N, flush=True)
S, N = 9000, 16
pos, vel = ragged_state(S, N)
shared = torch.zeros(3 * S * N * 4, dtype=torch.float32, device="cuda")
out = bb.PinnedBuffer(4 * S * N)
with bb.HairSim(S, N) as a, bb.HairSim(S, N) as b:
    for sim in (a, b):
        sim.configure(scale=1.45, sphere=SPHERE); sim.upload(pos, vel)
    b.register_device_buffer(shared.data_ptr(), shared.numel() * 4)
    for _ in range(3):
        a.step_readback(float(DT), 2, out.array); b.step(float(DT), 2)
    b.synchronize()
    assert_bit_equal(shared[:4 * S * N].cpu().numpy().reshape(-1, 4), out.array.reshape(-1, 4))
    b.unregister_device_buffer()
    root = pos.reshape(S, N, 4)[:, 0, :3].copy()
    G = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (4, 1)); G[:, 12] = [0.0, 0.01, 0.02, 0.03]
    dq = bb.dq_palette_from_matrices(G, np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (4, 1)))
    b.set_skin(root, np.tile(np.arange(4, dtype=np.int32), (S, 1)), np.full((S, 3), 0.25, np.float32)); b.skin_roots(dq); b.step(float(DT), 1)
    print("ok readback / device buffer / skin", b.buffer_map_stats(), flush=True)
out.free()
# ---- the latency-oriented wavefront kernel (bh_set_step_policy): ragged strand counts, capsules, passes
for (S, N, k) in [(449, 4, 4), (130, 16, 1), (37, 32, 2), (5, 1, 1)]:
    pos, vel = ragged_state(S, N)
    par = po.default_params(dt=float(np.float32(DT) / np.float32(k)), scale=1.45, sphere=SPHERE)
    rp, rv = pos.copy(), vel.copy()
    for _ in range(2 * k): po.step(rp, rv, S, N, par)
    with bb.HairSim(S, N) as sim:
        sim.configure(scale=1.45, sphere=SPHERE); sim.set_step_policy(bb.BH_POLICY_LATENCY); sim.upload(pos, vel)
        for _ in range(2): sim.step(float(DT), k)
        gp, gv, _ = sim.download()
    assert_bit_equal(gp, rp); assert_bit_equal(gv, rv)
    print("ok wave", S, N, k, flush=True)
