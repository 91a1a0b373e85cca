"""bh_group_*: one scalp sharded over several GPUs behind one handle and one host thread (SURVEY.md §8b "Threading", §8e) —
the multi-GPU form a single-threaded host like the reference's frame loop can call. Contiguous strand ranges, no per-step
exchange, peer-copy gather of a plane to the render GPU. Device lists with repeats put several shards on one GPU, which is
how a one-GPU box covers the sharding logic; a distinct-device case runs where two GPUs exist."""
import numpy as np
import pytest

import barbu_b200 as bb
from barbu_b200 import shard
from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, sphere_state


def test_group_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU: covered by the gpu tests")
    with pytest.raises(bb.BarbuHairError):
        bb.HairGroup([0, 0], 64, 8)


def _device_plane_to_numpy(ptr, nvertices, device):
    import torch
    return torch.as_tensor(shard._CudaPlane(ptr, nvertices), device=torch.device("cuda", device)).cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 1], [1, 0, 1, 0]])
@pytest.mark.parametrize("order", [0, 1])
def test_group_equals_single_device_run_and_oracle(devices, order):
    import torch
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs more GPUs than this box has")
    rows, cols, N, frames, substeps = 31, 33, 16, 3, 2                    # 1023 strands: ragged shards, ragged tiles
    S = rows * cols
    _, _, _, _, pos, vel = sphere_state(rows, cols, N, column_major=bool(order))
    h = float(np.float32(DT) / np.float32(substeps))
    par = po.default_params(dt=h, scale=1.45, sphere=SPHERE)
    for _ in range(frames * substeps):
        po.step(pos, vel, S, N, par)
    with bb.HairGroup(devices, S, N) as grp:
        assert grp.size == len(devices)
        ranges = [grp.shard_range(g) for g in range(grp.size)]
        assert ranges == [shard.shard_range(S, len(devices), g) for g in range(len(devices))]
        grp.configure(scale=1.45, sphere=SPHERE)
        grp.init_sphere_scalp(rows, cols, order=order, seed=1234)
        for _ in range(frames):
            grp.set_bounding_sphere(SPHERE)                                # Renderer::update: collider feed, then the step
            grp.step(float(DT), substeps)
        ptr, ms = grp.gather_plane(bb.hair.BH_PLANE_POSITION, dst_device=devices[0])
        gathered = _device_plane_to_numpy(ptr, S * N, devices[0])
        gp, gv, _ = grp.download()
        assert grp.launch_count >= frames * substeps * len(devices)
    assert ms >= 0.0
    assert_bit_equal(gp, pos, "positions of the sharded run")
    assert_bit_equal(gv, vel, "velocities of the sharded run")
    assert_bit_equal(gathered, pos, "position plane gathered on the render GPU")


@pytest.mark.gpu
def test_group_global_planes_upload_step_timed_download():
    rows, cols, N = 16, 24, 8
    S = rows * cols
    root_pos, root_nrm, _, rv, pos, vel = sphere_state(rows, cols, N)
    tan = po.init_tangents(root_nrm, N)
    par = po.default_params(dt=float(DT), scale=1.0, sphere=SPHERE)
    with bb.HairGroup([0, 0, 0], S, N) as grp:
        grp.configure(scale=1.0, sphere=SPHERE, math=bb.BH_MATH_EXACT)
        grp.init_strands(root_pos, root_nrm, rv)                           # global arrays, split by range inside
        p0, v0, _ = grp.download()
        assert_bit_equal(p0, pos, "device strand generation across shards")
        grp.upload(pos, vel, tan)
        mx, per = grp.step_timed(float(DT), 1, 4)
        gp, gv, gt = grp.download()
        # substep fusion on every shard: 3 substeps as 3 passes of one launch per shard, bit-identical
        launches = grp.launch_count
        grp.set_substep_fusion(True, always=True)
        grp.step(float(DT), 3); grp.synchronize()
        assert grp.launch_count - launches == 3, "one fused launch per shard"
        fp, fv, _ = grp.download()
    assert len(per) == 3 and mx == max(per) and min(per) > 0.0
    for _ in range(4):
        po.step(pos, vel, S, N, par)
    assert_bit_equal(gp, pos); assert_bit_equal(gv, vel); assert_bit_equal(gt, tan)
    par3 = po.default_params(dt=float(np.float32(DT) / np.float32(3)), scale=1.0, sphere=SPHERE)
    for _ in range(3):
        po.step(pos, vel, S, N, par3)
    assert_bit_equal(fp, pos, "fused group step"); assert_bit_equal(fv, vel)


@pytest.mark.gpu
def test_group_argument_errors():
    with pytest.raises(bb.BarbuHairError):
        bb.HairGroup([0, 99], 64, 8)                                       # no such device
    with pytest.raises(bb.BarbuHairError):
        bb.HairGroup([0, 0, 0], 2, 8)                                      # fewer strands than shards
    with bb.HairGroup([0, 0], 64, 8) as grp:
        with pytest.raises(bb.BarbuHairError):
            grp.step(float(DT), 1)                                         # no strand state yet
        with pytest.raises(bb.BarbuHairError):
            grp.init_sphere_scalp(7, 7)                                    # rows * cols != strands
        with pytest.raises(bb.BarbuHairError):
            grp.gather_to_gl()                                             # no GL buffer registered
        with pytest.raises(bb.BarbuHairError):
            grp.register_gl_buffer(1, 0)                                   # no GL context in this process: CUDA refuses, nothing is left registered
        with pytest.raises(bb.BarbuHairError):
            grp.gather_plane(0, 0)                                         # shards hold no state
