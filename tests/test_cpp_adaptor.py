"""The C++17 `Hair` adaptor (include/barbu_hair.hpp), driven like the reference's Renderer drives its Hair module."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from barbu_b200 import _build
from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, sphere_state


def run_driver(root_pos, root_nrm, tri, N, nframes, seed, dt, scale, sphere, math=0, devices=None):
    exe = _build.build_cpp_adaptor_driver()
    S, F = root_pos.shape[0], tri.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<qqiiIff4fi", S, F, N, nframes, seed, dt, scale, *sphere, math))
            f.write(np.ascontiguousarray(root_pos, np.float32).tobytes())
            f.write(np.ascontiguousarray(root_nrm, np.float32).tobytes())
            f.write(np.ascontiguousarray(tri, np.int32).tobytes())
        res = subprocess.run([exe, fin, fout] + ([devices] if devices else []), capture_output=True, text=True, timeout=300)
        if res.returncode != 0:
            return res, None
        raw = open(fout, "rb").read()
    V, nelems = struct.unpack_from("<qq", raw, 0)
    off = 16
    planes = []
    for _ in range(3):
        planes.append(np.frombuffer(raw, np.float32, 4 * V, off).reshape(V, 4).copy())
        off += 16 * V
    patch = np.frombuffer(raw, np.int32, nelems, off).copy()
    off += 4 * nelems
    (nstream,) = struct.unpack_from("<q", raw, off)
    stream = np.frombuffer(raw, np.float32, 4 * nstream, off + 8).reshape(nstream, 4).copy()
    return res, (planes[0], planes[1], planes[2], patch, stream)


def test_adaptor_driver_builds_and_fails_loudly_without_a_device():
    """CPU box: the driver compiles with g++ alone; with no CUDA device setup() must leave the module uninitialised
    (exit code 3) instead of computing anything on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU: covered by the gpu test")
    root_pos, root_nrm, tri, _, _, _ = sphere_state(4, 8, 4)
    res, out = run_driver(root_pos, root_nrm, tri, 4, 1, 1234, float(DT), 1.45, SPHERE)
    assert res.returncode == 3 and out is None
    assert "bh_create failed" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("N,scale,nframes", [(4, 1.45, 5), (16, 1.0, 3), (32, 1.45, 2)])
def test_adaptor_matches_oracle_bit_exact(N, scale, nframes):
    rows, cols = 16, 32
    root_pos, root_nrm, tri, rv, pos, vel = sphere_state(rows, cols, N)
    S = rows * cols
    res, out = run_driver(root_pos, root_nrm, tri, N, nframes, 1234, float(DT), scale, SPHERE)
    assert res.returncode == 0, res.stderr
    gp, gv, gt, patch, stream = out
    par = po.default_params(dt=float(DT), scale=scale, sphere=SPHERE)
    for _ in range(nframes):
        po.step(pos, vel, S, N, par)
    assert_bit_equal(gp, pos, "positions")
    assert_bit_equal(gv, vel, "velocities")
    assert_bit_equal(gt, po.init_tangents(root_nrm, N), "tangent plane")
    assert_bit_equal(patch, po.patch_indices(tri, N), "patch indices")
    # Hair::stream() = the tess-stream half of the reference's render(), default tessellation 3 x 2 x 16, seed = params.b200.seed
    assert_bit_equal(stream, po.tess_stream(pos, gt, patch, N, scale, 3, 2, 16, 1234), "tess-stream")


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,0,0", "0,1"])
def test_sharded_adaptor_matches_single_device_run(devices):
    """params().b200.devices: the same scalp sharded over several devices behind the same Hair interface (bh_group_*, one host
    thread) — positions, velocities, tangents and patch indices equal the oracle's single run bit for bit; the gathered
    position plane exists on the render GPU. "0,0": two shards on one GPU; "0,1": two GPUs (skipped on a one-GPU box)."""
    import torch
    if len(set(devices.split(","))) > torch.cuda.device_count():
        pytest.skip("needs more GPUs than this box has")
    rows, cols, N, nframes = 15, 31, 16, 4                                # 465 strands: ragged shards
    root_pos, root_nrm, tri, rv, pos, vel = sphere_state(rows, cols, N)
    S = rows * cols
    res, out = run_driver(root_pos, root_nrm, tri, N, nframes, 1234, float(DT), 1.45, SPHERE, devices=devices)
    assert res.returncode == 0, res.stderr
    gp, gv, gt, patch, stream = out
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(nframes):
        po.step(pos, vel, S, N, par)
    assert_bit_equal(gp, pos, "positions")
    assert_bit_equal(gv, vel, "velocities")
    assert_bit_equal(gt, po.init_tangents(root_nrm, N), "tangent plane")
    assert_bit_equal(patch, po.patch_indices(tri, N), "patch indices")
