"""Multi-GPU host logic: contiguous strand ranges with no per-step exchange, and the optional
all-gather of the position plane. The N>1 plumbing is exercised here on CPU with the gloo backend (world size 2 and 3;
the oracle stands in for the device step — it is the checker of what the gather must reassemble); the same code
runs over NCCL in the gpu-marked test and in bench.py."""
import os
import socket

import numpy as np
import pytest

import barbu_b200 as bb
from barbu_b200 import shard
from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, sphere_state


@pytest.mark.parametrize("S,world", [(1, 1), (10, 3), (1 << 20, 8), (4097, 8), (7, 8), (0, 2)])
def test_shard_ranges_tile_the_strand_set(S, world):
    ranges = [shard.shard_range(S, world, r) for r in range(world)]
    assert ranges[0][0] == 0 and sum(c for _, c in ranges) == S
    for (f0, c0), (f1, _) in zip(ranges, ranges[1:]):
        assert f0 + c0 == f1
    counts = shard.shard_counts(S, world)
    assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(S, world, world)


def test_local_patch_indices_and_seams():
    rows, cols = 6, 8
    _, _, tri = po.sphere_scalp(rows, cols)
    S = rows * cols
    seen = np.zeros(tri.shape[0], bool)
    for r in range(3):
        first, count = shard.shard_range(S, 3, r)
        local, seams = shard.local_patch_indices(tri, first, count)
        assert local.min() >= 0 and local.max() < count
        inside = ((tri >= first) & (tri < first + count)).all(axis=1)
        assert_bit_equal(local, (tri[inside] - first).astype(np.int32))
        seen |= inside
        seen[seams] = True
    assert seen.all(), "every triangle is local to a shard or reported as a seam"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows, cols, N, nsteps, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = rows * cols
        first, count = shard.shard_range(S, world, rank)
        # what rank r builds for its shard on the device in production: roots of its range, jitter slice, expansion
        root_pos, root_nrm, _ = po.sphere_scalp(rows, cols)
        rv = bb.random_values(1234, first, count)                     # host entry point of the product library
        pos, vel = po.init_strands(root_pos[first:first + count], root_nrm[first:first + count], rv, N)
        par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
        for _ in range(nsteps):                                       # no communication inside the loop
            po.step(pos, vel, count, N, par)
        counts = [c * N for c in shard.shard_counts(S, world)]
        full = shard.allgather_plane(torch.from_numpy(pos), counts)
        if rank == 0:
            np.save(out, full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rows,cols", [(2, 8, 16), (3, 5, 7)])
def test_sharded_run_plus_allgather_equals_single_run_gloo(world, rows, cols, tmp_path):
    import torch.multiprocessing as mp
    N, nsteps = 8, 3
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), rows, cols, N, nsteps, out), nprocs=world, join=True)
    _, _, _, _, pos, vel = sphere_state(rows, cols, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(nsteps):
        po.step(pos, vel, rows * cols, N, par)
    assert_bit_equal(np.load(out), pos, "gathered position plane == single-process run")


def test_column_major_scalp_is_the_same_mesh_renumbered():
    """BH_SCALP_COLUMN_MAJOR (strand c * rows + r): the product's host triangle list against the oracle's renumbered scalp,
    and the property that makes it the balanced partition — every contiguous shard holds every latitude row equally often."""
    rows, cols = 12, 16
    pos_r, _, tri_r = po.sphere_scalp(rows, cols)
    pos_c, _, tri_c = po.sphere_scalp(rows, cols, column_major=True)
    assert_bit_equal(bb.sphere_scalp_triangles(rows, cols, bb.BH_SCALP_COLUMN_MAJOR), tri_c, "column-major triangles")
    assert_bit_equal(bb.sphere_scalp_triangles(rows, cols, bb.BH_SCALP_ROW_MAJOR), tri_r, "row-major triangles")
    assert_bit_equal(pos_c[tri_c], pos_r[tri_r], "same faces, same corner positions, same face order")
    for world in (2, 4, 8):
        for rank in range(world):
            first, count = shard.shard_range(rows * cols, world, rank)
            lat = np.arange(first, first + count) % rows                 # latitude row of each strand of the shard
            assert (np.bincount(lat, minlength=rows) == cols // world).all()
    with pytest.raises(bb.BarbuHairError):
        bb.sphere_scalp_triangles(rows, cols, 7)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [0, 1])
def test_hair_shard_on_device_matches_global_run(order):
    """Two shards stepped independently on the device == the oracle stepping the whole scalp (zero exchange steps), for
    both strand orders (row-major: latitude bands; column-major: longitude wedges)."""
    rows, cols, N = 32, 64, 16
    _, _, _, _, pos, vel = sphere_state(rows, cols, N, column_major=bool(order))
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(3):
        po.step(pos, vel, rows * cols, N, par)
    got = []
    for rank in range(2):
        sh = shard.HairShard(rows, cols, N, 2, rank, device=0, order=order, scale=1.45, sphere=SPHERE)
        for _ in range(3):
            sh.step(float(DT), 1)
        sh.sim.synchronize()
        t = shard.plane_tensor(sh.sim, 0)                              # zero-copy view NCCL would send
        assert t.is_cuda and tuple(t.shape) == (sh.count * N, 4)
        got.append(t.cpu().numpy())
        sh.close()
    assert_bit_equal(np.concatenate(got), pos, "positions of shard 0 ++ shard 1")
