"""Strand-state files and device checksums (format in include/barbu_hair.h).

CPU side: header parsing and validation (bh_peek_state needs no device) against files written here from the format
description alone. GPU side: save/load round trips, bit-exact resume, checksum vs the numpy restatement, additivity over
shards ("a checksum of checksums"), corruption detection."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import barbu_b200 as bb
from barbu_b200 import hair
from oracle import pyoracle as po
from tests.util import DT, SPHERE, assert_bit_equal, host_checksum, sphere_state, write_state_file


def params_bytes(**kw):
    p = bb.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return bytes(p)


def test_format_constants():
    assert C.sizeof(bb.BhParams) == 292, "bh_params is part of the file format"
    assert struct.calcsize("<8sIIqiIqqqfIQQII") == 88


def test_peek_reads_a_file_written_from_the_format_description(lib, tmp_path):
    S, N = 6, 4
    rng = np.random.default_rng(0)
    planes = [rng.standard_normal((S * N, 4)).astype(np.float32) for _ in range(2)]
    path = tmp_path / "a.bhs"
    write_state_file(path, planes, S, N, params_bytes=params_bytes(scale=1.25, iterations=8), total=48, first=12, frame=77,
                     dt=float(DT), seed=1234)
    info = bb.peek_state(path)
    assert (info.nstrands, info.nverts, info.plane_mask) == (S, N, 3)
    assert (info.total_strands, info.first_strand, info.frame, info.seed) == (48, 12, 77, 1234)
    assert np.float32(info.dt) == DT and np.float32(info.params.scale) == np.float32(1.25)
    assert (int(info.checksum[0]), int(info.checksum[1])) == host_checksum(planes, N, 12, (0, 1))
    assert os.path.getsize(path) == 512 + 2 * S * N * 16


@pytest.mark.parametrize("breakage", ["magic", "version", "truncated", "trailing", "shape", "shard", "mask", "params", "missing"])
def test_peek_refuses_bad_files(lib, tmp_path, breakage):
    S, N = 3, 4
    planes = [np.ones((S * N, 4), np.float32)]
    path = tmp_path / "b.bhs"
    kw = dict(params_bytes=params_bytes())
    if breakage == "magic":
        kw["magic"] = b"BARBUHS0"
    if breakage == "version":
        kw["version"] = 9
    if breakage == "shard":
        kw.update(total=4, first=2)                       # 2 + 3 > 4
    if breakage == "params":
        kw["params_bytes"] = params_bytes(math=7)
    if breakage == "mask":
        kw["plane_mask"] = 1
    write_state_file(path, planes, 0 if breakage == "shape" else S, N, **kw)
    if breakage == "truncated":
        os.truncate(path, os.path.getsize(path) - 16)
    if breakage == "trailing":
        with open(path, "ab") as f:
            f.write(b"\0" * 16)
    if breakage == "mask":                                # header says two planes, payload holds one
        raw = bytearray(open(path, "rb").read()); raw[28] = 3; open(path, "wb").write(raw)
    if breakage == "missing":
        path = tmp_path / "nope.bhs"
    with pytest.raises(bb.BarbuHairError) as e:
        bb.peek_state(path)
    assert e.value.code == (hair.BH_ERR_UNSUPPORTED if breakage == "version" else hair.BH_ERR_INVALID)


def test_save_and_load_need_a_sim(lib, tmp_path):
    assert lib.bh_save_state(None, b"x", None) == hair.BH_ERR_INVALID
    assert lib.bh_load_state(None, b"x", None) == hair.BH_ERR_INVALID
    assert lib.bh_state_checksum(None, 7, 0, None) == hair.BH_ERR_INVALID


# ---- device ---------------------------------------------------------------------------------------------------------

def stepped_sim(rows, cols, N, steps):
    _, root_nrm, _, _, pos, vel = sphere_state(rows, cols, N)
    sim = bb.HairSim(rows * cols, N)
    sim.configure(scale=1.45, sphere=SPHERE)
    sim.upload(pos, vel, po.init_tangents(root_nrm, N))
    for _ in range(steps):
        sim.step(float(DT), 1)
    return sim


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,N", [(8, 16, 32), (5, 7, 4), (3, 1, 1)])
def test_device_checksum_matches_the_numpy_restatement(rows, cols, N):
    with stepped_sim(rows, cols, N, 2) as sim:
        planes = sim.download(True, True, True)
        assert sim.checksum(7) == host_checksum(planes, N)
        assert sim.checksum(1) == host_checksum(planes[:1], N, 0, (0,))
        assert sim.checksum(6, first_strand=1000) == host_checksum(planes[1:], N, 1000, (1, 2))


@pytest.mark.gpu
def test_checksum_of_a_scalp_is_the_sum_of_its_shard_checksums():
    rows, cols, N = 8, 12, 16
    S = rows * cols
    with stepped_sim(rows, cols, N, 3) as whole:
        want = whole.checksum(3)
        pos, vel, _ = whole.download()
    for cuts in [(0, 40, S), (0, 1, 33, 95, S)]:
        c0 = c1 = 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            with bb.HairSim(b - a, N) as shard:
                shard.upload(pos[a * N:b * N], vel[a * N:b * N])
                x, y = shard.checksum(3, first_strand=a)
            c0, c1 = (c0 + x) % 2**64, (c1 + y) % 2**64
        assert (c0, c1) == want
    # ... and it does see a swap of two vertices that a plain sum does not
    with bb.HairSim(S, N) as sim:
        pos2 = pos.copy(); pos2[[5, 9]] = pos2[[9, 5]]
        sim.upload(pos2, vel)
        got = sim.checksum(3)
    assert got[0] == want[0] and got[1] != want[1]


@pytest.mark.gpu
def test_round_trip_and_bit_exact_resume(tmp_path):
    rows, cols, N = 8, 16, 32
    path = tmp_path / "resume.bhs"
    with stepped_sim(rows, cols, N, 6) as straight:
        want = straight.download(True, True, True)
    with stepped_sim(rows, cols, N, 3) as first_half:
        first_half.configure(wind=(0.0, 0.0, 0.0), damp=0.8)
        first_half.save(path, frame=3, dt=float(DT), seed=1234)
        saved = first_half.download(True, True, True)
        saved_sum = first_half.checksum(7)
    info = bb.peek_state(path)
    assert (info.frame, info.nstrands, info.nverts, info.plane_mask) == (3, rows * cols, N, 7)
    assert (int(info.checksum[0]), int(info.checksum[1])) == saved_sum == host_checksum(saved, N)
    raw = np.fromfile(path, np.float32, offset=512).reshape(3, -1, 4)      # the payload IS the planes of buffer 0
    for p in range(3):
        assert_bit_equal(raw[p], saved[p], f"payload plane {p}")
    sim, info = bb.HairSim.from_state(path)
    with sim:
        assert np.float32(sim.get_params().scale) == np.float32(1.45) and tuple(sim.get_params().sphere) == tuple(np.float32(SPHERE))
        for a, b, what in zip(sim.download(True, True, True), saved, ("pos", "vel", "tan")):
            assert_bit_equal(a, b, "loaded " + what)
        for _ in range(info.frame, 6):
            sim.step(info.dt, 1)
        for a, b, what in zip(sim.download(True, True, True), want, ("pos", "vel", "tan")):
            assert_bit_equal(a, b, "resumed " + what)


@pytest.mark.gpu
def test_load_a_file_written_by_the_host_writer_and_by_the_oracle(tmp_path):
    """Oracle -> device exchange: a state stepped on the CPU oracle, written here, continues on the device bit for bit."""
    rows, cols, N = 4, 8, 16
    S = rows * cols
    _, _, _, _, pos, vel = sphere_state(rows, cols, N)
    par = po.default_params(dt=float(DT), scale=1.45, sphere=SPHERE)
    for _ in range(2):
        po.step(pos, vel, S, N, par)
    path = tmp_path / "oracle.bhs"
    pb = bb.default_params(); pb.scale = 1.45
    for i, x in enumerate(SPHERE):
        pb.sphere[i] = x
    write_state_file(path, [pos, vel], S, N, params_bytes=bytes(pb), frame=2, dt=float(DT))
    with bb.HairSim(S, N) as sim:
        info = sim.load(path)
        assert info.plane_mask == 3
        sim.step(float(DT), 1)
        po.step(pos, vel, S, N, par)
        gpos, gvel, _ = sim.download()
    assert_bit_equal(gpos, pos, "pos"); assert_bit_equal(gvel, vel, "vel")


@pytest.mark.gpu
def test_corruption_and_shape_mismatch_are_refused(tmp_path):
    rows, cols, N = 4, 8, 8
    path = tmp_path / "c.bhs"
    with stepped_sim(rows, cols, N, 1) as sim:
        sim.save(path)
        raw = bytearray(open(path, "rb").read())
        raw[512 + 1000] ^= 0x10                                             # one flipped payload bit
        open(path, "wb").write(raw)
        with pytest.raises(bb.BarbuHairError) as e:
            sim.load(path)
        assert e.value.code == hair.BH_ERR_INVALID and "checksum" in str(e.value)
        with pytest.raises(bb.BarbuHairError) as e:                         # the sim refuses to step garbage
            sim.step(float(DT), 1)
        assert e.value.code == hair.BH_ERR_NOT_INITIALIZED
    with bb.HairSim(rows * cols, N * 2) as other:
        with pytest.raises(bb.BarbuHairError) as e:
            other.load(path)
        assert e.value.code == hair.BH_ERR_INVALID
    with bb.HairSim(4, 4) as empty:
        with pytest.raises(bb.BarbuHairError) as e:
            empty.save(tmp_path / "d.bhs")
        assert e.value.code == hair.BH_ERR_NOT_INITIALIZED
