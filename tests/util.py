"""Shared helpers of the test-suite (oracle-side state builders, bitwise comparison)."""
import os

import numpy as np

from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = np.float32(1.0) / np.float32(90.0)          # core/global_clock.cc:160-162
SPHERE = (0.0, 0.0, 0.0, 0.98)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        # NaN payload/sign is the one thing IEEE-754 leaves open (x86 makes 0xFFC00000, sm_100 0x7FFFFFFF):
        # a NaN must meet a NaN, everything else must match bit for bit (so -0.0 != +0.0).
        neq = (bits(a) != bits(b)) & ~(np.isnan(a) & np.isnan(b))
    else:
        neq = a != b
    if neq.any():
        idx = np.argwhere(neq)[0]
        raise AssertionError(f"{what}: {int(neq.sum())} of {neq.size} words differ; first at {tuple(idx)}: "
                             f"{a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sphere_state(rows, cols, nverts, seed=1234, maxlength=0.5, column_major=False):
    """Oracle-built initial state of a rows x cols sphere scalp: (root_pos, root_nrm, tri, rv, pos, vel)."""
    root_pos, root_nrm, tri = po.sphere_scalp(rows, cols, column_major)
    rv = po.random_values(seed, rows * cols)
    pos, vel = po.init_strands(root_pos, root_nrm, rv, nverts, maxlength)
    return root_pos, root_nrm, tri, rv, pos, vel


def ragged_state(nstrands, nverts, seed=99):
    """nstrands not tied to a grid: first nstrands strands of a larger sphere scalp, jittered velocities."""
    rows = 8
    cols = (nstrands + rows - 1) // rows
    root_pos, root_nrm, _, rv, pos, vel = sphere_state(rows, max(cols, 1), nverts, seed)
    pos, vel = pos[:nstrands * nverts].copy(), vel[:nstrands * nverts].copy()
    rng = np.random.default_rng(seed)
    vel[:, :3] = (rng.standard_normal((nstrands * nverts, 3)) * 1e-3).astype(np.float32)
    return pos, vel


def rel_err(a, b):
    """max over vertices of |a-b| / |b| on xyz (SURVEY.md App. B error measure)."""
    a, b = np.asarray(a, np.float64)[:, :3], np.asarray(b, np.float64)[:, :3]
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)


def host_checksum(planes, nverts: int, first_strand: int = 0, plane_ids=(0, 1, 2)):
    """The state checksum include/barbu_hair.h defines, restated with numpy (checker: the product computes it on the device)."""
    G = np.uint64(0x9E3779B97F4A7C15)
    c0 = np.uint64(0)
    c1 = np.uint64(0)
    with np.errstate(over="ignore"):
        for pid, plane in zip(plane_ids, planes):
            w = np.ascontiguousarray(plane, dtype=np.float32).reshape(-1).view(np.uint32).astype(np.uint64)
            key = (np.uint64(pid) << np.uint64(60)) + np.uint64(4 * first_strand * nverts) + np.arange(w.size, dtype=np.uint64)
            c0 = c0 + w.sum(dtype=np.uint64)
            c1 = c1 + (w * ((key + np.uint64(1)) * G)).sum(dtype=np.uint64)
    return int(c0), int(c1)


def write_state_file(path, planes, nstrands, nverts, *, params_bytes, plane_mask=None, total=None, first=0, frame=0, dt=0.0, seed=0,
                     magic=b"BARBUHS1", version=1, checksum=None):
    """A BARBUHS1 state file written from the format description in include/barbu_hair.h alone (struct-packed here,
    independently of the C writer). `planes`: list of (V, 4) float32 in ascending plane order; `params_bytes`: bh_params."""
    import struct
    ids = [p for p in range(3) if (plane_mask if plane_mask is not None else (1 << len(planes)) - 1) >> p & 1]
    mask = sum(1 << p for p in ids)
    c = checksum if checksum is not None else host_checksum(planes, nverts, first, ids)
    head = struct.pack("<8sIIqiIqqqfIQQII", magic, version, 512, nstrands, nverts, mask, total if total is not None else nstrands,
                       first, frame, dt, seed, c[0], c[1], len(params_bytes), 0) + bytes(params_bytes)
    with open(path, "wb") as f:
        f.write(head.ljust(512, b"\0"))
        for pl in planes:
            f.write(np.ascontiguousarray(pl, np.float32).tobytes())
