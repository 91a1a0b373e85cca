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


def sphere_state(rows, cols, nverts, seed=1234, maxlength=0.5):
    """Oracle-built initial state of a rows x cols sphere scalp: (root_pos, root_nrm, tri, rv, pos, vel)."""
    root_pos, root_nrm, tri = po.sphere_scalp(rows, cols)
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
