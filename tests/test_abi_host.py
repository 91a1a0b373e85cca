"""CPU-side checks of the drop-in boundary: libbarbu_hair.so loads and exports exactly what include/barbu_hair.h
declares, the host-evaluated entry points agree with the oracle bit for bit, and every compute entry point fails
loudly (BH_ERR_CUDA) when there is no device — there is no CPU path to fall back to."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import barbu_b200 as bb
from barbu_b200 import hair
from oracle import pyoracle as po
from tests.util import assert_bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "barbu_hair.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/barbu_hair.h but not exported"
    assert sorted(hair.ABI_SYMBOLS) == declared, "barbu_b200.hair.ABI_SYMBOLS is out of sync with the header"
    assert lib.bh_version().decode().startswith("barbu_hair")


def test_params_struct_layout_matches_header(lib):
    p = bb.default_params()
    assert C.sizeof(hair.BhParams) == 4 * (1 + 4 + 1 + 3 + 1 + 1 + 1 + 3 + 1 + 1) + hair.BH_MAX_CAPSULES * 28
    assert (p.scale, p.iterations, p.force_coeff, p.damp, p.math) == (np.float32(1.45), 8, 20.0, np.float32(0.8), 0)
    assert tuple(p.sphere) == (0.0, 0.0, 0.0, 1.0) and tuple(p.gravity) == (0.0, np.float32(-9.81), 0.0)
    assert p.ncapsules == 0 and tuple(p.wind) == (0.0, 0.0, 0.0) and p.drag == 0.0


def test_random_values_match_oracle_and_slice():
    full = bb.random_values(1234, 0, 1000)
    assert_bit_equal(full, po.random_values(1234, 1000), "length jitter (hair.cc:273-275)")
    assert_bit_equal(bb.random_values(1234, 400, 350), full[400:750], "slice [400, 750)")
    assert bb.random_values(1234, 0, 0).size == 0
    assert not np.array_equal(bb.random_values(7, 0, 16), full[:16])
    assert (full > 0.9 - 1e-6).all() and (full < 1.1 + 1e-6).all()


@pytest.mark.parametrize("N", [2, 4, 16, 33])
def test_tangents_host_match_oracle_and_shard(N):
    _, nrm, _ = po.sphere_scalp(6, 10)
    S = nrm.shape[0]
    want = po.init_tangents(nrm, N)
    assert_bit_equal(bb.init_tangents_host(nrm, S, 0, N), want, "tangent plane (hair.cc:290-328)")
    first, count = 17, 29
    part = bb.init_tangents_host(nrm[first:first + count], S, first, N)
    assert_bit_equal(part, want[first * N:(first + count) * N], "tangent plane of a shard")


def test_sphere_scalp_triangles_match_oracle():
    for rows, cols in ((2, 3), (8, 16), (33, 17)):
        _, _, tri = po.sphere_scalp(rows, cols)
        assert_bit_equal(bb.sphere_scalp_triangles(rows, cols), tri, f"{rows}x{cols} triangle list")


def test_compute_entry_points_fail_loudly_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(bb.BarbuHairError) as e:
        bb.HairSim(64, 8)
    assert e.value.code == hair.BH_ERR_CUDA
    with pytest.raises(bb.BarbuHairError) as e:
        bb.build_patch_indices(np.array([[0, 1, 2]], np.int32), 4)
    assert e.value.code == hair.BH_ERR_CUDA
    with pytest.raises(bb.BarbuHairError):
        bb.selftest_math(0)
    # the mirror of the reference interface logs and stays uninitialised, like hair.cc:45-48
    h = bb.Hair()
    h.init()
    h.update(0.01)
    assert not h.initialized() and "without initialization" in h.log[-1]
    h.setup(None)
    assert not h.initialized() and "not found" in h.log[-1]


def test_argument_validation_needs_no_device(lib):
    out = C.c_void_p()
    assert lib.bh_create(C.byref(out), 0, 8, 0) == hair.BH_ERR_INVALID
    assert lib.bh_create(None, 8, 8, 0) == hair.BH_ERR_INVALID
    assert lib.bh_step(None, 0.01, 1) == hair.BH_ERR_INVALID
    assert lib.bh_random_values(1, -1, 4, None) == hair.BH_ERR_INVALID
    assert lib.bh_sphere_scalp_triangles(0, 4, None) == hair.BH_ERR_INVALID
    assert b"bh_sphere_scalp_triangles" in lib.bh_last_error()           # host-only entry points report through bh_last_error too
    assert lib.bh_step(None, 0.01, 1) == hair.BH_ERR_INVALID
    assert b"NULL" in lib.bh_last_error() or b"sim" in lib.bh_last_error()
    tri = np.array([[0, 1, 2]], np.int32)
    big = np.array([[0, 1, 2 ** 30]], np.int32)
    outbuf = np.zeros(6 * 3, np.int32)
    assert lib.bh_build_patch_indices(tri.ctypes.data_as(C.c_void_p), 1, 0, outbuf.ctypes.data_as(C.c_void_p), 0) == hair.BH_ERR_INVALID
    assert lib.bh_build_patch_indices(big.ctypes.data_as(C.c_void_p), 1, 4, outbuf.ctypes.data_as(C.c_void_p), 0) == hair.BH_ERR_OVERFLOW
