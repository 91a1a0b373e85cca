"""Marschner lookup tables: oracle vs the reference shader sources (golden, bit-exact on the CPU),
CUDA kernel vs oracle (tolerance: transcendental functions are CUDA's on the device and libm's in the oracle)."""
import ctypes as C

import numpy as np
import pytest

import barbu_b200 as bb
from barbu_b200 import hair
from oracle import pyoracle as po
from tests.util import assert_bit_equal, golden

CASES = ["default", "wet", "dense"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_bit_exact_vs_reference_shader_sources(case):
    g = golden("marschner")
    p = g[case + "_params"]
    m, n = po.marschner_luts(p, 32)
    assert_bit_equal(m, g[case + "_m32_res32"], "M fp32"); assert_bit_equal(n, g[case + "_n32_res32"], "N fp32")
    m, n = po.marschner_luts(p, 128)
    assert_bit_equal(po.float_to_half(m), g[case + "_m16"].view(np.uint16), "M rgba16f")
    assert_bit_equal(po.float_to_half(n), g[case + "_n16"].view(np.uint16), "N rgba16f")


@pytest.mark.skipif(not po.ref_marschner_available(), reason="oracle/_ref not built (no reference tree)")
def test_oracle_bit_exact_vs_reference_live():
    p = po.marschner_params(eta=1.7, absorption=0.33, ar=-3.0, br=7.5)
    for res in (16, 77):
        m, n = po.marschner_luts(p, res); rm, rn = po.ref_marschner_luts(p, res)
        assert_bit_equal(m, rm, "M"); assert_bit_equal(n, rn, "N")


def test_known_answers_from_the_shader_text():
    m, n = po.marschner_luts(po.marschner_params(), 128)
    assert (n[..., 3] == 1.0).all() and np.isfinite(m).all() and np.isfinite(n).all()
    assert n[..., :3].min() >= 0.0 and n[..., :3].max() <= 1.0            # Np ends in min(L, 1)
    # theta_i == theta_r  ->  theta_d = 0  ->  cos theta_d = 1
    assert (np.diagonal(m[..., 3]) == 1.0).all()
    # M is symmetric in (theta_i, theta_r) up to the sign of theta_d: lobes depend on theta_h only
    assert np.array_equal(m[..., :3], m[..., :3].transpose(1, 0, 2))
    # M_R peaks where degrees(theta_h) == ar: gaussian(br, 0) = 1 / (sqrt(2 pi) br)
    assert abs(m[..., 0].max() - 1.0 / (2.5066282 * 5.0)) < 2e-4
    h = po.float_to_half(np.array([1.0, -2.0, 65504.0, 1e-8, np.inf], np.float32))
    assert list(h) == [0x3C00, 0xC000, 0x7BFF, 0x0000, 0x7C00]


def test_abi_argument_checks(lib):
    from barbu_b200 import marschner
    l = marschner._lib()
    p = bb.BhMarschnerParams()
    l.bh_marschner_default_params(C.byref(p))
    assert (p.eta, p.ar, p.br) == (np.float32(1.55), -5.0, 5.0) and p.azimuthal_width == 10.0
    assert l.bh_marschner_generate(None, 128, 0, None, None, None, None) == hair.BH_ERR_INVALID
    assert l.bh_marschner_generate(C.byref(p), 0, 0, None, None, None, None) == hair.BH_ERR_INVALID


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_luts_match_the_oracle(case):
    g = golden("marschner")
    p = g[case + "_params"]
    om, on = po.marschner_luts(p, 128)
    m16, n16, m32, n32 = bb.generate_luts(bb.BhMarschnerParams(*[float(x) for x in p]), 128, full=True)
    # the device's own half store is the round-to-nearest-even conversion of its fp32 texels
    assert_bit_equal(m16.view(np.uint16), po.float_to_half(m32), "M half store")
    assert_bit_equal(n16.view(np.uint16), po.float_to_half(n32), "N half store")
    # M: smooth everywhere -> every texel within 1e-5 relative (+1e-7 absolute for the underflowing gaussian tails)
    assert np.all(np.abs(m32 - om) <= 1e-5 * np.abs(om) + 1e-7)
    # N: the cubic solver branches on D > 0 and |c| < eps and the Fresnel term on sin^2 > 1; a texel sitting on such a
    # boundary may take the other branch with a different libm. Stated tolerance: >= 99.5 % of the texels within 1e-4
    # relative (+1e-6), every texel inside [0, 1], and the stored half-floats equal to the reference's golden texture in
    # >= 99 % of the texels and within 1 half-ulp in >= 99.5 %.
    close = np.abs(n32 - on) <= 1e-4 * np.abs(on) + 1e-6
    assert close.all(axis=-1).mean() >= 0.995, f"only {close.all(axis=-1).mean():.4f} of the N texels within tolerance"
    assert np.isfinite(n32).all() and n32.min() >= 0.0 and n32.max() <= 1.0
    for got, want, what in ((m16, g[case + "_m16"], "M"), (n16, g[case + "_n16"], "N")):
        d = np.abs(got.view(np.uint16).astype(np.int32) - want.view(np.uint16).astype(np.int32))
        assert (d == 0).all(axis=-1).mean() >= 0.99 and (d <= 1).all(axis=-1).mean() >= 0.995, \
            f"{what}: equal {(d == 0).all(axis=-1).mean():.4f}, within 1 half-ulp {(d <= 1).all(axis=-1).mean():.4f}"


@pytest.mark.gpu
def test_marschner_module_mirror_regenerates_only_on_change():
    mod = bb.Marschner()
    mod.init()
    mod.update()
    assert mod.generations == 1 and mod.lut_m.shape == (128, 128, 4) and mod.lut_m.dtype == np.float16
    first = mod.lut_n.copy()
    mod.update()
    assert mod.generations == 1                                            # marschner.cc:28: parameters unchanged
    mod.params.shading.absorption = 0.5
    mod.update()
    assert mod.generations == 2 and not np.array_equal(first, mod.lut_n)
    mod.update(True)
    assert mod.generations == 3
    with pytest.raises(bb.BarbuHairError):
        bb.generate_luts(bb.BhMarschnerParams(), 128, device=99)
