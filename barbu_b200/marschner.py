"""Python mirror of the reference's `Marschner` module (src/fx/marschner.h:21-102) over the C ABI.

`init / update(force) / generate` and `Parameters.shading` keep the reference's names and behaviour: `update` regenerates
the two lookup tables only when the shading parameters changed (marschner.cc:27-32). The tables are produced by one CUDA
kernel (barbu_b200/csrc/hair_marschner.cu); `bindLUTs` stays with the reference's GL code, which uploads `lut_m` /
`lut_n` (RGBA16F, 128 x 128) as the two textures of marschner.cc:17-18. No CPU path."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace

import numpy as np

from .hair import _check, _ptr, load_library


class BhMarschnerParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("eta", "absorption", "eccentricity", "ar", "br", "glint_scale", "azimuthal_width",
                                         "delta_caustic", "delta_hm")]


def _lib():
    lib = load_library()
    if not getattr(lib, "_marschner_sig", False):
        lib.bh_marschner_default_params.argtypes, lib.bh_marschner_default_params.restype = [C.POINTER(BhMarschnerParams)], None
        lib.bh_marschner_generate.argtypes = [C.POINTER(BhMarschnerParams), C.c_int, C.c_int] + [C.c_void_p] * 4
        lib.bh_marschner_generate.restype = C.c_int
        lib._marschner_sig = True
    return lib


def generate_luts(params: BhMarschnerParams, resolution: int = 128, device: int = 0, full: bool = False):
    """(M, N) as (res, res, 4) float16 arrays — and the fp32 texels before the half store when `full`."""
    shape = (resolution, resolution, 4)
    m16, n16 = np.empty(shape, np.float16), np.empty(shape, np.float16)
    m32, n32 = (np.empty(shape, np.float32), np.empty(shape, np.float32)) if full else (None, None)
    _check(_lib().bh_marschner_generate(C.byref(params), resolution, device, _ptr(m16), _ptr(n16), _ptr(m32), _ptr(n32)))
    return (m16, n16, m32, n32) if full else (m16, n16)


class Marschner:
    kTextureResolution = 128             # marschner.h:28
    kNumLUTs = 2

    @dataclass
    class ShadingParameters:             # marschner.h:38-52
        eta: float = 1.55
        absorption: float = 0.20
        eccentricity: float = 0.85
        ar: float = -5.0
        br: float = 5.0
        glintScale: float = 0.5
        azimuthalWidth: float = 10.0
        deltaCaustic: float = 0.2
        deltaHm: float = 0.5

    @dataclass
    class Parameters:
        shading: "Marschner.ShadingParameters" = field(default_factory=lambda: Marschner.ShadingParameters())

    def __init__(self, device: int = 0):
        self.device = device
        self.params = Marschner.Parameters()
        self._previous = None
        self.lut_m = self.lut_n = None
        self.generations = 0

    def init(self):
        _lib()

    def update(self, bForceUpdate: bool = False):                            # marschner.cc:27-32
        if bForceUpdate or self._previous != self.params.shading:
            self.generate()
        self._previous = replace(self.params.shading)

    def generate(self):                                                      # marschner.cc:35-69
        s = self.params.shading
        p = BhMarschnerParams(s.eta, s.absorption, s.eccentricity, s.ar, s.br, s.glintScale, s.azimuthalWidth, s.deltaCaustic, s.deltaHm)
        self.lut_m, self.lut_n = generate_luts(p, self.kTextureResolution, self.device)
        self.generations += 1
