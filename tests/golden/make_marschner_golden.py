#!/usr/bin/env python3
"""Golden Marschner lookup tables from the reference shader SOURCES (oracle/_ref/libbarbu_ref_marschner.so, built by
oracle/Makefile from /root/reference/src/shaders/hair/marschner/*.glsl). Run in the container that has the reference
tree; the GPU box only reads tests/golden/marschner.npz.   Usage: python tests/golden/make_marschner_golden.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle as po

assert po.build_ref() and po.ref_marschner_available(), "needs /root/reference"
CASES = {"default": dict(), "wet": dict(eta=1.33, absorption=0.6, ar=-8.0, br=9.0), "dense": dict(eta=2.4, absorption=0.05)}
out = {}
for name, kw in CASES.items():
    p = po.marschner_params(**kw)
    m, n = po.ref_marschner_luts(p, 128)
    out[name + "_params"] = p
    out[name + "_m16"] = m.astype(np.float16)      # numpy's float32 -> float16 is round-to-nearest-even
    out[name + "_n16"] = n.astype(np.float16)
    # fp32 texels of a 32 x 32 table: bit-exact pin of the oracle restatement at full precision
    m, n = po.ref_marschner_luts(p, 32)
    out[name + "_m32_res32"], out[name + "_n32_res32"] = m, n
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "marschner.npz"), **out)
print({k: v.shape for k, v in out.items()})
