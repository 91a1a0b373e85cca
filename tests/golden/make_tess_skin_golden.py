#!/usr/bin/env python3
"""Generate tests/golden/tess_skin.npz by RUNNING THE REFERENCE SHADER SOURCES (oracle/_ref/libbarbu_ref_tess_skin.so:
the four 02_tess_stream stages + inc_maths.glsl functions, and apply_skinning / skinning_DQBS of inc_skinning.glsl, compiled
over the reference's vendored GLM — see oracle/make_ref.py, oracle/ref_tess_skin_harness.cpp).

Run in the build container only (it needs /root/reference):  python tests/golden/make_tess_skin_golden.py
  tess_*   : an 8x16 sphere scalp after 10 reference updates (hair_N4_s145 / hair_N16_s100 fixtures), streamed with the
             reference defaults (3 instances x 2 lines x 16 sub-segments, hair.h:33-35) and with (2, 5, 7); the random table
             is the oracle's seeded pairs for `seed` (its content is std::random_device noise in the reference).
  skin_*   : 256 random roots, 11-joint dual-quaternion palette with antipodal joints, weights around the Epsilon() early-out.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NPATCH = 90


def main():
    assert po.build_ref() and po.ref_tess_skin_available(), "reference tree missing"
    out = {}
    for name, (ninst, nlines, nsub), scale, seed in (("hair_N4_s145", (3, 2, 16), 1.45, 99), ("hair_N16_s100", (2, 5, 7), 1.0, 5)):
        g = np.load(os.path.join(OUT, name + ".npz"))
        N = int(g["nverts"])
        table = po.tess_random_table(seed)
        patch = g["patch"][:6 * NPATCH]                                   # a small fixture: the first NPATCH patches
        out[f"tess_{name}"] = po.ref_tess_stream(g["pos10"], g["tan0"], patch, N, scale, ninst, nlines, nsub, table)
        out[f"tess_{name}_args"] = np.array([ninst, nlines, nsub, seed], np.int64)
        out[f"tess_{name}_scale"] = np.float32(scale)
    rng = np.random.default_rng(2024)
    S, J = 256, 11
    q = rng.standard_normal((J, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.standard_normal((J, 3)) * 0.5
    x, y, z, w = q.T
    dual = 0.5 * np.stack([t[:, 0] * w + t[:, 1] * z - t[:, 2] * y, -t[:, 0] * z + t[:, 1] * w + t[:, 2] * x,
                           t[:, 0] * y - t[:, 1] * x + t[:, 2] * w, -t[:, 0] * x - t[:, 1] * y - t[:, 2] * z], axis=1)
    dq = np.concatenate([q, dual], axis=1)
    dq[rng.random(J) < 0.5] *= -1.0
    dq = dq.astype(np.float32)
    pos = (rng.standard_normal((S, 3)) * 2.0).astype(np.float32)
    nrm = rng.standard_normal((S, 3)).astype(np.float32)
    joints = rng.integers(0, J, (S, 4)).astype(np.int32)
    wts = rng.dirichlet([1.0, 1.0, 1.0, 1.0], S)[:, :3].astype(np.float32)
    wts[0] = (0.0, 0.5, 0.5); wts[1] = (1e-6, 0.5, 0.4); wts[2] = (1.0000001e-6, 0.0, 0.0); wts[3] = (1.0, 0.0, 0.0)
    wts[4] = (0.5, -0.25, 0.5)
    sp, sn = po.ref_skin_dq(pos, nrm, joints, wts, dq)
    out.update(skin_pos=pos, skin_nrm=nrm, skin_joints=joints, skin_weights=wts, skin_dq=dq, skin_out_pos=sp, skin_out_nrm=sn)
    np.savez_compressed(os.path.join(OUT, "tess_skin.npz"), **out)
    print("tess_skin.npz ok:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
