#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ by RUNNING THE REFERENCE SOURCES (oracle/_ref).

Run in the build container only (it needs /root/reference):  python tests/golden/make_golden.py
oracle/_ref = the reference's cs_simulation.glsl, Hair::init_simulation and Hair::init_mesh compiled
as C++ against the reference's vendored GLM (see oracle/make_ref.py, oracle/Makefile). The fixtures
let the GPU box — which has no /root/reference — check the oracle and the CUDA path against what
the reference code itself computes.

Each hair_N<n>_s<100*scale>.npz holds, for a 8x16 (=128 strands) sphere scalp and srand(1234):
  root_pos, root_nrm, tri           inputs
  pos0, vel0, tan0                  Hair::init_simulation output (3 planes)
  patch                             Hair::init_mesh element buffer
  dt, scale, sphere                 uniforms of the runs below
  pos1, vel1 / pos10, vel10         state after 1 / 10 Hair::update calls of the reference shader
  posw1, velw1                      after 1 more update starting from a "warm" state (pos10, vel10)
simplex.npz: glm::simplex(vec2) at 4096 seeded points.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ROWS, COLS, SEED = 8, 16, 1234
DT = np.float32(1.0) / np.float32(90.0)
SPHERE = (0.0, 0.0, 0.0, 0.98)


def main():
    assert po.build_ref(), "reference tree missing"
    root_pos, root_nrm, tri = po.sphere_scalp(ROWS, COLS)
    S = ROWS * COLS
    for N, scale in ((2, 1.0), (3, 1.45), (4, 1.45), (8, 1.45), (16, 1.0), (32, 1.0), (32, 1.45), (64, 1.0)):
        pos0, vel0, tan0 = po.ref_init_simulation(root_pos, root_nrm, SEED, N)
        patch = po.ref_patch_indices(tri, S, N)
        pos, vel = pos0.copy(), vel0.copy()
        snaps = {}
        for it in range(1, 11):
            po.ref_update(pos, vel, S, N, float(DT), scale, SPHERE)
            if it in (1, 10):
                snaps[f"pos{it}"], snaps[f"vel{it}"] = pos.copy(), vel.copy()
        po.ref_update(pos, vel, S, N, float(DT), scale, SPHERE)
        tag = f"hair_N{N}_s{int(round(scale * 100)):03d}"
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), root_pos=root_pos, root_nrm=root_nrm, tri=tri,
                            pos0=pos0, vel0=vel0, tan0=tan0, patch=patch, dt=DT, scale=np.float32(scale),
                            sphere=np.array(SPHERE, np.float32), posw1=pos, velw1=vel, nverts=N, seed=SEED, **snaps)
        print(tag, "ok", "nan" if np.isnan(pos).any() else "")
    rng = np.random.default_rng(7)
    pts = (rng.random((4096, 2), dtype=np.float32) * 40.0 - 20.0).astype(np.float32)
    val = np.array([po.ref_simplex2(float(x), float(y)) for x, y in pts], np.float32)
    np.savez_compressed(os.path.join(OUT, "simplex.npz"), pts=pts, val=val)
    print("simplex ok")


if __name__ == "__main__":
    main()
