#!/usr/bin/env python3
"""Known answers for the reference's own scalp asset (assets/models/InfiniteScan/Head_scalp.obj, read WHERE IT LIES
under /root/reference; nothing of it is copied): vertex / face counts and FNV-1a checksums of the arrays the oracle's
restatement of the reference loader produces, plus the hair state generated from them (N = 4, the reference default).
Run in the build container only:  python tests/golden/make_scalp_golden.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pyoracle as po

ASSET = "/root/reference/assets/models/InfiniteScan/Head_scalp.obj"
P, Nn, T = po.obj_scalp(ASSET)
S = P.shape[0]
rv = po.random_values(1234, S)
pos, vel = po.init_strands(P, Nn, rv, 4)
tan = po.init_tangents(Nn, 4)
patch = po.patch_indices(T, 4)
par = po.default_params(dt=float(np.float32(1.0) / np.float32(90.0)), scale=1.45, sphere=(0.0, 2.0, 0.9, 0.25))
p10, v10 = pos.copy(), vel.copy()
for _ in range(10):
    po.step(p10, v10, S, 4, par)
out = {"asset": "assets/models/InfiniteScan/Head_scalp.obj", "nvertices": int(S), "nfaces": int(T.shape[0]),
       "fnv_positions": po.fnv1a64(P), "fnv_normals": po.fnv1a64(Nn), "fnv_triangles": po.fnv1a64(T),
       "first_vertex": [float(x) for x in P[0]], "first_triangles": T[:3].tolist(),
       "N": 4, "seed": 1234, "fnv_pos0": po.fnv1a64(pos), "fnv_tan0": po.fnv1a64(tan), "fnv_patch": po.fnv1a64(patch),
       "sphere": [0.0, 2.0, 0.9, 0.25], "scale": 1.45, "fnv_pos10": po.fnv1a64(p10), "fnv_vel10": po.fnv1a64(v10)}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "head_scalp.json"), "w"), indent=1)
print(out)
