#!/usr/bin/env python3
"""Fixtures for the scalp INPUT rows (SURVEY.md §8 a-ext; VERDICT r1 "Inputs"):
  tests/golden/skinned_scalp.gltf   a small synthetic skinned scalp (sphere cap, 4-joint chain along y, JOINTS_0 / WEIGHTS_0,
                                    uint16 indices whose first-appearance order differs from the accessor order, a mesh node
                                    with a `matrix`), buffers embedded as base64 — the reference's own Head.glb is a git-LFS pointer.
  tests/golden/inputs.npz           known answers made by the REFERENCE's host code compiled here (oracle/_ref):
      recalc_*   RawMeshData::recalculateNormals (src/utils/raw_mesh_file.cc:11-50) on a triangle soup with zero-area faces;
      dq_*       SkeletonController::generate_skinning_datas (src/fx/animation/skeleton_controller.cc:248-265) on posed joints
                 covering all four branches of glm's dualquat_cast;
      gltf_*     what the file must load as (vertex order = first appearance in the index list), written by plain loops here.
Run in the build container only (needs /root/reference for oracle/_ref):  python tests/golden/make_inputs_golden.py"""
import base64, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from oracle import pyoracle as po

rng = np.random.default_rng(2024)

# ---- recalculated normals ----------------------------------------------------------------------------------------------
nv, nf = 90, 220
rpos = rng.standard_normal((nv, 3)).astype(np.float32)
rtri = rng.integers(0, nv - 2, (nf, 3)).astype(np.int32)
rtri[3] = [5, 5, 8]; rpos[20] = rpos[21]; rtri[4] = [20, 21, 22]
rnrm, ridx = po.ref_recalc_normals(rpos, rtri)
assert ridx.tolist() == list(range(3 * nf))

# ---- dual-quaternion palette ---------------------------------------------------------------------------------------------
def quat_matrix(q, t):
    x, y, z, w = q / np.linalg.norm(q)
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = t
    return M
quats = [np.array(q, float) for q in ([0, 0, 0, 1], [1, 0, 0, 0.05], [0, 1, 0, 0.05], [0, 0, 1, 0.05], [0.3, -0.2, 0.1, 0.9], [0.7, 0.7, 0.1, -0.1])]
quats += [rng.standard_normal(4) for _ in range(26)]
G = np.stack([quat_matrix(q, rng.standard_normal(3)) for q in quats]); B = np.stack([quat_matrix(rng.standard_normal(4), rng.standard_normal(3)) for _ in quats])
B[:6] = np.eye(4)                                                            # the first six: pose alone decides the branch
Gc = np.ascontiguousarray(G.transpose(0, 2, 1), np.float32).reshape(-1, 16); Bc = np.ascontiguousarray(B.transpose(0, 2, 1), np.float32).reshape(-1, 16)
dq = po.ref_dq_palette_from_matrices(Gc, Bc)

# ---- the glTF file -------------------------------------------------------------------------------------------------------
rows, cols, J = 7, 12, 4
th = np.linspace(0.15, 1.2, rows); ph = np.linspace(0, 2 * np.pi, cols, endpoint=False)
nrm = np.stack([np.outer(np.sin(th), np.cos(ph)), np.repeat(np.cos(th)[:, None], cols, 1), np.outer(np.sin(th), np.sin(ph))], -1).reshape(-1, 3).astype(np.float32)
pos = (nrm * np.float32(0.9)).astype(np.float32)
jy = np.linspace(0.2, 0.9, J).astype(np.float32)                             # joint heights
d = np.abs(pos[:, 1:2] - jy[None, :])
order = np.argsort(d, axis=1)
w = 1.0 / (np.take_along_axis(d, order, 1) + 0.05); w = (w / w.sum(1, keepdims=True)).astype(np.float32)
joints = order.astype(np.uint8); weights = w
faces = []
for r in range(rows - 1):
    for c in range(cols):
        c1 = (c + 1) % cols
        a, b, e, f = r * cols + c, (r + 1) * cols + c, (r + 1) * cols + c1, r * cols + c1
        faces += [[a, b, f], [f, b, e]]
faces = np.array(faces[::-1], np.uint16)                                     # last ring first: first-appearance order != accessor order
ibm = np.tile(np.eye(4, dtype=np.float32), (J, 1, 1))
for j in range(J): ibm[j, 3, 1] = -jy[j]                                     # column-major rows: [3] is the translation column; inverse of T(0, jy, 0)
chunks, views, accessors = [], [], []
def add(arr, ctype, atype, target=None, minmax=False):
    raw = np.ascontiguousarray(arr).tobytes(); off = sum(len(c) for c in chunks)
    pad = (-len(raw)) % 4
    chunks.append(raw + b"\0" * pad)
    v = {"buffer": 0, "byteOffset": off, "byteLength": len(raw)}
    if target: v["target"] = target
    views.append(v)
    a = {"bufferView": len(views) - 1, "componentType": ctype, "count": int(arr.shape[0]) if atype != "SCALAR" else int(arr.size), "type": atype}
    if minmax: a["min"], a["max"] = [float(x) for x in arr.min(0)], [float(x) for x in arr.max(0)]
    accessors.append(a)
    return len(accessors) - 1
a_pos = add(pos, 5126, "VEC3", 34962, True); a_nrm = add(nrm, 5126, "VEC3", 34962)
a_j = add(joints, 5121, "VEC4", 34962); a_w = add(weights, 5126, "VEC4", 34962)
a_idx = add(faces.reshape(-1), 5123, "SCALAR", 34963); a_ibm = add(ibm.reshape(J, 16), 5126, "MAT4")
blob = b"".join(chunks)
mesh_matrix = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.125, -0.25, 0.5, 1]     # column-major: a translation (exact in binary32)
nodes = [{"name": "scalp", "mesh": 0, "skin": 0, "matrix": mesh_matrix}]
for j in range(J):
    n = {"name": f"joint{j}", "translation": [0.0, float(jy[j] - (jy[j - 1] if j else 0.0)), 0.0]}
    if j + 1 < J: n["children"] = [j + 2]
    nodes.append(n)
doc = {"asset": {"version": "2.0", "generator": "barbu_b200 tests/golden/make_inputs_golden.py"}, "scene": 0, "scenes": [{"nodes": [0, 1]}], "nodes": nodes,
       "meshes": [{"name": "scalp", "primitives": [{"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "JOINTS_0": a_j, "WEIGHTS_0": a_w}, "indices": a_idx, "mode": 4}]}],
       "skins": [{"joints": list(range(1, J + 1)), "inverseBindMatrices": a_ibm, "skeleton": 1}],
       "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
       "bufferViews": views, "accessors": accessors}
json.dump(doc, open(os.path.join(HERE, "skinned_scalp.gltf"), "w"), indent=1)
# what it must load as: world transform (a translation: v + t with the zero products of GLM's mat * vec), first-appearance order
t = np.array(mesh_matrix[12:15], np.float32)
seen, new_of = [], {}
for i in faces.reshape(-1).tolist():
    if i not in new_of: new_of[i] = len(seen); seen.append(i)
gp = np.stack([((np.float32(1) * pos[i] + np.float32(0) * pos[i]) + (np.float32(0) * pos[i] + t * np.float32(1))) for i in seen]).astype(np.float32)
gn = np.stack([nrm[i] for i in seen]); gj = np.stack([joints[i] for i in seen]).astype(np.int32); gw = np.stack([weights[i] for i in seen])
gi = np.array([new_of[i] for i in faces.reshape(-1).tolist()], np.int32).reshape(-1, 3)
np.savez_compressed(os.path.join(HERE, "inputs.npz"), recalc_pos=rpos, recalc_tri=rtri, recalc_nrm=rnrm, dq_global=Gc, dq_inverse_bind=Bc, dq_palette=dq,
                    gltf_pos=gp, gltf_nrm=gn, gltf_joints=gj, gltf_weights=gw, gltf_indices=gi, gltf_inverse_bind=ibm.reshape(J, 16), gltf_joint_y=jy)
print("skinned_scalp.gltf", os.path.getsize(os.path.join(HERE, "skinned_scalp.gltf")), "bytes; inputs.npz", os.path.getsize(os.path.join(HERE, "inputs.npz")), "bytes;",
      len(seen), "vertices,", gi.shape[0], "faces; dq branches:", [("w" if (g[0] + g[5] + g[10]) > 0 else "xyz") for g in Gc[:6]])
