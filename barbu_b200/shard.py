"""Strand sharding across the GPUs of one node.

Strands never interact, so rank g of G owns the contiguous global strand range
[g*S/G, (g+1)*S/G) of the scalp and steps it with no per-step communication. The only exchange the
path knows is optional: an all-gather of the position plane so that the render GPU holds buffer 0 of
every shard (one `ncclAllGather` over NVLink; `gloo` on CPU tensors for the host-logic tests).

Host logic only — the simulation itself is `HairSim` (C ABI, CUDA). torch is imported lazily and is
used for process-group plumbing, never for compute.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(nstrands: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of rank's contiguous strand range; ranges tile [0, nstrands) and differ by at most one strand."""
    if world < 1 or not 0 <= rank < world or nstrands < 0:
        raise ValueError("need 0 <= rank < world and nstrands >= 0")
    first = (nstrands * rank) // world
    last = (nstrands * (rank + 1)) // world
    return first, last - first


def shard_counts(nstrands: int, world: int) -> List[int]:
    return [shard_range(nstrands, world, r)[1] for r in range(world)]


def local_patch_indices(tri_indices: np.ndarray, first: int, count: int) -> Tuple[np.ndarray, np.ndarray]:
    """Scalp triangles whose three roots all lie in [first, first+count), re-based to the shard, and the global ids
    of the triangles that straddle a seam (those only matter to rendering, which sees the gathered plane)."""
    tri = np.asarray(tri_indices, np.int64).reshape(-1, 3)
    inside = ((tri >= first) & (tri < first + count)).all(axis=1)
    any_in = ((tri >= first) & (tri < first + count)).any(axis=1)
    return (tri[inside] - first).astype(np.int32), np.nonzero(any_in & ~inside)[0]


def allgather_plane(local, counts_vertices: List[int], group=None):
    """All-gather one SoA plane ((V_local, 4) float32 torch tensor, CPU or CUDA) into the global (V, 4) plane in strand
    order: one all_gather_into_tensor (a single ncclAllGather). Ragged shards (strand count not a multiple of the
    world size: counts differ by one strand) are padded to the largest shard and compacted afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if len(counts_vertices) != world or local.shape[0] != counts_vertices[dist.get_rank(group)]:
        raise ValueError("counts_vertices must list every rank's vertex count, local must match this rank's")
    total = int(sum(counts_vertices))
    out = torch.empty((total, 4), dtype=local.dtype, device=local.device)
    if len(set(counts_vertices)) == 1:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        mx = int(max(counts_vertices))
        padded = torch.zeros((mx, 4), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        gathered = torch.empty((world * mx, 4), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        off = 0
        for r, c in enumerate(counts_vertices):
            out[off:off + c] = gathered[r * mx:r * mx + c]
            off += c
    return out


class _CudaPlane:
    """__cuda_array_interface__ view of a device plane owned by a bh_sim (no copy)."""

    def __init__(self, ptr: int, nvertices: int):
        self.__cuda_array_interface__ = {"shape": (nvertices, 4), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


def plane_tensor(sim, plane: int):
    """Zero-copy torch CUDA tensor (V, 4) over plane `plane` of the sim's buffer 0 (for NCCL / CUDA consumers)."""
    import torch
    ptr, _ = sim.device_plane(plane)
    return torch.as_tensor(_CudaPlane(ptr, sim.nvertices), device=torch.device("cuda", sim.device))


class HairShard:
    """This rank's shard of a global sphere-scalp hair system: HairSim over [first, first+count)."""

    def __init__(self, rows: int, cols: int, nverts: int, world: int, rank: int, device: int, seed: int = 1234,
                 maxlength: float = 0.5, order: int = 1, **params):
        """`order`: strand order of the scalp (hair.BH_SCALP_*). The default, column-major, makes every contiguous shard a
        longitude wedge with the whole latitude range — the same share of strands lying on the collider on every rank. With
        row-major order rank = latitude band, and the band over the pole (all strands in contact, the costliest kind of
        warp-step) sets the frame time of the whole job (measured: profiles/r02_straggler.txt)."""
        from . import hair
        self.rows, self.cols, self.nverts, self.world, self.rank, self.order = rows, cols, nverts, world, rank, order
        self.first, self.count = shard_range(rows * cols, world, rank)
        self.sim = hair.HairSim(self.count, nverts, device=device)
        if params:
            self.sim.configure(**params)
        self.sim.init_sphere_scalp(rows, cols, self.first, hair.random_values(seed, self.first, self.count), maxlength, order)

    def step(self, dt: float, substeps: int = 1):
        self.sim.step(dt, substeps)

    def gather_positions(self, group=None):
        """Global position plane on every rank (optional; off the per-step path)."""
        self.sim.synchronize()
        counts = [c * self.nverts for c in shard_counts(self.rows * self.cols, self.world)]
        return allgather_plane(plane_tensor(self.sim, 0), counts, group)

    def close(self):
        self.sim.close()
