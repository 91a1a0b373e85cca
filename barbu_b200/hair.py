"""Python mirror of the reference's `Hair` module interface, over the C ABI of libbarbu_hair.so.

Everything here is plumbing: argument marshalling into include/barbu_hair.h entry points. The
simulation itself only exists as CUDA kernels (barbu_b200/csrc); there is no CPU path, and loading
fails loudly when the shared library has not been built.

Reference interface mirrored (src/fx/hair.h:56-78):
    Hair::init / deinit / setup / update / set_bounding_sphere / initialized
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _build

BH_OK, BH_ERR_INVALID, BH_ERR_CUDA, BH_ERR_NOT_INITIALIZED, BH_ERR_UNSUPPORTED, BH_ERR_OVERFLOW = range(6)
BH_MATH_EXACT, BH_MATH_FAST = 0, 1
BH_POLICY_THROUGHPUT, BH_POLICY_LATENCY, BH_POLICY_AUTO = 0, 1, 2
BH_SCALP_ROW_MAJOR, BH_SCALP_COLUMN_MAJOR = 0, 1
BH_PLANE_POSITION, BH_PLANE_VELOCITY, BH_PLANE_TANGENT = 0, 1, 2
BH_MAX_CAPSULES = 8

# Every symbol include/barbu_hair.h declares (tests check the library exports all of them).
ABI_SYMBOLS = (
    "bh_create", "bh_destroy", "bh_set_stream", "bh_reset_stream", "bh_synchronize", "bh_default_params", "bh_set_params",
    "bh_get_params", "bh_set_bounding_sphere", "bh_upload", "bh_download", "bh_device_plane",
    "bh_random_values", "bh_init_strands", "bh_init_sphere_scalp", "bh_init_tangents_host",
    "bh_sphere_scalp_triangles", "bh_init_sphere_scalp_ordered", "bh_sphere_scalp_triangles_ordered", "bh_load_obj_scalp", "bh_free", "bh_build_patch_indices", "bh_step", "bh_set_substep_fusion", "bh_group_set_substep_fusion", "bh_set_step_policy", "bh_step_host", "bh_step_readback", "bh_host_alloc",
    "bh_host_free", "bh_tess_set_patches", "bh_tess_stream_count", "bh_tess_stream", "bh_tess_device_buffer", "bh_launch_count", "bh_step_kernel_kind", "bh_selftest_math", "bh_set_skin", "bh_skin_roots", "bh_dq_palette_from_matrices", "bh_register_gl_buffer",
    "bh_unregister_gl_buffer", "bh_register_device_buffer", "bh_unregister_device_buffer", "bh_buffer_map_stats", "bh_last_error", "bh_version",
    "bh_state_checksum", "bh_save_state", "bh_peek_state", "bh_load_state",
    "bh_marschner_default_params", "bh_marschner_generate",
    "bh_group_create", "bh_group_destroy", "bh_group_size", "bh_group_shard", "bh_group_shard_range", "bh_group_set_params",
    "bh_group_set_bounding_sphere", "bh_group_init_sphere_scalp", "bh_group_init_strands", "bh_group_upload", "bh_group_download",
    "bh_group_step", "bh_group_synchronize", "bh_group_step_timed", "bh_group_launch_count", "bh_group_gather_plane",
    "bh_group_register_gl_buffer", "bh_group_unregister_gl_buffer", "bh_group_gather_to_gl",
)


class BhCapsule(C.Structure):
    _fields_ = [("a", C.c_float * 3), ("b", C.c_float * 3), ("radius", C.c_float)]


class BhTessParams(C.Structure):
    _fields_ = [("ninstances", C.c_int), ("nlines", C.c_int), ("nsubsegments", C.c_int), ("seed", C.c_uint)]


class BhParams(C.Structure):
    """struct bh_params of include/barbu_hair.h."""
    _fields_ = [
        ("scale", C.c_float), ("sphere", C.c_float * 4), ("iterations", C.c_int),
        ("gravity", C.c_float * 3), ("force_coeff", C.c_float), ("damp", C.c_float), ("math", C.c_int),
        ("wind", C.c_float * 3), ("drag", C.c_float), ("ncapsules", C.c_int),
        ("capsules", BhCapsule * BH_MAX_CAPSULES),
    ]


class BhStateInfo(C.Structure):
    """struct bh_state_info of include/barbu_hair.h (header fields of a BARBUHS1 state file)."""
    _fields_ = [
        ("nstrands", C.c_int64), ("nverts", C.c_int), ("plane_mask", C.c_uint), ("total_strands", C.c_int64),
        ("first_strand", C.c_int64), ("frame", C.c_int64), ("dt", C.c_float), ("seed", C.c_uint),
        ("checksum", C.c_uint64 * 2), ("params", BhParams),
    ]


class BarbuHairError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"barbu_hair error {code}: {message}")
        self.code = code


_lib: Optional[C.CDLL] = None


def load_library(build_if_missing: bool = False) -> C.CDLL:
    """dlopen barbu_b200/lib/libbarbu_hair.so. Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if build_if_missing:
            _build.build()
        else:
            raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(the hair simulation has no CPU fallback)")
    lib = C.CDLL(path)
    vp, i64, f32 = C.c_void_p, C.c_int64, C.c_float
    sig = {
        "bh_create": ([C.POINTER(vp), i64, C.c_int, C.c_int], C.c_int),
        "bh_destroy": ([vp], C.c_int),
        "bh_set_stream": ([vp, vp], C.c_int),
        "bh_reset_stream": ([vp], C.c_int),
        "bh_synchronize": ([vp], C.c_int),
        "bh_default_params": ([C.POINTER(BhParams)], None),
        "bh_set_params": ([vp, C.POINTER(BhParams)], C.c_int),
        "bh_get_params": ([vp, C.POINTER(BhParams)], C.c_int),
        "bh_set_bounding_sphere": ([vp, C.POINTER(f32)], C.c_int),
        "bh_upload": ([vp, vp, vp, vp], C.c_int),
        "bh_download": ([vp, vp, vp, vp], C.c_int),
        "bh_device_plane": ([vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)], C.c_int),
        "bh_random_values": ([C.c_uint, i64, i64, vp], C.c_int),
        "bh_init_strands": ([vp, vp, vp, vp, f32], C.c_int),
        "bh_init_sphere_scalp": ([vp, C.c_int, C.c_int, i64, vp, f32], C.c_int),
        "bh_init_tangents_host": ([vp, i64, i64, i64, C.c_int, f32, vp], C.c_int),
        "bh_sphere_scalp_triangles": ([C.c_int, C.c_int, vp], C.c_int),
        "bh_init_sphere_scalp_ordered": ([vp, C.c_int, C.c_int, C.c_int, i64, vp, f32], C.c_int),
        "bh_sphere_scalp_triangles_ordered": ([C.c_int, C.c_int, C.c_int, vp], C.c_int),
        "bh_build_patch_indices": ([vp, i64, C.c_int, vp, C.c_int], C.c_int),
        "bh_load_obj_scalp": ([C.c_char_p, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64)], C.c_int),
        "bh_free": ([vp], None),
        "bh_step": ([vp, f32, C.c_int], C.c_int),
        "bh_step_host": ([vp, f32, C.c_int, vp, vp], C.c_int),
        "bh_set_substep_fusion": ([vp, C.c_int], C.c_int),
        "bh_step_readback": ([vp, f32, C.c_int, vp], C.c_int),
        "bh_dq_palette_from_matrices": ([vp, vp, C.c_int, vp], C.c_int),
        "bh_register_device_buffer": ([vp, vp, C.c_uint64], C.c_int),
        "bh_unregister_device_buffer": ([vp], C.c_int),
        "bh_buffer_map_stats": ([vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)], C.c_int),
        "bh_group_set_substep_fusion": ([vp, C.c_int], C.c_int),
        "bh_set_step_policy": ([vp, C.c_int], C.c_int),
        "bh_host_alloc": ([C.POINTER(vp), C.c_uint64], C.c_int),
        "bh_host_free": ([vp], C.c_int),
        "bh_tess_set_patches": ([vp, vp, i64], C.c_int),
        "bh_tess_stream_count": ([vp, C.POINTER(BhTessParams)], i64),
        "bh_tess_stream": ([vp, C.POINTER(BhTessParams), vp], C.c_int),
        "bh_tess_device_buffer": ([vp, C.POINTER(vp), C.POINTER(i64)], C.c_int),
        "bh_launch_count": ([vp], i64),
        "bh_step_kernel_kind": ([vp], C.c_int),
        "bh_selftest_math": ([C.c_int, C.POINTER(C.c_uint64)], C.c_int),
        "bh_set_skin": ([vp, vp, vp, vp], C.c_int),
        "bh_skin_roots": ([vp, vp, C.c_int], C.c_int),
        "bh_register_gl_buffer": ([vp, C.c_uint], C.c_int),
        "bh_unregister_gl_buffer": ([vp], C.c_int),
        "bh_state_checksum": ([vp, C.c_uint, i64, C.POINTER(C.c_uint64)], C.c_int),
        "bh_save_state": ([vp, C.c_char_p, C.POINTER(BhStateInfo)], C.c_int),
        "bh_peek_state": ([C.c_char_p, C.POINTER(BhStateInfo)], C.c_int),
        "bh_load_state": ([vp, C.c_char_p, C.POINTER(BhStateInfo)], C.c_int),
        "bh_group_create": ([C.POINTER(vp), C.POINTER(C.c_int), C.c_int, i64, C.c_int], C.c_int),
        "bh_group_destroy": ([vp], C.c_int),
        "bh_group_size": ([vp], C.c_int),
        "bh_group_shard": ([vp, C.c_int], vp),
        "bh_group_shard_range": ([vp, C.c_int, C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "bh_group_set_params": ([vp, C.POINTER(BhParams)], C.c_int),
        "bh_group_set_bounding_sphere": ([vp, C.POINTER(f32)], C.c_int),
        "bh_group_init_sphere_scalp": ([vp, C.c_int, C.c_int, C.c_int, C.c_uint, f32], C.c_int),
        "bh_group_init_strands": ([vp, vp, vp, vp, f32], C.c_int),
        "bh_group_upload": ([vp, vp, vp, vp], C.c_int),
        "bh_group_download": ([vp, vp, vp, vp], C.c_int),
        "bh_group_step": ([vp, f32, C.c_int], C.c_int),
        "bh_group_synchronize": ([vp], C.c_int),
        "bh_group_step_timed": ([vp, f32, C.c_int, C.c_int, C.POINTER(f32), C.POINTER(f32)], C.c_int),
        "bh_group_launch_count": ([vp], i64),
        "bh_group_gather_plane": ([vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(f32)], C.c_int),
        "bh_group_register_gl_buffer": ([vp, C.c_uint, C.c_int], C.c_int),
        "bh_group_unregister_gl_buffer": ([vp], C.c_int),
        "bh_group_gather_to_gl": ([vp, C.c_uint], C.c_int),
        "bh_last_error": ([], C.c_char_p),
        "bh_version": ([], C.c_char_p),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, restype
    _lib = lib
    return lib


def _apply_param_keywords(p, kw) -> None:
    for k, v in kw.items():
        cur = getattr(p, k)
        if isinstance(cur, C.Array) and not isinstance(v, C.Array):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)


def _check(rc: int) -> None:
    if rc != BH_OK:
        raise BarbuHairError(rc, load_library().bh_last_error().decode(errors="replace"))


def _f32(a, shape_last: Optional[int] = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and (a.ndim == 0 or a.shape[-1] != shape_last):
        raise ValueError(f"expected trailing dimension {shape_last}, got shape {a.shape}")
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params() -> BhParams:
    p = BhParams()
    load_library().bh_default_params(C.byref(p))
    return p


def random_values(seed: int, first: int, count: int) -> np.ndarray:
    """hair.cc:273-275 length jitter for global strands [first, first+count)."""
    out = np.empty(count, np.float32)
    _check(load_library().bh_random_values(seed, first, count, _ptr(out)))
    return out


def peek_state(path) -> BhStateInfo:
    """Header of a BARBUHS1 state file (bh_peek_state; host only)."""
    info = BhStateInfo()
    _check(load_library().bh_peek_state(os.fsencode(path), C.byref(info)))
    return info


def selftest_math(device: int = 0) -> int:
    """Mismatch count of the exhaustive 1/sqrt check (must be 0)."""
    bad = C.c_uint64()
    _check(load_library().bh_selftest_math(device, C.byref(bad)))
    return int(bad.value)


def sphere_scalp_triangles(rows: int, cols: int, order: int = 0) -> np.ndarray:
    """Triangle list of the synthetic sphere scalp; order: BH_SCALP_ROW_MAJOR (0) or BH_SCALP_COLUMN_MAJOR (1) vertex numbering."""
    tri = np.empty((2 * (rows - 1) * cols, 3), np.int32)
    _check(load_library().bh_sphere_scalp_triangles_ordered(rows, cols, order, _ptr(tri)))
    return tri


def build_patch_indices(tri_indices, nverts: int, device: int = 0) -> np.ndarray:
    """Hair::init_mesh element buffer (hair.cc:397-409), computed on the GPU."""
    tri = np.ascontiguousarray(tri_indices, dtype=np.int32).reshape(-1, 3)
    out = np.empty(6 * tri.shape[0] * max(nverts - 1, 0), np.int32)
    _check(load_library().bh_build_patch_indices(_ptr(tri), tri.shape[0], nverts, _ptr(out), device))
    return out


def init_tangents_host(root_nrm3, total: int, first: int, nverts: int, maxlength: float = 0.5) -> np.ndarray:
    nrm = _f32(root_nrm3, 3)
    out = np.empty((nrm.shape[0] * nverts, 4), np.float32)
    _check(load_library().bh_init_tangents_host(_ptr(nrm), total, first, nrm.shape[0], nverts, maxlength, _ptr(out)))
    return out


class PinnedBuffer:
    """Page-locked host array (cudaMallocHost) viewed as float32 numpy."""

    def __init__(self, nfloats: int):
        self._lib = load_library()
        self._ptr = C.c_void_p()
        _check(self._lib.bh_host_alloc(C.byref(self._ptr), nfloats * 4))
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_float)), shape=(nfloats,))

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.bh_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class HairSim:
    """Thin owner of one `bh_sim` handle (one strand shard on one GPU)."""

    def __init__(self, nstrands: int, nverts: int, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.nstrands, self.nverts, self.device = int(nstrands), int(nverts), int(device)
        _check(self._lib.bh_create(C.byref(self._h), self.nstrands, self.nverts, self.device))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._lib.bh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def nvertices(self) -> int:
        return self.nstrands * self.nverts

    # -- parameters --------------------------------------------------------------------------
    def get_params(self) -> BhParams:
        p = BhParams()
        _check(self._lib.bh_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, p: BhParams):
        _check(self._lib.bh_set_params(self._h, C.byref(p)))

    def configure(self, **kw):
        """Update named fields of bh_params (scale=..., math=..., sphere=(x,y,z,r), ...)."""
        p = self.get_params()
        _apply_param_keywords(p, kw)
        self.set_params(p)

    def set_step_policy(self, policy: int):
        """BH_POLICY_THROUGHPUT (default), BH_POLICY_LATENCY (wavefront kernel: small scalps), BH_POLICY_AUTO (by size)."""
        _check(self._lib.bh_set_step_policy(self._h, policy))

    def set_substep_fusion(self, enabled, always: bool = False):
        """step(dt, k > 1) as k passes of ONE launch (tiles re-read from L2 between substeps); bit-identical results.
        enabled: fuse where it pays (small shards keep their launches); always=True: wherever the shape allows it."""
        _check(self._lib.bh_set_substep_fusion(self._h, (2 if always else 1) if enabled else 0))

    def set_bounding_sphere(self, sphere: Sequence[float]):
        _check(self._lib.bh_set_bounding_sphere(self._h, (C.c_float * 4)(*sphere)))

    def set_stream(self, cuda_stream: int):
        """Launch on this cudaStream_t handle (0 = the CUDA default stream)."""
        _check(self._lib.bh_set_stream(self._h, C.c_void_p(cuda_stream or None)))

    def reset_stream(self):
        _check(self._lib.bh_reset_stream(self._h))

    def synchronize(self):
        _check(self._lib.bh_synchronize(self._h))

    # -- state -------------------------------------------------------------------------------
    def upload(self, pos4=None, vel4=None, tan4=None):
        arrs = [None if a is None else _f32(a).reshape(-1) for a in (pos4, vel4, tan4)]
        for a in arrs:
            if a is not None and a.size != 4 * self.nvertices:
                raise ValueError(f"plane must hold {self.nvertices} float4")
        _check(self._lib.bh_upload(self._h, *[_ptr(a) for a in arrs]))

    def download(self, pos=True, vel=True, tan=False):
        outs = [np.empty((self.nvertices, 4), np.float32) if want else None for want in (pos, vel, tan)]
        _check(self._lib.bh_download(self._h, *[_ptr(a) for a in outs]))
        return tuple(outs)

    def download_into(self, pos4=None, vel4=None, tan4=None):
        """bh_download into caller-owned float32 arrays (e.g. PinnedBuffer.array): no allocation, no staging copy."""
        for a in (pos4, vel4, tan4):
            if a is not None and (a.dtype != np.float32 or not a.flags.c_contiguous or a.size != 4 * self.nvertices):
                raise ValueError("download_into: need C-contiguous float32 arrays of 4 * nvertices elements")
        _check(self._lib.bh_download(self._h, _ptr(pos4), _ptr(vel4), _ptr(tan4)))

    def device_plane(self, plane: int):
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        _check(self._lib.bh_device_plane(self._h, plane, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def init_strands(self, root_pos3, root_nrm3, random_value, maxlength: float = 0.5):
        p, n, r = _f32(root_pos3, 3), _f32(root_nrm3, 3), _f32(random_value)
        if p.shape[0] != self.nstrands or n.shape[0] != self.nstrands or r.size != self.nstrands:
            raise ValueError("one root position, normal and jitter value per strand")
        _check(self._lib.bh_init_strands(self._h, _ptr(p), _ptr(n), _ptr(r), maxlength))

    def init_sphere_scalp(self, rows: int, cols: int, first: int, random_value, maxlength: float = 0.5, order: int = 0):
        """Strands [first, first + nstrands) of the rows x cols sphere scalp in strand order `order` (BH_SCALP_ROW_MAJOR:
        latitude circles, BH_SCALP_COLUMN_MAJOR: meridians — contiguous shards are then balanced longitude wedges)."""
        r = _f32(random_value)
        if r.size != self.nstrands:
            raise ValueError("one jitter value per strand of this shard")
        _check(self._lib.bh_init_sphere_scalp_ordered(self._h, rows, cols, order, first, _ptr(r), maxlength))

    # -- state files / checksums ---------------------------------------------------
    def checksum(self, plane_mask: int = 7, first_strand: int = 0):
        """(sum, keyed sum) of the planes in `plane_mask`, computed on the device (bh_state_checksum)."""
        out = (C.c_uint64 * 2)()
        _check(self._lib.bh_state_checksum(self._h, plane_mask, first_strand, out))
        return int(out[0]), int(out[1])

    def save(self, path, *, plane_mask: int = 7, total_strands: int = 0, first_strand: int = 0, frame: int = 0,
             dt: float = 0.0, seed: int = 0):
        info = BhStateInfo(plane_mask=plane_mask, total_strands=total_strands, first_strand=first_strand, frame=frame,
                           dt=dt, seed=seed)
        _check(self._lib.bh_save_state(self._h, os.fsencode(path), C.byref(info)))

    def load(self, path) -> BhStateInfo:
        """Load a state file of this sim's shape: planes + parameters; verified by the device checksum."""
        info = BhStateInfo()
        _check(self._lib.bh_load_state(self._h, os.fsencode(path), C.byref(info)))
        return info

    @classmethod
    def from_state(cls, path, device: int = 0):
        """bh_peek_state -> bh_create -> bh_load_state. Returns (sim, info)."""
        head = peek_state(path)
        sim = cls(head.nstrands, head.nverts, device)
        try:
            return sim, sim.load(path)
        except Exception:
            sim.close()
            raise

    # -- stepping ----------------------------------------------------------------------------
    def step(self, dt: float, substeps: int = 1):
        _check(self._lib.bh_step(self._h, dt, substeps))

    def step_host(self, dt: float, substeps: int, pos4: np.ndarray, vel4: np.ndarray):
        for a in (pos4, vel4):
            if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != 4 * self.nvertices:
                raise ValueError("host planes must be C-contiguous float32 with 4*V elements")
        _check(self._lib.bh_step_host(self._h, dt, substeps, _ptr(pos4), _ptr(vel4)))

    def step_readback(self, dt: float, substeps: int, pos4: np.ndarray):
        """Hair::update with the state resident on the device, then the position plane in `pos4` (bh_step_readback: the
        device->host copy of a slice of strands overlaps the step of the next slice)."""
        if pos4.dtype != np.float32 or not pos4.flags.c_contiguous or pos4.size != 4 * self.nvertices:
            raise ValueError("pos4 must be C-contiguous float32 with 4*V elements")
        _check(self._lib.bh_step_readback(self._h, dt, substeps, _ptr(pos4)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.bh_launch_count(self._h))

    @property
    def kernel_kind(self) -> int:
        """0 streaming, 1 per-strand pipelined, 2 generic (bh_step_kernel_kind)."""
        return int(self._lib.bh_step_kernel_kind(self._h))

    # -- next stage: tess-stream ---------------------------------------------------------------
    def tess_set_patches(self, patch_indices):
        idx = np.ascontiguousarray(patch_indices, dtype=np.int32).reshape(-1)
        _check(self._lib.bh_tess_set_patches(self._h, _ptr(idx), idx.size))

    def tess_stream(self, ninstances: int = 3, nlines: int = 2, nsubsegments: int = 16, seed: int = 0, download: bool = True):
        """GL_LINES vertex stream (count, 4) of the interpolated render strands (defaults: hair.h:33-35)."""
        t = BhTessParams(ninstances, nlines, nsubsegments, seed)
        n = self._lib.bh_tess_stream_count(self._h, C.byref(t))
        if n < 0:
            raise ValueError("ninstances, nlines and nsubsegments must be >= 1")
        out = np.empty((n, 4), np.float32) if download else None
        _check(self._lib.bh_tess_stream(self._h, C.byref(t), _ptr(out)))
        return out if download else n

    # -- extensions --------------------------------------------------------------------------
    def set_skin(self, rest_root_pos3, joints4, weights3):
        p, w = _f32(rest_root_pos3, 3), _f32(weights3, 3)
        j = np.ascontiguousarray(joints4, dtype=np.int32)
        _check(self._lib.bh_set_skin(self._h, _ptr(p), _ptr(j), _ptr(w)))

    def skin_roots(self, dq_palette):
        dq = _f32(dq_palette, 8)
        _check(self._lib.bh_skin_roots(self._h, _ptr(dq), dq.shape[0]))

    def register_gl_buffer(self, gl_buffer: int):
        _check(self._lib.bh_register_gl_buffer(self._h, gl_buffer))

    def unregister_gl_buffer(self):
        _check(self._lib.bh_unregister_gl_buffer(self._h))

    def register_device_buffer(self, device_ptr: int, nbytes: int):
        """Step in a device allocation the caller owns (the GL buffer's protocol without GL, bh_register_device_buffer)."""
        _check(self._lib.bh_register_device_buffer(self._h, C.c_void_p(device_ptr), nbytes))

    def unregister_device_buffer(self):
        _check(self._lib.bh_unregister_device_buffer(self._h))

    def buffer_map_stats(self):
        """(maps, unmaps, mapped_now) of the registered GL / caller-owned buffer."""
        m, u, n = C.c_int64(), C.c_int64(), C.c_int()
        _check(self._lib.bh_buffer_map_stats(self._h, C.byref(m), C.byref(u), C.byref(n)))
        return int(m.value), int(u.value), bool(n.value)


class HairGroup:
    """One scalp sharded over several GPUs behind one handle and one host thread (bh_group_*, include/barbu_hair.h): contiguous
    strand ranges, asynchronous per-device launches, no per-step exchange; optional peer-copy gather of a plane to one GPU."""

    def __init__(self, devices, nstrands: int, nverts: int):
        self._lib = load_library()
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        _check(self._lib.bh_group_create(C.byref(self._h), devs, len(devices), nstrands, nverts))
        self.devices, self.nstrands, self.nverts = list(devices), nstrands, nverts
        self.nvertices = nstrands * nverts

    def close(self):
        if self._h:
            self._lib.bh_group_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self) -> int:
        return self._lib.bh_group_size(self._h)

    def shard_range(self, g: int):
        first, count = C.c_int64(), C.c_int64()
        _check(self._lib.bh_group_shard_range(self._h, g, C.byref(first), C.byref(count)))
        return first.value, count.value

    def configure(self, **kw):
        """Same keywords as HairSim.configure, applied to every shard."""
        p = default_params()
        shard0 = self._lib.bh_group_shard(self._h, 0)
        _check(self._lib.bh_get_params(shard0, C.byref(p)))
        _apply_param_keywords(p, kw)
        _check(self._lib.bh_group_set_params(self._h, C.byref(p)))

    def set_bounding_sphere(self, sphere):
        _check(self._lib.bh_group_set_bounding_sphere(self._h, (C.c_float * 4)(*sphere)))

    def init_sphere_scalp(self, rows: int, cols: int, order: int = BH_SCALP_COLUMN_MAJOR, seed: int = 1234, maxlength: float = 0.5):
        _check(self._lib.bh_group_init_sphere_scalp(self._h, rows, cols, order, seed, maxlength))

    def set_substep_fusion(self, enabled, always: bool = False):
        """HairSim.set_substep_fusion on every shard."""
        _check(self._lib.bh_group_set_substep_fusion(self._h, (2 if always else 1) if enabled else 0))

    def init_strands(self, root_pos3, root_nrm3, random_value, maxlength: float = 0.5):
        p, n, r = _f32(root_pos3, 3), _f32(root_nrm3, 3), _f32(random_value)
        if p.shape[0] != self.nstrands or n.shape[0] != self.nstrands or r.size != self.nstrands:
            raise ValueError("one root, one normal and one jitter value per strand of the whole scalp")
        _check(self._lib.bh_group_init_strands(self._h, _ptr(p), _ptr(n), _ptr(r), maxlength))

    def upload(self, pos4=None, vel4=None, tan4=None):
        arrs = [None if a is None else _f32(a) for a in (pos4, vel4, tan4)]
        for a in arrs:
            if a is not None and a.size != 4 * self.nvertices:
                raise ValueError("global planes: 4 floats per vertex of the whole scalp")
        _check(self._lib.bh_group_upload(self._h, *[_ptr(a) for a in arrs]))

    def download(self):
        out = [np.empty((self.nvertices, 4), np.float32) for _ in range(3)]
        _check(self._lib.bh_group_download(self._h, *[_ptr(a) for a in out]))
        return tuple(out)

    def step(self, dt: float, substeps: int = 1):
        _check(self._lib.bh_group_step(self._h, dt, substeps))

    def synchronize(self):
        _check(self._lib.bh_group_synchronize(self._h))

    def step_timed(self, dt: float, substeps: int, frames: int):
        """(max ms, [ms per shard]) of `frames` frames, measured with CUDA events on each device."""
        mx = C.c_float()
        per = (C.c_float * self.size)()
        _check(self._lib.bh_group_step_timed(self._h, dt, substeps, frames, C.byref(mx), per))
        return mx.value, list(per)

    @property
    def launch_count(self) -> int:
        return self._lib.bh_group_launch_count(self._h)

    def gather_plane(self, plane: int = BH_PLANE_POSITION, dst_device: int = 0):
        """(device pointer, ms): plane of the whole scalp on dst_device, gathered by peer copies."""
        ptr, ms = C.c_void_p(), C.c_float()
        _check(self._lib.bh_group_gather_plane(self._h, plane, dst_device, C.byref(ptr), C.byref(ms)))
        return ptr.value, ms.value

    def register_gl_buffer(self, gl_buffer: int, render_device: int = 0):
        _check(self._lib.bh_group_register_gl_buffer(self._h, gl_buffer, render_device))

    def gather_to_gl(self, plane_mask: int = 1):
        _check(self._lib.bh_group_gather_to_gl(self._h, plane_mask))


@dataclass
class ScalpMesh:
    """The members of MeshData (src/memory/resources/mesh_data.h:62-92) the hair path reads."""
    positions: np.ndarray                      # (S, 3) vertices[j].position
    normals: np.ndarray                        # (S, 3) vertices[j].normal
    indices: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))  # triangle list
    # skinned scalps (glTF): JOINTS_0 / WEIGHTS_0 per vertex (mesh_data.h:69-72), and the skin that drives them
    joints: Optional[np.ndarray] = None        # (S, 4) int32
    weights: Optional[np.ndarray] = None       # (S, 4) float32
    joint_nodes: Optional[list] = None
    inverse_bind: Optional[np.ndarray] = None  # (J, 16) column-major
    joint_rest_global: Optional[np.ndarray] = None   # (J, 16): the joints' world matrices in the file's rest pose

    @property
    def nvertices(self) -> int:
        return int(self.positions.shape[0])

    @property
    def nfaces(self) -> int:
        return int(np.asarray(self.indices).reshape(-1, 3).shape[0])


def dq_palette_from_matrices(global_pose, inverse_bind) -> np.ndarray:
    """(J, 8) dual-quaternion palette for HairSim.skin_roots from the joints' global pose and inverse bind matrices, (J, 16)
    each in GLM's column-major layout — SkeletonController::generate_skinning_datas (bh_dq_palette_from_matrices)."""
    A = np.ascontiguousarray(global_pose, np.float32).reshape(-1, 16)
    B = np.ascontiguousarray(inverse_bind, np.float32).reshape(-1, 16)
    if A.shape != B.shape:
        raise ValueError("one global pose and one inverse bind matrix per joint")
    out = np.empty((A.shape[0], 8), np.float32)
    _check(load_library().bh_dq_palette_from_matrices(_ptr(A), _ptr(B), A.shape[0], _ptr(out)))
    return out


_GLTF_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_GLTF_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _glm_mat4_mul(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """A * B for 16-float column-major matrices, in GLM's order of operations (type_mat4x4.inl:643-646), binary32."""
    A, B = A.astype(np.float32).reshape(4, 4), B.astype(np.float32).reshape(4, 4)        # [column][row]
    R = np.empty((4, 4), np.float32)
    for c in range(4):
        R[c] = ((A[0] * B[c, 0] + A[1] * B[c, 1]) + A[2] * B[c, 2]) + A[3] * B[c, 3]
    return R.reshape(16)


def _glm_mat4_vec(M: np.ndarray, v: np.ndarray, w: float) -> np.ndarray:
    """vec3(M * vec4(v, w)) for (n, 3) points, GLM's order: (m0*x + m1*y) + (m2*z + m3*w) (type_mat4x4.inl mat * vec)."""
    m = M.astype(np.float32).reshape(4, 4)
    x, y, z = (v[:, k:k + 1].astype(np.float32) for k in range(3))
    r = (m[0][None, :] * x + m[1][None, :] * y) + (m[2][None, :] * z + m[3][None, :] * np.float32(w))
    return np.ascontiguousarray(r[:, :3], np.float32)


def load_gltf_scalp(path: str, mesh: int = 0) -> "ScalpMesh":
    """Skinned scalp from a .gltf file (JSON with embedded base64 or side-by-side .bin buffers) the way the reference's loader
    presents it to Hair::setup (src/memory/resources/mesh_data_manager.cc:755-880 + MeshData::setup, mesh_data.cc:384-436):
    POSITION / NORMAL through the node's world matrix (w = 1 / w = 0), JOINTS_0 as integers, WEIGHTS_0 as floats, the index
    list of all primitives of the mesh concatenated; vertices are then re-indexed in FIRST-APPEARANCE order over that index
    list (every corner is the triple (i, i, i)), and joints / weights follow their vertex. The skin's joints, their inverse
    bind matrices and the joints' global rest matrices come along for bh_dq_palette_from_matrices. Plain Python + numpy:
    the reference reads the same file through cgltf; glb containers, sparse accessors and morph targets are not read."""
    import base64, json
    doc = json.load(open(path, "r"))
    base = os.path.dirname(os.path.abspath(path))
    blobs = []
    for b in doc.get("buffers", []):
        uri = b.get("uri", "")
        if uri.startswith("data:"):
            blobs.append(base64.b64decode(uri.split(",", 1)[1]))
        else:
            blobs.append(open(os.path.join(base, uri), "rb").read())

    def accessor(i):
        a = doc["accessors"][i]
        if "sparse" in a:
            raise ValueError("sparse accessors are not read (neither does the reference)")
        bv = doc["bufferViews"][a["bufferView"]]
        dt, width = np.dtype(_GLTF_COMPONENT[a["componentType"]]), _GLTF_WIDTH[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * width
        raw = np.frombuffer(blobs[bv["buffer"]], np.uint8, offset=start, count=(a["count"] - 1) * stride + dt.itemsize * width)
        rows = np.lib.stride_tricks.as_strided(raw, shape=(a["count"], dt.itemsize * width), strides=(stride, 1))
        out = np.ascontiguousarray(rows).view(dt).reshape(a["count"], width)
        if a.get("normalized") and dt.kind in "iu":                          # cgltf_accessor_read_float normalises
            out = np.maximum(out.astype(np.float32) / np.float32(np.iinfo(dt).max), np.float32(-1.0))
        return out

    def local_matrix(n):
        if "matrix" in n:
            return np.array(n["matrix"], np.float32)
        t = np.array(n.get("translation", [0, 0, 0]), np.float32); q = np.array(n.get("rotation", [0, 0, 0, 1]), np.float32)
        sc = np.array(n.get("scale", [1, 1, 1]), np.float32)
        x, y, z, w = (np.float32(v) for v in q)
        one, two = np.float32(1), np.float32(2)
        R = np.array([[one - two * (y * y + z * z), two * (x * y + w * z), two * (x * z - w * y), 0],
                      [two * (x * y - w * z), one - two * (x * x + z * z), two * (y * z + w * x), 0],
                      [two * (x * z + w * y), two * (y * z - w * x), one - two * (x * x + y * y), 0], [0, 0, 0, 1]], np.float32)   # [column][row]
        R[0, :3] *= sc[0]; R[1, :3] *= sc[1]; R[2, :3] *= sc[2]
        R[3, :3] = t
        return R.reshape(16)

    nodes = doc.get("nodes", [])
    parent = {c: i for i, n in enumerate(nodes) for c in n.get("children", [])}

    def world_matrix(i):
        M = local_matrix(nodes[i])
        while i in parent:
            i = parent[i]
            M = _glm_mat4_mul(local_matrix(nodes[i]), M)
        return M

    node_id = next((i for i, n in enumerate(nodes) if n.get("mesh") == mesh), None)
    world = world_matrix(node_id) if node_id is not None else np.eye(4, dtype=np.float32).reshape(16)
    P, Nn, Jn, W, idx, last = [], [], [], [], [], 0
    for prim in doc["meshes"][mesh]["primitives"]:
        at = prim["attributes"]
        pos = _glm_mat4_vec(world, accessor(at["POSITION"]).astype(np.float32), 1.0)
        P.append(pos)
        if "NORMAL" in at: Nn.append(_glm_mat4_vec(world, accessor(at["NORMAL"]).astype(np.float32), 0.0))
        if "JOINTS_0" in at: Jn.append(accessor(at["JOINTS_0"]).astype(np.int32))
        if "WEIGHTS_0" in at: W.append(accessor(at["WEIGHTS_0"]).astype(np.float32))
        if "indices" not in prim:
            raise ValueError("a primitive without indices: the reference loads no faces for it")
        idx.append(accessor(prim["indices"]).reshape(-1).astype(np.int64) + last)
        last += pos.shape[0]
    P, idx = np.concatenate(P), np.concatenate(idx)
    if not Nn:
        raise ValueError("no NORMAL attribute: load the file's geometry through an OBJ, whose loader recalculates them")
    Nn = np.concatenate(Nn)
    # MeshData::setup: unique corner triples in first-appearance order (each corner is (i, i, i))
    _, first = np.unique(idx, return_index=True)
    order = idx[np.sort(first)]                                              # old vertex id of new vertex k
    remap = np.full(P.shape[0], -1, np.int64); remap[order] = np.arange(order.size)
    m = ScalpMesh(np.ascontiguousarray(P[order]), np.ascontiguousarray(Nn[order]), remap[idx].astype(np.int32).reshape(-1, 3))
    if Jn and W:
        m.joints, m.weights = np.ascontiguousarray(np.concatenate(Jn)[order]), np.ascontiguousarray(np.concatenate(W)[order])
    skin_id = nodes[node_id].get("skin") if node_id is not None else None
    if skin_id is not None:
        skin = doc["skins"][skin_id]
        m.joint_nodes = list(skin["joints"])
        m.inverse_bind = np.ascontiguousarray(accessor(skin["inverseBindMatrices"]).astype(np.float32)) if "inverseBindMatrices" in skin \
            else np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (len(m.joint_nodes), 1))
        m.joint_rest_global = np.stack([world_matrix(j) for j in m.joint_nodes]).astype(np.float32)
    return m


def load_obj_scalp(path: str) -> "ScalpMesh":
    """Scalp MeshData from a Wavefront OBJ, read by the reference's rules (bh_load_obj_scalp)."""
    lib = load_library()
    pos, nrm, tri = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nv, nf = C.c_int64(), C.c_int64()
    rc = lib.bh_load_obj_scalp(os.fsencode(path), C.byref(pos), C.byref(nrm), C.byref(nv), C.byref(tri), C.byref(nf))
    if rc != BH_OK:
        raise BarbuHairError(rc, f"cannot load scalp {path!r}")
    try:
        P = np.ctypeslib.as_array(C.cast(pos, C.POINTER(C.c_float)), shape=(nv.value, 3)).copy()
        Nn = np.ctypeslib.as_array(C.cast(nrm, C.POINTER(C.c_float)), shape=(nv.value, 3)).copy()
        T = np.ctypeslib.as_array(C.cast(tri, C.POINTER(C.c_int32)), shape=(nf.value, 3)).copy()
    finally:
        for ptr in (pos, nrm, tri):
            lib.bh_free(ptr)
    return ScalpMesh(P, Nn, T)


class Hair:
    """Same call surface as the reference's `class Hair` for the simulation path.

    init()/deinit(), setup(scalp), update(dt), set_bounding_sphere(vec4), initialized() behave as in
    src/fx/hair.cc:26-125: `setup` with an unusable scalp logs and leaves the module uninitialised,
    `update` before `setup` is a silent no-op. `stream()` is the tess-stream half of `render` (hair.cc:141-173);
    the draw itself stays with the reference's GL path, which reads plane 0 / plane 2 of buffer 0 (hair.cc:371-389).
    """

    @dataclass
    class Parameters:
        maxlength: float = 0.50          # sim.maxlength, hair.h:29
        length_scale: float = 1.450      # render.lengthScale -> uScaleFactor, hair.h:41, hair.cc:108
        ncontrol_points: int = 4         # HAIR_MAX_PARTICLE_PER_STRAND, interop.h:8 (runtime here)
        seed: int = 1234                 # stands in for srand(time(NULL)), app.cc:96-97
        substeps: int = 1                # extension: 1 == reference
        math: int = BH_MATH_EXACT
        ninstances: int = 3              # tess.ninstances, hair.h:33
        nlines: int = 2                  # tess.nlines, hair.h:34
        nsubsegments: int = 16           # tess.nsubsegments, hair.h:35

    def __init__(self, device: int = 0, params: Optional["Hair.Parameters"] = None):
        self.params = params or Hair.Parameters()
        self.device = device
        self.nroots = 0
        self.sim: Optional[HairSim] = None
        self.patch_indices: Optional[np.ndarray] = None
        self._sphere = None
        self._patches_uploaded = False
        self.log = []

    def init(self):
        load_library()

    def deinit(self):
        if self.sim is not None:
            self.sim.close()
        self.sim, self.nroots, self.patch_indices, self._patches_uploaded = None, 0, None, False

    def initialized(self) -> bool:
        return self.nroots != 0

    def setup(self, scalp):
        """`scalp`: a ScalpMesh, or the path of an OBJ scalp resource as in Application.cc:38-39."""
        self.deinit()                                                       # a second setup starts over, like barbu_hair.hpp
        if isinstance(scalp, (str, os.PathLike)):
            try:
                scalp = load_obj_scalp(os.fspath(scalp))
            except BarbuHairError:
                scalp = None
        if scalp is None or scalp.nvertices == 0:
            self.log.append("The scalp mesh resource was not found.")       # hair.cc:45-48
            return
        N = self.params.ncontrol_points
        S = scalp.nvertices                                                 # hair.cc:58
        self.sim = HairSim(S, N, self.device)
        self.sim.set_step_policy(BH_POLICY_AUTO)                            # small scalps (the reference's own: 448 x 4): latency kernel
        # init_simulation (hair.cc:236-361): device expansion + host tangents
        rv = random_values(self.params.seed, 0, S)
        self.sim.init_strands(scalp.positions, scalp.normals, rv, self.params.maxlength)
        tan = init_tangents_host(scalp.normals, S, 0, N, self.params.maxlength)
        self.sim.upload(tan4=tan)
        # init_mesh (hair.cc:363-417): element buffer for the tess patches
        if scalp.nfaces:
            self.patch_indices = build_patch_indices(scalp.indices, N, self.device)
        self.sim.configure(scale=self.params.length_scale, math=self.params.math)
        if self._sphere is not None:
            self.sim.set_bounding_sphere(self._sphere)
        self.nroots = S

    def set_bounding_sphere(self, sphere: Sequence[float]):
        self._sphere = tuple(float(x) for x in sphere)
        if self.sim is not None:
            self.sim.set_bounding_sphere(self._sphere)

    def update(self, dt: float):
        if not self.initialized():
            self.log.append("Calling Hair::update without initialization.")  # hair.cc:90-93
            return
        self.sim.configure(scale=self.params.length_scale)                   # uniform re-sent every frame, hair.cc:108
        self.sim.step(dt, self.params.substeps)

    def stream(self, download: bool = True):
        """The tess-stream half of Hair::render (hair.cc:141-173) for the tess parameters (hair.h:33-35): the GL_LINES
        vertex stream (count, 4) of the interpolated render strands, or its length when `download` is False."""
        if not self.initialized():
            self.log.append("Calling Hair::render without initialization.")  # hair.cc:128-131
            return None
        if self.patch_indices is None:
            self.log.append("The scalp has no faces: nothing to tessellate.")
            return None
        if not self._patches_uploaded:
            self.sim.tess_set_patches(self.patch_indices)
            self._patches_uploaded = True
        return self.sim.tess_stream(self.params.ninstances, self.params.nlines, self.params.nsubsegments, self.params.seed, download)
