"""barbu_b200 — B200-native hair-strand simulation behind Barbü's `Hair` interface.

The package holds only what the hot path needs: csrc/ (sm_100a kernels + the C ABI of
include/barbu_hair.h), the build recipe, and the host-side mirror of the reference interface.
"""
from .hair import (BH_MATH_EXACT, BH_MATH_FAST, BH_POLICY_THROUGHPUT, BH_POLICY_LATENCY, BH_POLICY_AUTO, BH_SCALP_ROW_MAJOR, BH_SCALP_COLUMN_MAJOR, BarbuHairError, BhParams, BhStateInfo, Hair, HairGroup, HairSim, PinnedBuffer, ScalpMesh,
                   build_patch_indices, default_params, init_tangents_host, load_library, load_obj_scalp, load_gltf_scalp, dq_palette_from_matrices, peek_state, random_values,
                   selftest_math, sphere_scalp_triangles)
from .marschner import BhMarschnerParams, Marschner, generate_luts

__all__ = ["BhMarschnerParams", "Marschner", "generate_luts", "BH_MATH_EXACT", "BH_MATH_FAST", "BH_POLICY_THROUGHPUT", "BH_POLICY_LATENCY", "BH_POLICY_AUTO", "BH_SCALP_ROW_MAJOR", "BH_SCALP_COLUMN_MAJOR", "BarbuHairError", "BhParams", "BhStateInfo", "Hair", "HairGroup", "HairSim", "PinnedBuffer",
           "ScalpMesh", "build_patch_indices", "default_params", "init_tangents_host", "load_library", "load_obj_scalp", "load_gltf_scalp", "dq_palette_from_matrices", "peek_state",
           "random_values", "selftest_math", "sphere_scalp_triangles"]
