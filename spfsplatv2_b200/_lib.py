"""ctypes binding of libspfsplat.so (include/spfsplat.h).  Fails loudly when the CUDA library is
missing: there is NO CPU / PyTorch fallback on the product path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspfsplat.so")

SPF_FLAG_SH_LAYOUT_CK = 1 << 0
SPF_FLAG_NO_COV_GRAD = 1 << 1
SPF_FLAG_NO_SH_GRAD = 1 << 2
SPF_FLAG_NO_TMA = 1 << 3
SPF_FLAG_QUAT_XYZW = 1 << 4
SPF_FLAG_BWD_V1 = 1 << 6

_fp = C.c_void_p


class SpfRasterDesc(C.Structure):
    _fields_ = [("n_scenes", C.c_int32), ("views_per_scene", C.c_int32), ("n_gaussians", C.c_int32),
                ("image_height", C.c_int32), ("image_width", C.c_int32), ("sh_degree", C.c_int32),
                ("flags", C.c_uint32), ("scale_modifier", C.c_float), ("dup_capacity", C.c_int64),
                ("ticket", C.c_int32), ("pair_capacity", C.c_int32)]


class SpfRasterIn(C.Structure):
    _fields_ = [("means3D", _fp), ("scales", _fp), ("rotations", _fp), ("opacities", _fp), ("shs", _fp),
                ("colors_precomp", _fp), ("sh_coeffs", C.c_int32), ("viewmatrix", _fp), ("projmatrix", _fp),
                ("tanfov", _fp), ("bg", _fp), ("pre_scale", _fp), ("raw_head", _fp), ("raw_stride", C.c_int32),
                ("raw_has_density", C.c_int32), ("raw_eps", C.c_float), ("opacity_exponent", C.c_float)]


class SpfRasterState(C.Structure):
    _fields_ = [("xy", _fp), ("depth", _fp), ("conic_opacity", _fp), ("rgb", _fp), ("radii", _fp),
                ("tiles_touched", _fp), ("dup_offset", _fp), ("control", _fp), ("bucket", _fp), ("slab", _fp),
                ("cullbox", _fp), ("tile_ranges", _fp), ("final_T", _fp), ("n_contrib", _fp), ("accum", _fp), ("pair_log", _fp), ("pair_count", _fp), ("host_counters", _fp)]


class SpfRasterOut(C.Structure):
    _fields_ = [("color", _fp), ("depth", _fp), ("alpha", _fp)]


class SpfRasterGradOut(C.Structure):
    _fields_ = [("dL_dcolor", _fp), ("dL_ddepth", _fp), ("dL_dalpha", _fp)]


class SpfRasterGradIn(C.Structure):
    _fields_ = [("dup_grad", _fp), ("pose_partial", _fp), ("dL_dmeans3D", _fp), ("dL_dscales", _fp),
                ("dL_drotations", _fp), ("dL_dopacities", _fp), ("dL_dshs", _fp), ("dL_dcolors", _fp),
                ("dL_dviewmatrix", _fp), ("dL_dmeans2D", _fp), ("dL_draw_head", _fp)]


EXPORTS = ("spf_version", "spf_last_error", "spf_raster_control_ints", "spf_raster_forward",
           "spf_raster_backward", "spf_raster_forward_stages", "spf_raster_backward_stages",
           "spf_raster_unpack_sorted", "spf_camera_forward", "spf_camera_backward", "spf_rope2d", "spf_rope2d_qk",
           "spf_image_mse", "spf_image_mse_blocks", "spf_adapter_forward", "spf_adapter_backward", "spf_head_forward", "spf_head_backward", "spf_ply_pack",
           "spf_multimem_allreduce_f32", "spf_multimem_allreduce_f32_fused")

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libspfsplat.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    res = subprocess.run(["bash", script], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libspfsplat.so failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout, file=sys.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or spfsplatv2_b200/csrc/build.sh). "
            "There is no CPU fallback.")
    l = C.CDLL(LIB_PATH)
    l.spf_version.restype = C.c_int
    l.spf_last_error.restype = C.c_char_p
    l.spf_raster_control_ints.restype = C.c_int64
    l.spf_raster_control_ints.argtypes = [C.POINTER(SpfRasterDesc)]
    l.spf_raster_forward.restype = C.c_int
    l.spf_raster_forward.argtypes = [C.POINTER(SpfRasterDesc), C.POINTER(SpfRasterIn), C.POINTER(SpfRasterState),
                                     C.POINTER(SpfRasterOut), C.c_void_p]
    l.spf_raster_backward.restype = C.c_int
    l.spf_raster_backward.argtypes = [C.POINTER(SpfRasterDesc), C.POINTER(SpfRasterIn), C.POINTER(SpfRasterState),
                                      C.POINTER(SpfRasterGradOut), C.POINTER(SpfRasterGradIn), C.c_void_p]
    l.spf_raster_forward_stages.restype = C.c_int
    l.spf_raster_forward_stages.argtypes = [C.POINTER(SpfRasterDesc), C.POINTER(SpfRasterIn),
                                            C.POINTER(SpfRasterState), C.POINTER(SpfRasterOut), C.c_uint32, C.c_void_p]
    l.spf_raster_backward_stages.restype = C.c_int
    l.spf_raster_backward_stages.argtypes = [C.POINTER(SpfRasterDesc), C.POINTER(SpfRasterIn),
                                             C.POINTER(SpfRasterState), C.POINTER(SpfRasterGradOut),
                                             C.POINTER(SpfRasterGradIn), C.c_uint32, C.c_void_p]
    l.spf_raster_unpack_sorted.restype = C.c_int
    l.spf_raster_unpack_sorted.argtypes = [C.POINTER(SpfRasterDesc), C.POINTER(SpfRasterState), C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    l.spf_camera_forward.restype = C.c_int
    l.spf_camera_forward.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 9
    l.spf_camera_backward.restype = C.c_int
    l.spf_camera_backward.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5
    l.spf_rope2d.restype = C.c_int
    l.spf_rope2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                             C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    l.spf_rope2d_qk.restype = C.c_int
    l.spf_rope2d_qk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    l.spf_image_mse_blocks.restype = C.c_int
    l.spf_image_mse_blocks.argtypes = [C.c_int64]
    l.spf_image_mse.restype = C.c_int
    l.spf_image_mse.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    l.spf_adapter_forward.restype = C.c_int
    l.spf_adapter_forward.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.spf_adapter_backward.restype = C.c_int
    l.spf_adapter_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                       C.c_void_p, C.c_void_p]
    l.spf_head_forward.restype = C.c_int
    l.spf_head_forward.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    l.spf_head_backward.restype = C.c_int
    l.spf_head_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p]
    l.spf_ply_pack.restype = C.c_int
    l.spf_ply_pack.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    l.spf_multimem_allreduce_f32.restype = C.c_int
    l.spf_multimem_allreduce_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    l.spf_multimem_allreduce_f32_fused.restype = C.c_int
    l.spf_multimem_allreduce_f32_fused.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                   C.c_int32, C.c_int32, C.c_void_p]
    _lib = l
    return l


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().spf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")
