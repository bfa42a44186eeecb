"""ctypes binding of liblighthand_b200.so (include/lighthand_b200.h).

There is no fallback: if the shared library is missing the import of any op raises, and every
op raises when CUDA is unavailable.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C lighthand_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIGHTHAND_B200_LIB") or os.path.join(_HERE, "liblighthand_b200.so")   # env override: kernel-variant experiments

LH_INTERP_CV2_U8, LH_INTERP_CV2_F32, LH_INTERP_TORCH = 0, 1, 2
LH_CHAN_RGB, LH_CHAN_BGR = 0, 1
LH_LAYOUT_CHW, LH_LAYOUT_HWC = 0, 1
LH_TARGET_HRNET, LH_TARGET_HOURGLASS = 0, 1
LH_PCK_PROPORTION, LH_PCK_MM = 0, 1
LH_COMM_ID_BYTES = 128
ABI_VERSION = 4

c_p = C.c_void_p
c_i = C.c_int
c_i64 = C.c_int64
c_f = C.c_float


class LighthandError(RuntimeError):
    pass


class HeatmapArgs(C.Structure):
    _fields_ = [
        ("pred", c_p), ("joints", c_p), ("gauss_table", c_p), ("target_in", c_p), ("inv_affine", c_p),
        ("target_out", c_p), ("target_weight_out", c_p), ("grad", c_p), ("loss_f64", c_p),
        ("loss_meter", c_p), ("loss_f32", c_p), ("coords", c_p), ("maxvals", c_p), ("argidx", c_p), ("ws", c_p),
        ("metric_counts", c_p), ("metric_sums", c_p),
        ("N", c_i64), ("J", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("joint_stride", C.c_int32),
        ("table_side", C.c_int32), ("style", C.c_int32), ("use_target_weight", C.c_int32),
        ("subpixel", C.c_int32), ("stride_x", c_f), ("stride_y", c_f), ("out_scale", c_f), ("pck_thr", c_f),
        ("pck_mode", C.c_int32), ("use_vis", C.c_int32),
    ]


class PipelineCfg(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("batch", C.c_int32), ("src_h", C.c_int32), ("src_w", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("J", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("interp", C.c_int32), ("train", C.c_int32), ("materialize_target", C.c_int32),
        ("subpixel", C.c_int32), ("exact_norm", C.c_int32), ("sigma", c_f), ("pck_T", c_f), ("mean_std", c_f * 6),
    ]


class PipelineIO(C.Structure):
    _fields_ = [
        ("images", c_p), ("fwd_affine", c_p), ("flip", c_p), ("joints", c_p), ("pred", c_p),
        ("loss", c_p), ("coords", c_p), ("maxvals", c_p), ("counts", c_p), ("sums", c_p),
    ]


class WarpArgs(C.Structure):
    _fields_ = [
        ("src", c_p), ("src_img_stride", c_i64), ("src_h", C.c_int32), ("src_w", C.c_int32), ("src_pitch", C.c_int32),
        ("fwd_affine", c_p), ("flip", c_p), ("chan_gain", c_p), ("mean_std", c_f * 6), ("has_mean_std", C.c_int32),
        ("lut", c_p), ("interp", C.c_int32), ("chan_order", C.c_int32), ("out_layout", C.c_int32),
        ("dst", c_p), ("N", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
    ]


class FreiDesc(C.Structure):
    _fields_ = [
        ("ulx", C.c_int32), ("uly", C.c_int32), ("cw", C.c_int32), ("ch", C.c_int32), ("pad", C.c_int32),
        ("rotated", C.c_int32), ("iM", C.c_double * 6), ("scale_x", C.c_double), ("scale_y", C.c_double), ("pn", c_f * 3), ("reserved", C.c_int32),
        ("scratch_off", c_i64),
    ]


class JitterParams(C.Structure):
    _fields_ = [("order", C.c_int32 * 4), ("factor", c_f * 4), ("one_minus", c_f * 4), ("apply", C.c_int32)]


LH_JITTER_PARTS = 32

# name -> (restype, argtypes); every symbol include/lighthand_b200.h declares
PROTOTYPES = {
    "lh_abi_version": (c_i, []),
    "lh_strerror": (C.c_char_p, [c_i]),
    "lh_device_query": (c_i, [c_i, C.POINTER(c_i), C.POINTER(c_i), C.POINTER(c_i), C.POINTER(C.c_uint64)]),
    "lh_norm_lut_build": (c_i, [c_p, c_i, c_p, c_p, c_p]),
    "lh_warp_normalize": (c_i, [c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p,
                                c_i, c_i, c_i, c_p]),
    "lh_gen_targets": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_i, c_i, c_p, c_i, c_p, c_p, c_p]),
    "lh_transform_joints": (c_i, [c_p, c_i, c_p, c_p, c_p, c_i, c_p, c_i64, c_i, c_i, c_i, c_p, c_i, c_p, c_p]),
    "lh_bbox_diag": (c_i, [c_p, c_i, c_i64, c_i, c_p, c_p]),
    "lh_heatmap_ws_bytes": (C.c_size_t, [c_i64, c_i]),
    "lh_jmse_loss": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "lh_decode": (c_i, [c_p, c_i64, c_i, c_i, c_i, c_f, c_i, c_p, c_p, c_p, c_p, c_p]),
    "lh_heatmap_fused": (c_i, [C.POINTER(HeatmapArgs), c_p]),
    "lh_pck_epe_batch": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_f, c_i, c_i, c_p, c_p, c_p]),
    "lh_pck_curve_accum": (c_i, [c_p, c_p, c_i, c_p, c_p, c_i, c_i64, c_i, c_p, c_i, c_i, c_i,
                                 c_p, c_p, c_p, c_p, c_p]),
    "lh_frei_rgb_processing": (c_i, [c_p, c_i64, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p]),
    "lh_frei_fused": (c_i, [c_p, c_i64, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p]),
    "lh_resize_normalize_f32": (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p]),
    "lh_color_jitter_normalize": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "lh_pipeline_create": (c_i, [C.POINTER(PipelineCfg), c_p, c_i, C.POINTER(c_p)]),
    "lh_pipeline_destroy": (None, [c_p]),
    "lh_pipeline_input_dev": (c_p, [c_p]),
    "lh_pipeline_pred_dev": (c_p, [c_p]),
    "lh_pipeline_grad_dev": (c_p, [c_p]),
    "lh_pipeline_target_dev": (c_p, [c_p]),
    "lh_pipeline_step_host": (c_i, [c_p, C.POINTER(PipelineIO)]),
    "lh_pipeline_submit": (c_i, [c_p, C.POINTER(PipelineIO), C.POINTER(c_i)]),
    "lh_pipeline_wait": (c_i, [c_p, c_i]),
    "lh_pipeline_submit_inputs": (c_i, [c_p, C.POINTER(PipelineIO), C.POINTER(c_i)]),
    "lh_pipeline_submit_loss": (c_i, [c_p, c_i]),
    "lh_pipeline_submit_inputs_jpeg": (c_i, [c_p, c_p, c_p, c_i, C.POINTER(PipelineIO), C.POINTER(c_i)]),
    "lh_pipeline_loss_stream": (c_p, [c_p]),
    "lh_pipeline_input_dev_slot": (c_p, [c_p, c_i]),
    "lh_pipeline_allreduce_metrics": (c_i, [c_p, c_p]),
    "lh_pipeline_read_metrics": (c_i, [c_p, c_p, c_p]),
    "lh_pipeline_reset_metrics": (c_i, [c_p]),
    "lh_pipeline_launches_per_step": (c_i, [c_p]),
    "lh_stage_ring_create": (c_i, [c_i, C.POINTER(c_p)]),
    "lh_stage_ring_destroy": (None, [c_p]),
    "lh_stage_ring_data_stream": (c_p, [c_p]),
    "lh_stage_ring_begin_data": (c_i, [c_p, c_i, c_p, c_p, c_i]),
    "lh_stage_ring_end_data": (c_i, [c_p, c_i]),
    "lh_stage_ring_begin_loss": (c_i, [c_p, c_i, c_p]),
    "lh_stage_ring_end_loss": (c_i, [c_p, c_i, c_p]),
    "lh_stage_ring_data": (c_i, [c_p, c_i, C.POINTER(WarpArgs)]),
    "lh_stage_ring_loss": (c_i, [c_p, c_i, C.POINTER(HeatmapArgs), c_p]),
    "lh_host_alloc": (c_p, [C.c_size_t]),
    "lh_host_free": (None, [c_p]),
    "lh_jpeg_create": (c_i, [c_i, c_i, c_i, C.POINTER(c_p)]),
    "lh_jpeg_destroy": (None, [c_p]),
    "lh_jpeg_decode_batch": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_i64, c_i, c_p]),
    "lh_comm_unique_id": (c_i, [c_p]),
    "lh_comm_init": (c_i, [c_p, c_i, c_i, c_i, C.POINTER(c_p)]),
    "lh_comm_destroy": (c_i, [c_p]),
    "lh_metrics_allreduce": (c_i, [c_p, c_p, c_i, c_p, c_i, c_p]),
    "lh_metrics_allreduce_multi": (c_i, [c_p, c_i, C.POINTER(c_p), C.POINTER(C.c_size_t), C.POINTER(c_i), c_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every prototype; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LighthandError(
            f"{LIB_PATH} not found: lighthand_b200 has no CPU fallback; build the sm_100a library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C lighthand_b200/csrc)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.lh_abi_version() != ABI_VERSION:
        raise LighthandError("liblighthand_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def strerror(code: int) -> str:
    return load().lh_strerror(int(code)).decode()


def check(code: int, what: str = "") -> None:
    if code != 0:
        raise LighthandError(f"{what or 'lighthand_b200 call'} failed: [{code}] {strerror(code)}")
