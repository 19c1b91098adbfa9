"""ctypes binding of libpaa_b200.so -- mirrors include/paa_b200.h field for field.

There is no fallback: if the library is missing or does not load, importing a product module that
needs it raises.  Build it with ``python -m paa_b200.build`` (needs nvcc) or
``__graft_entry__.build()``.
"""
import contextlib
import ctypes as C
import os

from paa_b200 import build as _build

MAX_LEVELS = 8
MAX_IMAGES = 256
MAX_PEERS = 16
PEER_BUFFER_DOUBLES = 256
MAX_CANDIDATES = 128
ABI_VERSION = 8

LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
ERR_BAD_ARGUMENT, ERR_WORKSPACE, ERR_EMPTY_TARGET, ERR_UNSUPPORTED = -1, -2, -3, -4

_fp = C.POINTER(C.c_float)


class PaaLevel(C.Structure):
    _fields_ = [("box_cls", C.c_void_p), ("box_regression", C.c_void_p), ("iou_pred", C.c_void_p),
                ("anchors", C.c_void_p), ("grad_box_cls", C.c_void_p), ("grad_box_regression", C.c_void_p),
                ("grad_iou_pred", C.c_void_p), ("hw", C.c_int32), ("grid_w", C.c_int32)]


class PaaLossArgs(C.Structure):
    _fields_ = [("num_images", C.c_int32), ("num_levels", C.c_int32), ("num_classes", C.c_int32),
                ("anchors_per_loc", C.c_int32), ("topk", C.c_int32), ("use_iou_pred", C.c_int32),
                ("world_size", C.c_int32), ("loss_flavour", C.c_int32),
                ("gamma", C.c_float), ("alpha", C.c_float), ("iou_threshold", C.c_float),
                ("reg_loss_weight", C.c_float), ("iou_loss_weight", C.c_float), ("bg_iou_threshold", C.c_float),
                ("anchor_image_stride", C.c_int64),
                ("levels", PaaLevel * MAX_LEVELS),
                ("gt_boxes", C.c_void_p), ("gt_labels", C.c_void_p),
                ("gt_offsets", C.c_int32 * (MAX_IMAGES + 1)),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("normalisers", C.c_void_p), ("losses", C.c_void_p), ("grad_losses", C.c_void_p),
                ("dbg_matched_idx", C.c_void_p), ("dbg_iou_labels", C.c_void_p),
                ("dbg_combined_loss", C.c_void_p), ("dbg_cand_idx", C.c_void_p),
                ("dbg_cand_cnt", C.c_void_p), ("dbg_num_pos", C.c_void_p), ("dbg_gmm", C.c_void_p),
                ("dbg_paa_labels", C.c_void_p), ("teacher_combined_loss", C.c_void_p),
                ("rank", C.c_int32), ("reserved2", C.c_int32), ("peer_norm", C.c_void_p * MAX_PEERS),
                ("box_code_weights", C.c_float * 4), ("smooth_l1_beta", C.c_float),
                ("reg_norm_weight", C.c_float), ("fcos_strides", C.c_float * MAX_LEVELS),
                ("fcos_center_radius", C.c_float), ("fcos_iou_loss_type", C.c_int32),
                ("fcos_norm_reg_targets", C.c_int32), ("atss_positive_type", C.c_int32),
                ("peer_timeout_s", C.c_float), ("reserved3", C.c_int32), ("peer_status", C.c_void_p),
                ("gt_offsets_dev", C.c_void_p), ("gt_capacity", C.c_int32), ("gt_per_image_capacity", C.c_int32),
                ("head_layout", C.c_int32), ("reserved4", C.c_int32)]


class PaaPostArgs(C.Structure):
    _fields_ = [("num_images", C.c_int32), ("num_levels", C.c_int32), ("num_classes", C.c_int32),
                ("anchors_per_loc", C.c_int32), ("pre_nms_top_n", C.c_int32),
                ("detections_per_img", C.c_int32), ("score_voting", C.c_int32), ("skip_nms", C.c_int32),
                ("pre_nms_thresh", C.c_float), ("nms_thresh", C.c_float),
                ("anchor_image_stride", C.c_int64),
                ("levels", PaaLevel * MAX_LEVELS),
                ("image_wh", (C.c_float * 2) * MAX_IMAGES),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("out_boxes", C.c_void_p), ("out_scores", C.c_void_p), ("out_labels", C.c_void_p),
                ("out_count", C.c_void_p),
                ("dbg_pre_boxes", C.c_void_p), ("dbg_pre_scores", C.c_void_p),
                ("dbg_pre_labels", C.c_void_p), ("dbg_pre_count", C.c_void_p),
                ("dbg_nms_keep", C.c_void_p),
                ("box_decode", C.c_int32), ("decode_weights", C.c_float * 4), ("decode_clip", C.c_float),
                ("head_layout", C.c_int32)]

DECODE_ATSS_BOX, DECODE_LEGACY, DECODE_LTRB = 0, 1, 2
LOSS_PAA, LOSS_ATSS, LOSS_RETINANET, LOSS_FCOS = 0, 1, 2, 3
class PaaRpnArgs(C.Structure):
    _fields_ = [("num_images", C.c_int32), ("num_levels", C.c_int32), ("anchors_per_loc", C.c_int32),
                ("head_layout", C.c_int32), ("anchor_image_stride", C.c_int64), ("levels", PaaLevel * MAX_LEVELS),
                ("gt_boxes", C.c_void_p), ("gt_offsets", C.c_int32 * (MAX_IMAGES + 1)), ("matched_idx", C.c_void_p),
                ("sampled", C.c_void_p), ("n_pos", C.c_int32), ("n_neg", C.c_int32),
                ("box_code_weights", C.c_float * 4), ("smooth_l1_beta", C.c_float), ("reserved", C.c_int32),
                ("losses", C.c_void_p), ("grad_losses", C.c_void_p)]


ATSS_POSITIVE_TYPES = {"ATSS": 0, "SSC": 1, "IoU": 2}
IOU_LOSS_TYPES = {"iou": 0, "linear_iou": 1, "giou": 2}


# name -> (restype, argtypes); every symbol include/paa_b200.h declares
SYMBOLS = {
    "paa_abi_version": (C.c_int, []),
    "paa_last_error": (C.c_char_p, []),
    "paa_loss_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "paa_postprocess_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "paa_assign": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p]),
    "paa_loss": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p]),
    "paa_atss_assign": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p]),
    "paa_retinanet_assign": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p]),
    "paa_fcos_assign": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p]),
    "paa_assign_loss": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p]),
    "paa_rescale_grads": (C.c_int, [C.POINTER(PaaLossArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    "paa_rpn_loss": (C.c_int, [C.POINTER(PaaRpnArgs), C.c_void_p]),
    "paa_postprocess": (C.c_int, [C.POINTER(PaaPostArgs), C.c_void_p]),
    "paa_ml_nms_workspace_bytes": (C.c_size_t, [C.c_int]),
    "paa_ml_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "paa_sigmoid_focal_loss_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                                 C.c_float, C.c_void_p, C.c_void_p]),
    "paa_sigmoid_focal_loss_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                  C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "paa_grid_anchors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "paa_anchor_visibility": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                        C.c_void_p]),
    "paa_boxlist_iou": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "paa_box_vote_workspace_bytes": (C.c_size_t, [C.c_int]),
    "paa_box_vote": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    "paa_selftest_roots": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "paa_kernel_timing_begin": (C.c_int, [C.c_int]),
    "paa_kernel_timing_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

KERNEL_IDS = dict(pass1=1, match_score=2, select_gmm=3, final_loss=4, positive_terms=5, finish_loss=6, norm_wait=7,
                  prep_step=8,
                  post_threshold=18, post_segments=19, post_candidates=10, post_filter=11,
                  post_select=12, post_rank=13, post_nms_mask=14, post_nms_scan=15, post_finish=16, post_vote=17,
                  post_nms_runs=20)

_lib = None


class PaaLibraryError(RuntimeError):
    pass


def library_path():
    return os.environ.get("PAA_B200_LIB", _build.LIB_PATH)


def load():
    """Loads the shared library once and types every entry point.  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise PaaLibraryError(
            "libpaa_b200.so not found at %s -- build it with `python -m paa_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.paa_abi_version() != ABI_VERSION:
        raise PaaLibraryError("libpaa_b200.so ABI %d != binding %d" % (lib.paa_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


_NO_GUARD = contextlib.nullcontext()


def stream_handle(device):
    """``cudaStream_t`` of torch's current stream on `device` as an integer.  Takes torch's raw-stream accessor
    (no ``torch.cuda.Stream`` object per call: this sits on the per-step host path) when it exists."""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None and device.index is not None:
        return raw(device.index)
    return torch.cuda.current_stream(device).cuda_stream


def device_guard(device):
    """Context that makes `device` the current CUDA device for the library call; free when it already is."""
    import torch
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def check(rc, what):
    """Turns a non-zero return code into the exception the reference would raise at that point."""
    if rc == 0:
        return
    msg = load().paa_last_error().decode("utf-8", "replace")
    if rc == ERR_EMPTY_TARGET:
        raise ValueError(msg)                      # matcher.py:53-58
    if rc in (ERR_BAD_ARGUMENT, ERR_UNSUPPORTED, ERR_WORKSPACE):
        raise RuntimeError("%s: %s" % (what, msg))
    raise RuntimeError("%s: CUDA error %d: %s" % (what, rc, msg))
