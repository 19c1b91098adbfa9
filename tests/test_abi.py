"""The C ABI: header, ctypes mirror and shared library agree (no GPU needed)."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from paa_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.paa_abi_version() == _lib.ABI_VERSION
    header = open(os.path.join(ROOT, "include", "paa_b200.h")).read()
    import re
    declared = set(re.findall(r"\b(paa_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    raw = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(raw, name), name


def test_struct_layout_matches_header(tmp_path):
    """Compiles a C program against include/paa_b200.h and compares sizeof/offsetof with ctypes."""
    from paa_b200 import _lib
    fields_loss = ["num_images", "gamma", "anchor_image_stride", "levels", "gt_boxes", "gt_offsets",
                   "workspace", "normalisers", "grad_losses", "dbg_matched_idx", "teacher_combined_loss", "rank",
                   "peer_norm", "bg_iou_threshold", "box_code_weights", "smooth_l1_beta", "reg_norm_weight", "fcos_strides",
                   "fcos_center_radius", "fcos_iou_loss_type", "fcos_norm_reg_targets", "atss_positive_type",
                   "peer_timeout_s", "peer_status", "gt_offsets_dev", "gt_capacity", "gt_per_image_capacity",
                   "head_layout"]
    fields_post = ["num_images", "pre_nms_thresh", "anchor_image_stride", "levels", "image_wh", "workspace",
                   "out_boxes", "out_count", "dbg_pre_boxes", "dbg_nms_keep", "box_decode", "decode_weights",
                   "decode_clip", "head_layout"]
    fields_rpn = ["num_images", "head_layout", "anchor_image_stride", "levels", "gt_boxes", "gt_offsets", "matched_idx",
                  "sampled", "n_pos", "box_code_weights", "smooth_l1_beta", "losses", "grad_losses"]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "paa_b200.h"', 'int main(void){',
           'printf("%zu %zu %zu\\n", sizeof(PaaLevel), sizeof(PaaLossArgs), sizeof(PaaPostArgs));']
    for f in fields_loss:
        src.append('printf("%%zu\\n", offsetof(PaaLossArgs, %s));' % f)
    for f in fields_post:
        src.append('printf("%%zu\\n", offsetof(PaaPostArgs, %s));' % f)
    src.append('printf("%zu\\n", sizeof(PaaRpnArgs));')
    for f in fields_rpn:
        src.append('printf("%%zu\\n", offsetof(PaaRpnArgs, %s));' % f)
    src.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [int(x) for x in out]
    assert sizes[0:3] == [ctypes.sizeof(_lib.PaaLevel), ctypes.sizeof(_lib.PaaLossArgs),
                          ctypes.sizeof(_lib.PaaPostArgs)]
    k = 3
    for f in fields_loss:
        assert sizes[k] == getattr(_lib.PaaLossArgs, f).offset, f
        k += 1
    for f in fields_post:
        assert sizes[k] == getattr(_lib.PaaPostArgs, f).offset, f
        k += 1
    assert sizes[k] == ctypes.sizeof(_lib.PaaRpnArgs)
    k += 1
    for f in fields_rpn:
        assert sizes[k] == getattr(_lib.PaaRpnArgs, f).offset, f
        k += 1


def test_workspace_size_is_monotonic_and_argument_errors_are_reported():
    from paa_b200 import _lib
    lib = _lib.load()
    a = lib.paa_loss_workspace_bytes(2, 22400, 40, 5, 9)
    b = lib.paa_loss_workspace_bytes(4, 22400, 80, 5, 9)
    assert 0 < a < b
    assert lib.paa_loss_workspace_bytes(0, 22400, 40, 5, 9) == 0
    args = _lib.PaaLossArgs()          # all zero: rejected before any CUDA call
    rc = lib.paa_assign_loss(ctypes.byref(args), None)
    assert rc == _lib.ERR_BAD_ARGUMENT
    assert b"num_images" in lib.paa_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "paa_assign_loss")


def test_missing_library_fails_loudly(monkeypatch):
    from paa_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("PAA_B200_LIB", "/nonexistent/libpaa_b200.so")
    with pytest.raises(_lib.PaaLibraryError):
        _lib.load()
