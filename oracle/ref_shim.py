"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (JunhoPark0314/PAA).

Imports the unmodified reference Python from ``/root/reference`` (present only in the
build container, never on the GPU box) so that

* ``oracle/make_golden.py`` can record golden vectors from the reference's own code, and
* ``tests/test_oracle_vs_reference.py`` can check the restatement in ``oracle/paa_oracle.py``
  against it.

Nothing here is copied from the reference; these are the import shims listed in SURVEY.md 8(c):

1. ``paa_core._C`` is a compiled extension that cannot be built against torch 2.11
   (``<THC/THC.h>`` is gone): a stub module is injected whose ``ml_nms`` is the CPU
   restatement in ``oracle/nms_oracle.py`` (the reference has no CPU ml_nms, ml_nms.h:26).
2. ``np.float`` (removed in numpy>=1.24) is used by anchor_generator.py:275,284.
3. yacs is absent: the cfg is a ``types.SimpleNamespace`` tree (see ``make_cfg``); gamma/alpha
   are 1-tuples because sigmoid_focal_loss.py:42-43 indexes them on the CPU path.
4. The post-processor needs channels-last head outputs under torch 2.11 (inference.py:49).
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("PAA_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "paa_core"))


class _RecordingC(types.ModuleType):
    """Stub for the reference's pybind module (csrc/vision.cpp:10-27)."""

    def __init__(self):
        super().__init__("paa_core._C")
        self.ml_nms_calls = []

    def ml_nms(self, dets, scores, labels, thresh):
        from oracle.nms_oracle import ml_nms_cpu
        keep = ml_nms_cpu(dets.detach().cpu().numpy(), scores.detach().cpu().numpy(),
                          labels.detach().cpu().numpy(), float(thresh))
        self.ml_nms_calls.append(keep.copy())
        return torch.from_numpy(keep).to(dets.device)

    def nms(self, dets, scores, thresh):
        from oracle.nms_oracle import ml_nms_cpu
        n = dets.shape[0]
        keep = ml_nms_cpu(dets.detach().cpu().numpy(), scores.detach().cpu().numpy(),
                          np.zeros(n, np.float32), float(thresh))
        return torch.from_numpy(keep).to(dets.device)


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's PAALossComputation, PAAPostProcessor,
    BoxCoder, BoxList, make_anchor_generator_paa and the `_C` stub."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "float"):
        np.float = float  # anchor_generator.py:275,284
    stub = _RecordingC()
    pkg = types.ModuleType("paa_core")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "paa_core")]
    pkg._C = stub
    sys.modules["paa_core"] = pkg
    sys.modules["paa_core._C"] = stub
    # yacs / apex are not needed by the modules below; import the path modules directly.
    from paa_core.structures.bounding_box import BoxList
    from paa_core.structures import boxlist_ops
    from paa_core.modeling.rpn.paa import loss as ref_loss
    from paa_core.modeling.rpn.paa import inference as ref_inference
    from paa_core.modeling.rpn.atss.atss import BoxCoder
    from paa_core.modeling.rpn.anchor_generator import make_anchor_generator_paa
    from paa_core.modeling.matcher import Matcher
    ns = types.SimpleNamespace(
        BoxList=BoxList, boxlist_ops=boxlist_ops, loss=ref_loss, inference=ref_inference,
        BoxCoder=BoxCoder, make_anchor_generator_paa=make_anchor_generator_paa,
        Matcher=Matcher, _C=stub)
    _loaded["ns"] = ns
    return ns


def make_cfg(**paa_overrides):
    """SimpleNamespace cfg with the PAA keys of config/defaults.py:292-331, :548."""
    paa = dict(
        NUM_CLASSES=81, ANCHOR_SIZES=(64, 128, 256, 512, 1024), ASPECT_RATIOS=(1.0,),
        ANCHOR_STRIDES=(8, 16, 32, 64, 128), STRADDLE_THRESH=0, OCTAVE=2.0,
        SCALES_PER_OCTAVE=1, LOSS_ALPHA=(0.25,), LOSS_GAMMA=(2.0,), IOU_THRESHOLD=0.1,
        TOPK=9, REG_LOSS_WEIGHT=1.3, PRIOR_PROB=0.01, INFERENCE_TH=0.05, NMS_TH=0.6,
        PRE_NMS_TOP_N=1000, USE_IOU_PRED=True, IOU_LOSS_WEIGHT=0.5,
        INFERENCE_SCORE_VOTING=True, REG_LOSS_TYPE="iou")
    paa.update(paa_overrides)
    ns = types.SimpleNamespace
    return ns(MODEL=ns(PAA=ns(**paa), ATSS=ns(REGRESSION_TYPE="BOX")),
              TEST=ns(DETECTIONS_PER_IMG=100, BBOX_AUG=ns(ENABLED=False, VOTE=False)))
