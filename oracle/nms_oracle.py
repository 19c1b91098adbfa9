"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's label-aware NMS.

The reference ships ``ml_nms`` for CUDA only (csrc/ml_nms.h:10-27 dispatches to
csrc/cuda/ml_nms.cu:75-136 and raises on CPU tensors), so the oracle restates its arithmetic:

* pair overlap, csrc/cuda/ml_nms.cu:13-24: boxes with different labels never suppress each other;
  otherwise IoU with the "+1" pixel convention, evaluated in float32 as
  ``inter / (Sa + Sb - inter)``;
* candidates are visited in descending score order (ml_nms.cu:79) and a box is dropped when a
  previously kept box overlaps it by *strictly more* than the threshold (ml_nms.cu:65, 116-128);
* the survivors are reported as indices into the unsorted input, ascending (ml_nms.cu:132-135).

Pinned by the reference's only golden vectors on this arithmetic, tests/test_nms.py:16-58 and
:65-217 (single-class NMS == ml_nms with uniform labels); see tests/test_oracle_nms_golden.py.
The reference's sort is unstable, so the visiting order among *equal scores* is undefined there;
this restatement breaks such ties by ascending input index (documented tie exemption (iii) of
SURVEY.md 8c).
"""
import numpy as np


def pair_iou_plus1(box, others):
    """float32 IoU(+1) of one box [4] against others [m,4]; op order of ml_nms.cu:17-23."""
    f = np.float32
    box = box.astype(f)
    others = others.astype(f)
    left = np.maximum(box[0], others[:, 0])
    right = np.minimum(box[2], others[:, 2])
    top = np.maximum(box[1], others[:, 1])
    bottom = np.minimum(box[3], others[:, 3])
    w = np.maximum(right - left + f(1), f(0))
    h = np.maximum(bottom - top + f(1), f(0))
    inter = w * h
    sa = (box[2] - box[0] + f(1)) * (box[3] - box[1] + f(1))
    sb = (others[:, 2] - others[:, 0] + f(1)) * (others[:, 3] - others[:, 1] + f(1))
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (sa + sb - inter)


def visiting_order(scores):
    """Descending score, ascending index among equals (ml_nms.cu:79; tie rule is ours)."""
    idx = np.arange(scores.shape[0])
    return np.lexsort((idx, -scores.astype(np.float64)))


def ml_nms_cpu(boxes, scores, labels, thresh):
    """boxes [n,4] f32 xyxy, scores [n] f32, labels [n] (any numeric) -> kept indices, ascending int64."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros(0, np.int64)
    order = visiting_order(scores)
    b = boxes[order].astype(np.float32)
    lab = np.asarray(labels)[order]
    dead = np.zeros(n, bool)
    kept = []
    thr = np.float32(thresh)
    for i in range(n):
        if dead[i]:
            continue
        kept.append(order[i])
        if i + 1 < n:
            ov = pair_iou_plus1(b[i], b[i + 1:])
            hit = (ov > thr) & (lab[i + 1:] == lab[i])
            dead[i + 1:] |= hit
    return np.sort(np.asarray(kept, np.int64))
