"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the test-time-augmentation box merging of the reference
(paa_core/engine/bbox_aug_vote.py:140-310), the step right after the post-processor in the multi-scale
configs.

Per class, detections pooled from all scales / flips are merged greedily in float32 (the reference converts
the float32 tensors to numpy and stays in float32 until the result rows are collected):

  bbox_vote       (:198-246)  repeat: the best remaining box is the pivot; every remaining box with
                              IoU(+1) >= vote_thresh to it (the pivot included) is removed; one box alone is
                              passed through, several are replaced by their score-weighted mean box carrying the
                              pivot's (= maximal) score.  Rows come out in pivot order.
  soft_bbox_vote  (:249-310)  same, and the removed boxes whose decayed score s * (1 - IoU) is still
                              >= score_thresh are kept as they are with that score; rows finally sorted by
                              score, descending.
  fewer than two boxes of a class: the class is passed through unchanged (:220-221 returns empties and
  boxlist_nms :190-192 then keeps the input).

merge (:140-176): classes 1..num_classes-1 in order, concatenated; if more than `max_detections` rows, those
with score >= the max_detections-th largest are kept (ties keep more), order preserved.
"""
import numpy as np

from oracle.nms_oracle import ml_nms_cpu


def _iou_to_first(det):
    """IoU(+1) of row 0 with every row, float32, the reference's operation order (:207-216)."""
    one = np.float32(1.0)
    area = (det[:, 2] - det[:, 0] + one) * (det[:, 3] - det[:, 1] + one)
    w = np.maximum(np.float32(0.0), np.minimum(det[0, 2], det[:, 2]) - np.maximum(det[0, 0], det[:, 0]) + one)
    h = np.maximum(np.float32(0.0), np.minimum(det[0, 3], det[:, 3]) - np.maximum(det[0, 1], det[:, 1]) + one)
    inter = w * h
    return inter / (area[0] + area - inter)


def vote_class(boxes, scores, vote_thresh, soft=False, score_thresh=0.05):
    """One class.  boxes [n,4] float32, scores [n] float32 -> (boxes [m,4] float32, scores [m] float32)."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    scores = np.asarray(scores, np.float32).reshape(-1)
    if boxes.shape[0] <= 1:
        return boxes, scores
    det = np.concatenate([boxes, scores[:, None]], axis=1)
    det = det[np.argsort(det[:, 4], kind="stable")[::-1]]           # best first (tie order: see tests)
    rows = []
    while det.shape[0] > 0:
        o = _iou_to_first(det)
        take = np.nonzero(o >= np.float32(vote_thresh))[0]
        group, group_iou = det[take], o[take]
        det = np.delete(det, take, axis=0)
        if take.shape[0] <= 1:
            rows.append(group.astype(np.float64))
            continue
        weighted = group[:, :4] * group[:, 4:5]                       # float32
        merged = np.zeros((1, 5), np.float64)
        merged[0, :4] = np.sum(weighted, axis=0) / np.sum(group[:, 4:5])
        merged[0, 4] = np.max(group[:, 4])
        rows.append(merged)
        if soft:
            decayed = group.copy()
            decayed[:, 4] = decayed[:, 4] * (np.float32(1.0) - group_iou)
            decayed = decayed[decayed[:, 4] >= score_thresh]
            if decayed.shape[0] > 0:
                rows.append(decayed.astype(np.float64))
    out = np.concatenate(rows, axis=0)
    if soft:
        out = out[np.argsort(out[:, 4], kind="stable")[::-1]]
    return out[:, :4].astype(np.float32), out[:, 4].astype(np.float32)


def merge_multi_scale(boxes, scores, labels, num_classes, merge_type="vote", vote_thresh=0.66, nms_thresh=0.6,
                      max_detections=1000, score_thresh=0.05):
    """merge_result_from_multi_scales for one image -> (boxes, scores, labels int64)."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    scores = np.asarray(scores, np.float32).reshape(-1)
    labels = np.asarray(labels).reshape(-1).astype(np.int64)
    ob, os_, ol = [], [], []
    for j in range(1, num_classes):
        idx = np.nonzero(labels == j)[0]
        b, s = boxes[idx], scores[idx]
        if merge_type == "nms":
            keep = ml_nms_cpu(b, s, np.zeros(len(idx), np.float32), float(nms_thresh))
            keep = keep[np.argsort(-s[keep], kind="stable")]          # _C.nms returns best first
            b, s = b[keep], s[keep]
        elif nms_thresh > 0:
            b, s = vote_class(b, s, vote_thresh, soft=(merge_type != "vote"), score_thresh=score_thresh)
        ob.append(b)
        os_.append(s)
        ol.append(np.full(len(s), j, np.int64))
    b, s, l = np.concatenate(ob), np.concatenate(os_), np.concatenate(ol)
    if max_detections > 0 and len(s) > max_detections:
        thr = np.sort(s)[len(s) - max_detections]                     # kthvalue(n - max + 1)
        keep = s >= thr
        b, s, l = b[keep], s[keep], l[keep]
    return b, s, l
