"""TEST INFRASTRUCTURE ONLY -- CPU restatement of PAAPostProcessor.

Follows paa_core/modeling/rpn/paa/inference.py: per-level candidate selection (:36-82), level
concatenation (:94-95), label-aware NMS + top-``detections_per_img`` cut (:105-122) and score voting
(:123-157), on plain CPU tensors.  ``ml_nms`` comes from ``oracle/nms_oracle.py`` because the
reference ships it for CUDA only (csrc/ml_nms.h:26).

Pinned by ``tests/golden/post_*.npz`` (recorded from the reference itself by
``oracle/make_golden.py``) and, for the suppression arithmetic, by the reference's known-answer
vectors tests/test_nms.py:16-58,:65-217.

Orders the reference leaves unspecified (``topk(sorted=False)`` inference.py:64, equal-score order
in the unstable sort ml_nms.cu:79) are compared as sets by the tests (tie exemption (iii)).
"""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import nms_oracle
from oracle.paa_oracle import decode, iou_matrix


def default_params(**kw):
    p = dict(pre_nms_thresh=0.05, pre_nms_top_n=1000, nms_thresh=0.6, detections_per_img=100,
             num_classes=81, score_voting=True, min_size=0, skip_nms=False,
             # "atss": rpn/atss/inference.py:33-80 -- per-level top-k on sigmoid(cls) * sigmoid(centerness),
             # square root taken on the selected scores; no score voting (set score_voting=False with it)
             # "retinanet": rpn/retinanet/inference.py:60-127 -- no third map, RPN BoxCoder (modeling/box_coder.py:51-95)
             # "fcos": rpn/fcos/inference.py:48-104 -- anchors are points (x, y, x, y), ltrb distances, product
             #         top-k then square root
             flavour="paa", decode_weights=(10.0, 10.0, 5.0, 5.0))
    p.update(kw)
    return SimpleNamespace(**p)


def decode_legacy(rel_codes, boxes, weights, clip=None):
    """modeling/box_coder.py:51-95 (TO_REMOVE = 1; x2 = cx + w/2 - 1)."""
    import math
    clip = math.log(1000.0 / 16) if clip is None else clip
    w = boxes[:, 2] - boxes[:, 0] + 1
    h = boxes[:, 3] - boxes[:, 1] + 1
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx, dy = rel_codes[:, 0] / wx, rel_codes[:, 1] / wy
    dw = torch.clamp(rel_codes[:, 2] / ww, max=clip)
    dh = torch.clamp(rel_codes[:, 3] / wh, max=clip)
    pcx, pcy = dx * w + cx, dy * h + cy
    pw, ph = torch.exp(dw) * w, torch.exp(dh) * h
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw - 1, pcy + 0.5 * ph - 1], dim=1)


def level_candidates(box_cls, box_regression, iou_pred, anchors, image_sizes, prm):
    """inference.py:36-82 for one level -> per image (boxes [k,4], scores [k], labels [k])."""
    N, _, H, W = box_cls.shape
    A = box_regression.shape[1] // 4                                            # anchors per location
    C = box_cls.shape[1] // A
    # permute_and_flatten (rpn/utils.py:10-14): [N, A*C, H, W] -> [N, H*W*A, C]
    prob = box_cls.view(N, A, C, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, C).sigmoid()   # :42-43
    reg = box_regression.view(N, A, 4, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, 4)       # :45-46
    cand = prob > prm.pre_nms_thresh                                            # :48
    k_per_im = cand.reshape(N, -1).sum(1).clamp(max=prm.pre_nms_top_n)          # :49-50
    flavour = getattr(prm, "flavour", "paa")
    atss = flavour in ("atss", "fcos")
    if iou_pred is not None:
        q = iou_pred.view(N, A, 1, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1).sigmoid()   # :54-55
        prob = prob * q[:, :, None]                                             # atss/inference.py:53
        if not atss:
            prob = prob.sqrt()                                                  # :56
    out = []
    for i in range(N):
        s = prob[i][cand[i]]                                                    # :62
        s, pick = s.topk(int(k_per_im[i]), sorted=False)                        # :64
        if atss:
            s = torch.sqrt(s)                                                   # atss/inference.py:75
        where = cand[i].nonzero()[pick, :]                                      # :66
        loc = where[:, 0]
        labels = where[:, 1] + 1                                                # :69
        if flavour == "retinanet":
            boxes = decode_legacy(reg[i][loc, :].view(-1, 4), anchors[loc, :].view(-1, 4), prm.decode_weights)
        elif flavour == "fcos":
            pts, d = anchors[loc, :].view(-1, 4), reg[i][loc, :].view(-1, 4)
            boxes = torch.stack([pts[:, 0] - d[:, 0], pts[:, 1] - d[:, 1], pts[:, 2] + d[:, 2], pts[:, 3] + d[:, 3]],
                                dim=1)                                          # fcos/inference.py:93-98
        else:
            boxes = decode(reg[i][loc, :].view(-1, 4), anchors[loc, :].view(-1, 4))  # :71-74
        w, h = image_sizes[i]
        boxes[:, 0].clamp_(min=0, max=w - 1)                                    # bounding_box.py:214-219
        boxes[:, 1].clamp_(min=0, max=h - 1)
        boxes[:, 2].clamp_(min=0, max=w - 1)
        boxes[:, 3].clamp_(min=0, max=h - 1)
        ws = boxes[:, 2] - boxes[:, 0] + 1                                      # boxlist_ops.py:62-76
        hs = boxes[:, 3] - boxes[:, 1] + 1
        keep = ((ws >= prm.min_size) & (hs >= prm.min_size)).nonzero().squeeze(1)
        out.append((boxes[keep], s[keep], labels[keep]))
    return out


def vote_boxes(det_boxes, det_labels, all_boxes, all_scores, all_labels, num_classes, sigma=0.025):
    """inference.py:123-157: every kept box becomes the p-weighted mean of all pre-NMS boxes of its
    class with IoU(+1) > 0.01, p = exp(-(1-IoU)^2/sigma) * score."""
    out = det_boxes.clone()
    for j in range(1, num_classes):
        src = (all_labels == j).nonzero().view(-1)
        dst = (det_labels == j).nonzero().view(-1)
        if dst.numel() == 0:
            continue
        bj = all_boxes[src, :].view(-1, 4)
        sj = all_scores[src]
        ious = iou_matrix(det_boxes[dst], bj)                                   # :137
        for r in range(dst.numel()):
            near = (ious[r] > 0.01).nonzero().squeeze(1)                        # :141
            p = (torch.exp(-(1 - ious[r][near]) ** 2 / sigma) * sj[near]).unsqueeze(1)   # :145
            out[dst[r]] = torch.sum(bj[near] * p, dim=0) / torch.sum(p, dim=0)  # :146
    return out


def select_over_all_levels(boxes, scores, labels, prm):
    """inference.py:105-159 for one image.  Returns (boxes, scores, labels, keep)."""
    keep = torch.from_numpy(nms_oracle.ml_nms_cpu(boxes.numpy(), scores.numpy(),
                                                  labels.float().numpy(), prm.nms_thresh))  # :110
    n_det = keep.numel()
    if n_det > prm.detections_per_img > 0:                                      # :114
        ks = scores[keep]
        thr, _ = torch.kthvalue(ks, n_det - prm.detections_per_img + 1)         # :116-119
        keep = keep[torch.nonzero(ks >= thr.item()).squeeze(1)]                 # :120-122
    db, ds, dl = boxes[keep], scores[keep], labels[keep]
    if prm.score_voting:
        db = vote_boxes(db, dl, boxes, scores, labels, prm.num_classes)
    return db, ds, dl, keep


def postprocess(box_cls, box_regression, iou_pred, anchors_per_level, image_sizes, params=None):
    """inference.py:84-99.  Returns per image a namespace with the pre-NMS candidate list (level-major)
    and the final detections."""
    prm = params or default_params()
    per_level = []
    for l in range(len(box_cls)):
        per_level.append(level_candidates(box_cls[l], box_regression[l],
                                          None if iou_pred is None else iou_pred[l],
                                          anchors_per_level[l], image_sizes, prm))
    results = []
    for i in range(box_cls[0].shape[0]):
        b = torch.cat([per_level[l][i][0] for l in range(len(box_cls))])
        s = torch.cat([per_level[l][i][1] for l in range(len(box_cls))])
        lab = torch.cat([per_level[l][i][2] for l in range(len(box_cls))])
        counts = [int(per_level[l][i][1].numel()) for l in range(len(box_cls))]
        if prm.skip_nms:                                                        # :96-97
            results.append(SimpleNamespace(pre_boxes=b, pre_scores=s, pre_labels=lab, level_counts=counts,
                                           boxes=b, scores=s, labels=lab, keep=torch.arange(s.numel())))
            continue
        db, ds, dl, keep = select_over_all_levels(b, s, lab, prm)
        results.append(SimpleNamespace(pre_boxes=b, pre_scores=s, pre_labels=lab, level_counts=counts,
                                       boxes=db, scores=ds, labels=dl, keep=keep))
    return results


def canonical_rows(boxes, scores, labels):
    """Order-independent view of a detection set: rows sorted by (label, score, x1, y1, x2, y2)."""
    b = np.asarray(boxes, np.float64).reshape(-1, 4)
    s = np.asarray(scores, np.float64).reshape(-1)
    l = np.asarray(labels, np.float64).reshape(-1)
    order = np.lexsort((b[:, 3], b[:, 2], b[:, 1], b[:, 0], s, l))
    return b[order], s[order], l[order].astype(np.int64)
