"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PAA assign + loss path.

This module is the parity oracle for ``paa_b200``'s CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; the product (``paa_b200``) never does and has no CPU fallback.

It restates, with torch CPU ops in the reference's own float32 operation order, what
``PAALossComputation.__call__`` (paa_core/modeling/rpn/paa/loss.py:267-359) computes, stage by
stage, and returns every intermediate so each CUDA kernel can be checked "teacher-forced" against
the previous stage's oracle output.  Each function cites the reference lines it
follows.  The per-GT Gaussian-mixture fit is third-party arithmetic (scikit-learn, unpinned in the
reference's requirements.txt:10; 1.9.0 in this image): see ``oracle/gmm_oracle.py``.

Pinning: the reference's own tests hold no golden vector for this path, so the
restatement is pinned against outputs of the reference itself, run in the build container with the
import shims of ``oracle/ref_shim.py`` and committed as ``tests/golden/loss_*.npz`` by
``oracle/make_golden.py``; ``tests/test_oracle_golden.py`` replays them (bit-exact for indices and
masks, <=1e-6 for float tensors) and ``tests/test_oracle_vs_reference.py`` re-runs the live
comparison whenever ``/root/reference`` is present.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

from oracle import gmm_oracle

INF = 100000000  # loss.py:15
BBOX_XFORM_CLIP = math.log(1000.0 / 16)  # atss.py:84-85


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------
def area_plus1(b):
    """bounding_box.py:226-231 (xyxy, TO_REMOVE=1)."""
    return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)


def iou_matrix(gt, anchors):
    """boxlist_ops.py:99-116: [G,4] x [A,4] -> [G,A], "+1" convention, float32."""
    a1 = area_plus1(gt)
    a2 = area_plus1(anchors)
    lt = torch.max(gt[:, None, :2], anchors[:, :2])
    rb = torch.min(gt[:, None, 2:], anchors[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (a1[:, None] + a2 - inter)


def match_anchors(iou, thr):
    """matcher.py:42-113 with high == low == thr and allow_low_quality_matches=True
    (constructed that way at loss.py:38-40).  Returns int64 [A] in {-1, 0..G-1}."""
    if iou.numel() == 0:
        raise ValueError("No ground-truth boxes available for one of the images during training"
                         if iou.shape[0] == 0 else
                         "No proposal boxes available for one of the images during training")
    best_val, best_gt = iou.max(dim=0)                       # matcher.py:66
    matched = best_gt.clone()
    matched[best_val < thr] = -1                             # :71,75 (between-set is empty: hi==lo)
    gt_best = iou.max(dim=1).values                          # :92
    is_a_gt_best = (iou == gt_best[:, None]).any(dim=0)      # :94-97, ties included
    matched[is_a_gt_best] = best_gt[is_a_gt_best]            # :112-113
    return matched


def encode(gt, anchors):
    """atss.py:33-50 (REGRESSION_TYPE 'BOX', weights 10,10,5,5)."""
    ew = anchors[:, 2] - anchors[:, 0] + 1
    eh = anchors[:, 3] - anchors[:, 1] + 1
    ecx = (anchors[:, 2] + anchors[:, 0]) / 2
    ecy = (anchors[:, 3] + anchors[:, 1]) / 2
    gw = gt[:, 2] - gt[:, 0] + 1
    gh = gt[:, 3] - gt[:, 1] + 1
    gcx = (gt[:, 2] + gt[:, 0]) / 2
    gcy = (gt[:, 3] + gt[:, 1]) / 2
    return torch.stack((10.0 * (gcx - ecx) / ew, 10.0 * (gcy - ecy) / eh,
                        5.0 * torch.log(gw / ew), 5.0 * torch.log(gh / eh)), dim=1)


def decode(deltas, anchors):
    """atss.py:68-96."""
    anchors = anchors.to(deltas.dtype)
    w = anchors[:, 2] - anchors[:, 0] + 1
    h = anchors[:, 3] - anchors[:, 1] + 1
    cx = (anchors[:, 2] + anchors[:, 0]) / 2
    cy = (anchors[:, 3] + anchors[:, 1]) / 2
    dx = deltas[:, 0] / 10.0
    dy = deltas[:, 1] / 10.0
    dw = torch.clamp(deltas[:, 2] / 5.0, max=BBOX_XFORM_CLIP)
    dh = torch.clamp(deltas[:, 3] / 5.0, max=BBOX_XFORM_CLIP)
    pcx = dx * w + cx
    pcy = dy * h + cy
    pw = torch.exp(dw) * w
    ph = torch.exp(dh) * h
    return torch.stack((pcx - 0.5 * (pw - 1), pcy - 0.5 * (ph - 1),
                        pcx + 0.5 * (pw - 1), pcy + 0.5 * (ph - 1)), dim=1)


def aligned_iou_plus1(b1, b2):
    """loss.py:258-265."""
    a1 = area_plus1(b1)
    a2 = area_plus1(b2)
    lt = torch.max(b1[:, :2], b2[:, :2])
    rb = torch.min(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    return inter / (a1 + a2 - inter)


def giou_loss(pred_deltas, target_deltas, anchors, weight=None):
    """loss.py:46-87: 1 - GIoU between decode(pred) and decode(target), no "+1"."""
    p = decode(pred_deltas.view(-1, 4), anchors.view(-1, 4))
    px1, py1 = p[:, 0], p[:, 1]
    px2 = torch.max(px1, p[:, 2])
    py2 = torch.max(py1, p[:, 3])
    p_area = (px2 - px1) * (py2 - py1)
    t = decode(target_deltas.view(-1, 4), anchors.view(-1, 4))
    tx1, ty1, tx2, ty2 = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    t_area = (tx2 - tx1) * (ty2 - ty1)
    ix1 = torch.max(px1, tx1)
    iy1 = torch.max(py1, ty1)
    ix2 = torch.min(px2, tx2)
    iy2 = torch.min(py2, ty2)
    overlap = (iy2 > iy1) & (ix2 > ix1)
    inter = torch.where(overlap, (ix2 - ix1) * (iy2 - iy1), torch.zeros_like(ix1))
    ex1 = torch.min(px1, tx1)
    ey1 = torch.min(py1, ty1)
    ex2 = torch.max(px2, tx2)
    ey2 = torch.max(py2, ty2)
    enclosing = (ex2 - ex1) * (ey2 - ey1) + 1e-7
    union = p_area + t_area - inter + 1e-7
    iou = inter / union
    giou = iou - (enclosing - union) / enclosing
    loss = 1 - giou
    if weight is not None and weight.sum() > 0:                # loss.py:83-84
        return loss * weight
    return loss


# ----------------------------------------------------------------------------------------------
# focal loss and head layout
# ----------------------------------------------------------------------------------------------
def focal_loss_cpu(logits, targets, gamma, alpha):
    """sigmoid_focal_loss.py:40-52, the formula the reference uses on CPU tensors:
    raw log(p) / log(1-p) in float32.  targets int32 [n]: class c>0 is positive for column c-1,
    t>=0 marks the other columns negative, t<0 would be ignored."""
    C = logits.shape[1]
    cols = torch.arange(1, C + 1, dtype=targets.dtype).unsqueeze(0)
    t = targets.unsqueeze(1)
    p = torch.sigmoid(logits)
    pos_term = (1 - p) ** gamma * torch.log(p)
    neg_term = p ** gamma * torch.log(1 - p)
    is_pos = (t == cols).float()
    is_neg = ((t != cols) * (t >= 0)).float()
    return -is_pos * pos_term * alpha - is_neg * neg_term * (1 - alpha)


def flatten_level(x, channels):
    """rpn/utils.py:10-14 for one anchor per location: [N,C,H,W] -> [N,H*W,C]."""
    n = x.shape[0]
    return x.permute(0, 2, 3, 1).reshape(n, -1, channels)


def flatten_heads(box_cls, box_regression, iou_pred):
    """rpn/utils.py:17-45 and loss.py:283-284: level-concatenated [N*A,C], [N*A,4], [N*A]."""
    C = box_cls[0].shape[1]
    cls = torch.cat([flatten_level(x, C) for x in box_cls], dim=1).reshape(-1, C)
    reg = torch.cat([flatten_level(x, 4) for x in box_regression], dim=1).reshape(-1, 4)
    iou = None
    if iou_pred is not None:
        n = iou_pred[0].shape[0]
        iou = torch.cat([x.permute(0, 2, 3, 1).reshape(n, -1, 1) for x in iou_pred], dim=1).reshape(-1)
    return cls, reg, iou


# ----------------------------------------------------------------------------------------------
# stage 1: IoU-based pre-assignment (loss.py:89-126)
# ----------------------------------------------------------------------------------------------
def iou_based_targets(gt_boxes, gt_labels, anchors_cat, thr):
    labels, reg_targets, matched = [], [], []
    for boxes, cls in zip(gt_boxes, gt_labels):
        q = iou_matrix(boxes, anchors_cat)
        m = match_anchors(q, thr)
        lab = cls[m.clamp(min=0)].to(torch.float32)
        lab[m == -1] = 0                                        # loss.py:112-113
        labels.append(lab)
        matched.append(m)
        reg_targets.append(encode(boxes[m.clamp(min=0)], anchors_cat))
    return labels, reg_targets, matched


# ----------------------------------------------------------------------------------------------
# stage 3: candidate selection (loss.py:151-178)
# ----------------------------------------------------------------------------------------------
def select_candidates(loss_im, label_im, matched_im, level_sizes, num_gt, topk):
    """Per GT, per level: the (at most) ``topk`` smallest-loss anchors among those matched to the GT
    with a positive IoU-label.  Returns a list (len G) of int64 index tensors (level-major concat,
    within a level in ``torch.topk`` order) or None."""
    out = []
    for g in range(num_gt):
        parts = []
        start = 0
        for n_l in level_sizes:
            sl = slice(start, start + n_l)
            hit = torch.nonzero((matched_im[sl] == g) & (label_im[sl] > 0), as_tuple=False)[:, 0]
            if hit.numel() > 0:
                _, pick = loss_im[sl][hit].topk(min(hit.numel(), topk), largest=False)
                parts.append(hit[pick] + start)
            start += n_l
        out.append(torch.cat(parts) if parts else None)
    return out


# ----------------------------------------------------------------------------------------------
# stage 4: per-GT mixture fit and labelling (loss.py:180-236)
# ----------------------------------------------------------------------------------------------
def paa_labels_for_image(cands, loss_im, matched_im, boxes, cls, n_anchors, gmm_impl="sklearn",
                         record=None):
    """Returns (labels int64 [A], matched_gt_boxes [A,4]).  ``record`` (a list) receives one dict per
    GT with the sorted candidate losses / indices, GMM parameters and the positive count."""
    labels = torch.zeros(n_anchors, dtype=torch.long)            # loss.py:182
    matched_boxes = torch.zeros((n_anchors, 4), dtype=boxes.dtype)
    fg = matched_im >= 0
    matched_boxes[fg] = boxes[matched_im[fg]]                    # :183-185
    for g, cand in enumerate(cands):
        rec = dict(gt=g, n=0)
        if cand is not None:
            if cand.numel() > 1:
                vals, order = loss_im[cand].sort()               # :190-191
                x = vals.view(-1, 1).numpy()
                fit = gmm_oracle.fit_two_component(x, impl=gmm_impl)   # :193-203
                n_pos = gmm_oracle.positive_prefix_length(fit)   # :206-217
                pos = cand[order[:n_pos]]
                labels[cand] = 0                                 # :212,225-227 (fg|bgs == everything)
                rec.update(n=int(cand.numel()), sorted_idx=cand[order].numpy().copy(),
                           sorted_loss=vals.numpy().copy(), n_pos=int(n_pos), fit=fit)
            else:
                pos = cand[0:1]                                  # :219 (is_pos = 0)
                rec.update(n=1, sorted_idx=cand.numpy().copy(),
                           sorted_loss=loss_im[cand].numpy().copy(), n_pos=1, fit=None)
            labels[pos] = cls[g]                                 # :228-229
            matched_boxes[pos] = boxes[g]                        # :230
        if record is not None:
            record.append(rec)
    return labels, matched_boxes


# ----------------------------------------------------------------------------------------------
# whole path
# ----------------------------------------------------------------------------------------------
def default_params(**kw):
    p = dict(gamma=2.0, alpha=0.25, iou_threshold=0.1, topk=9, reg_loss_weight=1.3,
             iou_loss_weight=0.5, use_iou_pred=True)
    p.update(kw)
    return SimpleNamespace(**p)


def assign(box_cls, box_regression, iou_pred, gt_boxes, gt_labels, anchors_per_level, params=None,
           gmm_impl="sklearn", combined_loss_override=None):
    """Stages 1-4 for the images of one rank: everything up to the PAA labels and this rank's
    partial normalisers.  Inputs are plain CPU tensors; ``anchors_per_level`` is the per-level anchor
    list shared by all images (anchor_generator.py:112-125 hands every image the same tensors).
    ``combined_loss_override`` ([N,A]) teacher-forces stage 2's output."""
    prm = params or default_params()
    anchors_cat = torch.cat(list(anchors_per_level), dim=0)      # loss.py:100,145
    level_sizes = [a.shape[0] for a in anchors_per_level]
    A = anchors_cat.shape[0]
    N = len(gt_boxes)
    with torch.no_grad():
        iou_labels, iou_reg_targets, matched = iou_based_targets(gt_boxes, gt_labels, anchors_cat,
                                                                 prm.iou_threshold)
        matched_all = torch.stack(matched, dim=0)                # loss.py:273
        iou_labels_flat = torch.cat(iou_labels).int()            # :276
        iou_reg_flat = torch.cat(iou_reg_targets)
        cls_flat, reg_flat, iou_flat = flatten_heads(box_cls, box_regression,
                                                     iou_pred if prm.use_iou_pred else None)
        anchors_flat = anchors_cat.repeat(N, 1)                  # :280-281
        pos = torch.nonzero(iou_labels_flat > 0, as_tuple=False).squeeze(1)
        if pos.numel() == 0:
            raise RuntimeError("no IoU-positive anchor in the batch (loss.py:351-356 NameError)")
        cls_score = focal_loss_cpu(cls_flat.detach(), iou_labels_flat, prm.gamma, prm.alpha)  # :293
        reg_score = giou_loss(reg_flat.detach(), iou_reg_flat, anchors_flat)[iou_labels_flat > 0]  # :296,256
        reg_full = torch.full((cls_score.shape[0],), float(INF), dtype=cls_score.dtype)
        reg_full[pos] = reg_score.view(-1, 1).mean(1)            # :301-305
        combined = cls_score.sum(dim=1) + reg_full               # :306
        assert not torch.isnan(combined).any()                   # :307
        combined = combined.view(N, A)
        used_loss = combined if combined_loss_override is None else combined_loss_override
        labels, matched_boxes, records, cand_lists = [], [], [], []
        for i in range(N):
            cands = select_candidates(used_loss[i], iou_labels_flat.view(N, A)[i], matched_all[i],
                                      level_sizes, gt_boxes[i].shape[0], prm.topk)
            rec = []
            lab, mb = paa_labels_for_image(cands, used_loss[i], matched_all[i], gt_boxes[i],
                                           gt_labels[i], A, gmm_impl=gmm_impl, record=rec)
            labels.append(lab)
            matched_boxes.append(mb)
            records.append(rec)
            cand_lists.append(cands)
        labels_flat = torch.cat(labels).int()                    # :318
        reg_targets_flat = torch.cat([encode(mb, anchors_cat) for mb in matched_boxes])  # :232,319
        pos_inds = torch.nonzero(labels_flat > 0, as_tuple=False).squeeze(1)            # :320
        if prm.use_iou_pred and pos_inds.numel() > 0:
            gt_dec = decode(reg_targets_flat[pos_inds], anchors_flat[pos_inds])         # :331
            pr_dec = decode(reg_flat[pos_inds], anchors_flat[pos_inds])                 # :332
            ious = aligned_iou_plus1(gt_dec, pr_dec)                                   # :333
        else:
            ious = torch.zeros(0)
    return SimpleNamespace(
        N=N, A=A, level_sizes=level_sizes, anchors_cat=anchors_cat, anchors_flat=anchors_flat,
        matched_idx=matched_all, iou_labels=iou_labels_flat.view(N, A), combined_loss=combined,
        candidates=cand_lists, gmm_records=records, paa_labels=labels_flat.view(N, A),
        reg_targets=reg_targets_flat, pos_inds=pos_inds, pos_ious=ious,
        num_pos=int(pos_inds.numel()), sum_iou=float(ious.sum()) if ious.numel() else 0.0,
        params=prm)


def with_labels(asg, paa_labels, box_regression, gt_boxes):
    """Teacher-forces stage 4's output: a copy of the assignment `asg` whose PAA labels are `paa_labels`
    ([N, A] integer array / tensor) instead of the oracle's own, with everything stage 5 derives from the
    labels rebuilt the way loss.py:230-232,318-333 does (a positive's regression target is its matched GT --
    candidates come from `matched == g`, loss.py:163 --, `pos_inds`, the aligned IoUs and both normalisers).
    Lets a test compare losses / gradients when a documented tie flipped a positive set on the device."""
    prm = asg.params
    N, A = asg.N, asg.A
    labels = torch.as_tensor(np.asarray(paa_labels)).reshape(N, A).to(torch.int32)
    with torch.no_grad():
        matched_boxes = [gt_boxes[i][asg.matched_idx[i].clamp(min=0)] for i in range(N)]          # :183-185,230
        reg_targets_flat = torch.cat([encode(mb, asg.anchors_cat) for mb in matched_boxes])       # :232,319
        labels_flat = labels.reshape(-1)
        pos_inds = torch.nonzero(labels_flat > 0, as_tuple=False).squeeze(1)
        if prm.use_iou_pred and pos_inds.numel() > 0:
            reg_flat = torch.cat([flatten_level(x, 4) for x in box_regression], dim=1).reshape(-1, 4)
            gt_dec = decode(reg_targets_flat[pos_inds], asg.anchors_flat[pos_inds])
            pr_dec = decode(reg_flat.detach()[pos_inds], asg.anchors_flat[pos_inds])
            ious = aligned_iou_plus1(gt_dec, pr_dec)
        else:
            ious = torch.zeros(0)
    out = SimpleNamespace(**vars(asg))
    out.paa_labels = labels
    out.reg_targets = reg_targets_flat
    out.pos_inds = pos_inds
    out.pos_ious = ious
    out.num_pos = int(pos_inds.numel())
    out.sum_iou = float(ious.sum()) if ious.numel() else 0.0
    return out


def losses_and_grads(box_cls, box_regression, iou_pred, asg):
    """Stage 5 + backward of the summed losses on fresh leaves; returns (losses, grads namespace)."""
    box_cls = [x.detach().clone().requires_grad_(True) for x in box_cls]
    box_regression = [x.detach().clone().requires_grad_(True) for x in box_regression]
    if iou_pred is not None:
        iou_pred = [x.detach().clone().requires_grad_(True) for x in iou_pred]
    ls = losses(box_cls, box_regression, iou_pred, asg)
    sum(ls).backward()
    z = lambda xs: [x.grad if x.grad is not None else torch.zeros_like(x) for x in xs]       # noqa: E731
    grads = SimpleNamespace(box_cls=z(box_cls), box_regression=z(box_regression),
                            iou_pred=None if iou_pred is None else z(iou_pred))
    return [l.detach() for l in ls], grads


def losses(box_cls, box_regression, iou_pred, asg, total_num_pos=None, total_sum_iou=None,
           world_size=1):
    """Stage 5 (loss.py:317-358).  ``total_*`` are the all-reduced normalisers (loss.py:321,338);
    with world_size == 1 they default to this rank's own.  The returned losses carry autograd
    graphs to the head outputs, exactly like the reference's."""
    prm = asg.params
    cls_flat, reg_flat, iou_flat = flatten_heads(box_cls, box_regression,
                                                 iou_pred if prm.use_iou_pred else None)
    total_num_pos = asg.num_pos if total_num_pos is None else total_num_pos
    num_pos_avg = max(total_num_pos / float(world_size), 1.0)    # :322
    pos = asg.pos_inds
    labels_flat = asg.paa_labels.reshape(-1)
    reg_p = reg_flat[pos]
    tgt_p = asg.reg_targets[pos]
    anc_p = asg.anchors_flat[pos]
    out = []
    if prm.use_iou_pred:
        ious = asg.pos_ious
        bce = torch.nn.functional.binary_cross_entropy_with_logits(iou_flat[pos], ious, reduction="sum")
        iou_pred_loss = bce / num_pos_avg * prm.iou_loss_weight  # :336-337
        total_sum_iou = asg.sum_iou if total_sum_iou is None else total_sum_iou
        reg_norm = total_sum_iou / float(world_size)             # :338,354
        weight = ious
    else:
        reg_norm = num_pos_avg
        weight = None
    reg_loss = giou_loss(reg_p, tgt_p, anc_p, weight=weight)[labels_flat[pos] > 0].view(-1)  # :345
    cls_loss = focal_loss_cpu(cls_flat, labels_flat.int(), prm.gamma, prm.alpha)           # :350
    out.append(cls_loss.sum() / num_pos_avg)                     # :355
    out.append(reg_loss.sum() / reg_norm * prm.reg_loss_weight)  # :356
    if prm.use_iou_pred:
        out.append(iou_pred_loss)
    return out


def assign_and_loss(box_cls, box_regression, iou_pred, gt_boxes, gt_labels, anchors_per_level,
                    params=None, gmm_impl="sklearn", with_grad=True):
    """Single-rank forward (+ backward of the summed losses, as the trainer does,
    engine/trainer.py:71-80).  Returns (losses, grads-or-None, assignment)."""
    leaves = None
    if with_grad:
        box_cls = [x.detach().clone().requires_grad_(True) for x in box_cls]
        box_regression = [x.detach().clone().requires_grad_(True) for x in box_regression]
        if iou_pred is not None:
            iou_pred = [x.detach().clone().requires_grad_(True) for x in iou_pred]
        leaves = (box_cls, box_regression, iou_pred)
    asg = assign(box_cls, box_regression, iou_pred, gt_boxes, gt_labels, anchors_per_level,
                 params=params, gmm_impl=gmm_impl)
    ls = losses(box_cls, box_regression, iou_pred, asg)
    grads = None
    if with_grad:
        sum(ls).backward()
        grads = SimpleNamespace(
            box_cls=[x.grad if x.grad is not None else torch.zeros_like(x) for x in leaves[0]],
            box_regression=[x.grad if x.grad is not None else torch.zeros_like(x) for x in leaves[1]],
            iou_pred=None if leaves[2] is None else
            [x.grad if x.grad is not None else torch.zeros_like(x) for x in leaves[2]])
    return [l.detach() for l in ls], grads, asg
