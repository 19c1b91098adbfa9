"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the plain RPN training loss of the reference
(paa_core/modeling/rpn/loss.py:21-157 with paa_core/modeling/balanced_positive_negative_sampler.py:5-68), the sibling
of the PAA / RetinaNet losses that shares their IoU matching and Matcher.

  prepare_targets (rpn/loss.py:56-95)  per image: boxlist_iou, Matcher(FG, BG, allow_low_quality_matches=True);
      label = 1 where an anchor is matched, 0 below BG, then -1 for anchors that are not "visible" (outside the image by
      more than the straddle threshold, anchor_generator.py:100-110) and -1 between the thresholds -- in that order;
      regression target = BoxCoder(1, 1, 1, 1).encode(matched GT, anchor) (box_coder.py:22-50).
  sampler (balanced_positive_negative_sampler.py:35-68)  per image at most BATCH_SIZE_PER_IMAGE * POSITIVE_FRACTION
      positives and the rest negatives, chosen by torch.randperm -- TWO randperm calls per image, positives first; the
      restatement makes the same calls in the same order, so under the same torch seed it draws the same sample.
  losses (rpn/loss.py:98-137)  smooth-L1(beta = 1/9) summed over the sampled positives / number of sampled anchors;
      binary cross-entropy with logits, mean over the sampled anchors.

Parity pin: tests/test_oracle_rpn_vs_reference.py runs the reference's own RPNLossComputation (imported through
oracle/ref_shim.py) under the same seed and finds labels, sample, losses and gradients identical.
"""
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from oracle import paa_oracle as P
from oracle.retinanet_oracle import (BELOW_LOW_THRESHOLD, BETWEEN_THRESHOLDS, encode_legacy, flatten_heads,
                                     match_anchors, smooth_l1_sum)


def default_params(**kw):
    p = dict(fg_iou_threshold=0.7, bg_iou_threshold=0.3, batch_size_per_image=256, positive_fraction=0.5,
             weights=(1.0, 1.0, 1.0, 1.0), beta=1.0 / 9)
    p.update(kw)
    return SimpleNamespace(**p)


def assign(gt_boxes, anchors_per_level, visibility, params=None):
    """Labels (float32: 1 / 0 / -1), Matcher results and regression targets per image (rpn/loss.py:56-95).
    ``visibility``: bool [A] (shared by the images) or a list of per-image bool [A] tensors."""
    prm = params or default_params()
    anchors_cat = torch.cat(list(anchors_per_level), dim=0)
    labels, matched, reg_t = [], [], []
    for i, gb in enumerate(gt_boxes):
        vis = visibility[i] if isinstance(visibility, (list, tuple)) else visibility
        m = match_anchors(P.iou_matrix(gb, anchors_cat), prm.fg_iou_threshold, prm.bg_iou_threshold)
        lab = (m >= 0).to(torch.float32)                         # generate_rpn_labels, :140-143
        lab[m == BELOW_LOW_THRESHOLD] = 0                        # :71-72
        lab[~vis.bool()] = -1                                    # :75-76
        lab[m == BETWEEN_THRESHOLDS] = -1                        # :79-81
        labels.append(lab)
        matched.append(m)
        reg_t.append(encode_legacy(gb[m.clamp(min=0)], anchors_cat, prm.weights))      # :84-86
    return SimpleNamespace(N=len(gt_boxes), A=anchors_cat.shape[0], labels=labels, matched=matched,
                           reg_targets=torch.cat(reg_t), params=prm)


def sample(labels, params):
    """balanced_positive_negative_sampler.py:35-68 + rpn/loss.py:112-116: global indices of the sampled positives and
    negatives (ascending, the order torch.nonzero gives the masks).  Consumes torch's global CPU generator exactly like
    the reference: randperm(#positives) then randperm(#negatives) per image."""
    pos_masks, neg_masks = [], []
    for lab in labels:
        positive = torch.nonzero(lab >= 1).squeeze(1)
        negative = torch.nonzero(lab == 0).squeeze(1)
        num_pos = min(positive.numel(), int(params.batch_size_per_image * params.positive_fraction))
        num_neg = min(negative.numel(), params.batch_size_per_image - num_pos)
        perm1 = torch.randperm(positive.numel())[:num_pos]
        perm2 = torch.randperm(negative.numel())[:num_neg]
        pm = torch.zeros_like(lab, dtype=torch.uint8)
        nm = torch.zeros_like(lab, dtype=torch.uint8)
        pm[positive[perm1]] = 1
        nm[negative[perm2]] = 1
        pos_masks.append(pm)
        neg_masks.append(nm)
    pos = torch.nonzero(torch.cat(pos_masks)).squeeze(1)
    neg = torch.nonzero(torch.cat(neg_masks)).squeeze(1)
    return pos, neg


def losses(objectness, box_regression, asg, sampled_pos, sampled_neg):
    """rpn/loss.py:118-137 -> [objectness_loss, box_loss] with autograd graphs."""
    prm = asg.params
    obj_flat, reg_flat = flatten_heads(objectness, box_regression)
    obj_flat = obj_flat.squeeze()
    labels = torch.cat(asg.labels)
    sampled = torch.cat([sampled_pos, sampled_neg])
    box_loss = smooth_l1_sum(reg_flat[sampled_pos], asg.reg_targets[sampled_pos], prm.beta) / sampled.numel()
    obj_loss = F.binary_cross_entropy_with_logits(obj_flat[sampled], labels[sampled])
    return [obj_loss, box_loss]


def assign_and_loss(objectness, box_regression, gt_boxes, anchors_per_level, visibility, params=None, with_grad=True,
                    sampled=None):
    """The whole call.  ``sampled`` = (pos, neg) replays a given sample instead of drawing one."""
    leaves = None
    if with_grad:
        objectness = [x.detach().clone().requires_grad_(True) for x in objectness]
        box_regression = [x.detach().clone().requires_grad_(True) for x in box_regression]
        leaves = (objectness, box_regression)
    asg = assign(gt_boxes, anchors_per_level, visibility, params)
    pos, neg = sampled if sampled is not None else sample(asg.labels, asg.params)
    asg.sampled_pos, asg.sampled_neg = pos, neg
    ls = losses(objectness, box_regression, asg, pos, neg)
    grads = None
    if with_grad:
        sum(ls).backward()
        grads = SimpleNamespace(objectness=[x.grad for x in leaves[0]], box_regression=[x.grad for x in leaves[1]])
    return [l.detach() for l in ls], grads, asg
