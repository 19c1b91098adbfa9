"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the RetinaNet training path of the reference
(paa_core/modeling/rpn/retinanet/loss.py on top of rpn/loss.py:41-88), SURVEY.md 8f-2: the loss that shares the
PAA path's IoU matching and focal machinery but labels anchors straight from the Matcher.

  prepare_targets (rpn/loss.py:56-88)  per image: boxlist_iou, Matcher(FG, BG, allow_low_quality_matches=True)
      (matcher.py:42-113); label = the matched GT's class, 0 below BG, -1 (ignored) between the thresholds;
      regression target = BoxCoder(10, 10, 5, 5).encode(matched GT, anchor) (box_coder.py:22-50).
  losses (retinanet/loss.py:45-81)  smooth-L1(beta) summed over the positives / max(1, num_pos * BBOX_REG_WEIGHT);
      sigmoid focal loss over all anchors (label -1 contributes nothing) / (num_pos + N).
"""
from types import SimpleNamespace

import torch

from oracle import paa_oracle as P

BELOW_LOW_THRESHOLD = -1
BETWEEN_THRESHOLDS = -2


def default_params(**kw):
    p = dict(gamma=2.0, alpha=0.25, fg_iou_threshold=0.5, bg_iou_threshold=0.4, bbox_reg_beta=0.11,
             bbox_reg_weight=4.0, weights=(10.0, 10.0, 5.0, 5.0))
    p.update(kw)
    return SimpleNamespace(**p)


def match_anchors(iou, high, low):
    """matcher.py:42-113 with allow_low_quality_matches=True -> int64 [A] in {-2, -1, 0..G-1}."""
    if iou.numel() == 0:
        raise ValueError("No ground-truth boxes available for one of the images during training"
                         if iou.shape[0] == 0 else
                         "No proposal boxes available for one of the images during training")
    best_val, best_gt = iou.max(dim=0)                                   # :66
    matched = best_gt.clone()
    below = best_val < low                                               # :71
    between = (best_val >= low) & (best_val < high)                      # :72-74
    matched[below] = BELOW_LOW_THRESHOLD
    matched[between] = BETWEEN_THRESHOLDS
    gt_best = iou.max(dim=1).values                                      # :92
    is_a_gt_best = (iou == gt_best[:, None]).any(dim=0)                  # :94-97, ties included
    matched[is_a_gt_best] = best_gt[is_a_gt_best]                        # :112-113
    return matched


def encode_legacy(gt, anchors, weights):
    """box_coder.py:22-50."""
    ew = anchors[:, 2] - anchors[:, 0] + 1
    eh = anchors[:, 3] - anchors[:, 1] + 1
    ecx = anchors[:, 0] + 0.5 * ew
    ecy = anchors[:, 1] + 0.5 * eh
    gw = gt[:, 2] - gt[:, 0] + 1
    gh = gt[:, 3] - gt[:, 1] + 1
    gcx = gt[:, 0] + 0.5 * gw
    gcy = gt[:, 1] + 0.5 * gh
    wx, wy, ww, wh = weights
    return torch.stack((wx * (gcx - ecx) / ew, wy * (gcy - ecy) / eh, ww * torch.log(gw / ew),
                        wh * torch.log(gh / eh)), dim=1)


def smooth_l1_sum(x, t, beta):
    """layers/smooth_l1_loss.py:6-17 with size_average=False."""
    n = torch.abs(x - t)
    return torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta).sum()


def flatten_heads(box_cls, box_regression):
    """rpn/utils.py:10-45 for A anchors per location: [N, A*C, H, W] -> [N*sum(H*W*A), C] (location-major,
    anchor inner), same for the regression with 4 channels."""
    cls_flat, reg_flat = [], []
    for c, r in zip(box_cls, box_regression):
        n, axc, h, w = c.shape
        a = r.shape[1] // 4
        ch = axc // a
        cls_flat.append(c.view(n, a, ch, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, ch))
        reg_flat.append(r.view(n, a, 4, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, 4))
    ch = cls_flat[0].shape[2]
    return torch.cat(cls_flat, dim=1).reshape(-1, ch), torch.cat(reg_flat, dim=1).reshape(-1, 4)


def assign(gt_boxes, gt_labels, anchors_per_level, params=None):
    prm = params or default_params()
    anchors_cat = torch.cat(list(anchors_per_level), dim=0)
    labels, matched, reg_t = [], [], []
    for gb, gl in zip(gt_boxes, gt_labels):
        m = match_anchors(P.iou_matrix(gb, anchors_cat), prm.fg_iou_threshold, prm.bg_iou_threshold)
        mc = m.clamp(min=0)
        lab = gl[mc].to(torch.float32)
        lab[m == BELOW_LOW_THRESHOLD] = 0
        lab[m == BETWEEN_THRESHOLDS] = -1
        labels.append(lab)
        matched.append(m)
        reg_t.append(encode_legacy(gb[mc], anchors_cat, prm.weights))
    labels_flat = torch.cat(labels)
    pos = torch.nonzero(labels_flat > 0).squeeze(1)
    return SimpleNamespace(N=len(gt_boxes), A=anchors_cat.shape[0], labels=torch.stack(labels).long(),
                           matched=torch.stack(matched), reg_targets=torch.cat(reg_t), pos_inds=pos,
                           num_pos=int(pos.numel()), params=prm)


def losses(box_cls, box_regression, asg):
    """retinanet/loss.py:58-81 -> [cls, reg] with autograd graphs."""
    prm = asg.params
    cls_flat, reg_flat = flatten_heads(box_cls, box_regression)
    pos = asg.pos_inds
    reg_loss = smooth_l1_sum(reg_flat[pos], asg.reg_targets[pos], prm.bbox_reg_beta) / \
        max(1, pos.numel() * prm.bbox_reg_weight)
    cls_loss = P.focal_loss_cpu(cls_flat, asg.labels.reshape(-1).int(), prm.gamma, prm.alpha).sum() / \
        (pos.numel() + asg.N)
    return [cls_loss, reg_loss]


def assign_and_loss(box_cls, box_regression, gt_boxes, gt_labels, anchors_per_level, params=None, with_grad=True):
    leaves = None
    if with_grad:
        box_cls = [x.detach().clone().requires_grad_(True) for x in box_cls]
        box_regression = [x.detach().clone().requires_grad_(True) for x in box_regression]
        leaves = (box_cls, box_regression)
    asg = assign(gt_boxes, gt_labels, anchors_per_level, params)
    ls = losses(box_cls, box_regression, asg)
    grads = None
    if with_grad:
        sum(ls).backward()
        grads = SimpleNamespace(box_cls=[x.grad for x in leaves[0]], box_regression=[x.grad for x in leaves[1]])
    return [l.detach() for l in ls], grads, asg
