"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the ATSS training path of the reference
(paa_core/modeling/rpn/atss/loss.py, POSITIVE_TYPE 'ATSS'), SURVEY.md 8f-2: the loss that shares the PAA path's
focal / GIoU / BCE machinery but assigns anchors by the ATSS rule.

  prepare_targets (:139-197)  per image: per GT and level the TOPK anchors whose centres are nearest to the GT
      centre; the GT's IoU threshold is mean + (unbiased) std of its candidates' IoUs; a candidate is positive
      if its IoU reaches the threshold and its centre lies inside the GT (by more than 0.01); an anchor that is
      positive for several GTs goes to the one with the largest IoU (first on ties).
  losses (:247-279)  focal over all anchors / num_pos_avg; GIoU of the positives weighted by their centerness
      targets / (sum of centerness targets / world) * REG_LOSS_WEIGHT; BCE(centerness, target) / num_pos_avg.
"""
from types import SimpleNamespace

import torch

from oracle import paa_oracle as P

NEG = -100000000.0


def default_params(**kw):
    p = dict(gamma=2.0, alpha=0.25, topk=9, reg_loss_weight=2.0, positive_type="ATSS", fg_iou_threshold=0.5,
             bg_iou_threshold=0.4)
    p.update(kw)
    return SimpleNamespace(**p)


def assign_image(gt_boxes, gt_labels, anchors_per_level, topk):
    """-> labels [A] int64 (0 = background), matched gt index [A], candidate index matrix [L*topk, G],
    per-GT thresholds."""
    anchors = torch.cat(list(anchors_per_level), dim=0)
    G = gt_boxes.shape[0]
    ious = P.iou_matrix(gt_boxes, anchors).t().contiguous()                    # [A, G]  (:141)
    gcx = (gt_boxes[:, 2] + gt_boxes[:, 0]) / 2.0
    gcy = (gt_boxes[:, 3] + gt_boxes[:, 1]) / 2.0
    acx = (anchors[:, 2] + anchors[:, 0]) / 2.0
    acy = (anchors[:, 3] + anchors[:, 1]) / 2.0
    dist = ((acx[:, None] - gcx[None, :]).pow(2) + (acy[:, None] - gcy[None, :]).pow(2)).sqrt()   # :151
    cand, start = [], 0
    for a in anchors_per_level:
        n = a.shape[0]
        _, idx = dist[start:start + n].topk(topk, dim=0, largest=False)      # :159
        cand.append(idx + start)
        start += n
    cand = torch.cat(cand, dim=0)                                              # [L*topk, G]
    cand_iou = ious[cand, torch.arange(G)]                                     # :165
    thr = cand_iou.mean(0) + cand_iou.std(0)                                   # :166-168
    is_pos = cand_iou >= thr[None, :]
    l = acx[cand] - gt_boxes[:, 0][None, :]
    t = acy[cand] - gt_boxes[:, 1][None, :]
    r = gt_boxes[:, 2][None, :] - acx[cand]
    b = gt_boxes[:, 3][None, :] - acy[cand]
    inside = torch.stack([l, t, r, b], dim=0).min(dim=0).values > 0.01         # :175-180
    is_pos = is_pos & inside
    masked = torch.full_like(ious, NEG)                                        # [A, G]
    gsel = torch.arange(G)[None, :].expand_as(cand)
    masked[cand[is_pos], gsel[is_pos]] = ious[cand[is_pos], gsel[is_pos]]      # :183-187
    val, arg = masked.max(dim=1)                                               # :189 (first maximum)
    labels = gt_labels[arg].clone()
    labels[val == NEG] = 0
    return labels, arg, cand, thr, cand_iou


INF = 100000000
SIZE_RANGES = ((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF))


def assign_image_ssc(gt_boxes, gt_labels, anchors_per_level):
    """POSITIVE_TYPE 'SSC' (atss/loss.py:89-128): FCOS's rule on the anchor centres -- inside the GT by more than
    0.01, max(l, t, r, b) in the level's size range, smallest area wins (first on ties)."""
    anchors = torch.cat(list(anchors_per_level), dim=0)
    sizes = torch.cat([torch.tensor(SIZE_RANGES[l], dtype=torch.float32)[None].expand(len(a), -1)
                       for l, a in enumerate(anchors_per_level)], dim=0)
    xs = (anchors[:, 2] + anchors[:, 0]) / 2.0
    ys = (anchors[:, 3] + anchors[:, 1]) / 2.0
    area = P.area_plus1(gt_boxes)
    l = xs[:, None] - gt_boxes[:, 0][None]
    t = ys[:, None] - gt_boxes[:, 1][None]
    r = gt_boxes[:, 2][None] - xs[:, None]
    b = gt_boxes[:, 3][None] - ys[:, None]
    reg = torch.stack([l, t, r, b], dim=2)
    inside = reg.min(dim=2)[0] > 0.01
    mx = reg.max(dim=2)[0]
    cared = (mx >= sizes[:, [0]]) & (mx <= sizes[:, [1]])
    a = area[None].repeat(len(anchors), 1)
    a[inside == 0] = INF
    a[cared == 0] = INF
    amin, arg = a.min(dim=1)
    labels = gt_labels[arg].clone()
    labels[amin == INF] = 0
    return labels, arg


def assign_image_iou(gt_boxes, gt_labels, anchors_per_level, high, low):
    """POSITIVE_TYPE 'IoU' (atss/loss.py:199-226): Matcher(high, low, allow_low_quality_matches) labels (-1 between
    the thresholds), then positives whose centre is not inside the matched GT by more than 0.01 become -1 too."""
    from oracle import retinanet_oracle as R
    anchors = torch.cat(list(anchors_per_level), dim=0)
    m = R.match_anchors(P.iou_matrix(gt_boxes, anchors), high, low)
    mc = m.clamp(min=0)
    labels = gt_labels[mc].to(torch.float32)
    labels[m == R.BELOW_LOW_THRESHOLD] = 0
    labels[m == R.BETWEEN_THRESHOLDS] = -1
    g = gt_boxes[mc]
    pos = torch.nonzero(labels > 0).squeeze(1)
    cx = (anchors[pos, 2] + anchors[pos, 0]) / 2.0
    cy = (anchors[pos, 3] + anchors[pos, 1]) / 2.0
    inside = torch.stack([cx - g[pos, 0], cy - g[pos, 1], g[pos, 2] - cx, g[pos, 3] - cy], dim=1).min(dim=1)[0] > 0.01
    labels[pos[inside == 0]] = -1
    return labels.long(), mc, m


def centerness_targets(reg_targets, anchors):
    """loss.py:233-245."""
    g = P.decode(reg_targets, anchors)
    cx = (anchors[:, 2] + anchors[:, 0]) / 2
    cy = (anchors[:, 3] + anchors[:, 1]) / 2
    l, t, r, b = cx - g[:, 0], cy - g[:, 1], g[:, 2] - cx, g[:, 3] - cy
    lr = torch.stack([l, r], dim=1)
    tb = torch.stack([t, b], dim=1)
    return torch.sqrt((lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0]))


def assign(gt_boxes, gt_labels, anchors_per_level, params=None):
    prm = params or default_params()
    anchors_cat = torch.cat(list(anchors_per_level), dim=0)
    N = len(gt_boxes)
    labels, matched, cands, thrs, reg_t = [], [], [], [], []
    for i in range(N):
        if prm.positive_type == "SSC":
            lab, arg = assign_image_ssc(gt_boxes[i], gt_labels[i], anchors_per_level)
            cand, thr = None, None
        elif prm.positive_type == "IoU":
            lab, arg, _ = assign_image_iou(gt_boxes[i], gt_labels[i], anchors_per_level, prm.fg_iou_threshold,
                                           prm.bg_iou_threshold)
            cand, thr = None, None
        else:
            lab, arg, cand, thr, _ = assign_image(gt_boxes[i], gt_labels[i], anchors_per_level, prm.topk)
        labels.append(lab)
        matched.append(arg)
        cands.append(cand)
        thrs.append(thr)
        reg_t.append(P.encode(gt_boxes[i][arg], anchors_cat))                  # :229
    labels_flat = torch.cat(labels)
    reg_targets = torch.cat(reg_t)
    anchors_flat = anchors_cat.repeat(N, 1)
    pos = torch.nonzero(labels_flat > 0).squeeze(1)
    ctr = centerness_targets(reg_targets[pos], anchors_flat[pos]) if pos.numel() else torch.zeros(0)
    return SimpleNamespace(N=N, A=anchors_cat.shape[0], labels=torch.stack(labels), matched=torch.stack(matched),
                           candidates=cands, thresholds=thrs, reg_targets=reg_targets, anchors_flat=anchors_flat,
                           pos_inds=pos, centerness=ctr, num_pos=int(pos.numel()),
                           sum_centerness=float(ctr.sum()) if pos.numel() else 0.0, params=prm)


def losses(box_cls, box_regression, centerness, asg, total_num_pos=None, total_sum_centerness=None, world_size=1):
    """loss.py:247-279 -> [cls, reg * REG_LOSS_WEIGHT, centerness] with autograd graphs."""
    prm = asg.params
    cls_flat, reg_flat, ctr_flat = P.flatten_heads(box_cls, box_regression, centerness)
    total_num_pos = asg.num_pos if total_num_pos is None else total_num_pos
    num_pos_avg = max(total_num_pos / float(world_size), 1.0)
    cls_loss = P.focal_loss_cpu(cls_flat, asg.labels.reshape(-1).int(), prm.gamma, prm.alpha).sum() / num_pos_avg
    pos = asg.pos_inds
    total_sum = asg.sum_centerness if total_sum_centerness is None else total_sum_centerness
    reg = P.giou_loss(reg_flat[pos], asg.reg_targets[pos], asg.anchors_flat[pos], weight=asg.centerness).sum()
    reg_loss = reg / (total_sum / float(world_size))
    ctr_loss = torch.nn.functional.binary_cross_entropy_with_logits(ctr_flat[pos], asg.centerness,
                                                                    reduction="sum") / num_pos_avg
    return [cls_loss, reg_loss * prm.reg_loss_weight, ctr_loss]


def assign_and_loss(box_cls, box_regression, centerness, gt_boxes, gt_labels, anchors_per_level, params=None,
                    with_grad=True):
    leaves = None
    if with_grad:
        box_cls = [x.detach().clone().requires_grad_(True) for x in box_cls]
        box_regression = [x.detach().clone().requires_grad_(True) for x in box_regression]
        centerness = [x.detach().clone().requires_grad_(True) for x in centerness]
        leaves = (box_cls, box_regression, centerness)
    asg = assign(gt_boxes, gt_labels, anchors_per_level, params)
    ls = losses(box_cls, box_regression, centerness, asg)
    grads = None
    if with_grad:
        sum(ls).backward()
        grads = SimpleNamespace(box_cls=[x.grad for x in leaves[0]], box_regression=[x.grad for x in leaves[1]],
                                centerness=[x.grad for x in leaves[2]])
    return [l.detach() for l in ls], grads, asg
