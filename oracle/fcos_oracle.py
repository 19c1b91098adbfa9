"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the FCOS training path of the reference
(paa_core/modeling/rpn/fcos/loss.py, layers/iou_loss.py), SURVEY.md 8f-2: anchor-free assignment + focal /
IoU / centerness losses.

  compute_targets_for_locations (:153-201)  per image and location: distances (l, t, r, b) to every GT; a GT
      is a candidate if the location lies inside it (or inside its centre region of radius * stride,
      get_sample_region :55-103) and max(l, t, r, b) falls in the level's size range; the candidate with the
      smallest (+1 convention) area wins, first on ties; no candidate = background.
  __call__ (:211-281)  focal over all locations / num_pos_avg; IOULoss(pred ltrb, target ltrb) weighted by the
      centerness targets / (sum of centerness targets / world); BCE(centerness, target) / num_pos_avg.
"""
from types import SimpleNamespace

import torch

from oracle import paa_oracle as P

INF = 100000000
SIZE_RANGES = ((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF))


def default_params(**kw):
    p = dict(gamma=2.0, alpha=0.25, fpn_strides=(8, 16, 32, 64, 128), center_sampling_radius=0.0,
             iou_loss_type="iou", norm_reg_targets=False)
    p.update(kw)
    return SimpleNamespace(**p)


def sample_region(gt, strides, points_per_level, xs, ys, radius):
    """get_sample_region, fcos/loss.py:55-103."""
    K, G = xs.shape[0], gt.shape[0]
    g = gt[None].expand(K, G, 4)
    cx = (g[..., 0] + g[..., 2]) / 2
    cy = (g[..., 1] + g[..., 3]) / 2
    if cx[..., 0].sum() == 0:                                            # :68-69 "no gt"
        return torch.zeros((K, G), dtype=torch.bool)
    cg = torch.zeros((K, G, 4))
    beg = 0
    for level, n_p in enumerate(points_per_level):
        end = beg + n_p
        s = strides[level] * radius
        xmin, ymin, xmax, ymax = cx[beg:end] - s, cy[beg:end] - s, cx[beg:end] + s, cy[beg:end] + s
        cg[beg:end, :, 0] = torch.where(xmin > g[beg:end, :, 0], xmin, g[beg:end, :, 0])
        cg[beg:end, :, 1] = torch.where(ymin > g[beg:end, :, 1], ymin, g[beg:end, :, 1])
        cg[beg:end, :, 2] = torch.where(xmax > g[beg:end, :, 2], g[beg:end, :, 2], xmax)
        cg[beg:end, :, 3] = torch.where(ymax > g[beg:end, :, 3], g[beg:end, :, 3], ymax)
        beg = end
    left = xs[:, None] - cg[..., 0]
    right = cg[..., 2] - xs[:, None]
    top = ys[:, None] - cg[..., 1]
    bottom = cg[..., 3] - ys[:, None]
    return torch.stack((left, top, right, bottom), -1).min(-1)[0] > 0


def assign_image(gt_boxes, gt_labels, locations_per_level, prm):
    """-> labels [K] int64, matched GT [K], reg targets [K, 4] (not yet divided by the stride)."""
    pts = torch.cat(list(locations_per_level), dim=0)
    npl = [len(p) for p in locations_per_level]
    xs, ys = pts[:, 0], pts[:, 1]
    sizes = torch.cat([torch.tensor(SIZE_RANGES[l], dtype=torch.float32)[None].expand(n, -1)
                       for l, n in enumerate(npl)], dim=0)
    area = P.area_plus1(gt_boxes)                                          # BoxList.area(), bounding_box.py:226-231
    l = xs[:, None] - gt_boxes[:, 0][None]
    t = ys[:, None] - gt_boxes[:, 1][None]
    r = gt_boxes[:, 2][None] - xs[:, None]
    b = gt_boxes[:, 3][None] - ys[:, None]
    reg = torch.stack([l, t, r, b], dim=2)
    if prm.center_sampling_radius > 0:
        inside = sample_region(gt_boxes, prm.fpn_strides, npl, xs, ys, prm.center_sampling_radius)
    else:
        inside = reg.min(dim=2)[0] > 0
    mx = reg.max(dim=2)[0]
    cared = (mx >= sizes[:, [0]]) & (mx <= sizes[:, [1]])
    a = area[None].repeat(len(pts), 1)
    a[inside == 0] = INF
    a[cared == 0] = INF
    amin, arg = a.min(dim=1)
    reg = reg[range(len(pts)), arg]
    labels = gt_labels[arg].clone()
    labels[amin == INF] = 0
    return labels, arg, reg


def centerness_targets(reg):
    lr, tb = reg[:, [0, 2]], reg[:, [1, 3]]
    return torch.sqrt((lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0]))


def iou_loss(pred, target, weight, loss_type):
    """layers/iou_loss.py:12-51."""
    pl, pt, pr, pb = pred[:, 0], pred[:, 1], pred[:, 2], pred[:, 3]
    tl, tt, tr, tb = target[:, 0], target[:, 1], target[:, 2], target[:, 3]
    target_area = (tl + tr) * (tt + tb)
    pred_area = (pl + pr) * (pt + pb)
    w_int = torch.min(pl, tl) + torch.min(pr, tr)
    g_w = torch.max(pl, tl) + torch.max(pr, tr)
    h_int = torch.min(pb, tb) + torch.min(pt, tt)
    g_h = torch.max(pb, tb) + torch.max(pt, tt)
    ac = g_w * g_h + 1e-7
    inter = w_int * h_int
    union = target_area + pred_area - inter
    ious = (inter + 1.0) / (union + 1.0)
    gious = ious - (ac - union) / ac
    if loss_type == "iou":
        losses = -torch.log(ious)
    elif loss_type == "linear_iou":
        losses = 1 - ious
    elif loss_type == "giou":
        losses = 1 - gious
    else:
        raise NotImplementedError
    if weight is not None and weight.sum() > 0:
        return (losses * weight).sum()
    assert losses.numel() != 0
    return losses.sum()


def assign(gt_boxes, gt_labels, locations_per_level, params=None):
    prm = params or default_params()
    npl = [len(p) for p in locations_per_level]
    labels, matched, regs = [], [], []
    for gb, gl in zip(gt_boxes, gt_labels):
        lab, arg, reg = assign_image(gb, gl, locations_per_level, prm)
        if prm.norm_reg_targets:                                            # :147-148
            reg = torch.cat([r / prm.fpn_strides[l] for l, r in enumerate(torch.split(reg, npl, dim=0))], dim=0)
        labels.append(lab)
        matched.append(arg)
        regs.append(reg)
    labels, matched, regs = torch.stack(labels), torch.stack(matched), torch.stack(regs)   # image-major [N, K(,4)]
    pos = labels > 0
    ctr = centerness_targets(regs[pos]) if pos.any() else torch.zeros(0)
    return SimpleNamespace(N=len(gt_boxes), K=sum(npl), labels=labels, matched=matched, reg_targets=regs,
                           pos_mask=pos, centerness=ctr, num_pos=int(pos.sum()),
                           sum_centerness=float(ctr.sum()) if pos.any() else 0.0, params=prm)


def level_first(x, points_per_level):
    """[N, K, ...] image-major -> the reference's flattening (fcos/loss.py:235-246): per level all images, levels
    concatenated."""
    return torch.cat([t.reshape((-1,) + tuple(x.shape[2:])) for t in torch.split(x, points_per_level, dim=1)], dim=0)


def losses(box_cls, box_regression, centerness, asg, total_num_pos=None, total_sum_centerness=None, world_size=1):
    """fcos/loss.py:226-281 -> [cls, reg, centerness] with autograd graphs, in the reference's level-first
    flattening so that the float32 sums add up in the same order."""
    prm = asg.params
    C = box_cls[0].shape[1]
    npl = [t.shape[2] * t.shape[3] for t in box_cls]
    cls_flat = torch.cat([t.permute(0, 2, 3, 1).reshape(-1, C) for t in box_cls], dim=0)
    reg_flat = torch.cat([t.permute(0, 2, 3, 1).reshape(-1, 4) for t in box_regression], dim=0)
    ctr_flat = torch.cat([t.reshape(-1) for t in centerness], dim=0)
    labels = level_first(asg.labels, npl)
    reg_targets = level_first(asg.reg_targets, npl)
    pos = torch.nonzero(labels > 0).squeeze(1)
    total_num_pos = asg.num_pos if total_num_pos is None else total_num_pos
    num_pos_avg = max(total_num_pos / float(world_size), 1.0)
    cls_loss = P.focal_loss_cpu(cls_flat, labels.int(), prm.gamma, prm.alpha).sum() / num_pos_avg
    if pos.numel() > 0:
        ctr_t = centerness_targets(reg_targets[pos])
        total_sum = float(ctr_t.sum()) if total_sum_centerness is None else total_sum_centerness
        reg_loss = iou_loss(reg_flat[pos], reg_targets[pos], ctr_t, prm.iou_loss_type) / (total_sum / float(world_size))
        ctr_loss = torch.nn.functional.binary_cross_entropy_with_logits(ctr_flat[pos], ctr_t,
                                                                        reduction="sum") / num_pos_avg
    else:
        reg_loss = reg_flat[pos].sum()
        ctr_loss = ctr_flat[pos].sum()
    return [cls_loss, reg_loss, ctr_loss]


def assign_and_loss(box_cls, box_regression, centerness, gt_boxes, gt_labels, locations_per_level, params=None,
                    with_grad=True):
    leaves = None
    if with_grad:
        box_cls = [x.detach().clone().requires_grad_(True) for x in box_cls]
        box_regression = [x.detach().clone().requires_grad_(True) for x in box_regression]
        centerness = [x.detach().clone().requires_grad_(True) for x in centerness]
        leaves = (box_cls, box_regression, centerness)
    asg = assign(gt_boxes, gt_labels, locations_per_level, params)
    ls = losses(box_cls, box_regression, centerness, asg)
    grads = None
    if with_grad:
        sum(ls).backward()
        zero = lambda x: torch.zeros_like(x) if x.grad is None else x.grad
        grads = SimpleNamespace(box_cls=[zero(x) for x in leaves[0]], box_regression=[zero(x) for x in leaves[1]],
                                centerness=[zero(x) for x in leaves[2]])
    return [l.detach() for l in ls], grads, asg
