"""Deterministic synthetic inputs for the PAA hot path.

There is no dataset and no checkpoint in this environment, so tests, the bench and the golden
fixtures all draw their inputs from here: anchors laid out exactly like the reference's
``make_anchor_generator_paa`` (modeling/rpn/anchor_generator.py:73-95,192-212 -- one square anchor of
side 8*stride per location, rows outer / columns inner), ground-truth boxes inside the image, and
head outputs shaped like ``PAAHead.forward`` (modeling/rpn/paa/paa.py:90-108):
``box_cls[l] [N,80,H_l,W_l]``, ``box_regression[l] [N,4,H_l,W_l]``, ``iou_pred[l] [N,1,H_l,W_l]``.

Everything is generated on the CPU with a seeded ``torch.Generator`` so that the same seed gives
the same bytes here and on the GPU box.
"""
import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch

STRIDES = (8, 16, 32, 64, 128)
NUM_FG_CLASSES = 80


def padded_size(h: int, w: int, divisor: int = 32) -> Tuple[int, int]:
    """Batch padding of structures/image_list.py:54-61 (SIZE_DIVISIBILITY 32)."""
    return (int(math.ceil(h / divisor) * divisor), int(math.ceil(w / divisor) * divisor))


def level_grids(hp: int, wp: int, strides=STRIDES) -> List[Tuple[int, int]]:
    """Feature-map sizes of P3..P7 for a padded input: each level halves (ceil) the previous one."""
    grids = []
    h, w = hp, wp
    s_prev = 1
    for s in strides:
        while s_prev < s:
            h, w = (h + 1) // 2, (w + 1) // 2
            s_prev *= 2
        grids.append((h, w))
    return grids


def level_anchors(grid: Tuple[int, int], stride: int, scale: float = 8.0) -> torch.Tensor:
    """[H*W,4] xyxy anchors of one level. The cell anchor of stride s and size 8s is the square
    centred on s/2 - 0.5 + 0.5 with side 8s in the reference's "+1" convention, i.e.
    [-3.5s+0.5, -3.5s+0.5, 4.5s-0.5, 4.5s-0.5] (for s=8: [-27.5,-27.5,35.5,35.5])."""
    h, w = grid
    side = scale * stride
    ctr = 0.5 * stride
    lo = ctr - 0.5 * (side - 1.0)
    hi = ctr + 0.5 * (side - 1.0)
    ys = torch.arange(h, dtype=torch.float32) * stride
    xs = torch.arange(w, dtype=torch.float32) * stride
    yy = ys.view(h, 1).expand(h, w).reshape(-1)
    xx = xs.view(1, w).expand(h, w).reshape(-1)
    return torch.stack((xx + lo, yy + lo, xx + hi, yy + hi), dim=1).contiguous()


@dataclass
class SyntheticBatch:
    """Plain tensors (CPU, float32/int64). ``to_boxlists`` in tests wraps them per API."""
    image_sizes: List[Tuple[int, int]]          # (w, h) per image, BoxList.size convention
    grids: List[Tuple[int, int]]                # (H_l, W_l) per level
    anchors: List[torch.Tensor]                 # per level [H_l*W_l, 4]
    gt_boxes: List[torch.Tensor]                # per image [G,4]
    gt_labels: List[torch.Tensor]               # per image [G] int64 in 1..80
    box_cls: List[torch.Tensor]                 # per level [N,80,H,W]
    box_regression: List[torch.Tensor]          # per level [N,4,H,W]
    iou_pred: List[torch.Tensor]                # per level [N,1,H,W]
    meta: dict = field(default_factory=dict)

    @property
    def num_images(self):
        return int(self.box_cls[0].shape[0])

    @property
    def num_anchors(self):
        return sum(a.shape[0] for a in self.anchors)


def make_gt(gen: torch.Generator, n_gt: int, w: int, h: int):
    """wh ~ 16 + U[0,1]*(0.3W, 0.375H); top-left uniform so the box stays inside the image."""
    u = torch.rand((n_gt, 4), generator=gen)
    bw = 16.0 + u[:, 0] * 0.3 * w
    bh = 16.0 + u[:, 1] * 0.375 * h
    x1 = u[:, 2] * (w - 1 - bw).clamp(min=0)
    y1 = u[:, 3] * (h - 1 - bh).clamp(min=0)
    boxes = torch.stack((x1, y1, (x1 + bw).clamp(max=w - 1), (y1 + bh).clamp(max=h - 1)), dim=1)
    # pixel-ish coordinates with a quarter-pixel fraction, like resized COCO boxes
    boxes = (boxes * 4).round() / 4
    labels = torch.randint(1, NUM_FG_CLASSES + 1, (n_gt,), generator=gen, dtype=torch.int64)
    return boxes.contiguous(), labels


def _pair_iou(gt: torch.Tensor, anc: torch.Tensor) -> torch.Tensor:
    lt = torch.max(gt[:, None, :2], anc[None, :, :2])
    rb = torch.min(gt[:, None, 2:], anc[None, :, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    ag = (gt[:, 2] - gt[:, 0] + 1) * (gt[:, 3] - gt[:, 1] + 1)
    aa = (anc[:, 2] - anc[:, 0] + 1) * (anc[:, 3] - anc[:, 1] + 1)
    return inter / (ag[:, None] + aa[None, :] - inter)


def make_batch(seed: int, num_images: int, image_hw: Tuple[int, int] = (800, 1333),
               gt_per_image=20, trained_like: bool = True, cls_mean: float = -4.0,
               cls_std: float = 1.0, strides=STRIDES, per_image_hw=None) -> SyntheticBatch:
    """Training-shaped batch. ``gt_per_image`` is an int (exact count) or a (lo, hi) range.

    ``trained_like`` raises the GT-class logit and tightens the regression around the matched GT
    for well-overlapping anchors, so the per-GT candidate losses are bimodal like a partly trained
    detector's and the GMM has something to separate.
    """
    gen = torch.Generator().manual_seed(seed)
    if per_image_hw is None:
        per_image_hw = [image_hw] * num_images
    hp, wp = padded_size(max(h for h, _ in per_image_hw), max(w for _, w in per_image_hw))
    grids = level_grids(hp, wp, strides)
    anchors = [level_anchors(g, s) for g, s in zip(grids, strides)]
    all_anchors = torch.cat(anchors, dim=0)
    A = all_anchors.shape[0]
    gt_boxes, gt_labels = [], []
    for (h, w) in per_image_hw:
        if isinstance(gt_per_image, int):
            n_gt = gt_per_image
        else:
            n_gt = int(torch.randint(gt_per_image[0], gt_per_image[1] + 1, (1,), generator=gen))
        b, l = make_gt(gen, n_gt, w, h)
        gt_boxes.append(b)
        gt_labels.append(l)
    N = num_images
    cls = (torch.randn((N, A, NUM_FG_CLASSES), generator=gen) * cls_std + cls_mean)
    reg = torch.randn((N, A, 4), generator=gen) * 0.5
    iou = torch.randn((N, A), generator=gen)
    if trained_like:
        aw = all_anchors[:, 2] - all_anchors[:, 0] + 1
        ah = all_anchors[:, 3] - all_anchors[:, 1] + 1
        acx = (all_anchors[:, 2] + all_anchors[:, 0]) / 2
        acy = (all_anchors[:, 3] + all_anchors[:, 1]) / 2
        for i in range(N):
            q = _pair_iou(gt_boxes[i], all_anchors)
            best, arg = q.max(dim=0)
            good = best > 0.5
            if good.any():
                idx = good.nonzero().squeeze(1)
                g = gt_boxes[i][arg[idx]]
                gw = g[:, 2] - g[:, 0] + 1
                gh = g[:, 3] - g[:, 1] + 1
                tgt = torch.stack((10 * ((g[:, 0] + g[:, 2]) / 2 - acx[idx]) / aw[idx],
                                   10 * ((g[:, 1] + g[:, 3]) / 2 - acy[idx]) / ah[idx],
                                   5 * torch.log(gw / aw[idx]), 5 * torch.log(gh / ah[idx])), dim=1)
                reg[i, idx] = tgt + 0.15 * reg[i, idx]
                cls[i, idx, gt_labels[i][arg[idx]] - 1] += 6.0
                iou[i, idx] += 1.5
    cls.clamp_(-12.0, 12.0)
    box_cls, box_reg, iou_pred = [], [], []
    off = 0
    for (gh_, gw_) in grids:
        n = gh_ * gw_
        box_cls.append(cls[:, off:off + n].reshape(N, gh_, gw_, NUM_FG_CLASSES)
                       .permute(0, 3, 1, 2).contiguous())
        box_reg.append(reg[:, off:off + n].reshape(N, gh_, gw_, 4).permute(0, 3, 1, 2).contiguous())
        iou_pred.append(iou[:, off:off + n].reshape(N, gh_, gw_, 1).permute(0, 3, 1, 2).contiguous())
        off += n
    return SyntheticBatch(image_sizes=[(w, h) for (h, w) in per_image_hw], grids=grids,
                          anchors=anchors, gt_boxes=gt_boxes, gt_labels=gt_labels, box_cls=box_cls,
                          box_regression=box_reg, iou_pred=iou_pred,
                          meta=dict(seed=seed, padded_hw=(hp, wp), trained_like=trained_like))


def make_inference_batch(seed: int, num_images: int, image_hw: Tuple[int, int] = (800, 1333),
                         cls_mean: float = -3.0, cls_std: float = 1.5, n_objects: int = 12,
                         strides=STRIDES, candidates_per_level=None) -> SyntheticBatch:
    """Post-processing-shaped batch (config C4): dense candidates on every level plus a few
    object-like clusters so that NMS has overlapping same-class boxes to suppress and score voting
    has neighbours to average."""
    b = make_batch(seed, num_images, image_hw, gt_per_image=n_objects, trained_like=True,
                   cls_mean=cls_mean, cls_std=cls_std, strides=strides)
    b.meta["inference"] = True
    if candidates_per_level is not None:
        # detector-like sparsity: shift every level so that about `candidates_per_level` of its
        # (location, class) logits pass sigmoid(x) > 0.05 -- a few thousand on P3 (0.3 % of 1.3 M) up to
        # most of P7 -- instead of the same dense fraction everywhere
        gate = math.log(0.05 / 0.95)
        for l, t in enumerate(b.box_cls):
            n_el = t.shape[1] * t.shape[2] * t.shape[3]
            frac = min(0.9, candidates_per_level / float(n_el))
            z = float(torch.special.ndtri(torch.tensor(1.0 - frac, dtype=torch.float64)))
            t.add_((gate - z * cls_std) - cls_mean).clamp_(-12.0, 12.0)
        b.meta["candidates_per_level"] = candidates_per_level
    return b


def multiscale_hw(seed: int, num_images: int, max_long: int = 1333):
    """Config C5: short side U{640..800}, aspect U[1.2,1.7], long side capped at 1333."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_images):
        s = int(torch.randint(640, 801, (1,), generator=gen))
        r = 1.2 + 0.5 * float(torch.rand((1,), generator=gen))
        out.append((s, min(int(round(s * r)), max_long)))
    return out


def make_retinanet_batch(seed: int, num_images: int, image_hw: Tuple[int, int] = (320, 416),
                         cls_mean: float = -3.5, cls_std: float = 1.5, strides=STRIDES,
                         aspect_ratios=(0.5, 1.0, 2.0), scales_per_octave: int = 3, octave: float = 2.0,
                         gt_per_image=None):
    """RetinaNet-shaped inputs (``gt_per_image`` -- an int or a (lo, hi) range -- adds ground truth for the loss) (rpn/retinanet/retinanet.py): A = ratios x scales anchors per location,
    ``box_cls[l] [N, A*80, H, W]``, ``box_regression[l] [N, A*4, H, W]``, anchors ``[H*W*A, 4]`` (location-major,
    anchor inner, rpn/utils.py:10-14) built like make_anchor_generator_retinanet."""
    import numpy as np
    from paa_b200.anchor_generator import generate_cell_anchors
    gen = torch.Generator().manual_seed(seed)
    hp, wp = padded_size(*image_hw)
    grids = level_grids(hp, wp, strides)
    A = len(aspect_ratios) * scales_per_octave
    anchors, box_cls, box_reg = [], [], []
    for (h, w), stride in zip(grids, strides):
        sizes = tuple(octave ** (k / float(scales_per_octave)) * 4.0 * stride for k in range(scales_per_octave))
        cell = torch.from_numpy(generate_cell_anchors(stride, sizes, aspect_ratios)).float()       # [A, 4]
        ys = torch.arange(h, dtype=torch.float32) * stride
        xs = torch.arange(w, dtype=torch.float32) * stride
        shift = torch.stack((xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w),
                             xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)), dim=2).reshape(-1, 1, 4)
        anchors.append((shift + cell.view(1, A, 4)).reshape(-1, 4).contiguous())
        box_cls.append((torch.randn((num_images, A * NUM_FG_CLASSES, h, w), generator=gen) * cls_std + cls_mean)
                       .clamp_(-12.0, 12.0))
        box_reg.append(torch.randn((num_images, A * 4, h, w), generator=gen) * 0.5)
    gt_boxes, gt_labels = [], []
    if gt_per_image is not None:
        for _ in range(num_images):
            n_gt = gt_per_image if isinstance(gt_per_image, int) else \
                int(torch.randint(gt_per_image[0], gt_per_image[1] + 1, (1,), generator=gen))
            b, l = make_gt(gen, n_gt, image_hw[1], image_hw[0])
            gt_boxes.append(b)
            gt_labels.append(l)
    return SyntheticBatch(image_sizes=[(image_hw[1], image_hw[0])] * num_images, grids=grids, anchors=anchors,
                          gt_boxes=gt_boxes, gt_labels=gt_labels, box_cls=box_cls, box_regression=box_reg, iou_pred=None,
                          meta=dict(seed=seed, padded_hw=(hp, wp), anchors_per_loc=A))


def fcos_locations(grids, strides=STRIDES):
    """fcos.py compute_locations: (x, y) = (col * stride + stride // 2, row * stride + stride // 2), row-major."""
    out = []
    for (h, w), stride in zip(grids, strides):
        ys = torch.arange(0, h * stride, step=stride, dtype=torch.float32)
        xs = torch.arange(0, w * stride, step=stride, dtype=torch.float32)
        yy = ys.view(h, 1).expand(h, w).reshape(-1)
        xx = xs.view(1, w).expand(h, w).reshape(-1)
        out.append(torch.stack((xx, yy), dim=1) + stride // 2)
    return out


def to_device_inputs(batch, device="cuda", requires_grad=False, channels_last=False):
    """SyntheticBatch -> (box_cls, box_regression, iou_pred, targets, anchors) in the reference's API
    shapes, on the device, with the per-level anchor tensors shared by all images like
    anchor_generator.py:112-125 does.  `channels_last`: the head tensors in torch.channels_last memory format (the
    same logical [N, C, H, W] values), as an AMP / cuDNN NHWC pipeline hands them over."""
    import torch
    from paa_b200.structures import BoxList
    fmt = torch.channels_last if channels_last else torch.contiguous_format

    def head(t):
        return t.to(device).contiguous(memory_format=fmt).requires_grad_(requires_grad)
    cls = [head(t) for t in batch.box_cls]
    reg = [head(t) for t in batch.box_regression]
    iou = None if batch.iou_pred is None else [head(t) for t in batch.iou_pred]
    anc = [a.to(device) for a in batch.anchors]
    targets, anchors = [], []
    for i in range(batch.num_images):
        t = BoxList(batch.gt_boxes[i].to(device), batch.image_sizes[i], mode="xyxy")
        t.add_field("labels", batch.gt_labels[i].to(device))
        targets.append(t)
        anchors.append([BoxList(a, batch.image_sizes[i], mode="xyxy") for a in anc])
    return cls, reg, iou, targets, anchors


def slice_batch(batch, start, stop):
    """Images [start, stop) of a batch as a batch of its own (one rank's share of a global batch: the reference's
    DistributedSampler split, data/build.py:110-115).  Anchors / grids are shared, head tensors are contiguous
    copies."""
    sl = slice(start, stop)
    return SyntheticBatch(image_sizes=batch.image_sizes[sl], grids=batch.grids, anchors=batch.anchors,
                          gt_boxes=batch.gt_boxes[sl], gt_labels=batch.gt_labels[sl],
                          box_cls=[t[sl].contiguous() for t in batch.box_cls],
                          box_regression=[t[sl].contiguous() for t in batch.box_regression],
                          iou_pred=None if batch.iou_pred is None else [t[sl].contiguous() for t in batch.iou_pred],
                          meta=dict(batch.meta, slice=(start, stop)))
