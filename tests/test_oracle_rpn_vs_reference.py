"""CPU, only where the reference tree is mounted: oracle/rpn_oracle.py against the reference's own RPNLossComputation
(paa_core/modeling/rpn/loss.py with its BalancedPositiveNegativeSampler) under the same torch seed -- labels,
regression targets of the sampled positives, the sample itself, losses and gradients bit for bit."""
import types

import pytest
import torch

from oracle import ref_shim, rpn_oracle
from paa_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")

OTHER = dict(fg_iou_threshold=0.6, bg_iou_threshold=0.4, batch_size_per_image=64, positive_fraction=0.25)


def rpn_batch(seed, hw, gt, num_images=2, straddle=0.0):
    """RPN-shaped inputs: A anchors per location (3 ratios x 3 scales as make_retinanet_batch builds them), one
    objectness logit and four regression outputs per anchor, the anchors' visibility (anchor_generator.py:97-110)."""
    b = synthetic.make_retinanet_batch(seed=seed, num_images=num_images, image_hw=hw, gt_per_image=gt)
    objectness = []
    for c, r in zip(b.box_cls, b.box_regression):
        n, axc, h, w = c.shape
        a = r.shape[1] // 4
        objectness.append(c.view(n, a, axc // a, h, w)[:, :, 0].contiguous() + 3.0)      # around 0: both classes live
    anchors = torch.cat(list(b.anchors))
    vis = []
    for (w, h) in b.image_sizes:
        vis.append((anchors[:, 0] >= -straddle) & (anchors[:, 1] >= -straddle) & (anchors[:, 2] < w + straddle) &
                   (anchors[:, 3] < h + straddle))
    return b, objectness, vis


@pytest.mark.parametrize("seed,hw,gt,other", [(171, (320, 416), (2, 7), {}), (172, (384, 512), (3, 12), {}),
                                              (173, (384, 512), (130, 150), OTHER), (174, (320, 416), 1, OTHER)],
                         ids=["defaults-a", "defaults-b", "other-parameters-crowded", "other-parameters-one-gt"])
def test_rpn_oracle_is_the_reference(seed, hw, gt, other):
    ref = ref_shim.load_reference()
    from paa_core.modeling.box_coder import BoxCoder
    from paa_core.modeling.rpn import loss as rloss
    ns = types.SimpleNamespace
    prm = rpn_oracle.default_params(**other)
    cfg = ns(MODEL=ns(RPN=ns(FG_IOU_THRESHOLD=prm.fg_iou_threshold, BG_IOU_THRESHOLD=prm.bg_iou_threshold,
                             BATCH_SIZE_PER_IMAGE=prm.batch_size_per_image, POSITIVE_FRACTION=prm.positive_fraction)))
    ev = rloss.make_rpn_loss_evaluator(cfg, BoxCoder(weights=(1.0, 1.0, 1.0, 1.0)))
    b, objectness, vis = rpn_batch(seed, hw, gt)
    obj = [t.clone().requires_grad_(True) for t in objectness]
    reg = [t.clone().requires_grad_(True) for t in b.box_regression]
    targets = [ref.BoxList(b.gt_boxes[i], b.image_sizes[i]) for i in range(b.num_images)]
    anchors = []
    for i in range(b.num_images):
        per_level, o = [], 0
        for a in b.anchors:
            bl = ref.BoxList(a, b.image_sizes[i])
            bl.add_field("visibility", vis[i][o:o + a.shape[0]])
            o += a.shape[0]
            per_level.append(bl)
        anchors.append(per_level)
    from paa_core.structures.boxlist_ops import cat_boxlist
    labels, reg_targets = ev.prepare_targets([cat_boxlist(a) for a in anchors], targets)
    torch.manual_seed(seed)
    rl = ev(anchors, obj, reg, targets)
    sum(rl).backward()
    torch.manual_seed(seed)
    ol, og, asg = rpn_oracle.assign_and_loss(objectness, b.box_regression, b.gt_boxes, b.anchors, vis, prm)
    for i in range(b.num_images):
        assert torch.equal(labels[i], asg.labels[i])
    assert (torch.cat(asg.labels) == -1).any() and (torch.cat(asg.labels) == 1).any()
    pos = asg.sampled_pos
    assert pos.numel() > 0 and asg.sampled_neg.numel() > 0
    assert torch.equal(torch.cat(reg_targets)[pos], asg.reg_targets[pos])
    assert [float(x) for x in rl] == [float(x) for x in ol]
    for a, g in zip(obj + reg, og.objectness + og.box_regression):
        assert torch.equal(a.grad, g)
    # the gradients live on the sampled anchors only
    flat = torch.cat([g.view(g.shape[0], g.shape[1], -1).permute(0, 2, 1).reshape(-1) for g in og.objectness])
    assert int((flat != 0).sum()) <= pos.numel() + asg.sampled_neg.numel()
