"""CPU, only where the reference tree is mounted: oracle/retinanet_oracle.py against the reference's own
RetinaNetLossComputation (paa_core/modeling/rpn/retinanet/loss.py) -- labels, regression targets, losses and
gradients bit for bit."""
import types

import pytest
import torch

from oracle import ref_shim, retinanet_oracle
from paa_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


OTHER = dict(fg_iou_threshold=0.6, bg_iou_threshold=0.3, bbox_reg_beta=0.25, bbox_reg_weight=1.0, gamma=1.5, alpha=0.4)


@pytest.mark.parametrize("seed,hw,gt,other", [(71, (320, 416), (2, 7), {}), (72, (384, 512), (3, 12), {}),
                                              (73, (384, 512), (130, 150), OTHER), (74, (320, 416), (2, 7), OTHER)],
                         ids=["defaults-a", "defaults-b", "other-parameters-crowded", "other-parameters"])
def test_retinanet_oracle_is_the_reference(seed, hw, gt, other):
    """`other`: the thresholds / beta / normaliser / focal parameters (and the > 128 GT images) that
    tests/test_gpu_retinanet_loss.py runs the kernels with."""
    ref = ref_shim.load_reference()
    from paa_core.modeling.box_coder import BoxCoder
    from paa_core.modeling.rpn.retinanet import loss as rloss
    ns = types.SimpleNamespace
    prm = retinanet_oracle.default_params(**other)
    cfg = ns(MODEL=ns(RETINANET=ns(LOSS_GAMMA=(prm.gamma,), LOSS_ALPHA=(prm.alpha,),
                                   FG_IOU_THRESHOLD=prm.fg_iou_threshold, BG_IOU_THRESHOLD=prm.bg_iou_threshold,
                                   BBOX_REG_BETA=prm.bbox_reg_beta, BBOX_REG_WEIGHT=prm.bbox_reg_weight)))
    ev = rloss.make_retinanet_loss_evaluator(cfg, BoxCoder(weights=(10.0, 10.0, 5.0, 5.0)))
    b = synthetic.make_retinanet_batch(seed=seed, num_images=2, image_hw=hw, gt_per_image=gt)
    cls = [t.clone().requires_grad_(True) for t in b.box_cls]
    reg = [t.clone().requires_grad_(True) for t in b.box_regression]
    targets = []
    for i in range(b.num_images):
        t = ref.BoxList(b.gt_boxes[i], b.image_sizes[i])
        t.add_field("labels", b.gt_labels[i])
        targets.append(t)
    anchors = [[ref.BoxList(a, b.image_sizes[i]) for a in b.anchors] for i in range(b.num_images)]
    from paa_core.structures.boxlist_ops import cat_boxlist
    labels, reg_targets = ev.prepare_targets([cat_boxlist(a) for a in anchors], targets)
    rl = ev(anchors, cls, reg, targets)
    sum(rl).backward()
    ol, og, asg = retinanet_oracle.assign_and_loss(b.box_cls, b.box_regression, b.gt_boxes, b.gt_labels, b.anchors,
                                                   prm)
    assert (asg.labels == -1).any() and (asg.labels > 0).any()
    for i in range(b.num_images):
        assert torch.equal(labels[i].long(), asg.labels[i])
    pos = asg.pos_inds
    assert torch.equal(torch.cat(reg_targets)[pos], asg.reg_targets[pos])
    assert [float(x) for x in rl] == [float(x) for x in ol]
    for a, g in zip(cls + reg, og.box_cls + og.box_regression):
        assert torch.equal(a.grad, g)
