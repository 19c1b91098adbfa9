"""CPU, only where the reference tree is mounted: oracle/atss_oracle.py against the reference's own
ATSSLossComputation (paa_core/modeling/rpn/atss/loss.py) -- labels, losses and gradients bit for bit."""
import types

import pytest
import torch

from oracle import atss_oracle, ref_shim
from paa_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("seed,hw,gt,ptype,other", [
    (61, (384, 512), (2, 7), "ATSS", {}), (62, (512, 640), (3, 12), "ATSS", {}),
    (63, (384, 512), (2, 7), "SSC", {}), (64, (512, 640), (3, 12), "IoU", {}),
    # what tests/test_gpu_atss_loss.py runs the kernels with beyond the defaults
    (68, (512, 640), (5, 40), "IoU", dict(fg_iou_threshold=0.3, bg_iou_threshold=0.2)),
    (69, (384, 512), (130, 150), "IoU", {}),
    (70, (384, 512), (3, 12), "ATSS", dict(topk=5, reg_loss_weight=1.0, gamma=1.5, alpha=0.4)),
])
def test_atss_oracle_is_the_reference(seed, hw, gt, ptype, other):
    ref = ref_shim.load_reference()
    from paa_core.modeling.rpn.atss import loss as aloss
    ns = types.SimpleNamespace
    prm = atss_oracle.default_params(positive_type=ptype, **other)
    cfg = ns(MODEL=ns(ATSS=ns(LOSS_GAMMA=(prm.gamma,), LOSS_ALPHA=(prm.alpha,), FG_IOU_THRESHOLD=prm.fg_iou_threshold,
                              BG_IOU_THRESHOLD=prm.bg_iou_threshold, POSITIVE_TYPE=ptype, TOPK=prm.topk,
                              REG_LOSS_WEIGHT=prm.reg_loss_weight, REGRESSION_TYPE="BOX")))
    ev = aloss.ATSSLossComputation(cfg, ref.BoxCoder(cfg))
    b = synthetic.make_batch(seed=seed, num_images=2, image_hw=hw, gt_per_image=gt)
    cls = [t.clone().requires_grad_(True) for t in b.box_cls]
    reg = [t.clone().requires_grad_(True) for t in b.box_regression]
    ctr = [t.clone().requires_grad_(True) for t in b.iou_pred]
    targets = []
    for i in range(b.num_images):
        t = ref.BoxList(b.gt_boxes[i], b.image_sizes[i])
        t.add_field("labels", b.gt_labels[i])
        targets.append(t)
    anchors = [[ref.BoxList(a, b.image_sizes[i]) for a in b.anchors] for i in range(b.num_images)]
    labels, _ = ev.prepare_targets(targets, anchors)
    rl = ev(cls, reg, ctr, targets, anchors)
    sum(rl).backward()
    ol, og, asg = atss_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                              b.anchors, prm)
    assert asg.num_pos > 0 and (ptype != "IoU" or (asg.labels == -1).any())
    for i in range(b.num_images):
        assert torch.equal(labels[i].long(), asg.labels[i])
    assert [float(x) for x in rl] == [float(x) for x in ol]
    for a, g in zip(cls + reg + ctr, og.box_cls + og.box_regression + og.centerness):
        assert torch.equal(a.grad, g)
