"""CPU, only where the reference tree is mounted: oracle/atss_oracle.py against the reference's own
ATSSLossComputation (paa_core/modeling/rpn/atss/loss.py) -- labels, losses and gradients bit for bit."""
import types

import pytest
import torch

from oracle import atss_oracle, ref_shim
from paa_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("seed,hw,gt,ptype", [(61, (384, 512), (2, 7), "ATSS"), (62, (512, 640), (3, 12), "ATSS"),
                                              (63, (384, 512), (2, 7), "SSC"), (64, (512, 640), (3, 12), "IoU")])
def test_atss_oracle_is_the_reference(seed, hw, gt, ptype):
    ref = ref_shim.load_reference()
    from paa_core.modeling.rpn.atss import loss as aloss
    ns = types.SimpleNamespace
    cfg = ns(MODEL=ns(ATSS=ns(LOSS_GAMMA=(2.0,), LOSS_ALPHA=(0.25,), FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4,
                              POSITIVE_TYPE=ptype, TOPK=9, REG_LOSS_WEIGHT=2.0, REGRESSION_TYPE="BOX")))
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
                                              b.anchors, atss_oracle.default_params(positive_type=ptype))
    assert asg.num_pos > 0 and (ptype != "IoU" or (asg.labels == -1).any())
    for i in range(b.num_images):
        assert torch.equal(labels[i].long(), asg.labels[i])
    assert [float(x) for x in rl] == [float(x) for x in ol]
    for a, g in zip(cls + reg + ctr, og.box_cls + og.box_regression + og.centerness):
        assert torch.equal(a.grad, g)
