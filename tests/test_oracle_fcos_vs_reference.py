"""CPU, only where the reference tree is mounted: oracle/fcos_oracle.py against the reference's own
FCOSLossComputation (paa_core/modeling/rpn/fcos/loss.py) -- labels, regression targets, losses and gradients
bit for bit, for the plain and the centre-sampling assignment and the three IoU loss types."""
import types

import pytest
import torch

from oracle import fcos_oracle, ref_shim
from paa_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def fcos_batch(seed, hw, gt, num_images=2):
    """PAA-shaped synthetic heads with FCOS semantics: the regression maps are positive distances (the FCOS head
    applies exp / relu, fcos.py:87-95)."""
    b = synthetic.make_batch(seed=seed, num_images=num_images, image_hw=hw, gt_per_image=gt, trained_like=False)
    b.box_regression = [(t.abs() * 40.0 + 1.0) for t in b.box_regression]
    return b, synthetic.fcos_locations(b.grids)


@pytest.mark.parametrize("seed,hw,gt,radius,loss_type,norm", [
    (81, (320, 416), (2, 7), 0.0, "iou", False),
    (82, (384, 512), (3, 12), 1.5, "giou", True),
    (83, (320, 416), (2, 7), 0.0, "linear_iou", False),
    # the other shapes tests/test_gpu_fcos_loss.py runs the kernels on: crowded images, centre sampling at full size
    (83, (384, 512), (130, 150), 0.0, "linear_iou", False),
    (82, (800, 1333), (5, 40), 1.5, "giou", True),
])
def test_fcos_oracle_is_the_reference(seed, hw, gt, radius, loss_type, norm):
    ref_shim.load_reference()
    from paa_core.modeling.rpn.fcos import loss as floss
    from paa_core.structures.bounding_box import BoxList
    ns = types.SimpleNamespace
    cfg = ns(MODEL=ns(FCOS=ns(LOSS_GAMMA=(2.0,), LOSS_ALPHA=(0.25,), FPN_STRIDES=[8, 16, 32, 64, 128],
                              CENTER_SAMPLING_RADIUS=radius, IOU_LOSS_TYPE=loss_type, NORM_REG_TARGETS=norm)))
    ev = floss.make_fcos_loss_evaluator(cfg)
    b, locations = fcos_batch(seed, hw, gt)
    cls = [t.clone().requires_grad_(True) for t in b.box_cls]
    reg = [t.clone().requires_grad_(True) for t in b.box_regression]
    ctr = [t.clone().requires_grad_(True) for t in b.iou_pred]
    targets = []
    for i in range(b.num_images):
        t = BoxList(b.gt_boxes[i], b.image_sizes[i])
        t.add_field("labels", b.gt_labels[i])
        targets.append(t)
    labels, reg_targets = ev.prepare_targets(locations, targets)     # level-first: [N * K_l], [N * K_l, 4]
    rl = ev(locations, cls, reg, ctr, targets)
    sum(rl).backward()
    prm = fcos_oracle.default_params(center_sampling_radius=radius, iou_loss_type=loss_type, norm_reg_targets=norm)
    ol, og, asg = fcos_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                              locations, prm)
    assert asg.num_pos > 0
    npl = [len(p) for p in locations]
    want_labels = torch.split(asg.labels, npl, dim=1)
    want_regs = torch.split(asg.reg_targets, npl, dim=1)
    for l in range(len(npl)):
        assert torch.equal(labels[l], want_labels[l].reshape(-1))
        pos = labels[l] > 0
        assert torch.equal(reg_targets[l][pos], want_regs[l].reshape(-1, 4)[pos])
    assert [float(x) for x in rl] == [float(x) for x in ol]
    for a, g in zip(cls + reg + ctr, og.box_cls + og.box_regression + og.centerness):
        assert torch.equal(a.grad if a.grad is not None else torch.zeros_like(a), g)


def test_fcos_oracle_without_positives_is_the_reference():
    """A GT that contains no location: every label is background and the regression / centerness losses are the
    empty sums of fcos/loss.py:274-277 -- the case tests/test_gpu_fcos_loss.py::test_fcos_loss_without_positives runs."""
    ref_shim.load_reference()
    from paa_core.modeling.rpn.fcos import loss as floss
    from paa_core.structures.bounding_box import BoxList
    ns = types.SimpleNamespace
    cfg = ns(MODEL=ns(FCOS=ns(LOSS_GAMMA=(2.0,), LOSS_ALPHA=(0.25,), FPN_STRIDES=[8, 16, 32, 64, 128],
                              CENTER_SAMPLING_RADIUS=0.0, IOU_LOSS_TYPE="iou", NORM_REG_TARGETS=False)))
    ev = floss.make_fcos_loss_evaluator(cfg)
    b, locations = fcos_batch(85, (320, 416), 1, num_images=1)
    b.gt_boxes[0] = torch.tensor([[13.0, 13.0, 18.0, 18.0]])
    cls = [t.clone().requires_grad_(True) for t in b.box_cls]
    reg = [t.clone().requires_grad_(True) for t in b.box_regression]
    ctr = [t.clone().requires_grad_(True) for t in b.iou_pred]
    t = BoxList(b.gt_boxes[0], b.image_sizes[0])
    t.add_field("labels", b.gt_labels[0])
    rl = ev(locations, cls, reg, ctr, [t])
    sum(rl).backward()
    ol, og, asg = fcos_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                              locations, fcos_oracle.default_params())
    assert asg.num_pos == 0
    assert [float(x.detach()) for x in rl] == [float(x) for x in ol]
    for a, g in zip(cls, og.box_cls):
        assert torch.equal(a.grad, g)
