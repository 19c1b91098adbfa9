"""GPU parity of the FCOS training path (anchor-free assignment + the shared loss pass with IOULoss / centerness)
against oracle/fcos_oracle.py, which tests/test_oracle_fcos_vs_reference.py pins bit-exactly to the reference's
FCOSLossComputation.  Labels and chosen GTs bit-exact; losses and gradients 1e-4 relative."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import fcos_oracle
from paa_b200 import synthetic
from tests.test_oracle_fcos_vs_reference import fcos_batch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _cfg(radius=0.0, loss_type="iou", norm=False, gamma=2.0, alpha=0.25):
    return NS(MODEL=NS(FCOS=NS(LOSS_GAMMA=gamma, LOSS_ALPHA=alpha, FPN_STRIDES=[8, 16, 32, 64, 128],
                               CENTER_SAMPLING_RADIUS=radius, IOU_LOSS_TYPE=loss_type, NORM_REG_TARGETS=norm)))


def _check(b, locations, radius, loss_type, norm):
    import paa_b200
    prm = fcos_oracle.default_params(center_sampling_radius=radius, iou_loss_type=loss_type, norm_reg_targets=norm)
    ref_losses, ref_grads, asg = fcos_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes,
                                                             b.gt_labels, locations, prm)
    ev = paa_b200.make_fcos_loss_evaluator(_cfg(radius, loss_type, norm))
    ev.debug = True
    cls, reg, ctr, targets, _ = synthetic.to_device_inputs(b, requires_grad=True)
    locs = [p.cuda() for p in locations]
    losses = ev(locs, cls, reg, ctr, targets)
    assert len(losses) == 3
    sum(losses).backward()
    torch.cuda.synchronize()
    d = ev.last_debug
    assert np.array_equal(d["paa_labels"].cpu().numpy(), asg.labels.numpy())
    pos = asg.labels.numpy() > 0
    assert np.array_equal(d["matched_idx"].cpu().numpy()[pos], asg.matched.numpy()[pos])
    assert (d["matched_idx"].cpu().numpy()[~pos] == -1).all()
    np.testing.assert_allclose(d["normalisers"].cpu().numpy(), [asg.num_pos, asg.sum_centerness], rtol=1e-6)
    np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)
    for got, want in zip(cls + reg + ctr, ref_grads.box_cls + ref_grads.box_regression + ref_grads.centerness):
        np.testing.assert_allclose(got.grad.cpu().numpy(), want.numpy(), rtol=RTOL, atol=1e-9)
    return asg


@pytest.mark.parametrize("seed,hw,gt,radius,loss_type,norm", [
    (81, (320, 416), (2, 7), 0.0, "iou", False),
    (82, (800, 1333), (5, 40), 1.5, "giou", True),
    (83, (384, 512), (130, 150), 0.0, "linear_iou", False),
    (84, (800, 1333), (5, 40), 0.0, "giou", False),
])
def test_fcos_loss_against_oracle(seed, hw, gt, radius, loss_type, norm):
    b, locations = fcos_batch(seed, hw, gt)
    asg = _check(b, locations, radius, loss_type, norm)
    assert asg.num_pos > 0


def test_fcos_loss_without_positives():
    """A GT smaller than the grid spacing that contains no location: everything is background, the regression
    and centerness losses are empty sums (fcos/loss.py:274-277)."""
    b, locations = fcos_batch(85, (320, 416), 1, num_images=1)
    b.gt_boxes[0] = torch.tensor([[13.0, 13.0, 18.0, 18.0]])
    asg = _check(b, locations, 0.0, "iou", False)
    assert asg.num_pos == 0


def test_fcos_more_than_five_levels_is_rejected():
    import paa_b200
    b, locations = fcos_batch(86, (320, 416), 2)
    ev = paa_b200.make_fcos_loss_evaluator(_cfg())
    cls, reg, ctr, targets, _ = synthetic.to_device_inputs(b)
    locs = [p.cuda() for p in locations]
    with pytest.raises(IndexError):
        ev(locs + locs[:1], cls + cls[:1], reg + reg[:1], ctr + ctr[:1], targets)
