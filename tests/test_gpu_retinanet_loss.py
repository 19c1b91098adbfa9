"""GPU parity of the RetinaNet training path (IoU matching + Matcher labelling + focal / smooth-L1 pass) against
oracle/retinanet_oracle.py, which tests/test_oracle_retinanet_vs_reference.py pins bit-exactly to the reference's
RetinaNetLossComputation.  Labels and Matcher results bit-exact; losses and gradients 1e-4 relative."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import retinanet_oracle
from paa_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _cfg(**kw):
    rn = dict(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4, BBOX_REG_BETA=0.11,
              BBOX_REG_WEIGHT=4.0)
    rn.update(kw)
    return NS(MODEL=NS(RETINANET=NS(**rn)))


def _device_inputs(b, requires_grad=True):
    cls, reg, _, targets, anchors = synthetic.to_device_inputs(b, requires_grad=requires_grad)
    return cls, reg, targets, anchors


def _check(b, cfg, prm):
    import paa_b200
    ref_losses, ref_grads, asg = retinanet_oracle.assign_and_loss(b.box_cls, b.box_regression, b.gt_boxes,
                                                                  b.gt_labels, b.anchors, prm)
    ev = paa_b200.make_retinanet_loss_evaluator(cfg, NS(weights=prm.weights))
    ev.debug = True
    cls, reg, targets, anchors = _device_inputs(b)
    losses = ev(anchors, cls, reg, targets)
    assert len(losses) == 2
    sum(losses).backward()
    torch.cuda.synchronize()
    d = ev.last_debug
    assert np.array_equal(d["matched_idx"].cpu().numpy(), asg.matched.numpy())
    assert np.array_equal(d["paa_labels"].cpu().numpy(), asg.labels.numpy())
    assert float(d["normalisers"][0]) == asg.num_pos
    np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)
    for got, want in zip(cls, ref_grads.box_cls):
        np.testing.assert_allclose(got.grad.cpu().numpy(), want.numpy(), rtol=RTOL, atol=1e-9)
    for got, want in zip(reg, ref_grads.box_regression):
        np.testing.assert_allclose(got.grad.cpu().numpy(), want.numpy(), rtol=RTOL, atol=1e-9)
    return asg


@pytest.mark.parametrize("seed,hw,gt", [(71, (320, 416), (2, 7)), (72, (800, 1333), (5, 40))])
def test_retinanet_loss_against_oracle(seed, hw, gt):
    b = synthetic.make_retinanet_batch(seed=seed, num_images=2, image_hw=hw, gt_per_image=gt)
    asg = _check(b, _cfg(), retinanet_oracle.default_params())
    assert (asg.labels == -1).any() and (asg.labels > 0).any()


def test_retinanet_loss_other_parameters_and_crowded_image():
    """Other thresholds / beta / normaliser / focal parameters, and > 128 GTs in an image (the coarse levels' GT
    list is then matched in parts that meet in an atomicMax)."""
    b = synthetic.make_retinanet_batch(seed=73, num_images=2, image_hw=(384, 512), gt_per_image=(130, 150))
    cfg = _cfg(FG_IOU_THRESHOLD=0.6, BG_IOU_THRESHOLD=0.3, BBOX_REG_BETA=0.25, BBOX_REG_WEIGHT=1.0, LOSS_GAMMA=1.5,
               LOSS_ALPHA=0.4)
    prm = retinanet_oracle.default_params(fg_iou_threshold=0.6, bg_iou_threshold=0.3, bbox_reg_beta=0.25,
                                          bbox_reg_weight=1.0, gamma=1.5, alpha=0.4)
    _check(b, cfg, prm)


def test_retinanet_fallback_that_patches_ignored_anchors_one_by_one(monkeypatch):
    """The loss pass normally skips ignored anchors in stream through a bitmap; the fallback (used when a level's
    tensor has 2^32 elements or more) takes their terms back per anchor instead."""
    monkeypatch.setenv("PAA_RETINA_PATCH", "1")
    b = synthetic.make_retinanet_batch(seed=76, num_images=2, image_hw=(320, 416), gt_per_image=(3, 9))
    asg = _check(b, _cfg(), retinanet_oracle.default_params())
    assert (asg.labels == -1).any()


def test_retinanet_positives_from_the_low_quality_restore_only():
    """One GT far smaller than every anchor: no IoU reaches the thresholds, the only positives are the GT's best
    anchors restored by allow_low_quality_matches (matcher.py:83-113)."""
    b = synthetic.make_retinanet_batch(seed=74, num_images=1, image_hw=(320, 416), gt_per_image=1)
    b.gt_boxes[0] = torch.tensor([[100.0, 100.0, 103.0, 102.0]])
    asg = _check(b, _cfg(), retinanet_oracle.default_params())
    assert asg.num_pos >= 1 and not (asg.labels == -1).any()


def test_retinanet_image_without_ground_truth_raises():
    import paa_b200
    b = synthetic.make_retinanet_batch(seed=75, num_images=2, image_hw=(320, 416), gt_per_image=2)
    b.gt_boxes[1] = b.gt_boxes[1][:0]
    b.gt_labels[1] = b.gt_labels[1][:0]
    ev = paa_b200.make_retinanet_loss_evaluator(_cfg(), NS(weights=(10.0, 10.0, 5.0, 5.0)))
    cls, reg, targets, anchors = _device_inputs(b, requires_grad=False)
    with pytest.raises(ValueError, match="No ground-truth boxes"):
        ev(anchors, cls, reg, targets)
