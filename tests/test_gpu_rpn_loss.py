"""GPU parity of the plain RPN training loss (paa_b200.rpn_loss: Matcher labelling on the RetinaNet path's kernels,
torch.randperm sampler, paa_rpn_loss) against oracle/rpn_oracle.py, which tests/test_oracle_rpn_vs_reference.py pins
bit for bit to the reference's RPNLossComputation under the same seed.  Matcher results and labels bit-exact; losses
and gradients 1e-4 relative on the oracle's own sample (a random draw has no parity beyond that)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import rpn_oracle
from tests.test_oracle_rpn_vs_reference import OTHER, rpn_batch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _cfg(prm):
    return NS(MODEL=NS(RPN=NS(FG_IOU_THRESHOLD=prm.fg_iou_threshold, BG_IOU_THRESHOLD=prm.bg_iou_threshold,
                              BATCH_SIZE_PER_IMAGE=prm.batch_size_per_image, POSITIVE_FRACTION=prm.positive_fraction)))


def _device_inputs(b, objectness, vis, channels_last=False):
    from paa_b200.structures import BoxList
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    obj = [t.cuda().contiguous(memory_format=fmt).requires_grad_(True) for t in objectness]
    reg = [t.cuda().contiguous(memory_format=fmt).requires_grad_(True) for t in b.box_regression]
    anc = [a.cuda() for a in b.anchors]
    targets, anchors = [], []
    for i in range(b.num_images):
        targets.append(BoxList(b.gt_boxes[i].cuda(), b.image_sizes[i], mode="xyxy"))
        per_level, o = [], 0
        for a in anc:
            bl = BoxList(a, b.image_sizes[i], mode="xyxy")
            bl.add_field("visibility", vis[i][o:o + a.shape[0]].cuda())
            o += a.shape[0]
            per_level.append(bl)
        anchors.append(per_level)
    return obj, reg, targets, anchors


def _evaluator(prm):
    import paa_b200
    return paa_b200.make_rpn_loss_evaluator(_cfg(prm), NS(weights=prm.weights))


@pytest.mark.parametrize("seed,hw,gt,other", [(171, (320, 416), (2, 7), {}), (172, (800, 1333), (5, 40), {}),
                                              (173, (384, 512), (130, 150), OTHER), (174, (320, 416), 1, OTHER)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_rpn_loss_against_oracle(seed, hw, gt, other, channels_last):
    prm = rpn_oracle.default_params(**other)
    b, objectness, vis = rpn_batch(seed, hw, gt)
    torch.manual_seed(seed)
    ref_losses, ref_grads, asg = rpn_oracle.assign_and_loss(objectness, b.box_regression, b.gt_boxes, b.anchors, vis,
                                                            prm)
    ev = _evaluator(prm)
    ev.sample_override = (asg.sampled_pos, asg.sampled_neg)          # the oracle's draw
    obj, reg, targets, anchors = _device_inputs(b, objectness, vis, channels_last)
    losses = ev(anchors, obj, reg, targets)
    assert len(losses) == 2
    sum(losses).backward()
    torch.cuda.synchronize()
    d = ev.last_debug
    assert np.array_equal(d["matched_idx"].cpu().numpy(), torch.stack(asg.matched).numpy())
    assert np.array_equal(d["labels"].cpu().numpy(), torch.stack(asg.labels).numpy())
    np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)
    for got, want in zip(obj + reg, ref_grads.objectness + ref_grads.box_regression):
        assert got.grad.shape == want.shape
        if channels_last:
            assert got.grad.is_contiguous(memory_format=torch.channels_last)
        np.testing.assert_allclose(got.grad.cpu().numpy(), want.numpy(), rtol=RTOL, atol=1e-9)


def test_rpn_loss_with_its_own_random_sample():
    """The evaluator's own draw (torch.randperm on the device): the sample obeys the sampler's rules, and the losses /
    gradients are the oracle's for that very sample."""
    prm = rpn_oracle.default_params()
    b, objectness, vis = rpn_batch(175, (384, 512), (3, 12), num_images=3)
    ev = _evaluator(prm)
    obj, reg, targets, anchors = _device_inputs(b, objectness, vis)
    torch.manual_seed(5)
    losses = ev(anchors, obj, reg, targets)
    (2.0 * losses[0] + 0.5 * losses[1]).backward()                   # upstream gradients other than one
    torch.cuda.synchronize()
    d = ev.last_debug
    labels = d["labels"].cpu()
    A = labels.shape[1]
    pos, neg = d["sampled_pos"].cpu(), d["sampled_neg"].cpu()
    assert (labels.reshape(-1)[pos] == 1).all() and (labels.reshape(-1)[neg] == 0).all()
    for i in range(b.num_images):
        n_pos = int(((pos // A) == i).sum())
        n_neg = int(((neg // A) == i).sum())
        want_pos = min(int((labels[i] >= 1).sum()), int(prm.batch_size_per_image * prm.positive_fraction))
        assert n_pos == want_pos
        assert n_neg == min(int((labels[i] == 0).sum()), prm.batch_size_per_image - n_pos)
    ref_losses, ref_grads, _ = rpn_oracle.assign_and_loss(objectness, b.box_regression, b.gt_boxes, b.anchors, vis, prm,
                                                          sampled=(pos, neg), with_grad=False)
    np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)
    o2 = [t.detach().clone().requires_grad_(True) for t in objectness]
    r2 = [t.detach().clone().requires_grad_(True) for t in b.box_regression]
    asg = rpn_oracle.assign(b.gt_boxes, b.anchors, vis, prm)
    ls = rpn_oracle.losses(o2, r2, asg, pos, neg)
    (2.0 * ls[0] + 0.5 * ls[1]).backward()
    for got, want in zip(obj + reg, o2 + r2):
        np.testing.assert_allclose(got.grad.cpu().numpy(), want.grad.numpy(), rtol=RTOL, atol=1e-9)
    # a second call draws another sample
    l2 = ev(anchors, obj, reg, targets)
    assert not torch.equal(ev.last_debug["sampled_neg"].cpu(), neg)
    assert all(torch.isfinite(x) for x in l2)


def test_rpn_loss_error_behaviour_and_no_grad():
    from paa_b200.structures import BoxList
    prm = rpn_oracle.default_params()
    b, objectness, vis = rpn_batch(176, (320, 416), (2, 7))
    ev = _evaluator(prm)
    obj, reg, targets, anchors = _device_inputs(b, objectness, vis)
    with torch.no_grad():
        l0 = ev(anchors, [t.detach() for t in obj], [t.detach() for t in reg], targets)
    assert all(torch.isfinite(x) for x in l0) and not l0[0].requires_grad
    empty = [targets[0], BoxList(torch.zeros((0, 4), device="cuda"), b.image_sizes[1], mode="xyxy")]
    with pytest.raises(ValueError):                                  # matcher.py:53-58
        ev(anchors, obj, reg, empty)
    with pytest.raises(RuntimeError):                                # objectness must have one channel per anchor
        ev(anchors, reg, reg, targets)
