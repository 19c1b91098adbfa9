"""Channels-last (NHWC) head tensors are consumed in place (PaaLossArgs / PaaPostArgs `head_layout`): the same
logical [N, C, H, W] values in `torch.channels_last` memory format must give what the NCHW call gives -- the oracle's
assignment (bit-exact apart from the counted tie exemptions), losses and gradients within 1e-4, gradients returned in
the heads' own memory format, the post-processor's detections bit for bit -- and no copy of the logits.  The reference
is layout-blind here: it permutes and flattens whatever it is given (rpn/utils.py:10-14)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import paa_oracle
from paa_b200 import synthetic
from tests.helpers import (check_losses_and_grads_against_oracle, gmm_tie_exempt, to_device_inputs, topk_tie_exempt)

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _evaluator(flavour="paa", **kw):
    import paa_b200
    cfg = paa_b200.default_cfg(**kw)
    return paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))


def _step(ev, b, channels_last, use_iou=True):
    cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=True, channels_last=channels_last)
    with warnings.catch_warnings():
        warnings.simplefilter("error")              # a layout copy would warn: it must not happen
        losses = ev(cls, reg, iou if use_iou else None, targets, anchors, None)
    sum(losses).backward()
    torch.cuda.synchronize()
    return losses, cls, reg, (iou if use_iou else None)


@pytest.mark.parametrize("shape", [dict(seed=311, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8)),
                                   dict(seed=312, num_images=3, image_hw=(512, 672), gt_per_image=(1, 30)),
                                   dict(seed=1000, num_images=2, image_hw=(800, 1333), gt_per_image=20)])
def test_channels_last_loss_against_oracle(shape):
    b = synthetic.make_batch(**shape)
    ref_losses, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                                    b.anchors, with_grad=False)
    ev = _evaluator()
    ev.debug = True
    losses, cls, reg, iou = _step(ev, b, channels_last=True)
    d = ev.last_debug
    assert np.array_equal(d["matched_idx"].cpu().numpy().astype(np.int64), asg.matched_idx.numpy())
    pos = asg.iou_labels.numpy() > 0
    np.testing.assert_allclose(d["combined_loss"].cpu().numpy()[pos], asg.combined_loss.numpy()[pos], rtol=RTOL)
    got = d["paa_labels"].cpu().numpy()
    exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
    diff = {(i, int(asg.matched_idx[i][a])) for i, a in zip(*np.nonzero(got != asg.paa_labels.numpy()))}
    assert diff <= exempt and len(diff) <= 2, diff - exempt
    # the gradients come back in the heads' own memory format (what autograd hands to a channels-last conv backward)
    for t in cls + reg:
        assert t.grad.shape == t.shape
        assert t.grad.is_contiguous(memory_format=torch.channels_last), (t.shape, t.grad.stride())
    check_losses_and_grads_against_oracle(b, asg, got, losses, cls, reg, iou)
    if not diff:
        np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)


def test_channels_last_equals_nchw_call():
    """Same batch through both layouts: identical Matcher result and labels, losses / gradients to float32 rounding
    (the class sums of an anchor are added in a different order)."""
    b = synthetic.make_batch(seed=313, num_images=4, image_hw=(640, 800), gt_per_image=(3, 40))
    outs = []
    for cl in (False, True):
        ev = _evaluator()
        ev.debug = True
        losses, cls, reg, iou = _step(ev, b, channels_last=cl)
        outs.append((ev.last_debug, [float(x) for x in losses], [t.grad.contiguous().cpu() for t in cls],
                     [t.grad.contiguous().cpu() for t in reg], [t.grad.contiguous().cpu() for t in iou]))
    (d0, l0, gc0, gr0, gi0), (d1, l1, gc1, gr1, gi1) = outs
    assert torch.equal(d0["matched_idx"], d1["matched_idx"])
    assert torch.equal(d0["iou_labels"], d1["iou_labels"])
    same = torch.equal(d0["paa_labels"], d1["paa_labels"])
    pos = (d0["iou_labels"] > 0)
    # the 80 class terms of an anchor are added in a different order; where the labelled class dominates the sum its
    # negative term is swapped out again (negsum - term), which leaves the ordering noise of the big term on a small
    # rest: both layouts are within the contract's 1e-4 of the oracle, and of each other
    torch.testing.assert_close(d0["combined_loss"][pos], d1["combined_loss"][pos], rtol=1e-4, atol=0)
    if same:
        np.testing.assert_allclose(l0, l1, rtol=1e-5)
        for a, c in zip(gc0 + gr0 + gi0, gc1 + gr1 + gi1):
            torch.testing.assert_close(a, c, rtol=1e-4, atol=1e-9)
    else:       # a rounding-level tie may flip a positive set: bounded
        assert int((d0["paa_labels"] != d1["paa_labels"]).sum()) <= 45


def test_channels_last_without_iou_pred_and_no_grad():
    b = synthetic.make_batch(seed=314, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8))
    cfg_kw = dict(USE_IOU_PRED=False)
    ev = _evaluator(**cfg_kw)
    ev.debug = True
    losses, cls, reg, _ = _step(ev, b, channels_last=True, use_iou=False)
    assert len(losses) == 2
    ref_losses, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, None, b.gt_boxes, b.gt_labels,
                                                    b.anchors, with_grad=False,
                                                    params=paa_oracle.default_params(use_iou_pred=False))
    got = ev.last_debug["paa_labels"].cpu().numpy()
    exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
    diff = {(i, int(asg.matched_idx[i][a])) for i, a in zip(*np.nonzero(got != asg.paa_labels.numpy()))}
    assert diff <= exempt and len(diff) <= 1
    if not diff:
        np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)
    # no gradients requested
    cls2, reg2, iou2, targets, anchors = to_device_inputs(b, requires_grad=False, channels_last=True)
    with torch.no_grad():
        l2 = ev(cls2, reg2, None, targets, anchors, None)
    np.testing.assert_allclose([float(x) for x in l2], [float(x.detach()) for x in losses], rtol=1e-6)


def test_mixed_layout_is_copied_with_a_warning():
    """Logits channels-last, regression NCHW: the call's layout follows the logits, the odd tensors are copied and the
    user is told once."""
    from paa_b200 import loss as loss_mod
    loss_mod._warned_layout.clear()
    b = synthetic.make_batch(seed=315, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8))
    cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=False, channels_last=True)
    reg = [t.contiguous() for t in reg]
    ev = _evaluator()
    with pytest.warns(UserWarning, match="not dense in the call's layout"):
        with torch.no_grad():
            l_mixed = ev(cls, reg, iou, targets, anchors, None)
    cls0, reg0, iou0, targets0, anchors0 = to_device_inputs(b, requires_grad=False)
    with torch.no_grad():
        l_ref = _evaluator()(cls0, reg0, iou0, targets0, anchors0, None)
    np.testing.assert_allclose([float(x) for x in l_mixed], [float(x) for x in l_ref], rtol=1e-5)


@pytest.mark.parametrize("flavour", ["atss", "atss_iou", "retinanet", "fcos"])
def test_channels_last_sibling_losses_equal_nchw(flavour):
    """ATSS / RetinaNet (9 anchors per location, ignored anchors) / FCOS evaluators share the loss kernels: both
    layouts must agree (their NCHW results are pinned to the oracles in their own test files)."""
    import paa_b200
    from types import SimpleNamespace as NS
    res = []
    for cl in (False, True):
        if flavour in ("atss", "atss_iou"):
            from tests.test_gpu_atss_loss import _cfg
            b = synthetic.make_batch(seed=316, num_images=2, image_hw=(384, 512), gt_per_image=(2, 12))
            cfg = _cfg("IoU" if flavour == "atss_iou" else "ATSS")
            ev = paa_b200.make_atss_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
            cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=True, channels_last=cl)
            out = ev(cls, reg, iou, targets, anchors)
            heads = cls + reg + iou
        elif flavour == "retinanet":
            from oracle import retinanet_oracle
            from tests.test_gpu_retinanet_loss import _cfg
            b = synthetic.make_retinanet_batch(seed=317, num_images=2, image_hw=(320, 416), gt_per_image=(2, 7))
            ev = paa_b200.make_retinanet_loss_evaluator(_cfg(), NS(weights=retinanet_oracle.default_params().weights))
            cls, reg, _, targets, anchors = to_device_inputs(b, requires_grad=True, channels_last=cl)
            out = ev(anchors, cls, reg, targets)
            heads = cls + reg
        else:
            from tests.test_gpu_fcos_loss import _cfg
            from tests.test_oracle_fcos_vs_reference import fcos_batch
            b, locations = fcos_batch(318, (384, 512), (2, 9))
            ev = paa_b200.make_fcos_loss_evaluator(_cfg(1.5, "giou", True))
            cls, reg, iou, targets, _ = to_device_inputs(b, requires_grad=True, channels_last=cl)
            out = ev([p.cuda() for p in locations], cls, reg, iou, targets)
            heads = cls + reg + iou
        out = list(out)
        sum(out).backward()
        torch.cuda.synchronize()
        if cl:
            assert all(t.grad.is_contiguous(memory_format=torch.channels_last) for t in heads)
        res.append(([float(x) for x in out], [t.grad.contiguous().cpu() for t in heads]))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-5)
    for a, c in zip(res[0][1], res[1][1]):
        torch.testing.assert_close(a, c, rtol=1e-4, atol=1e-9)


def test_channels_last_postprocessor_equals_nchw():
    import paa_b200
    b = synthetic.make_inference_batch(seed=4100, num_images=3, image_hw=(512, 672), candidates_per_level=1500)
    cfg = paa_b200.default_cfg()
    outs = []
    for cl in (False, True):
        pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
        cls, reg, iou, _, anchors = to_device_inputs(b, channels_last=cl)
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            with torch.no_grad():
                outs.append(pp(cls, reg, iou, anchors))
        torch.cuda.synchronize()
    for a, c in zip(*outs):
        assert torch.equal(a.bbox, c.bbox)
        assert torch.equal(a.get_field("scores"), c.get_field("scores"))
        assert torch.equal(a.get_field("labels"), c.get_field("labels"))
    assert sum(len(a.bbox) for a in outs[0]) > 0


def test_channels_last_graph_mode_replays_on_new_targets():
    """Graph mode (one captured step replayed for every new target set) on channels-last heads: bit-identical to the
    eager launches, gradients in the heads' memory format."""
    import dataclasses
    import paa_b200
    kw = dict(num_images=2, image_hw=(384, 512))
    heads = synthetic.make_batch(seed=821, gt_per_image=(3, 9), **kw)
    other = synthetic.make_batch(seed=822, gt_per_image=(5, 30), **kw)
    batches = [heads, dataclasses.replace(heads, gt_boxes=other.gt_boxes, gt_labels=other.gt_labels)]
    cfg = paa_b200.default_cfg()
    ev_g = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    ev_g.use_graph = True
    ev_e = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    cls, reg, iou, _, anchors = to_device_inputs(heads, channels_last=True)
    for b in batches:
        _, _, _, targets, _ = to_device_inputs(b)
        lg, gg = ev_g.forward_backward(cls, reg, iou, targets, anchors)
        lg = lg.clone()
        gg = [t.clone() for t in gg["cls"] + gg["reg"] + gg["iou"]]
        le, ge = ev_e.forward_backward(cls, reg, iou, targets, anchors)
        torch.cuda.synchronize()
        assert torch.equal(lg, le)
        for a, c in zip(gg, ge["cls"] + ge["reg"] + ge["iou"]):
            assert a.shape == c.shape and torch.equal(a, c)
        assert all(t.is_contiguous(memory_format=torch.channels_last) for t in ge["cls"] + ge["reg"])
    assert len(ev_g._graphs) == 1
