"""GPU: one captured CUDA graph of the assign+loss step serves new targets.  The ground-truth
ranges live in device memory (PaaLossArgs.gt_offsets_dev), grids are sized by capacities, so the graph captured on the
first batch is replayed for every later batch on the same head tensors -- bit-identical to the eager launches and
correct against the oracle."""
import numpy as np
import pytest
import torch

from oracle import paa_oracle
from paa_b200 import synthetic
from tests.helpers import check_losses_and_grads_against_oracle, gmm_tie_exempt, to_device_inputs, topk_tie_exempt

pytestmark = pytest.mark.gpu


def _evaluator(graph):
    import paa_b200
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    ev.use_graph = graph
    return ev


def _with_targets(heads_batch, targets_batch):
    """The head outputs of one batch with the ground truth of another (same image sizes)."""
    import dataclasses
    return dataclasses.replace(heads_batch, gt_boxes=targets_batch.gt_boxes, gt_labels=targets_batch.gt_labels)


def test_one_graph_serves_different_target_sets_and_matches_eager_and_oracle():
    kw = dict(num_images=3, image_hw=(384, 512))
    heads = synthetic.make_batch(seed=811, gt_per_image=(3, 9), **kw)
    batches = [heads,
               _with_targets(heads, synthetic.make_batch(seed=812, gt_per_image=(1, 20), **kw)),
               _with_targets(heads, synthetic.make_batch(seed=813, gt_per_image=(10, 40), **kw))]
    ev_g, ev_e = _evaluator(True), _evaluator(False)
    cls, reg, iou, _, anchors = to_device_inputs(heads)
    results = []
    for b in batches:
        _, _, _, targets, _ = to_device_inputs(b)
        lg, gg = ev_g.forward_backward(cls, reg, iou, targets, anchors)
        lg = lg.clone()
        gg = [t.clone() for t in gg["cls"] + gg["reg"] + gg["iou"]]       # the graph's buffers are reused next call
        le, ge = ev_e.forward_backward(cls, reg, iou, targets, anchors)
        torch.cuda.synchronize()
        assert torch.equal(lg, le), (lg, le)
        for a, c in zip(gg, ge["cls"] + ge["reg"] + ge["iou"]):
            assert torch.equal(a, c)
        results.append(lg.cpu().numpy())
    assert len(ev_g._graphs) == 1                                     # ONE captured graph for the three target sets
    step = next(iter(ev_g._graphs.values()))
    assert step.graph is not None and step.calls == 3
    assert not np.array_equal(results[0], results[1]) and not np.array_equal(results[1], results[2])
    # the replayed batches against the oracle, through the reference-facing call + autograd of the graph evaluator
    for b in batches[1:]:
        cls_g, reg_g, iou_g, targets, anchors_g = to_device_inputs(b, requires_grad=True)
        ev = _evaluator(True)
        # first call captures on ANOTHER target set, the second replays on this one
        _, _, _, other, _ = to_device_inputs(batches[0])
        ev(cls_g, reg_g, iou_g, other, anchors_g, None)
        losses = ev(cls_g, reg_g, iou_g, targets, anchors_g, None)
        sum(losses).backward()
        torch.cuda.synchronize()
        assert next(iter(ev._graphs.values())).calls == 2
        ev_d = _evaluator(False)
        ev_d.debug = True
        ev_d.forward_backward(cls_g, reg_g, iou_g, targets, anchors_g)
        got_labels = ev_d.last_debug["paa_labels"].cpu().numpy()
        _, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                               b.anchors, with_grad=False)
        exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
        diff = {(i, int(asg.matched_idx[i][a])) for i, a in zip(*np.nonzero(got_labels != asg.paa_labels.numpy()))}
        assert diff <= exempt and len(diff) <= 2, diff - exempt
        check_losses_and_grads_against_oracle(b, asg, got_labels, losses, cls_g, reg_g, iou_g)


def test_graph_mode_capacity_fallback_and_errors():
    import paa_b200
    b = synthetic.make_batch(seed=821, num_images=2, image_hw=(256, 320), gt_per_image=(4, 6))
    big = synthetic.make_batch(seed=822, num_images=2, image_hw=(256, 320), gt_per_image=(12, 14))
    cls, reg, iou, targets, anchors = to_device_inputs(b)
    _, _, _, big_targets, _ = to_device_inputs(big)
    ev = _evaluator(True)
    ev.gt_per_image_capacity = 8
    l1, _ = ev.forward_backward(cls, reg, iou, targets, anchors)
    l1 = l1.clone()
    # more ground truth than the graph was planned for: the call takes the eager path and is still right
    l2, _ = ev.forward_backward(cls, reg, iou, big_targets, anchors)
    want, _ = _evaluator(False).forward_backward(cls, reg, iou, big_targets, anchors)
    assert torch.equal(l2, want)
    assert next(iter(ev._graphs.values())).calls == 1
    l3, _ = ev.forward_backward(cls, reg, iou, targets, anchors)
    assert torch.equal(l3, l1)
    empty = paa_b200.BoxList(torch.zeros((0, 4), device="cuda"), targets[0].size)
    empty.add_field("labels", torch.zeros(0, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):          # matcher.py:53-58, checked on the host before the replay
        ev.forward_backward(cls, reg, iou, [empty, targets[1]], anchors)
    wrong = paa_b200.BoxList(targets[0].bbox, (999, 999))
    wrong.add_field("labels", targets[0].get_field("labels"))
    with pytest.raises(RuntimeError):        # boxlist_ops.py:95-97
        ev.forward_backward(cls, reg, iou, [wrong, targets[1]], anchors)


def test_device_offsets_path_matches_host_offsets_path_on_the_crowded_split():
    """> 128 GT per image: the coarse tiles' GT list is cut into parts sized by the per-image CAPACITY, not by the
    batch -- labels and losses must not depend on that plan."""
    b = synthetic.make_batch(seed=831, num_images=2, image_hw=(416, 512), gt_per_image=(130, 150))
    cls, reg, iou, targets, anchors = to_device_inputs(b)
    ev = _evaluator(True)
    ev.gt_per_image_capacity = 400            # 4 parts, where the eager plan takes 2
    lg, gg = ev.forward_backward(cls, reg, iou, targets, anchors)
    le, ge = _evaluator(False).forward_backward(cls, reg, iou, targets, anchors)
    assert torch.equal(lg, le)
    for a, c in zip(gg["cls"] + gg["reg"], ge["cls"] + ge["reg"]):
        assert torch.equal(a, c)
