"""The label-forced variant of the oracle's stage 5 (oracle.paa_oracle.with_labels) reproduces the oracle's own
assignment when given the oracle's own labels, and reacts to a flipped positive."""
import numpy as np
import torch

from oracle import paa_oracle
from paa_b200 import synthetic


def test_with_labels_is_identity_on_own_labels_and_tracks_a_flip():
    b = synthetic.make_batch(seed=321, num_images=2, image_hw=(256, 320), gt_per_image=(2, 6))
    ls, grads, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                                b.anchors)
    forced = paa_oracle.with_labels(asg, asg.paa_labels.numpy(), b.box_regression, b.gt_boxes)
    assert torch.equal(forced.pos_inds, asg.pos_inds)
    assert torch.equal(forced.pos_ious, asg.pos_ious)
    pos = asg.pos_inds
    assert torch.equal(forced.reg_targets[pos], asg.reg_targets[pos])
    assert forced.num_pos == asg.num_pos and forced.sum_iou == asg.sum_iou
    ls2, grads2 = paa_oracle.losses_and_grads(b.box_cls, b.box_regression, b.iou_pred, forced)
    for a, c in zip(ls, ls2):
        assert float(a) == float(c)
    for x, y in zip(grads.box_cls + grads.box_regression + grads.iou_pred,
                    grads2.box_cls + grads2.box_regression + grads2.iou_pred):
        assert torch.equal(x, y)
    # drop one positive: num_pos falls by one and the classification loss changes
    labels = asg.paa_labels.numpy().copy()
    i, a = np.argwhere(labels > 0)[0]
    labels[i, a] = 0
    flipped = paa_oracle.with_labels(asg, labels, b.box_regression, b.gt_boxes)
    assert flipped.num_pos == asg.num_pos - 1
    ls3, _ = paa_oracle.losses_and_grads(b.box_cls, b.box_regression, b.iou_pred, flipped)
    assert float(ls3[0]) != float(ls[0])
