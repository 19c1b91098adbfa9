"""Measurement aid: element-wise deviation of the regression / IoU-prediction gradients from the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paa_b200
from oracle import paa_oracle
from paa_b200 import synthetic
from paa_b200.synthetic import to_device_inputs
from tests.helpers import flat_levels

for seed, n, hw, gt in ((1000, 2, (800, 1333), 20), (2000, 4, (800, 1333), (1, 100)), (91, 2, (320, 416), (2, 8))):
    b = synthetic.make_batch(seed=seed, num_images=n, image_hw=hw, gt_per_image=gt)
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    ev.debug = True
    cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=True)
    losses = ev(cls, reg, iou, targets, anchors, None)
    sum(losses).backward()
    torch.cuda.synchronize()
    _, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels, b.anchors,
                                           with_grad=False)
    got_labels = ev.last_debug["paa_labels"].cpu().numpy()
    forced = paa_oracle.with_labels(asg, got_labels, b.box_regression, b.gt_boxes)
    ref_losses, ref = paa_oracle.losses_and_grads(b.box_cls, b.box_regression, b.iou_pred, forced)
    print("seed", seed, "labels equal", np.array_equal(got_labels, asg.paa_labels.numpy()),
          "losses", [float(x) for x in losses], [float(x) for x in ref_losses])
    for name, g, r in (("reg", flat_levels([t.grad for t in reg]), flat_levels(ref.box_regression)),
                       ("iou", flat_levels([t.grad for t in iou]), flat_levels(ref.iou_pred)),
                       ("cls", flat_levels([t.grad for t in cls]), flat_levels(ref.box_cls))):
        err = np.abs(g - r)
        rel = err / np.maximum(np.abs(r), 1e-30)
        nz = np.abs(r) > 0
        row_scale = np.abs(r).max(axis=-1, keepdims=True)
        rel_row = err / np.maximum(row_scale, 1e-30)
        bad = nz & (rel > 1e-4)
        print(" ", name, "nonzero", int(nz.sum()), "max rel", float(rel[nz].max()), "viol@1e-4", int(bad.sum()),
              "max err/rowmax", float(rel_row[nz.any(axis=-1)].max()), "max |ref| among violators",
              float(np.abs(r)[bad].max()) if bad.any() else 0.0, "global max |ref|", float(np.abs(r).max()))
        if bad.any():
            idx = np.argwhere(bad)[:5]
            for i in idx:
                print("     ", tuple(i), "got", g[tuple(i)], "ref", r[tuple(i)], "row", r[tuple(i[:-1])])
