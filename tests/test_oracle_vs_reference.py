"""Live comparison of the restatement with the unmodified reference; only where the reference tree
is mounted (the build container).  Skipped on the GPU box."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def test_loss_path_live():
    from oracle import make_golden, paa_oracle
    from paa_b200 import synthetic
    from tests.helpers import flat_levels
    b = synthetic.make_batch(seed=31, num_images=2, image_hw=(288, 352), gt_per_image=(2, 9))
    ref = make_golden.run_reference_loss(b)
    losses, grads, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred,
                                                    b.gt_boxes, b.gt_labels, b.anchors)
    assert np.array_equal(asg.matched_idx.numpy(), ref["matched_idx"])
    assert np.array_equal(asg.paa_labels.numpy(), ref["paa_labels"])
    assert np.array_equal(asg.combined_loss.numpy(), ref["combined_loss"])
    np.testing.assert_allclose([float(x) for x in losses], ref["losses"], rtol=1e-7)
    np.testing.assert_allclose(flat_levels(grads.box_cls), ref["grad_cls"], rtol=1e-6, atol=1e-10)


def test_loss_path_live_without_iou_pred():
    from oracle import make_golden, paa_oracle
    from paa_b200 import synthetic
    b = synthetic.make_batch(seed=32, num_images=1, image_hw=(256, 256), gt_per_image=5)
    ref = make_golden.run_reference_loss(b, use_iou_pred=False)
    losses, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, None, b.gt_boxes,
                                                b.gt_labels, b.anchors,
                                                params=paa_oracle.default_params(use_iou_pred=False))
    assert len(losses) == 2
    assert np.array_equal(asg.paa_labels.numpy(), ref["paa_labels"])
    np.testing.assert_allclose([float(x) for x in losses], ref["losses"], rtol=1e-7)


def test_post_path_live():
    from oracle import make_golden, post_oracle
    from paa_b200 import synthetic
    b = synthetic.make_inference_batch(seed=33, num_images=1, image_hw=(320, 320), n_objects=6)
    ref = make_golden.run_reference_post(b, pre_nms_top_n=150)
    res = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes,
                                  post_oracle.default_params(pre_nms_top_n=150))
    db, ds, dl = post_oracle.canonical_rows(res[0].boxes, res[0].scores, res[0].labels)
    eb, es, el = post_oracle.canonical_rows(ref["det_boxes_0"], ref["det_scores_0"], ref["det_labels_0"])
    assert np.array_equal(dl, el)
    np.testing.assert_allclose(ds, es, rtol=1e-6)
    np.testing.assert_allclose(db, eb, rtol=1e-5, atol=1e-3)
