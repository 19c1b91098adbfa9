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


def _compare_loss_path(batch, oracle_kw, ref_kw, grad_rtol=1e-6):
    from oracle import make_golden, paa_oracle
    from tests.helpers import flat_levels
    ref = make_golden.run_reference_loss(batch, **ref_kw)
    losses, grads, asg = paa_oracle.assign_and_loss(batch.box_cls, batch.box_regression, batch.iou_pred,
                                                    batch.gt_boxes, batch.gt_labels, batch.anchors,
                                                    params=paa_oracle.default_params(**oracle_kw))
    assert np.array_equal(asg.matched_idx.numpy(), ref["matched_idx"])
    assert np.array_equal(asg.combined_loss.numpy(), ref["combined_loss"])
    assert np.array_equal(asg.paa_labels.numpy(), ref["paa_labels"])
    np.testing.assert_allclose([float(x) for x in losses], ref["losses"], rtol=1e-7)
    np.testing.assert_allclose(flat_levels(grads.box_cls), ref["grad_cls"], rtol=grad_rtol, atol=1e-10)
    np.testing.assert_allclose(flat_levels(grads.box_regression), ref["grad_reg"], rtol=grad_rtol, atol=1e-10)
    return ref


@pytest.mark.parametrize("oracle_kw,ref_kw", [
    (dict(topk=3), dict(TOPK=3)),
    (dict(topk=20), dict(TOPK=20)),
    (dict(iou_threshold=0.3), dict(IOU_THRESHOLD=0.3)),
    (dict(gamma=1.5, alpha=0.4), dict(LOSS_GAMMA=(1.5,), LOSS_ALPHA=(0.4,))),
    (dict(reg_loss_weight=2.0, iou_loss_weight=1.0), dict(REG_LOSS_WEIGHT=2.0, IOU_LOSS_WEIGHT=1.0)),
], ids=["topk3", "topk20", "iou_thr0.3", "gamma1.5_alpha0.4", "loss_weights"])
def test_loss_path_live_other_parameters(oracle_kw, ref_kw):
    """The parameters the GPU parity tests vary (tests/test_gpu_loss.py: TOPK 3 / 20, gamma / alpha) and the other
    cfg keys the evaluator reads (loss.py:34-47): the restatement follows the reference for each of them."""
    from paa_b200 import synthetic
    b = synthetic.make_batch(seed=34, num_images=2, image_hw=(288, 352), gt_per_image=(3, 10))
    _compare_loss_path(b, oracle_kw, ref_kw)


def test_loss_path_live_crowded_image():
    """More ground-truth boxes in one image than the kernels keep in one shared-memory GT list (128): the recorded
    reference still agrees with the restatement bit for bit (ties between overlapping GTs go to the first maximum,
    matcher.py:71 / loss.py:120)."""
    from paa_b200 import synthetic
    b = synthetic.make_batch(seed=35, num_images=1, image_hw=(416, 544), gt_per_image=140)
    ref = _compare_loss_path(b, {}, {})
    assert ref["gmm_n"].shape[0] > 0


def test_loss_path_live_c1_full_resolution():
    """BASELINE.json configs[0] (C1): 2 images of 800x1333 (22 400 anchors), 20 GT each, on the CPU."""
    from paa_b200 import synthetic
    b = synthetic.make_batch(seed=1000, num_images=2, image_hw=(800, 1333), gt_per_image=20)
    assert b.num_anchors == 22400
    _compare_loss_path(b, {}, {})


@pytest.mark.parametrize("ref_kw,oracle_kw", [
    (dict(score_voting=False), dict(score_voting=False)),
    (dict(use_iou_pred=False), dict()),
    (dict(detections_per_img=7), dict(detections_per_img=7)),
    (dict(pre_nms_top_n=40), dict(pre_nms_top_n=40)),
], ids=["no_voting", "no_iou_pred", "cut_to_7", "top40_per_level"])
def test_post_path_live_variants(ref_kw, oracle_kw):
    """PAAPostProcessor without score voting, without the IoU-prediction map, with a tight detections-per-image cut
    (kthvalue rule, inference.py:130-142) and a tight per-level cap: the restatement against the live reference."""
    from oracle import make_golden, post_oracle
    from paa_b200 import synthetic
    b = synthetic.make_inference_batch(seed=36, num_images=2, image_hw=(320, 384), n_objects=7)
    kw = dict(pre_nms_top_n=150)
    kw.update(ref_kw)
    ref = make_golden.run_reference_post(b, **kw)
    okw = dict(pre_nms_top_n=150)
    okw.update(oracle_kw)
    iou = b.iou_pred if kw.get("use_iou_pred", True) else None
    res = post_oracle.postprocess(b.box_cls, b.box_regression, iou, b.anchors, b.image_sizes,
                                  post_oracle.default_params(**okw))
    for i in range(2):
        db, ds, dl = post_oracle.canonical_rows(res[i].boxes, res[i].scores, res[i].labels)
        eb, es, el = post_oracle.canonical_rows(ref["det_boxes_%d" % i], ref["det_scores_%d" % i],
                                                ref["det_labels_%d" % i])
        assert np.array_equal(dl, el)
        np.testing.assert_allclose(ds, es, rtol=1e-6)
        np.testing.assert_allclose(db, eb, rtol=1e-5, atol=1e-3)


def test_loss_path_live_random_shapes():
    """Seeded sweep over ragged shapes (image sides that are not multiples of the strides, 1-3 images, 1-40 GT per
    image, trained-like and untrained heads): matched indices, anchor scores and PAA labels bit for bit, losses and
    gradients to rounding.  (A 300-case run of the same sweep found no difference.)"""
    from paa_b200 import synthetic
    rng = np.random.default_rng(2024)
    for _ in range(24):
        kw = dict(seed=int(rng.integers(0, 1 << 30)), num_images=int(rng.integers(1, 4)),
                  image_hw=(int(rng.integers(4, 16)) * 32 - int(rng.integers(0, 31)),
                            int(rng.integers(4, 16)) * 32 - int(rng.integers(0, 31))),
                  gt_per_image=(1, int(rng.choice([1, 2, 5, 15, 40]))), trained_like=bool(rng.integers(0, 2)))
        _compare_loss_path(synthetic.make_batch(**kw), {}, {})


def test_post_path_live_random_shapes():
    """Seeded sweep of the post-processor restatement against the live reference: ragged image sizes, sparse and
    dense candidate sets, per-level caps of 20 / 100 / 1000, cuts to 5 / 100 detections, voting on and off.
    (A 340-case run of the same sweep found no difference.)"""
    from oracle import make_golden, post_oracle
    from paa_b200 import synthetic
    rng = np.random.default_rng(2025)
    for _ in range(20):
        kw = dict(seed=int(rng.integers(0, 1 << 30)), num_images=int(rng.integers(1, 3)),
                  image_hw=(int(rng.integers(4, 14)) * 32 - int(rng.integers(0, 31)),
                            int(rng.integers(4, 14)) * 32 - int(rng.integers(0, 31))),
                  n_objects=int(rng.integers(1, 12)), cls_mean=float(rng.choice([-5.0, -4.0, -3.5, -3.0])))
        topn, dets = int(rng.choice([20, 100, 1000])), int(rng.choice([5, 100]))
        voting = bool(rng.integers(0, 2))
        b = synthetic.make_inference_batch(**kw)
        ref = make_golden.run_reference_post(b, pre_nms_top_n=topn, score_voting=voting, detections_per_img=dets)
        res = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes,
                                      post_oracle.default_params(pre_nms_top_n=topn, score_voting=voting,
                                                                 detections_per_img=dets))
        for i in range(b.num_images):
            db, ds, dl = post_oracle.canonical_rows(res[i].boxes, res[i].scores, res[i].labels)
            eb, es, el = post_oracle.canonical_rows(ref["det_boxes_%d" % i], ref["det_scores_%d" % i],
                                                    ref["det_labels_%d" % i])
            assert np.array_equal(dl, el), kw
            np.testing.assert_allclose(ds, es, rtol=1e-6)
            np.testing.assert_allclose(db, eb, rtol=1e-5, atol=1e-3)
