"""The oracle restatement replayed against vectors recorded from the reference itself
(tests/golden/*.npz, written by oracle/make_golden.py in the build container) and against the
reference's own NMS known-answer tests.  Runs anywhere (no GPU, no /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import gmm_oracle, nms_oracle, paa_oracle, post_oracle, make_golden
from tests.helpers import flat_levels, load_golden, loss_case_batch, post_case_batch


@pytest.mark.parametrize("name", [c[0] for c in make_golden.LOSS_CASES])
def test_loss_restatement_matches_recorded_reference(name):
    ref = load_golden(name)
    b = loss_case_batch(name)
    losses, grads, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred,
                                                    b.gt_boxes, b.gt_labels, b.anchors)
    # index / mask outputs: bit-exact
    assert np.array_equal(asg.matched_idx.numpy(), ref["matched_idx"])
    assert np.array_equal(asg.iou_labels.numpy(), ref["iou_labels"])
    assert np.array_equal(asg.paa_labels.numpy(), ref["paa_labels"])
    # float outputs: same torch ops in the same order -> tight tolerance
    np.testing.assert_allclose(asg.combined_loss.numpy(), ref["combined_loss"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(asg.reg_targets.numpy(), ref["reg_targets"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose([float(x) for x in losses], ref["losses"], rtol=1e-6)
    np.testing.assert_allclose(flat_levels(grads.box_cls), ref["grad_cls"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(flat_levels(grads.box_regression), ref["grad_reg"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(flat_levels(grads.iou_pred)[..., 0], ref["grad_iou"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", [c[0] for c in make_golden.LOSS_CASES])
def test_recorded_gmm_fits_replay(name):
    """Every sklearn fit recorded inside the reference's compute_paa is reproduced by both oracle
    implementations: same iteration count, parameters to 1e-9 relative."""
    ref = load_golden(name)
    off = 0
    for i, n in enumerate(ref["gmm_n"]):
        x = ref["gmm_x"][off:off + n]
        off += n
        for impl in ("sklearn", "numpy"):
            fit = gmm_oracle.fit_two_component(x, impl=impl)
            assert fit["n_iter"] == ref["gmm_n_iter"][i], (impl, i)
            np.testing.assert_allclose(fit["weights"], ref["gmm_w"][i], rtol=1e-9)
            np.testing.assert_allclose(fit["means"], ref["gmm_mu"][i], rtol=1e-9)
            np.testing.assert_allclose(fit["variances"], ref["gmm_var"][i], rtol=1e-6)


def test_nms_known_answers_from_reference_tests():
    """tests/test_nms.py:16-58 (5 boxes x 5 thresholds) and :65-217 (53 boxes) of the reference."""
    kat = load_golden("nms_kat")
    assert int(kat["n_cases"]) == 6
    for i in range(int(kat["n_cases"])):
        b, s = kat["boxes_%d" % i], kat["scores_%d" % i]
        keep = nms_oracle.ml_nms_cpu(b, s, np.zeros(len(s), np.float32), float(kat["thresh_%d" % i]))
        assert np.array_equal(keep, kat["keep_%d" % i])


def test_nms_labels_isolate_classes():
    kat = load_golden("nms_kat")
    b, s = kat["boxes_5"], kat["scores_5"]
    labels = (np.arange(len(s)) % 3).astype(np.float32)
    keep = nms_oracle.ml_nms_cpu(b, s, labels, 0.5)
    per_class = np.concatenate([np.nonzero(labels == c)[0][nms_oracle.ml_nms_cpu(
        b[labels == c], s[labels == c], np.zeros(int((labels == c).sum())), 0.5)] for c in range(3)])
    assert np.array_equal(keep, np.sort(per_class))
    assert nms_oracle.ml_nms_cpu(b[:0], s[:0], labels[:0], 0.5).shape == (0,)


@pytest.mark.parametrize("name", [c[0] for c in make_golden.POST_CASES])
def test_post_restatement_matches_recorded_reference(name):
    ref = load_golden(name)
    b = post_case_batch(name)
    prm = post_oracle.default_params(pre_nms_top_n=make_golden.POST_TOPN)
    res = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes, prm)
    assert len(res) == int(ref["n_images"])
    for i, r in enumerate(res):
        pb, ps, pl = post_oracle.canonical_rows(r.pre_boxes, r.pre_scores, r.pre_labels)
        rb, rs, rl = post_oracle.canonical_rows(ref["pre_boxes_%d" % i], ref["pre_scores_%d" % i],
                                                ref["pre_labels_%d" % i])
        assert np.array_equal(pl, rl)
        np.testing.assert_allclose(ps, rs, rtol=1e-6)
        np.testing.assert_allclose(pb, rb, rtol=1e-6, atol=1e-4)
        db, ds, dl = post_oracle.canonical_rows(r.boxes, r.scores, r.labels)
        eb, es, el = post_oracle.canonical_rows(ref["det_boxes_%d" % i], ref["det_scores_%d" % i],
                                                ref["det_labels_%d" % i])
        assert np.array_equal(dl, el)
        np.testing.assert_allclose(ds, es, rtol=1e-6)
        np.testing.assert_allclose(db, eb, rtol=1e-5, atol=1e-3)


def test_matcher_edge_cases():
    """matcher.py:53-62 raises on empty inputs; :83-113 restores a GT's best anchor below threshold."""
    with pytest.raises(ValueError):
        paa_oracle.match_anchors(torch.zeros((0, 5)), 0.1)
    q = torch.tensor([[0.05, 0.02, 0.0], [0.04, 0.5, 0.0]])
    m = paa_oracle.match_anchors(q, 0.1)
    assert m.tolist() == [0, 1, -1]          # anchor 0 is GT 0's best although 0.05 < 0.1
    q = torch.tensor([[0.3, 0.3, 0.0]])      # tie for the GT's best: both restored/kept
    assert paa_oracle.match_anchors(q, 0.1).tolist() == [0, 0, -1]
