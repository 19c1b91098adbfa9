"""GPU parity of the CUDA post-processing path (through the C ABI) against the oracle, the golden
vectors recorded from the reference, and the reference's own NMS known-answer tests.

Orders the reference leaves unspecified are compared as sets (tie exemption (iii)):
detection rows are canonicalised by (label, score, box).  Bit-exact: labels, NMS keep sets, counts.
Floats: scores 1e-6 relative (one sigmoid/sqrt ulp), boxes 1e-4 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import make_golden, nms_oracle, post_oracle
from paa_b200 import synthetic
from tests.helpers import load_golden, post_case_batch, to_device_inputs

pytestmark = pytest.mark.gpu


def _postprocessor(**kw):
    import paa_b200
    test_kw = {}
    if "DETECTIONS_PER_IMG" in kw:
        test_kw["DETECTIONS_PER_IMG"] = kw.pop("DETECTIONS_PER_IMG")
    cfg = paa_b200.default_cfg(**kw)
    for k, v in test_kw.items():
        setattr(cfg.TEST, k, v)
    return paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))


def _run(pp, batch, use_iou=True):
    cls, reg, iou, _, anchors = to_device_inputs(batch)
    with torch.no_grad():
        out = pp(cls, reg, iou if use_iou else None, anchors)
    torch.cuda.synchronize()
    return out


def _assert_rows_match(got, ref, score_rtol=1e-6, box_rtol=1e-4, box_atol=2e-3):
    gb, gs, gl = post_oracle.canonical_rows(*got)
    rb, rs, rl = post_oracle.canonical_rows(*ref)
    assert gl.shape == rl.shape, (gl.shape, rl.shape)
    assert np.array_equal(gl, rl)
    np.testing.assert_allclose(gs, rs, rtol=score_rtol)
    np.testing.assert_allclose(gb, rb, rtol=box_rtol, atol=box_atol)


def _pre_lists(pp, i):
    d = pp.last_debug
    n = int(d["pre_count"][i].sum())
    return (d["pre_boxes"][i, :n].cpu().numpy(), d["pre_scores"][i, :n].cpu().numpy(),
            d["pre_labels"][i, :n].cpu().numpy())


@pytest.mark.parametrize("name", [c[0] for c in make_golden.POST_CASES])
def test_against_recorded_reference(name):
    ref = load_golden(name)
    b = post_case_batch(name)
    pp = _postprocessor(PRE_NMS_TOP_N=make_golden.POST_TOPN)
    pp.debug = True
    out = _run(pp, b)
    assert len(out) == int(ref["n_images"])
    for i, r in enumerate(out):
        # stage 1: per-level candidates (set comparison)
        _assert_rows_match(_pre_lists(pp, i),
                           (ref["pre_boxes_%d" % i], ref["pre_scores_%d" % i], ref["pre_labels_%d" % i]))
        # final detections incl. score voting
        assert r.mode == "xyxy" and tuple(r.size) == tuple(b.image_sizes[i])
        assert r.get_field("labels").dtype == torch.int64
        _assert_rows_match((r.bbox.cpu().numpy(), r.get_field("scores").cpu().numpy(),
                            r.get_field("labels").cpu().numpy()),
                           (ref["det_boxes_%d" % i], ref["det_scores_%d" % i], ref["det_labels_%d" % i]))


@pytest.mark.parametrize("name", [c[0] for c in make_golden.POST_CASES])
def test_nms_keep_set_teacher_forced(name):
    """The device NMS on the reference's own pre-NMS boxes reproduces the recorded keep indices."""
    from paa_b200.inference import ml_nms
    ref = load_golden(name)
    for i in range(int(ref["n_images"])):
        keep = ml_nms(torch.from_numpy(ref["pre_boxes_%d" % i]).cuda(),
                      torch.from_numpy(ref["pre_scores_%d" % i]).cuda(),
                      torch.from_numpy(ref["pre_labels_%d" % i]).cuda().float(), 0.6)
        assert np.array_equal(keep.cpu().numpy(), np.sort(ref["nms_keep_%d" % i]))


def test_nms_known_answers_from_reference_tests():
    """tests/test_nms.py:16-58 and :65-217 of the reference through the device kernel."""
    from paa_b200.inference import ml_nms
    kat = load_golden("nms_kat")
    for i in range(int(kat["n_cases"])):
        b = torch.from_numpy(kat["boxes_%d" % i]).cuda()
        s = torch.from_numpy(kat["scores_%d" % i]).cuda()
        keep = ml_nms(b, s, torch.zeros(len(s), device="cuda"), float(kat["thresh_%d" % i]))
        assert np.array_equal(keep.cpu().numpy(), kat["keep_%d" % i])
    assert ml_nms(torch.zeros((0, 4), device="cuda"), torch.zeros(0, device="cuda"),
                  torch.zeros(0, device="cuda"), 0.5).numel() == 0


def test_nms_random_against_oracle_including_single_class_crowd():
    """Label runs on both sides of the fused kernel's limit (post_nms_runs_kernel: one warp per run of up to 256
    boxes, longer runs flag the call for the mask + scan pair), sparse and crowded (most boxes suppressed, so the
    kept-list test of the fused kernel does the work)."""
    from paa_b200.inference import ml_nms
    rng = np.random.default_rng(3)
    cases = [(1, 1, 400), (63, 3, 400), (64, 1, 400), (65, 80, 400), (700, 5, 400), (3000, 1, 400),
             (255, 1, 400), (256, 1, 300), (257, 1, 300), (1000, 4, 150), (1500, 6, 80), (33, 1, 40), (512, 2, 60)]
    for n, n_cls, spread in cases:
        ctr = rng.uniform(0, spread, (n, 2))
        wh = rng.uniform(10, 120, (n, 2))
        boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], axis=1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32) / n          # distinct scores: no tie ambiguity
        labels = rng.integers(1, n_cls + 1, n).astype(np.float32)
        if n == 1500:                                               # one long run among short ones
            labels[:700] = 1.0
        want = nms_oracle.ml_nms_cpu(boxes, scores, labels, 0.5)
        got = ml_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                     torch.from_numpy(labels).cuda(), 0.5)
        assert np.array_equal(got.cpu().numpy(), want), (n, n_cls)


def test_postprocess_with_one_long_label_run():
    """One image whose candidates are dominated by a single class (a run of more than 256 boxes: the mask + scan
    pair) next to an ordinary image (fused per-run kernel) in the same call, against the oracle."""
    b = synthetic.make_inference_batch(seed=4200, num_images=2, image_hw=(384, 512), candidates_per_level=1500)
    for t in b.box_cls[:2]:
        t[0, 7] += 4.0                                              # image 0: class 8 everywhere on P3 / P4
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes)
    assert int((want[0].pre_labels == 8).sum()) > 256 and int(np.bincount(want[1].pre_labels.numpy()).max()) <= 256
    pp = _postprocessor()
    pp.debug = True
    out = _run(pp, b)
    for i, r in enumerate(out):
        w = want[i]
        _assert_rows_match(_pre_lists(pp, i), (w.pre_boxes.numpy(), w.pre_scores.numpy(), w.pre_labels.numpy()))
        _assert_rows_match((r.bbox.cpu().numpy(), r.get_field("scores").cpu().numpy(),
                            r.get_field("labels").cpu().numpy()),
                           (w.boxes.numpy(), w.scores.numpy(), w.labels.numpy()))


@pytest.mark.parametrize("per_level", [None, 4000])
def test_c4_against_oracle_full_resolution(per_level):
    """Config C4 shape (800x1333, 1000 candidates/level, voting on) for 2 images against the oracle,
    dense (49 % of logits are candidates) and detector-like sparse (~4000 per level)."""
    b = synthetic.make_inference_batch(seed=4000, num_images=2, image_hw=(800, 1333),
                                       candidates_per_level=per_level)
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes)
    pp = _postprocessor()
    pp.debug = True
    out = _run(pp, b)
    for i, r in enumerate(out):
        w = want[i]
        assert pp.last_debug["pre_count"][i].cpu().tolist() == w.level_counts
        _assert_rows_match(_pre_lists(pp, i), (w.pre_boxes.numpy(), w.pre_scores.numpy(), w.pre_labels.numpy()))
        _assert_rows_match((r.bbox.cpu().numpy(), r.get_field("scores").cpu().numpy(),
                            r.get_field("labels").cpu().numpy()),
                           (w.boxes.numpy(), w.scores.numpy(), w.labels.numpy()))


def test_variants_without_iou_pred_without_voting_and_skip_nms():
    b = synthetic.make_inference_batch(seed=41, num_images=2, image_hw=(256, 320), n_objects=5)
    for kw, okw, use_iou in (
            (dict(INFERENCE_SCORE_VOTING=False), dict(score_voting=False), True),
            (dict(), dict(), False),
            (dict(DETECTIONS_PER_IMG=20), dict(detections_per_img=20), True)):
        want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred if use_iou else None, b.anchors,
                                       b.image_sizes, post_oracle.default_params(pre_nms_top_n=150, **okw))
        out = _run(_postprocessor(PRE_NMS_TOP_N=150, **kw), b, use_iou=use_iou)
        for i, r in enumerate(out):
            _assert_rows_match((r.bbox.cpu().numpy(), r.get_field("scores").cpu().numpy(),
                                r.get_field("labels").cpu().numpy()),
                               (want[i].boxes.numpy(), want[i].scores.numpy(), want[i].labels.numpy()))
    # bbox_aug_enabled and not bbox_aug_vote: NMS skipped (inference.py:96-97)
    import paa_b200
    cfg = paa_b200.default_cfg(PRE_NMS_TOP_N=150)
    cfg.TEST.BBOX_AUG.ENABLED = True
    pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes,
                                   post_oracle.default_params(pre_nms_top_n=150, skip_nms=True))
    out = _run(pp, b)
    for i, r in enumerate(out):
        _assert_rows_match((r.bbox.cpu().numpy(), r.get_field("scores").cpu().numpy(),
                            r.get_field("labels").cpu().numpy()),
                           (want[i].boxes.numpy(), want[i].scores.numpy(), want[i].labels.numpy()))


def test_no_candidates_and_few_candidates():
    b = synthetic.make_inference_batch(seed=42, num_images=2, image_hw=(160, 192), n_objects=2)
    for t in b.box_cls:
        t[0].fill_(-9.0)                 # image 0: nothing passes the 0.05 threshold
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes)
    out = _run(_postprocessor(), b)
    assert len(out[0]) == 0 and out[0].bbox.shape == (0, 4)
    _assert_rows_match((out[1].bbox.cpu().numpy(), out[1].get_field("scores").cpu().numpy(),
                        out[1].get_field("labels").cpu().numpy()),
                       (want[1].boxes.numpy(), want[1].scores.numpy(), want[1].labels.numpy()))


def test_full_size_properties_c4():
    """Batch of 8 at full resolution: deterministic, count rule, survivors are mutually non-suppressing
    (NMS is idempotent), every survivor carries its pre-NMS score."""
    from paa_b200.inference import ml_nms
    b = synthetic.make_inference_batch(seed=4001, num_images=8, image_hw=(800, 1333))
    pp = _postprocessor(INFERENCE_SCORE_VOTING=False)
    pp.debug = True
    o1 = _run(pp, b)
    d1 = {k: v.clone() for k, v in pp.last_debug.items()}
    o2 = _run(pp, b)
    for a, c in zip(o1, o2):
        assert torch.equal(a.bbox, c.bbox) and torch.equal(a.get_field("scores"), c.get_field("scores"))
    for i, r in enumerate(o1):
        n_pre = int(d1["pre_count"][i].sum())
        assert d1["pre_count"][i].max() <= 1000
        kept = int(d1["nms_keep"][i, :n_pre].sum())
        s = r.get_field("scores")
        if kept > 100:
            assert len(r) >= 100
            ks = d1["pre_scores"][i, :n_pre][d1["nms_keep"][i, :n_pre].bool()]
            thr = torch.sort(ks, descending=True).values[99]
            assert len(r) == int((ks >= thr).sum())
        else:
            assert len(r) == kept
        again = ml_nms(r.bbox, s, r.get_field("labels").float(), 0.6)
        assert again.numel() == len(r)


def smoke_post():
    """Used by __graft_entry__.smoke(): one small post-processing call checked against the oracle."""
    b = synthetic.make_inference_batch(seed=78, num_images=2, image_hw=(256, 320), n_objects=4)
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes,
                                   post_oracle.default_params(pre_nms_top_n=100))
    out = _run(_postprocessor(PRE_NMS_TOP_N=100), b)
    for i, r in enumerate(out):
        _assert_rows_match((r.bbox.cpu().numpy(), r.get_field("scores").cpu().numpy(),
                            r.get_field("labels").cpu().numpy()),
                           (want[i].boxes.numpy(), want[i].scores.numpy(), want[i].labels.numpy()))
    print("smoke post: detections", [len(r) for r in out])


def test_nms_degenerate_thresholds_take_the_exact_path():
    """A threshold of zero or below switches the division-free IoU gate off (every pair goes through the reference's own
    quotient test): threshold 0 suppresses every overlapping pair, a negative one everything behind a class's best box."""
    from paa_b200.inference import ml_nms
    rng = np.random.default_rng(11)
    for n, n_cls, thr in ((200, 3, 0.0), (300, 2, -0.5), (90, 1, 1.0), (150, 4, 0.999)):
        ctr = rng.uniform(0, 200, (n, 2))
        wh = rng.uniform(10, 80, (n, 2))
        boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], axis=1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32) / n
        labels = rng.integers(1, n_cls + 1, n).astype(np.float32)
        want = nms_oracle.ml_nms_cpu(boxes, scores, labels, thr)
        got = ml_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), thr)
        assert np.array_equal(got.cpu().numpy(), want), (n, n_cls, thr)


def test_bulk_copy_candidate_kernel_matches_the_default(monkeypatch):
    """The opt-in cp.async.bulk ring variant of the candidate pass (PAA_POST_RING=1; measured slower, kept as a switch)
    selects exactly the same detections as the register-staged kernel."""
    b = synthetic.make_inference_batch(seed=4300, num_images=3, image_hw=(512, 672), candidates_per_level=2500)
    outs = []
    for ring in (False, True):
        if ring:
            monkeypatch.setenv("PAA_POST_RING", "1")
        else:
            monkeypatch.delenv("PAA_POST_RING", raising=False)
        outs.append(_run(_postprocessor(), b))
    for a, c in zip(*outs):
        assert torch.equal(a.bbox, c.bbox)
        assert torch.equal(a.get_field("scores"), c.get_field("scores"))
        assert torch.equal(a.get_field("labels"), c.get_field("labels"))
    assert sum(len(a) for a in outs[0]) > 0
