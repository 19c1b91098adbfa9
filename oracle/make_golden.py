"""TEST INFRASTRUCTURE ONLY -- records golden vectors from the *real* reference.

Run in the build container (where ``/root/reference`` is mounted):

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

It imports the unmodified reference through ``oracle/ref_shim.py``, feeds it the synthetic batches
of ``paa_b200/synthetic.py`` and stores inputs' seeds plus the reference's outputs and the
intermediates that its own methods return (prepare_iou_based_targets loss.py:89-126, compute_paa
:128-236, forward_for_single_feature_map inference.py:36-82, select_over_all_levels :105-159).
Internals that are not returned are captured by instrumentation that still runs the reference's
code: a recording subclass of ``sklearn.mixture.GaussianMixture`` (fit inputs = sorted candidate
losses, fitted parameters, n_iter) and the ``_C.ml_nms`` stub (keep indices).

The fixtures are small (a few hundred KB) and are what ``tests/test_oracle_golden.py`` replays on
any machine, including the GPU box where the reference tree does not exist.
"""
import os
import sys

import numpy as np
import torch

from oracle import ref_shim  # noqa: E402
from paa_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, kwargs for synthetic.make_batch); kept tiny so the whole CPU suite stays in minutes.
LOSS_CASES = [
    ("loss_small", dict(seed=11, num_images=2, image_hw=(256, 320), gt_per_image=6)),
    ("loss_ragged", dict(seed=12, num_images=3, image_hw=(320, 416), gt_per_image=(1, 12))),
    ("loss_untrained", dict(seed=13, num_images=1, image_hw=(224, 224), gt_per_image=4,
                            trained_like=False)),
]
POST_CASES = [
    ("post_small", dict(seed=21, num_images=2, image_hw=(256, 320), n_objects=5)),
    ("post_tall", dict(seed=22, num_images=1, image_hw=(416, 288), n_objects=8, cls_mean=-3.5)),
]
POST_TOPN = 200   # PRE_NMS_TOP_N used for the (small) post fixtures so that the cap is exercised


def to_ref_inputs(ref, batch):
    targets, anchors = [], []
    for i in range(batch.num_images):
        t = ref.BoxList(batch.gt_boxes[i].clone(), batch.image_sizes[i], mode="xyxy")
        t.add_field("labels", batch.gt_labels[i].clone())
        targets.append(t)
        anchors.append([ref.BoxList(a, batch.image_sizes[i], mode="xyxy") for a in batch.anchors])
    return targets, anchors


def run_reference_loss(batch, use_iou_pred=True, with_grad=True, **paa_overrides):
    """Runs the reference's PAALossComputation on CPU; returns a dict of numpy arrays.  ``paa_overrides`` are
    MODEL.PAA keys (TOPK, IOU_THRESHOLD, LOSS_GAMMA / LOSS_ALPHA as 1-tuples, REG_LOSS_WEIGHT, ...)."""
    ref = ref_shim.load_reference()
    cfg = ref_shim.make_cfg(USE_IOU_PRED=use_iou_pred, **paa_overrides)
    ev = ref.loss.make_paa_loss_evaluator(cfg, ref.BoxCoder(cfg))
    targets, anchors = to_ref_inputs(ref, batch)
    fits = []
    base = ref.loss.skm.GaussianMixture

    class Recording(base):
        def fit(self, X, y=None):
            out = super().fit(X, y)
            fits.append(dict(x=np.array(X, np.float32).reshape(-1), w=np.array(self.weights_).reshape(2),
                             mu=np.array(self.means_).reshape(2),
                             var=np.array(self.covariances_, np.float64).reshape(2),
                             n_iter=int(self.n_iter_)))
            return out

    captured = {}
    orig_paa = ev.compute_paa

    def spy_paa(tg, an, labels_all, loss_all, matched_all):
        captured["iou_labels"] = labels_all.clone()
        captured["combined_loss"] = loss_all.clone()
        captured["matched_idx"] = matched_all.clone()
        out = orig_paa(tg, an, labels_all, loss_all, matched_all)
        captured["paa_labels"] = torch.stack(out[0], dim=0)
        captured["reg_targets"] = torch.cat(out[1], dim=0)
        return out

    ev.compute_paa = spy_paa
    ref.loss.skm.GaussianMixture = Recording
    try:
        cls = [x.clone().requires_grad_(with_grad) for x in batch.box_cls]
        reg = [x.clone().requires_grad_(with_grad) for x in batch.box_regression]
        iou = [x.clone().requires_grad_(with_grad) for x in batch.iou_pred] if use_iou_pred else None
        losses = ev(cls, reg, iou, targets, anchors, None)
        if with_grad:
            sum(losses).backward()
    finally:
        ref.loss.skm.GaussianMixture = base
    out = dict(losses=np.array([float(l.detach()) for l in losses], np.float64),
               matched_idx=captured["matched_idx"].numpy(),
               iou_labels=captured["iou_labels"].numpy(),
               combined_loss=captured["combined_loss"].numpy(),
               paa_labels=captured["paa_labels"].numpy(),
               reg_targets=captured["reg_targets"].numpy(),
               gmm_n=np.array([f["x"].shape[0] for f in fits], np.int64),
               gmm_x=np.concatenate([f["x"] for f in fits]) if fits else np.zeros(0, np.float32),
               gmm_w=np.array([f["w"] for f in fits]).reshape(-1, 2),
               gmm_mu=np.array([f["mu"] for f in fits]).reshape(-1, 2),
               gmm_var=np.array([f["var"] for f in fits]).reshape(-1, 2),
               gmm_n_iter=np.array([f["n_iter"] for f in fits], np.int64))
    if with_grad:
        def g(ts):
            return np.concatenate([(t.grad if t.grad is not None else torch.zeros_like(t)).permute(0, 2, 3, 1)
                                   .reshape(t.shape[0], -1, t.shape[1]).numpy() for t in ts], axis=1)
        out["grad_cls"] = g(cls)
        out["grad_reg"] = g(reg)
        if use_iou_pred:
            out["grad_iou"] = g(iou)[..., 0]
    return out


def run_reference_post(batch, pre_nms_top_n=1000, score_voting=True, use_iou_pred=True,
                       detections_per_img=100):
    """Runs the reference's PAAPostProcessor on CPU (channels-last inputs, shim 4)."""
    ref = ref_shim.load_reference()
    cfg = ref_shim.make_cfg(PRE_NMS_TOP_N=pre_nms_top_n, INFERENCE_SCORE_VOTING=score_voting,
                            USE_IOU_PRED=use_iou_pred)
    cfg.TEST.DETECTIONS_PER_IMG = detections_per_img
    pp = ref.inference.make_paa_postprocessor(cfg, ref.BoxCoder(cfg))
    _, anchors = to_ref_inputs(ref, batch)
    cl = lambda ts: [t.clone().contiguous(memory_format=torch.channels_last) for t in ts]
    pre = {}
    orig = pp.select_over_all_levels

    def spy(boxlists):
        pre["lists"] = [(b.bbox.clone(), b.get_field("scores").clone(), b.get_field("labels").clone())
                        for b in boxlists]
        return orig(boxlists)

    pp.select_over_all_levels = spy
    ref._C.ml_nms_calls.clear()
    with torch.no_grad():
        res = pp(cl(batch.box_cls), cl(batch.box_regression),
                 cl(batch.iou_pred) if use_iou_pred else None, anchors)
    out = dict(n_images=np.int64(batch.num_images))
    for i, r in enumerate(res):
        out["det_boxes_%d" % i] = r.bbox.numpy()
        out["det_scores_%d" % i] = r.get_field("scores").numpy()
        out["det_labels_%d" % i] = r.get_field("labels").numpy()
        b, s, l = pre["lists"][i]
        out["pre_boxes_%d" % i] = b.numpy()
        out["pre_scores_%d" % i] = s.numpy()
        out["pre_labels_%d" % i] = l.numpy()
        out["nms_keep_%d" % i] = ref._C.ml_nms_calls[i]
    return out


def record_nms_kats():
    """Replays the reference's own known-answer tests (tests/test_nms.py:11-58 and :60-217, Caffe2's
    UtilsNMSTest vectors) with ``box_nms`` pointing at a recorder: the test's inputs and the keep
    lists it expects are captured (no number is retyped by hand) and stored as nms_kat.npz."""
    import importlib.util
    ref = ref_shim.load_reference()
    spec = importlib.util.spec_from_file_location(
        "ref_test_nms", os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_nms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    calls, expected = [], []

    def recorder(boxes, scores, thresh):
        calls.append((boxes.numpy().copy(), scores.numpy().copy(), float(thresh)))
        return ref._C.nms(boxes, scores, thresh).numpy()

    orig_assert = np.testing.assert_array_equal

    def capture(actual, desired, *a, **k):
        expected.append(np.asarray(desired, np.int64))
        return orig_assert(actual, desired, *a, **k)     # also checks oracle/nms_oracle.py right here

    mod.box_nms = recorder
    np.testing.assert_array_equal = capture
    try:
        case = mod.TestNMS()
        case.test_nms_cpu()
        case.test_nms1_cpu()
    finally:
        np.testing.assert_array_equal = orig_assert
    assert len(calls) == len(expected) == 6
    out = dict(n_cases=np.int64(len(calls)))
    for i, ((b, s, t), e) in enumerate(zip(calls, expected)):
        out["boxes_%d" % i] = b
        out["scores_%d" % i] = s
        out["thresh_%d" % i] = np.float64(t)
        out["keep_%d" % i] = e
    return out


ANCHOR_CASES = [
    # name, PAA cfg overrides, padded (H, W) of the batch, per-image (h, w)
    ("paa_default", dict(), (96, 128), [(96, 128), (80, 100)]),
    ("multi_ratio_scale", dict(ASPECT_RATIOS=(0.5, 1.0, 2.0), SCALES_PER_OCTAVE=3, STRADDLE_THRESH=8,
                               ANCHOR_SIZES=(32, 64, 128), ANCHOR_STRIDES=(8, 16, 32)), (64, 96), [(64, 96), (50, 70)]),
    ("no_straddle_check", dict(STRADDLE_THRESH=-1), (64, 64), [(64, 64)]),
]


def record_anchor_cases():
    """Anchors (and their visibility fields) produced by the reference's own AnchorGenerator
    (modeling/rpn/anchor_generator.py:112-125,192-212) on CPU feature maps of the given sizes."""
    from types import SimpleNamespace
    ref = ref_shim.load_reference()
    out = {}
    for name, over, padded, sizes in ANCHOR_CASES:
        cfg = ref_shim.make_cfg(**over)
        gen = ref.make_anchor_generator_paa(cfg)
        strides = cfg.MODEL.PAA.ANCHOR_STRIDES
        grids = synthetic.level_grids(padded[0], padded[1], strides)
        fmaps = [torch.zeros((len(sizes), 1, h, w)) for (h, w) in grids]
        anchors = gen(SimpleNamespace(image_sizes=sizes), fmaps)
        out[name + "_grids"] = np.asarray(grids, np.int64)
        for i, per_image in enumerate(anchors):
            for l, bl in enumerate(per_image):
                if i == 0:
                    out["%s_l%d" % (name, l)] = bl.bbox.numpy().astype(np.float32)
                out["%s_vis_i%d_l%d" % (name, i, l)] = bl.get_field("visibility").numpy().astype(np.uint8)
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if "--anchors-only" in sys.argv:
        a = record_anchor_cases()
        np.savez_compressed(os.path.join(GOLDEN_DIR, "anchors.npz"), **a)
        print("anchors", len(a), "arrays")
        return 0
    a = record_anchor_cases()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "anchors.npz"), **a)
    print("anchors", len(a), "arrays")
    kats = record_nms_kats()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "nms_kat.npz"), **kats)
    print("nms_kat", int(kats["n_cases"]), "cases")
    for name, kw in LOSS_CASES:
        batch = synthetic.make_batch(**kw)
        out = run_reference_loss(batch)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(name, "losses", out["losses"], "fits", out["gmm_n"].shape[0],
              "positives", int((out["paa_labels"] > 0).sum()))
    for name, kw in POST_CASES:
        batch = synthetic.make_inference_batch(**kw)
        out = run_reference_post(batch, pre_nms_top_n=POST_TOPN)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(name, "detections", [out["det_scores_%d" % i].shape[0] for i in range(batch.num_images)],
              "pre-NMS", [out["pre_scores_%d" % i].shape[0] for i in range(batch.num_images)])


if __name__ == "__main__":
    sys.exit(main())
