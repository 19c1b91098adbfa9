"""GPU parity of the operators next to the hot path: device anchor generation against
anchors recorded from the reference's own AnchorGenerator, boxlist_iou against the oracle bit for bit, and the
ATSS post-processor (same kernels) against the oracle's ATSS flavour."""
import numpy as np
import pytest
import torch

from oracle import make_golden, paa_oracle, post_oracle, ref_shim
from paa_b200 import synthetic
from tests.helpers import load_golden, to_device_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", make_golden.ANCHOR_CASES, ids=[c[0] for c in make_golden.ANCHOR_CASES])
def test_anchor_generator_against_recorded_reference(case):
    import paa_b200
    from paa_b200.anchor_generator import make_anchor_generator_paa
    name, over, padded, sizes = case
    gold = load_golden("anchors")
    over = dict(over)
    cfg = paa_b200.default_cfg(**over)
    gen = make_anchor_generator_paa(cfg)
    grids = [tuple(int(v) for v in g) for g in gold[name + "_grids"]]
    fmaps = [torch.zeros((len(sizes), 1, h, w), device="cuda") for (h, w) in grids]
    anchors = gen(sizes, fmaps)
    assert len(anchors) == len(sizes)
    for i, per_image in enumerate(anchors):
        assert len(per_image) == len(grids)
        for l, bl in enumerate(per_image):
            assert bl.mode == "xyxy" and tuple(bl.size) == (sizes[i][1], sizes[i][0])
            np.testing.assert_array_equal(bl.bbox.cpu().numpy(), gold["%s_l%d" % (name, l)])
            np.testing.assert_array_equal(bl.get_field("visibility").cpu().numpy().astype(np.uint8),
                                          gold["%s_vis_i%d_l%d" % (name, i, l)])
    # the generated anchors are the ones the synthetic batches (and the bench) use
    if name == "paa_default":
        b = synthetic.make_batch(seed=1, num_images=1, image_hw=padded, gt_per_image=1)
        for l, a in enumerate(b.anchors):
            np.testing.assert_array_equal(anchors[0][l].bbox.cpu().numpy(), a.numpy())
    # second call: served from the cache, same tensors
    again = gen(sizes, fmaps)
    assert again[0][0].bbox.data_ptr() == anchors[0][0].bbox.data_ptr()


def test_full_size_anchor_grid_feeds_the_loss():
    """800x1333 grid generated on the device drives the evaluator exactly like the host-built anchors."""
    import paa_b200
    from paa_b200.anchor_generator import make_anchor_generator_paa
    b = synthetic.make_batch(seed=1000, num_images=2, image_hw=(800, 1333), gt_per_image=20)
    cfg = paa_b200.default_cfg()
    gen = make_anchor_generator_paa(cfg)
    cls, reg, iou, targets, anchors_host = to_device_inputs(b)
    anchors_dev = gen([(h, w) for (w, h) in b.image_sizes], cls)
    for l, a in enumerate(b.anchors):
        assert torch.equal(anchors_dev[0][l].bbox.cpu(), a)
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    l1 = [float(x) for x in ev(cls, reg, iou, targets, anchors_host, None)]
    l2 = [float(x) for x in ev(cls, reg, iou, targets, anchors_dev, None)]
    assert l1 == l2


@pytest.mark.parametrize("n1,n2", [(1, 1), (7, 300), (100, 5000), (333, 22400), (0, 5), (5, 0)])
def test_boxlist_iou_bit_exact(n1, n2):
    import paa_b200
    g = torch.Generator().manual_seed(n1 * 1000 + n2)
    def boxes(n):
        xy = torch.rand((n, 2), generator=g) * 600
        wh = torch.rand((n, 2), generator=g) * 300
        return torch.cat([xy, xy + wh], dim=1)
    b1, b2 = boxes(n1), boxes(n2)
    if n1 and n2:
        b2[0] = b1[0]                          # identical boxes -> IoU exactly 1
    ref = paa_oracle.iou_matrix(b1, b2) if n1 and n2 else torch.zeros((n1, n2))
    got = paa_b200.boxlist_iou(paa_b200.BoxList(b1.cuda(), (900, 900)), paa_b200.BoxList(b2.cuda(), (900, 900)))
    assert got.shape == (n1, n2)
    assert torch.equal(got.cpu(), ref)
    with pytest.raises(RuntimeError):
        paa_b200.boxlist_iou(paa_b200.BoxList(b1.cuda(), (900, 900)), paa_b200.BoxList(b2.cuda(), (901, 900)))


def test_boxlist_iou_against_the_reference_function():
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    import paa_b200
    ref = ref_shim.load_reference()
    b = synthetic.make_batch(seed=3, num_images=1, image_hw=(256, 320), gt_per_image=12)
    anchors = torch.cat(b.anchors)
    want = ref.boxlist_ops.boxlist_iou(ref.BoxList(b.gt_boxes[0], (320, 256)), ref.BoxList(anchors, (320, 256)))
    got = paa_b200.boxlist_iou(paa_b200.BoxList(b.gt_boxes[0].cuda(), (320, 256)),
                               paa_b200.BoxList(anchors.cuda(), (320, 256)))
    assert torch.equal(got.cpu(), want)


def test_atss_postprocessor_against_oracle():
    import paa_b200
    from types import SimpleNamespace as NS
    from paa_b200.inference import make_atss_postprocessor
    b = synthetic.make_inference_batch(seed=41, num_images=2, image_hw=(320, 416), candidates_per_level=600)
    cfg = NS(MODEL=NS(ATSS=NS(INFERENCE_TH=0.05, PRE_NMS_TOP_N=200, NMS_TH=0.6, NUM_CLASSES=81,
                              REGRESSION_TYPE="BOX")),
             TEST=NS(DETECTIONS_PER_IMG=100, BBOX_AUG=NS(ENABLED=False, VOTE=False)))
    pp = make_atss_postprocessor(cfg, paa_b200.BoxCoder(cfg))
    cls, reg, ctr, _, anchors = to_device_inputs(b)
    got = pp(cls, reg, ctr, anchors)
    prm = post_oracle.default_params(pre_nms_top_n=200, score_voting=False, flavour="atss")
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes, prm)
    for i in range(b.num_images):
        gb, gs, gl = post_oracle.canonical_rows(got[i].bbox.cpu(), got[i].get_field("scores").cpu(),
                                                got[i].get_field("labels").cpu())
        wb, ws, wl = post_oracle.canonical_rows(want[i].boxes, want[i].scores, want[i].labels)
        assert gl.shape == wl.shape and np.array_equal(gl, wl)
        np.testing.assert_allclose(gs, ws, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(gb, wb, rtol=1e-4, atol=1e-3)


def _same_detections(got, want, i):
    gb, gs, gl = post_oracle.canonical_rows(got[i].bbox.cpu(), got[i].get_field("scores").cpu(),
                                            got[i].get_field("labels").cpu())
    wb, ws, wl = post_oracle.canonical_rows(want[i].boxes, want[i].scores, want[i].labels)
    assert gl.shape == wl.shape and np.array_equal(gl, wl)
    np.testing.assert_allclose(gs, ws, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(gb, wb, rtol=1e-4, atol=1e-3)


def test_retinanet_postprocessor_against_oracle():
    """Nine anchors per location, RPN BoxCoder decode, no third head output."""
    import paa_b200
    from types import SimpleNamespace as NS
    from paa_b200.inference import make_retinanet_postprocessor
    b = synthetic.make_retinanet_batch(seed=43, num_images=2, image_hw=(256, 320))
    cfg = NS(MODEL=NS(RETINANET=NS(INFERENCE_TH=0.05, PRE_NMS_TOP_N=300, NMS_TH=0.4, NUM_CLASSES=81)),
             TEST=NS(DETECTIONS_PER_IMG=100))
    pp = make_retinanet_postprocessor(cfg, NS(weights=(10.0, 10.0, 5.0, 5.0), bbox_xform_clip=float(np.log(1000.0 / 16))))
    anc = [a.cuda() for a in b.anchors]
    n_img = b.box_cls[0].shape[0]
    anchors = [[paa_b200.BoxList(a, b.image_sizes[i]) for a in anc] for i in range(n_img)]
    got = pp(anchors, [t.cuda() for t in b.box_cls], [t.cuda() for t in b.box_regression])
    prm = post_oracle.default_params(pre_nms_top_n=300, nms_thresh=0.4, score_voting=False, flavour="retinanet")
    want = post_oracle.postprocess(b.box_cls, b.box_regression, None, b.anchors, b.image_sizes, prm)
    for i in range(n_img):
        assert len(got[i]) > 0
        _same_detections(got, want, i)


def test_fcos_postprocessor_against_oracle():
    """Anchor-free: points + distances to the four edges, centerness in the score."""
    import paa_b200
    from types import SimpleNamespace as NS
    from paa_b200.inference import make_fcos_postprocessor
    b = synthetic.make_inference_batch(seed=44, num_images=2, image_hw=(256, 320), candidates_per_level=500)
    locs = synthetic.fcos_locations(b.grids)
    reg = [torch.exp(t * 0.5) * 8.0 * s for t, s in zip(b.box_regression, synthetic.STRIDES)]   # positive distances
    cfg = NS(MODEL=NS(FCOS=NS(INFERENCE_TH=0.05, PRE_NMS_TOP_N=200, NMS_TH=0.6, NUM_CLASSES=81)),
             TEST=NS(DETECTIONS_PER_IMG=100, BBOX_AUG=NS(ENABLED=False)))
    pp = make_fcos_postprocessor(cfg)
    sizes = [(h, w) for (w, h) in b.image_sizes]
    got = pp([l.cuda() for l in locs], [t.cuda() for t in b.box_cls], [t.cuda() for t in reg],
             [t.cuda() for t in b.iou_pred], sizes)
    prm = post_oracle.default_params(pre_nms_top_n=200, score_voting=False, flavour="fcos")
    want = post_oracle.postprocess(b.box_cls, reg, b.iou_pred, [torch.cat([l, l], 1) for l in locs], b.image_sizes, prm)
    for i in range(b.num_images):
        assert len(got[i]) > 0
        _same_detections(got, want, i)


def _tta_detections(seed, n_classes_used=6, per_class=(0, 1, 2, 40, 150, 400)):
    """Pooled multi-scale detections: clusters of overlapping boxes per class, distinct scores."""
    g = torch.Generator().manual_seed(seed)
    boxes, labels = [], []
    for j, n in enumerate(per_class[:n_classes_used]):
        if n == 0:
            continue
        n_obj = max(1, n // 7)
        ctr = torch.rand((n_obj, 2), generator=g) * 500
        size = 40 + torch.rand((n_obj, 2), generator=g) * 120
        which = torch.randint(0, n_obj, (n,), generator=g)
        xy = ctr[which] + torch.randn((n, 2), generator=g) * 5
        wh = size[which] * (1 + 0.08 * torch.randn((n, 2), generator=g))
        boxes.append(torch.cat([xy, xy + wh], 1))
        labels.append(torch.full((n,), 3 * j + 1, dtype=torch.int64))
    boxes, labels = torch.cat(boxes), torch.cat(labels)
    scores = torch.randperm(boxes.shape[0], generator=g).float() / boxes.shape[0] * 0.9 + 0.06   # all distinct
    perm = torch.randperm(boxes.shape[0], generator=g)
    return boxes[perm], scores[perm], labels[perm]


@pytest.mark.parametrize("nms_type", ["nms", "vote", "soft-vote"])
@pytest.mark.parametrize("max_det", [1000, 60])
def test_tta_merge_against_oracle(nms_type, max_det):
    import paa_b200
    from types import SimpleNamespace as NS
    from oracle import vote_oracle
    from paa_b200.bbox_aug_vote import merge_result_from_multi_scales
    cfg = NS(MODEL=NS(RETINANET=NS(NUM_CLASSES=81, INFERENCE_TH=0.05), ATSS=NS(NMS_TH=0.6, PRE_NMS_TOP_N=max_det)))
    lists, want = [], []
    for seed in (11, 12):
        b, s, l = _tta_detections(seed)
        bl = paa_b200.BoxList(b.cuda(), (640, 640))
        bl.add_field("scores", s.cuda())
        bl.add_field("labels", l.cuda())
        lists.append(bl)
        want.append(vote_oracle.merge_multi_scale(b.numpy(), s.numpy(), l.numpy(), 81, merge_type=nms_type,
                                                  vote_thresh=0.66, nms_thresh=0.6, max_detections=max_det,
                                                  score_thresh=0.05))
    got = merge_result_from_multi_scales(lists, cfg, nms_type=nms_type, vote_thresh=0.66)
    for r, (wb, ws, wl) in zip(got, want):
        assert len(r) == wb.shape[0]
        np.testing.assert_array_equal(r.get_field("labels").cpu().numpy(), wl)
        np.testing.assert_array_equal(r.get_field("scores").cpu().numpy(), ws)
        np.testing.assert_array_equal(r.bbox.cpu().numpy(), wb)          # float32 numpy arithmetic, bit for bit
