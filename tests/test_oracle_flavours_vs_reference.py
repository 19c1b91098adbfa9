"""CPU, only where the reference tree is mounted: the oracle's ATSS / RetinaNet / FCOS post-processing flavours
(oracle/post_oracle.py) against the reference's own ATSSPostProcessor / RetinaNetPostProcessor /
FCOSPostProcessor run with the import shims of oracle/ref_shim.py."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import post_oracle, ref_shim
from paa_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def _cl(ts):
    """channels-last copies: inference code of the reference .view()s a permuted tensor (SURVEY.md 8c shim 4)."""
    return [t.contiguous(memory_format=torch.channels_last) for t in ts]


def _same(ref_boxlists, want, n):
    for i in range(n):
        gb, gs, gl = post_oracle.canonical_rows(ref_boxlists[i].bbox, ref_boxlists[i].get_field("scores"),
                                                ref_boxlists[i].get_field("labels"))
        wb, ws, wl = post_oracle.canonical_rows(want[i].boxes, want[i].scores, want[i].labels)
        assert np.array_equal(gl, wl)
        np.testing.assert_allclose(gs, ws, rtol=0, atol=0)
        np.testing.assert_allclose(gb, wb, rtol=0, atol=0)


def test_atss_flavour():
    ref = ref_shim.load_reference()
    from paa_core.modeling.rpn.atss import inference as ainf
    b = synthetic.make_inference_batch(seed=41, num_images=2, image_hw=(320, 416), candidates_per_level=600)
    cfg = ref_shim.make_cfg()
    pp = ainf.ATSSPostProcessor(0.05, 200, 0.6, 100, 0, 81, ref.BoxCoder(cfg))
    anchors = [[ref.BoxList(a, b.image_sizes[i]) for a in b.anchors] for i in range(b.num_images)]
    got = pp(_cl(b.box_cls), _cl(b.box_regression), _cl(b.iou_pred), anchors)
    prm = post_oracle.default_params(pre_nms_top_n=200, score_voting=False, flavour="atss")
    want = post_oracle.postprocess(b.box_cls, b.box_regression, b.iou_pred, b.anchors, b.image_sizes, prm)
    _same(got, want, b.num_images)


def test_retinanet_flavour():
    ref = ref_shim.load_reference()
    from paa_core.modeling.rpn.retinanet import inference as rinf
    from paa_core.modeling.box_coder import BoxCoder as RpnCoder
    b = synthetic.make_retinanet_batch(seed=43, num_images=2, image_hw=(256, 320))
    n = b.box_cls[0].shape[0]
    pp = rinf.RetinaNetPostProcessor(0.05, 300, 0.4, 100, 0, 81, RpnCoder(weights=(10.0, 10.0, 5.0, 5.0)))
    anchors = [[ref.BoxList(a, b.image_sizes[i]) for a in b.anchors] for i in range(n)]
    got = pp(anchors, b.box_cls, b.box_regression)
    prm = post_oracle.default_params(pre_nms_top_n=300, nms_thresh=0.4, score_voting=False, flavour="retinanet")
    want = post_oracle.postprocess(b.box_cls, b.box_regression, None, b.anchors, b.image_sizes, prm)
    _same(got, want, n)


def test_fcos_flavour():
    ref_shim.load_reference()
    from paa_core.modeling.rpn.fcos import inference as finf
    b = synthetic.make_inference_batch(seed=44, num_images=2, image_hw=(256, 320), candidates_per_level=500)
    locs = synthetic.fcos_locations(b.grids)
    reg = [torch.exp(t * 0.5) * 8.0 * s for t, s in zip(b.box_regression, synthetic.STRIDES)]
    pp = finf.FCOSPostProcessor(0.05, 200, 0.6, 100, 0, 81)
    sizes = [(h, w) for (w, h) in b.image_sizes]
    got = pp(locs, _cl(b.box_cls), _cl(reg), _cl(b.iou_pred), sizes)
    prm = post_oracle.default_params(pre_nms_top_n=200, score_voting=False, flavour="fcos")
    want = post_oracle.postprocess(b.box_cls, reg, b.iou_pred, [torch.cat([l, l], 1) for l in locs], b.image_sizes, prm)
    _same(got, want, b.num_images)
