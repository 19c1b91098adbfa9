"""CPU: the two-line integration of INTEGRATION.md, applied to the reference's own module.

`paa_core/modeling/rpn/paa/paa.py` (the unmodified reference, imported from /root/reference with the shims of
oracle/ref_shim.py) is given this package's two factories in place of its own (`paa.py:6-7`).  `PAAModule` must then
build with them from the reference's cfg and BoxCoder, and its `_forward_train` / `_forward_test` (`paa.py:137-152`)
must reach this package's evaluators with the objects the reference produces -- its head outputs, its `BoxList`
targets, the anchors of its `make_anchor_generator_paa`, its locations.  Without a GPU the calls stop at the first
thing the evaluators do after accepting those arguments: refuse CPU tensors.  (The numerical side of the seam is what
the `-m gpu` parity tests cover; the reference tree does not exist on the GPU box.)"""
import importlib

import pytest
import torch

import paa_b200
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture()
def patched_module(monkeypatch):
    ref_shim.load_reference()
    paa_mod = importlib.import_module("paa_core.modeling.rpn.paa.paa")
    monkeypatch.setattr(paa_mod, "make_paa_loss_evaluator", paa_b200.make_paa_loss_evaluator)
    monkeypatch.setattr(paa_mod, "make_paa_postprocessor", paa_b200.make_paa_postprocessor)
    cfg = ref_shim.make_cfg(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25)       # floats, as in the reference's yaml configs
    cfg.MODEL.PAA.NUM_CONVS = 1
    cfg.MODEL.PAA.USE_DCN_IN_TOWER = False
    return paa_mod, cfg


def _inputs(cfg):
    ref = ref_shim.load_reference()
    from paa_core.structures.image_list import ImageList
    n, h, w = 2, 128, 160
    images = ImageList(torch.zeros(n, 3, h, w), [(h, w - 8), (h - 5, w)])
    features = [torch.randn(n, 32, -(-h // s), -(-w // s)) for s in cfg.MODEL.PAA.ANCHOR_STRIDES]
    targets = []
    for ih, iw in images.image_sizes:
        t = ref.BoxList(torch.tensor([[10., 20., 90., 100.], [30., 5., 70., 60.]]), (iw, ih), mode="xyxy")
        t.add_field("labels", torch.tensor([3, 17]))
        targets.append(t)
    return images, features, targets


def test_reference_module_builds_with_this_packages_factories(patched_module):
    paa_mod, cfg = patched_module
    module = paa_mod.PAAModule(cfg, 32)
    assert isinstance(module.loss_evaluator, paa_b200.PAALossComputation)
    assert isinstance(module.box_selector_test, paa_b200.PAAPostProcessor)
    ev, post = module.loss_evaluator, module.box_selector_test
    assert (ev.gamma, ev.alpha, ev.topk, ev.iou_threshold) == (2.0, 0.25, 9, 0.1)
    assert (ev.reg_loss_weight, ev.iou_loss_weight) == (1.3, 0.5)
    assert (post.pre_nms_thresh, post.pre_nms_top_n, post.nms_thresh, post.fpn_post_nms_top_n) == (0.05, 1000, 0.6, 100)
    assert post.score_voting is True and post.num_classes == 81


def test_reference_forward_reaches_this_packages_evaluators(patched_module):
    paa_mod, cfg = patched_module
    module = paa_mod.PAAModule(cfg, 32)
    images, features, targets = _inputs(cfg)
    module.train()
    with pytest.raises(RuntimeError, match="paa_b200 has no CPU path: box_cls"):
        module(images, features, targets)                   # paa.py:123-135 -> _forward_train -> loss_evaluator
    module.eval()
    with pytest.raises(RuntimeError, match="paa_b200 has no CPU path: box_cls"):
        module(images, features)                            # -> _forward_test -> box_selector_test


def test_reference_objects_pass_the_argument_checks(patched_module):
    """The reference's anchors (one BoxList per image and level around shared tensors) and targets go through the
    evaluators' validation as they are: shared anchor storage is recognised, a target whose image size differs from
    its anchors' is refused the way boxlist_iou refuses it (boxlist_ops.py:95-97)."""
    from paa_b200 import loss as L
    paa_mod, cfg = patched_module
    module = paa_mod.PAAModule(cfg, 32)
    images, features, targets = _inputs(cfg)
    anchors = module.anchor_generator(images, features)
    assert len(anchors) == 2 and len(anchors[0]) == 5
    assert L._anchors_shared(anchors, 2, 5)
    assert [tuple(t.size) for t in targets] == [tuple(a[0].size) for a in anchors]
