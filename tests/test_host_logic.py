"""CPU: the host mirrors' argument handling -- constructor validation of the four loss evaluators, head / anchor
validation (gather_levels), and the refusal to run on CPU tensors (there is no CPU path).  No kernel is launched."""
from types import SimpleNamespace as NS

import pytest
import torch

import paa_b200
from paa_b200 import loss as L
from paa_b200 import synthetic


def _atss_cfg(**kw):
    a = dict(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, TOPK=9, REG_LOSS_WEIGHT=2.0, POSITIVE_TYPE="ATSS", REGRESSION_TYPE="BOX")
    a.update(kw)
    return NS(MODEL=NS(ATSS=NS(**a)))


def test_paa_evaluator_rejects_what_the_reference_cannot_run():
    cfg = paa_b200.default_cfg(REG_LOSS_TYPE="smoothl1")          # dead code in the reference (loss.py:246-251)
    with pytest.raises(NotImplementedError):
        paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    cfg = paa_b200.default_cfg()
    cfg.MODEL.ATSS.REGRESSION_TYPE = "POINT"
    with pytest.raises(NotImplementedError):
        paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))


def test_atss_positive_types():
    for ptype, code in (("ATSS", 0), ("SSC", 1), ("IoU", 2)):
        cfg = _atss_cfg(POSITIVE_TYPE=ptype)
        ev = paa_b200.make_atss_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
        assert L._lib.ATSS_POSITIVE_TYPES[ev.positive_type] == code
        assert (ev.iou_threshold, ev.bg_iou_threshold) == (0.5, 0.4)      # defaults.py FG / BG when the cfg omits them
    with pytest.raises(NotImplementedError):                               # atss/loss.py:227-228
        cfg = _atss_cfg(POSITIVE_TYPE="TOPK")
        paa_b200.make_atss_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))


def test_retinanet_evaluator_mirrors_the_reference_constructor():
    cfg = NS(MODEL=NS(RETINANET=NS(LOSS_GAMMA=(2.0,), LOSS_ALPHA=(0.25,), FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4,
                                   BBOX_REG_BETA=0.11, BBOX_REG_WEIGHT=4.0)))
    ev = paa_b200.make_retinanet_loss_evaluator(cfg, NS(weights=(10.0, 10.0, 5.0, 5.0)))
    assert (ev.gamma, ev.alpha, ev.iou_threshold, ev.bg_iou_threshold) == (2.0, 0.25, 0.5, 0.4)
    assert (ev.bbox_reg_beta, ev.regress_norm, ev.box_code_weights) == (0.11, 4.0, (10.0, 10.0, 5.0, 5.0))
    assert ev.copied_fields == ["labels"] and ev.discard_cases == ["between_thresholds"]
    matcher = NS(high_threshold=0.5, low_threshold=0.4, allow_low_quality_matches=False)
    with pytest.raises(NotImplementedError):
        L.RetinaNetLossComputation(matcher, NS(weights=(1.0, 1.0, 1.0, 1.0)), L.generate_retinanet_labels,
                                   NS(gamma=2.0, alpha=0.25))
    matcher.allow_low_quality_matches = True
    with pytest.raises(NotImplementedError):
        L.RetinaNetLossComputation(matcher, NS(weights=(1.0, 1.0, 1.0, 1.0)), lambda m: m, NS(gamma=2.0, alpha=0.25))


def test_fcos_evaluator_config():
    f = dict(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FPN_STRIDES=[8, 16, 32, 64, 128], CENTER_SAMPLING_RADIUS=1.5,
             IOU_LOSS_TYPE="giou", NORM_REG_TARGETS=True)
    ev = paa_b200.make_fcos_loss_evaluator(NS(MODEL=NS(FCOS=NS(**f))))
    assert ev.center_sampling_radius == 1.5 and ev.norm_reg_targets and ev.iou_loss_type == "giou"
    f["IOU_LOSS_TYPE"] = "diou"
    with pytest.raises(NotImplementedError):                               # layers/iou_loss.py:43-44
        paa_b200.make_fcos_loss_evaluator(NS(MODEL=NS(FCOS=NS(**f))))


def test_there_is_no_cpu_path():
    b = synthetic.make_batch(seed=5, num_images=1, image_hw=(128, 160), gt_per_image=2)
    cls, reg, iou, targets, anchors = synthetic.to_device_inputs(b, device="cpu")
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ev(cls, reg, iou, targets, anchors)


def test_head_and_anchor_validation():
    b = synthetic.make_batch(seed=6, num_images=2, image_hw=(128, 160), gt_per_image=2)
    cls, reg, iou, targets, anchors = synthetic.to_device_inputs(b, device="cpu")
    with pytest.raises(RuntimeError, match="same levels"):
        L.gather_levels(cls, reg[:-1], iou, anchors)
    with pytest.raises(RuntimeError, match="anchors lists"):
        L.gather_levels(cls, reg, iou, anchors[:1])


class _RecordingOwner(object):
    """Stands in for an evaluator behind `_PAALossFunction`: fixed losses / gradients, records the rescale vector."""

    def __init__(self):
        self.scaled = None

    def _run(self, cls, reg, iou, targets, anchors, need_grad):
        grads = None
        if need_grad:
            grads = dict(cls=[torch.full_like(t, 1.0) for t in cls], reg=[torch.full_like(t, 2.0) for t in reg],
                         iou=None if iou is None else [torch.full_like(t, 3.0) for t in iou])
        return torch.tensor([1.0, 2.0, 3.0]), grads, dict(device=None)

    def _rescale(self, call, grad_losses):
        self.scaled = grad_losses.tolist()


def test_autograd_wiring_of_the_loss_function():
    """The three losses leave the autograd function as separate 0-dim outputs; backward hands the rescale kernel
    the upstream gradients of all three (zero for a loss the caller did not use) and returns one gradient per head."""
    owner = _RecordingOwner()
    heads = [torch.randn(2, 3, requires_grad=True) for _ in range(6)]
    out = L._PAALossFunction.apply(owner, None, None, 2, True, *heads)
    assert len(out) == 3 and all(t.dim() == 0 and t.requires_grad for t in out)
    assert [float(t.detach()) for t in out] == [1.0, 2.0, 3.0]
    grads = torch.autograd.grad(out[0] * 2 + out[1], heads)                 # the third loss is unused
    assert owner.scaled == [2.0, 1.0, 0.0]
    assert [float(g.flatten()[0]) for g in grads] == [1.0, 1.0, 2.0, 2.0, 3.0, 3.0]
    out = L._PAALossFunction.apply(owner, None, None, 2, True, *heads)
    (out[0] + out[1] + out[2]).backward()
    assert owner.scaled == [1.0, 1.0, 1.0] and all(h.grad is not None for h in heads)
    out = L._PAALossFunction.apply(owner, None, None, 2, False, *heads[:4])  # USE_IOU_PRED = False: two head lists
    out[1].backward()
    assert owner.scaled == [0.0, 1.0, 0.0]
    with torch.no_grad():
        out = L._PAALossFunction.apply(owner, None, None, 2, True, *heads)
    assert not any(t.requires_grad for t in out)
    # the outputs behave like op results, not like views: a caller may weight a loss in place
    out = L._PAALossFunction.apply(owner, None, None, 2, True, *heads)
    weighted = out[0]
    weighted *= 0.5
    (weighted + out[1]).backward()
    assert owner.scaled == [0.5, 1.0, 0.0] and float(weighted.detach()) == 0.5


def test_anchor_sharing_is_detected_by_identity_or_address():
    b = synthetic.make_batch(seed=7, num_images=3, image_hw=(128, 160), gt_per_image=2)
    _, _, _, _, anchors = synthetic.to_device_inputs(b, device="cpu")
    n_levels = len(anchors[0])
    assert L._anchors_shared(anchors, 3, n_levels)
    views = [[paa_b200.BoxList(a.bbox[:], a.size) for a in per_image] for per_image in anchors]   # same storage
    assert L._anchors_shared(views, 3, n_levels)
    views[2][1] = paa_b200.BoxList(anchors[2][1].bbox.clone(), anchors[2][1].size)               # own copy
    assert not L._anchors_shared(views, 3, n_levels)


def test_fcos_points_cache_is_validated_and_bounded():
    """a cache keyed by address returned stale points when the allocator reused an address; entries are
    now tied to the tensor object and its version, and the cache is bounded."""
    import torch
    from paa_b200.loss import points_of, _POINTS_CACHE_ENTRIES
    cache = {}
    a = torch.rand(10, 2)
    p1 = points_of(cache, [a])[0]
    assert points_of(cache, [a])[0] is p1                 # same tensor, same version: reused
    a.add_(1.0)                                           # in-place update bumps the version
    p2 = points_of(cache, [a])[0]
    assert p2 is not p1 and torch.equal(p2[:, :2], a) and torch.equal(p2[:, 2:], a)
    seen = set()
    for _ in range(3 * _POINTS_CACHE_ENTRIES):            # a new grid every forward, like fcos.py:185-209
        loc = torch.rand(3, 2)
        q = points_of(cache, [loc])[0]
        assert torch.equal(q[:, :2], loc)
        seen.add(id(loc))
    assert len(cache) <= _POINTS_CACHE_ENTRIES


def test_head_layouts_consumed_in_place_and_odd_ones_copied_with_a_warning():
    """a channels_last head used to be `.contiguous()`-copied without a word.  Now a call whose logits
    are channels-last runs in that layout (no copy); only a tensor that is dense in neither the call's layout is copied,
    and the user is told once."""
    import warnings
    import torch
    from paa_b200 import _lib, loss as paa_loss

    class FakeCuda(torch.Tensor):
        is_cuda = True

    nhwc = [torch.zeros(2, 8, 4, 6).to(memory_format=torch.channels_last).as_subclass(FakeCuda),
            torch.zeros(2, 8, 1, 1).as_subclass(FakeCuda)]          # a 1x1 map is dense both ways
    nchw = [torch.zeros(2, 8, 4, 6).as_subclass(FakeCuda), torch.zeros(2, 8, 1, 1).as_subclass(FakeCuda)]
    assert paa_loss.call_layout(nhwc) == _lib.LAYOUT_NHWC
    assert paa_loss.call_layout(nchw) == _lib.LAYOUT_NCHW
    assert paa_loss.call_layout([nchw[0], nhwc[0]]) == _lib.LAYOUT_NCHW       # mixed: the reference's layout
    assert paa_loss.call_layout([nchw[1]]) == _lib.LAYOUT_NCHW
    paa_loss._warned_layout.clear()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for t in nhwc:
            assert paa_loss._head(t, "box_cls", _lib.LAYOUT_NHWC) is t         # consumed in place
        for t in nchw:
            assert paa_loss._head(t, "box_cls", _lib.LAYOUT_NCHW) is t
        assert not w
        out = paa_loss._head(nhwc[0], "box_cls", _lib.LAYOUT_NCHW)
        paa_loss._head(nhwc[0], "box_cls", _lib.LAYOUT_NCHW)                   # warned once per head name
        back = paa_loss._head(nchw[0], "box_regression", _lib.LAYOUT_NHWC)
        sliced = paa_loss._head(torch.zeros(2, 16, 4, 6)[:, ::2].as_subclass(FakeCuda), "iou_pred", _lib.LAYOUT_NCHW)
    assert out.is_contiguous() and sliced.is_contiguous()
    assert back.is_contiguous(memory_format=torch.channels_last)
    assert len([x for x in w if "not dense in the call's layout" in str(x.message)]) == 3
    c = nchw[0]
    assert paa_loss._head(c, "box_cls") is c               # the usual case costs nothing
