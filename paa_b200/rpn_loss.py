"""RPNLossComputation -- the reference's plain RPN training loss (paa_core/modeling/rpn/loss.py:21-157) on the kernels
of libpaa_b200.so: the sibling of the PAA / RetinaNet evaluators that samples its anchors.

Same constructor ``(proposal_matcher, fg_bg_sampler, box_coder, generate_labels_func)``, same call
``(anchors, objectness, box_regression, targets)`` -> ``(objectness_loss, box_loss)``, same factory
``make_rpn_loss_evaluator(cfg, box_coder)``.  What runs where:

* IoU matching + Matcher(FG, BG, allow_low_quality_matches=True) + labels 1 / 0 / -1: `paa_retinanet_assign`
  (the RetinaNet path's kernels with unit GT labels; rpn/loss.py:41-81).  The "visibility" field of the anchors
  (anchor_generator.py:97-110) turns anchors outside the image into -1 afterwards, in the reference's order.
* The balanced sampler (balanced_positive_negative_sampler.py:35-68) is `torch.randperm` on the anchors' device, the
  reference's own choice of generator -- a random draw has no parity beyond its distribution; the parity tests inject
  the oracle's sample through `sample_override`.
* Both losses and their gradients over the sampled anchors: `paa_rpn_loss` (csrc/rpn.cu), regression targets
  encoded on the fly (box_coder.py:22-50).

There is no CPU path.
"""
import ctypes as C

import torch

from paa_b200 import _lib
from paa_b200.loss import PAALossComputation, gather_levels


class BalancedPositiveNegativeSampler(object):
    """balanced_positive_negative_sampler.py:5-68: per image at most ``batch_size_per_image * positive_fraction``
    positives (label >= 1) and the rest negatives (label == 0), chosen by ``torch.randperm``; returns two lists of
    uint8 masks."""

    def __init__(self, batch_size_per_image, positive_fraction):
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction

    def __call__(self, matched_idxs):
        pos_idx, neg_idx = [], []
        for labels in matched_idxs:
            positive = torch.nonzero(labels >= 1).squeeze(1)
            negative = torch.nonzero(labels == 0).squeeze(1)
            num_pos = min(positive.numel(), int(self.batch_size_per_image * self.positive_fraction))
            num_neg = min(negative.numel(), self.batch_size_per_image - num_pos)
            perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
            perm2 = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
            pm = torch.zeros_like(labels, dtype=torch.uint8)
            nm = torch.zeros_like(labels, dtype=torch.uint8)
            pm[positive[perm1]] = 1
            nm[negative[perm2]] = 1
            pos_idx.append(pm)
            neg_idx.append(nm)
        return pos_idx, neg_idx


def generate_rpn_labels(matched_targets):
    """rpn/loss.py:140-143."""
    return matched_targets.get_field("matched_idxs") >= 0


class _MatcherPass(PAALossComputation):
    """The assignment half of the RetinaNet flavour, run on its own: Matcher results and labels per anchor."""

    def __init__(self, high, low):
        self.iou_threshold, self.bg_iou_threshold = float(high), float(low)
        self.gamma, self.alpha = 2.0, 0.25                       # unused by the assignment
        self.topk, self.iou_loss_weight, self.reg_loss_weight = 1, 0.0, 1.0
        self.box_code_weights, self.bbox_reg_beta, self.regress_norm = (1.0, 1.0, 1.0, 1.0), 1.0, 1.0
        self._init_runtime()
        self._flavour = _lib.LOSS_RETINANET

    def match(self, lv, targets, anchors):
        """-> (matched [N, A] int32 in {-2, -1, 0..G-1}, labels [N, A] int32 in {1, 0, -1}, gt_boxes, offsets)."""
        N, L, A = lv["N"], lv["L"], lv["A"]
        device = lv["cls"][0].device
        offsets, boxes, _, sum_g = self._collect_rpn_targets(targets, anchors, N)
        gt_boxes = torch.cat(boxes, dim=0).to(device=device, dtype=torch.float32)
        gt_labels = torch.ones(sum_g, dtype=torch.int64, device=device)
        args = self._make_args(lv, False, 1, None)
        args.gt_boxes, args.gt_labels = gt_boxes.data_ptr(), gt_labels.data_ptr()
        args.gt_offsets[:N + 1] = offsets
        nbytes = self._lib.paa_loss_workspace_bytes(N, A, sum_g, L, self.topk)
        ws = self._workspace_for(device, nbytes)
        base = (ws.data_ptr() + 255) // 256 * 256
        args.workspace, args.workspace_bytes = base, ws.numel() - (base - ws.data_ptr())
        normalisers = torch.empty(2, dtype=torch.float64, device=device)
        losses = torch.empty(3, dtype=torch.float32, device=device)
        args.normalisers, args.losses = normalisers.data_ptr(), losses.data_ptr()
        matched = torch.empty((N, A), dtype=torch.int32, device=device)
        labels = torch.empty((N, A), dtype=torch.int32, device=device)
        args.dbg_matched_idx, args.dbg_paa_labels = matched.data_ptr(), labels.data_ptr()
        with _lib.device_guard(device):
            _lib.check(self._lib.paa_retinanet_assign(C.byref(args), _lib.stream_handle(device)),
                       "paa_retinanet_assign")
        return matched, labels, gt_boxes, offsets

    @staticmethod
    def _collect_rpn_targets(targets, anchors, N):
        if len(targets) != N:
            raise RuntimeError("targets lists %d images, heads have batch %d" % (len(targets), N))
        if N > _lib.MAX_IMAGES:
            raise RuntimeError("at most %d images per call" % _lib.MAX_IMAGES)
        offsets, boxes, sum_g = [0], [], 0
        for i, t in enumerate(targets):
            if tuple(t.size) != tuple(anchors[i][0].size):            # boxlist_ops.py:95-97
                raise RuntimeError("boxlists should have same image size, got {}, {}".format(t, anchors[i][0]))
            if t.bbox.shape[0] == 0:                                  # matcher.py:53-58
                raise ValueError("No ground-truth boxes available for one of the images during training")
            sum_g += t.bbox.shape[0]
            offsets.append(sum_g)
            boxes.append(t.bbox)
        return offsets, boxes, None, sum_g


class _RpnLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, targets, anchors, n_levels, *heads):
        need_grad = any(ctx.needs_input_grad[4:])
        losses, grads = owner._run(list(heads[:n_levels]), list(heads[n_levels:]), targets, anchors, need_grad)
        ctx.grads, ctx.n_levels = grads, n_levels
        ctx.set_materialize_grads(False)
        return losses[0].detach(), losses[1].detach()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_obj, g_box):
        n_in = 4 + 2 * ctx.n_levels
        if ctx.grads is None or (g_obj is None and g_box is None):
            return (None,) * n_in
        out = [None if g_obj is None else t * g_obj for t in ctx.grads["cls"]] + \
              [None if g_box is None else t * g_box for t in ctx.grads["reg"]]
        return (None, None, None, None) + tuple(out)


class RPNLossComputation(object):
    """Drop-in for paa_core.modeling.rpn.loss.RPNLossComputation (rpn/loss.py:21-137)."""

    def __init__(self, proposal_matcher, fg_bg_sampler, box_coder, generate_labels_func=generate_rpn_labels):
        if not getattr(proposal_matcher, "allow_low_quality_matches", True):
            raise NotImplementedError("only Matcher(..., allow_low_quality_matches=True) is supported")
        if getattr(generate_labels_func, "__name__", "") != "generate_rpn_labels":
            raise NotImplementedError("only generate_rpn_labels is supported (RetinaNet has its own evaluator)")
        self.proposal_matcher = proposal_matcher
        self.fg_bg_sampler = fg_bg_sampler
        self.box_coder = box_coder
        self.copied_fields = []
        self.generate_labels_func = generate_labels_func
        self.discard_cases = ["not_visibility", "between_thresholds"]
        self.box_code_weights = tuple(float(w) for w in getattr(box_coder, "weights", (1.0, 1.0, 1.0, 1.0)))
        self.smooth_l1_beta = 1.0 / 9                                  # rpn/loss.py:126
        self._matcher = _MatcherPass(proposal_matcher.high_threshold, proposal_matcher.low_threshold)
        self._lib = self._matcher._lib
        self.sample_override = None      # (sampled_pos, sampled_neg) global indices: replaces the random draw (tests)
        self.last_debug = None

    def _labels(self, lv, targets, anchors):
        """rpn/loss.py:56-95 without the regression targets (the loss kernel encodes them where it needs them)."""
        matched, labels, gt_boxes, offsets = self._matcher.match(lv, targets, anchors)
        labels = labels.to(torch.float32)
        if "not_visibility" in self.discard_cases:
            for i, per_image in enumerate(anchors):
                if all(a.has_field("visibility") for a in per_image):
                    vis = torch.cat([a.get_field("visibility").to(torch.bool) for a in per_image])
                    labels[i][~vis] = -1                               # rpn/loss.py:75-76
        return matched, labels, gt_boxes, offsets

    def _run(self, objectness, box_regression, targets, anchors, need_grad):
        lv = gather_levels(objectness, box_regression, None, anchors)
        if lv["C"] != 1:
            raise RuntimeError("objectness must have one channel per anchor, got %d" % lv["C"])
        N, L, A = lv["N"], lv["L"], lv["A"]
        device = lv["cls"][0].device
        matched, labels, gt_boxes, offsets = self._labels(lv, targets, anchors)
        if self.sample_override is not None:
            pos, neg = (t.to(device=device, dtype=torch.int64) for t in self.sample_override)
        else:
            pos_masks, neg_masks = self.fg_bg_sampler([labels[i] for i in range(N)])   # rpn/loss.py:110
            pos = torch.nonzero(torch.cat(pos_masks, dim=0)).squeeze(1)
            neg = torch.nonzero(torch.cat(neg_masks, dim=0)).squeeze(1)
        sampled = torch.cat([pos, neg], dim=0).contiguous()
        grads = PAALossComputation._alloc_grads(lv, False) if need_grad else None
        args = _lib.PaaRpnArgs()
        args.num_images, args.num_levels, args.anchors_per_loc = N, L, lv["apl"]
        args.head_layout, args.anchor_image_stride = lv["layout"], lv["anchor_stride"]
        for l in range(L):
            s = args.levels[l]
            s.box_cls, s.box_regression = lv["cls"][l].data_ptr(), lv["reg"][l].data_ptr()
            s.anchors, s.hw, s.grid_w = lv["anchor_ptrs"][l], lv["hw"][l], lv["grid_w"][l]
            if grads is not None:
                s.grad_box_cls, s.grad_box_regression = grads["cls"][l].data_ptr(), grads["reg"][l].data_ptr()
        args.gt_boxes = gt_boxes.data_ptr()
        args.gt_offsets[:N + 1] = offsets
        args.matched_idx, args.sampled = matched.data_ptr(), sampled.data_ptr()
        args.n_pos, args.n_neg = int(pos.numel()), int(neg.numel())
        for k in range(4):
            args.box_code_weights[k] = self.box_code_weights[k]
        args.smooth_l1_beta = self.smooth_l1_beta
        losses = torch.empty(2, dtype=torch.float32, device=device)
        args.losses, args.grad_losses = losses.data_ptr(), None
        with _lib.device_guard(device):
            _lib.check(self._lib.paa_rpn_loss(C.byref(args), _lib.stream_handle(device)), "paa_rpn_loss")
        self.last_debug = dict(matched_idx=matched, labels=labels, sampled_pos=pos, sampled_neg=neg)
        return losses, grads

    def __call__(self, anchors, objectness, box_regression, targets):
        heads = list(objectness) + list(box_regression)
        return _RpnLossFunction.apply(self, targets, anchors, len(objectness), *heads)


def make_rpn_loss_evaluator(cfg, box_coder):
    """rpn/loss.py:146-157."""
    from types import SimpleNamespace
    rpn = cfg.MODEL.RPN
    matcher = SimpleNamespace(high_threshold=rpn.FG_IOU_THRESHOLD, low_threshold=rpn.BG_IOU_THRESHOLD,
                              allow_low_quality_matches=True)
    sampler = BalancedPositiveNegativeSampler(rpn.BATCH_SIZE_PER_IMAGE, rpn.POSITIVE_FRACTION)
    return RPNLossComputation(matcher, sampler, box_coder, generate_rpn_labels)
