"""PAALossComputation -- the reference's training-side entry point for the PAA hot path, backed by
libpaa_b200.so.

Mirror of paa_core/modeling/rpn/paa/loss.py: same constructor ``(cfg, box_coder)``, same call
signature ``(box_cls, box_regression, iou_pred, targets, anchors, locations)``, same return value
(a list of 0-dim float32 losses ``[loss_cls, loss_reg(, loss_iou_pred)]`` that back-propagate into
the three prediction lists), same exceptions (``ValueError`` for an image without ground truth,
matcher.py:53-58; ``RuntimeError`` when a target's image size differs from its anchors',
boxlist_ops.py:95-97).  The Python here only extracts device pointers and sizes; all arithmetic is
in the CUDA kernels reached through the C ABI of include/paa_b200.h.  There is no CPU path.

Multi-GPU: ranks own disjoint images.  Like the reference (loss.py:18-28) the world size is the
``WORLD_SIZE`` environment variable; with more than one rank the two normalisers are summed with a
single 2-element all-reduce between ``paa_assign`` and ``paa_loss`` (the reference issues two
all-reduces followed by ``.item()`` host syncs, loss.py:321,338).
"""
import ctypes as C
import os

import torch

from paa_b200 import _lib
from paa_b200.box_coder import coder_regression_type
from paa_b200.config import scalar


def get_num_gpus():
    return int(os.environ["WORLD_SIZE"]) if "WORLD_SIZE" in os.environ else 1   # loss.py:18-19


def reduce_normalisers(normalisers):
    """The path's only exchange step: sums the 2-element ``{num_pos, sum IoU}`` tensor over ranks in
    place (loss.py:22-28 applied once to both normalisers, loss.py:321,338).  Stream-ordered; no host
    synchronisation.  A no-op for a single process."""
    if get_num_gpus() <= 1:
        return normalisers
    import torch.distributed as dist
    dist.all_reduce(normalisers, op=dist.ReduceOp.SUM)
    return normalisers


class PeerNormExchange(object):
    """Per-device state of the peer-memory exchange of the two loss normalisers (include/paa_b200.h,
    ``PaaLossArgs.peer_norm``): one small symmetric-memory buffer per rank, mapped into every rank of the
    default process group over NVLink.  With it the step's only cross-rank traffic -- 16 bytes per rank --
    is written by the last block of the assignment kernel straight into the peers' memory and read by the
    first kernel of the loss pass: no collective launch on the stream (an NCCL all-reduce of this size costs
    ~14 us of a ~165 us step on B200).  Falls back to ``dist.all_reduce`` where symmetric memory is not
    available (``PAA_NORM_EXCHANGE=nccl`` forces the fallback)."""
    _by_device = {}

    def __init__(self, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError("more than %d ranks" % _lib.MAX_PEERS)
        self.buffer = symm_mem.empty(_lib.PEER_BUFFER_DOUBLES, dtype=torch.float64, device=device)
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, group=dist.group.WORLD.group_name)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        torch.cuda.synchronize(device)      # zeroed before the collective in get() lets anyone write into it

    @classmethod
    def get(cls, device):
        """The exchange for `device`, or None when it cannot be used (then the caller all-reduces).  The
        decision is collective: if the set-up fails on any rank, every rank falls back."""
        key = (device.type, device.index)
        if key not in cls._by_device:
            state = None
            import torch.distributed as dist
            usable = (os.environ.get("PAA_NORM_EXCHANGE", "peer") != "nccl" and dist.is_available()
                      and dist.is_initialized() and dist.get_world_size() == get_num_gpus()
                      and dist.get_world_size() <= _lib.MAX_PEERS)
            if usable:
                why = ""
                try:
                    state = cls(device)
                except Exception as e:  # noqa: BLE001 - symmetric memory unavailable on this rank
                    why = str(e)
                    state = None
                ok = torch.tensor([1 if state is not None else 0], dtype=torch.int32, device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 0:
                    if state is None:
                        import warnings
                        warnings.warn("paa_b200: peer-memory normaliser exchange unavailable (%s); "
                                      "using all_reduce" % (why,))
                    state = None
            cls._by_device[key] = state
        return cls._by_device[key]


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("paa_b200 has no CPU path: %s is on %s" % (name, t.device))


def _head(t, name):
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _anchors_shared(anchors, N, L):
    """True when every image lists the same anchor storage (the reference's generator wraps one tensor per level
    into a BoxList per image, anchor_generator.py:114-124)."""
    first = anchors[0]
    for i in range(1, N):
        per_image = anchors[i]
        if per_image is first:
            continue
        for l in range(L):
            b, b0 = per_image[l].bbox, first[l].bbox
            if b is not b0 and b.data_ptr() != b0.data_ptr():
                return False
    return True


def gather_levels(box_cls, box_regression, iou_pred, anchors):
    """Validates the head lists / anchor lists and returns contiguous tensors plus layout facts.
    anchors: list[N] of list[L] BoxList (anchor_generator.py:112-125).  This runs on every call of the
    evaluators, ahead of the kernel launches: plain loops, no per-element tensor calls beyond the checks."""
    L = len(box_cls)
    if L == 0 or len(box_regression) != L or (iou_pred is not None and len(iou_pred) != L):
        raise RuntimeError("box_cls / box_regression / iou_pred must list the same levels")
    if L > _lib.MAX_LEVELS:
        raise RuntimeError("at most %d levels are supported" % _lib.MAX_LEVELS)
    N = box_cls[0].shape[0]
    if len(anchors) != N:
        raise RuntimeError("anchors lists %d images, heads have batch %d" % (len(anchors), N))
    apl = box_regression[0].shape[1] // 4
    num_classes = box_cls[0].shape[1] // apl
    cls = [_head(t, "box_cls") for t in box_cls]
    reg = [_head(t, "box_regression") for t in box_regression]
    iou = None if iou_pred is None else [_head(t, "iou_pred") for t in iou_pred]
    hw, grid_w = [], []
    first = anchors[0]
    for l in range(L):
        n, ch, h, w = cls[l].shape
        if n != N or ch != apl * num_classes or reg[l].shape != (N, apl * 4, h, w) or \
                (iou is not None and iou[l].shape != (N, apl, h, w)):
            raise RuntimeError("level %d: inconsistent head shapes" % l)
        if first[l].bbox.shape[0] != h * w * apl:
            raise RuntimeError("level %d: %d anchors for a %dx%d map with %d per location"
                               % (l, first[l].bbox.shape[0], h, w, apl))
        hw.append(h * w)
        grid_w.append(w)
    A = sum(hw) * apl
    # every image normally shares the anchor tensors of the batch (anchor_generator.py:114-124)
    if _anchors_shared(anchors, N, L):
        anc = [first[l].bbox for l in range(L)]
        anc = [a if (a.dtype == torch.float32 and a.is_contiguous()) else a.contiguous().float() for a in anc]
        for a in anc:
            _require_cuda(a, "anchors")
        level_ptrs = [a.data_ptr() for a in anc]
        stride = 0
        keep = anc
    else:
        stacked = torch.stack([torch.cat([anchors[i][l].bbox for l in range(L)], dim=0)
                               for i in range(N)], dim=0).float().contiguous()      # [N, A, 4]
        _require_cuda(stacked, "anchors")
        offs, o = [], 0
        for l in range(L):
            offs.append(o)
            o += hw[l] * apl
        level_ptrs = [stacked.data_ptr() + 16 * off for off in offs]
        stride = A * 4
        keep = [stacked]
    return dict(L=L, N=N, apl=apl, C=num_classes, hw=hw, grid_w=grid_w, A=A, cls=cls, reg=reg, iou=iou,
                anchor_ptrs=level_ptrs, anchor_stride=stride, keep=keep)


class _PAALossFunction(torch.autograd.Function):
    """(loss_0, loss_1, loss_2) = f(heads); the gradients are produced by the same kernel pass as the losses and
    handed out in backward (rescaled on the device if the upstream gradients are not ones).  The three losses
    leave as separate 0-dim outputs (aliases of the kernel's 3-element result made here, outside autograd's
    recording): indexing a 3-vector output instead costs three SelectBackward nodes, i.e. ~8 tiny launches per
    step on the host path."""

    @staticmethod
    def forward(ctx, owner, targets, anchors, n_levels, has_iou, *heads):
        box_cls = list(heads[:n_levels])
        box_reg = list(heads[n_levels:2 * n_levels])
        iou_pred = list(heads[2 * n_levels:3 * n_levels]) if has_iou else None
        need_grad = any(ctx.needs_input_grad[5:])
        losses, grads, call = owner._run(box_cls, box_reg, iou_pred, targets, anchors, need_grad)
        ctx.owner = owner
        ctx.call = call
        ctx.grads = grads
        ctx.n_levels = n_levels
        ctx.has_iou = has_iou
        ctx.set_materialize_grads(False)
        # detached aliases, not autograd views: the caller may scale a loss in place like any op result
        return losses[0].detach(), losses[1].detach(), losses[2].detach()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grad_losses):
        grads = ctx.grads
        n_in = 5 + ctx.n_levels * (3 if ctx.has_iou else 2)
        if grads is None or all(g is None for g in grad_losses):
            return (None,) * n_in
        if any(g is None for g in grad_losses):          # a loss the caller did not use contributes nothing
            zero = next(g for g in grad_losses if g is not None).new_zeros(())
            grad_losses = [zero if g is None else g for g in grad_losses]
        g = torch.stack(grad_losses)
        ctx.owner._rescale(ctx.call, g if g.dtype == torch.float32 else g.float())
        out = list(grads["cls"]) + list(grads["reg"]) + (list(grads["iou"]) if ctx.has_iou else [])
        return (None, None, None, None, None) + tuple(out)


class PAALossComputation(object):
    """Drop-in for paa_core.modeling.rpn.paa.loss.PAALossComputation (loss.py:31-359)."""

    def __init__(self, cfg, box_coder):
        self.cfg = cfg
        paa = cfg.MODEL.PAA
        self.gamma = scalar(paa.LOSS_GAMMA)
        self.alpha = scalar(paa.LOSS_ALPHA)
        self.iou_threshold = float(paa.IOU_THRESHOLD)
        self.topk = int(paa.TOPK)
        self.box_coder = box_coder
        self.fpn_strides = [8, 16, 32, 64, 128]
        self.reg_loss_type = paa.REG_LOSS_TYPE
        self.iou_loss_weight = float(paa.IOU_LOSS_WEIGHT)
        self.reg_loss_weight = float(paa.REG_LOSS_WEIGHT)
        if "iou" not in self.reg_loss_type:
            # the reference's only other branch ('smoothl1') is dead code (loss.py:246-251 passes
            # arguments smooth_l1_loss does not take)
            raise NotImplementedError("REG_LOSS_TYPE %r: only 'iou' is supported" % (self.reg_loss_type,))
        if coder_regression_type(box_coder) != "BOX":
            raise NotImplementedError("only the 'BOX' BoxCoder regression type is supported")
        self._lib = _lib.load()          # raises if the CUDA library is missing
        self.debug = False               # True: keep per-stage parity outputs of the last call
        self.last_debug = None
        self.teacher_combined_loss = None   # [N, A] float32 cuda tensor: stage-wise parity protocol
        self._workspace = None
        self._ones = None
        self._flavour = _lib.LOSS_PAA       # which assignment / weighting the kernels apply (ATSS subclass)

    # -- plumbing -------------------------------------------------------------------------------
    def _workspace_for(self, device, nbytes):
        ws = self._workspace
        if ws is None or ws.device != device or ws.numel() < nbytes:
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    def _run(self, box_cls, box_reg, iou_pred, targets, anchors, need_grad):
        lv = gather_levels(box_cls, box_reg, iou_pred, anchors)
        N, L, A = lv["N"], lv["L"], lv["A"]
        device = lv["cls"][0].device
        if len(targets) != N:
            raise RuntimeError("targets lists %d images, heads have batch %d" % (len(targets), N))
        if N > _lib.MAX_IMAGES:
            raise RuntimeError("at most %d images per call" % _lib.MAX_IMAGES)
        offsets, boxes, labels, sum_g = [0], [], [], 0
        for i, t in enumerate(targets):
            assert t.mode == "xyxy"                                   # loss.py:97
            if tuple(t.size) != tuple(anchors[i][0].size):            # boxlist_ops.py:95-97
                raise RuntimeError("boxlists should have same image size, got {}, {}".format(t, anchors[i][0]))
            b = t.bbox
            sum_g += b.shape[0]
            offsets.append(sum_g)
            boxes.append(b)
            labels.append(t.get_field("labels"))
        gt_boxes = torch.cat(boxes, dim=0)
        if gt_boxes.dtype != torch.float32 or gt_boxes.device != device:
            gt_boxes = gt_boxes.to(device=device, dtype=torch.float32)
        gt_labels = torch.cat(labels, dim=0)
        if gt_labels.dtype != torch.int64 or gt_labels.device != device:
            gt_labels = gt_labels.to(device=device, dtype=torch.int64)
        has_iou = iou_pred is not None
        # the RetinaNet loss normalises by this rank's own counts (retinanet/loss.py:70,79): no exchange
        world = 1 if self._flavour == _lib.LOSS_RETINANET else get_num_gpus()

        args = _lib.PaaLossArgs()
        args.num_images, args.num_levels, args.num_classes = N, L, lv["C"]
        args.anchors_per_loc, args.topk = lv["apl"], self.topk
        args.use_iou_pred, args.world_size = int(has_iou), world
        args.loss_flavour = self._flavour
        args.gamma, args.alpha, args.iou_threshold = self.gamma, self.alpha, self.iou_threshold
        args.reg_loss_weight, args.iou_loss_weight = self.reg_loss_weight, self.iou_loss_weight
        if self._flavour == _lib.LOSS_RETINANET:
            args.bg_iou_threshold = self.bg_iou_threshold
            for k in range(4):
                args.box_code_weights[k] = self.box_code_weights[k]
            args.smooth_l1_beta, args.reg_norm_weight = self.bbox_reg_beta, self.regress_norm
        if self._flavour == _lib.LOSS_ATSS:
            args.atss_positive_type = _lib.ATSS_POSITIVE_TYPES[self.positive_type]
            args.bg_iou_threshold = self.bg_iou_threshold
        if self._flavour == _lib.LOSS_FCOS:
            for l, stride in enumerate(self.fpn_strides[:L]):
                args.fcos_strides[l] = float(stride)
            args.fcos_center_radius = self.center_sampling_radius
            args.fcos_iou_loss_type = _lib.IOU_LOSS_TYPES[self.iou_loss_type]
            args.fcos_norm_reg_targets = int(self.norm_reg_targets)
        args.anchor_image_stride = lv["anchor_stride"]
        grads = None
        if need_grad:
            grads = dict(cls=[torch.empty_like(t) for t in lv["cls"]],
                         reg=[torch.empty_like(t) for t in lv["reg"]],
                         iou=[torch.empty_like(t) for t in lv["iou"]] if has_iou else None)
        cls_l, reg_l, iou_l = lv["cls"], lv["reg"], lv["iou"]
        anchor_ptrs, hw_l, grid_w = lv["anchor_ptrs"], lv["hw"], lv["grid_w"]
        levels = args.levels
        for l in range(L):
            s = levels[l]
            s.box_cls, s.box_regression = cls_l[l].data_ptr(), reg_l[l].data_ptr()
            s.anchors, s.hw, s.grid_w = anchor_ptrs[l], hw_l[l], grid_w[l]
            if has_iou:
                s.iou_pred = iou_l[l].data_ptr()
            if grads is not None:
                s.grad_box_cls, s.grad_box_regression = grads["cls"][l].data_ptr(), grads["reg"][l].data_ptr()
                if has_iou:
                    s.grad_iou_pred = grads["iou"][l].data_ptr()
        args.gt_boxes, args.gt_labels = gt_boxes.data_ptr(), gt_labels.data_ptr()
        args.gt_offsets[:N + 1] = offsets
        nbytes = self._lib.paa_loss_workspace_bytes(N, A, sum_g, L, self.topk)
        ws = self._workspace_for(device, nbytes)
        base = (ws.data_ptr() + 255) // 256 * 256
        args.workspace, args.workspace_bytes = base, ws.numel() - (base - ws.data_ptr())
        normalisers = torch.empty(2, dtype=torch.float64, device=device)
        losses = torch.empty(3, dtype=torch.float32, device=device)
        args.normalisers, args.losses, args.grad_losses = normalisers.data_ptr(), losses.data_ptr(), None
        dbg = None
        if self.debug:
            cap = L * self.topk
            dbg = dict(matched_idx=torch.empty((N, A), dtype=torch.int32, device=device),
                       iou_labels=torch.empty((N, A), dtype=torch.int32, device=device),
                       combined_loss=torch.empty((N, A), dtype=torch.float32, device=device),
                       cand_idx=torch.full((sum_g, cap), -1, dtype=torch.int32, device=device),
                       cand_cnt=torch.zeros(sum_g, dtype=torch.int32, device=device),
                       num_pos=torch.zeros(sum_g, dtype=torch.int32, device=device),
                       gmm=torch.zeros((sum_g, 8), dtype=torch.float64, device=device),
                       paa_labels=torch.empty((N, A), dtype=torch.int32, device=device))
            args.dbg_matched_idx, args.dbg_iou_labels = dbg["matched_idx"].data_ptr(), dbg["iou_labels"].data_ptr()
            args.dbg_combined_loss, args.dbg_cand_idx = dbg["combined_loss"].data_ptr(), dbg["cand_idx"].data_ptr()
            args.dbg_cand_cnt, args.dbg_num_pos = dbg["cand_cnt"].data_ptr(), dbg["num_pos"].data_ptr()
            args.dbg_gmm, args.dbg_paa_labels = dbg["gmm"].data_ptr(), dbg["paa_labels"].data_ptr()
        teacher = self.teacher_combined_loss
        if teacher is not None:
            teacher = teacher.to(device=device, dtype=torch.float32).contiguous()
            assert teacher.shape == (N, A)
            args.teacher_combined_loss = teacher.data_ptr()
        stream = _lib.stream_handle(device)
        assign_name = {_lib.LOSS_PAA: "paa_assign", _lib.LOSS_ATSS: "paa_atss_assign",
                       _lib.LOSS_RETINANET: "paa_retinanet_assign", _lib.LOSS_FCOS: "paa_fcos_assign"}[self._flavour]
        assign = getattr(self._lib, assign_name)
        with _lib.device_guard(device):
            peer = PeerNormExchange.get(device) if world > 1 else None
            if peer is not None:
                args.rank = peer.rank
                for r, ptr in enumerate(peer.ptrs):
                    args.peer_norm[r] = ptr
            _lib.check(assign(C.byref(args), stream), assign_name)
            if world > 1 and peer is None:
                reduce_normalisers(normalisers)                         # loss.py:321,338 in one message
            _lib.check(self._lib.paa_loss(C.byref(args), stream), "paa_loss")
        if dbg is not None:
            dbg["normalisers"] = normalisers
            dbg["gt_offsets"] = offsets
            self.last_debug = dbg
        # keep every tensor whose pointer the kernels use alive until the stream work is enqueued
        call = dict(args=args, keep=(lv, gt_boxes, gt_labels, ws, normalisers, teacher), device=device)
        return losses, grads, call

    def _rescale(self, call, grad_losses):
        """Brings the gradients written by the forward pass (upstream gradients of one) to `grad_losses`.  A call
        whose gradients are taken again (``retain_graph=True``) is rescaled from the factors applied last time; a
        loss that was left out then (factor 0) cannot come back -- its gradients turn NaN rather than silently 0."""
        device = call["device"]
        old = call.get("applied")
        if old is None:
            if self._ones is None or self._ones.device != device:
                self._ones = torch.ones(3, dtype=torch.float32, device=device)
            old = self._ones
        stream = _lib.stream_handle(device)
        with _lib.device_guard(device):
            _lib.check(self._lib.paa_rescale_grads(C.byref(call["args"]), old.data_ptr(),
                                                   grad_losses.data_ptr(), stream), "paa_rescale_grads")
        call["applied"] = grad_losses

    def forward_backward(self, box_cls, box_regression, iou_pred, targets, anchors, grad_losses=None):
        """Fused training step without autograd bookkeeping: returns ``(losses[3], grads)`` where
        ``grads`` is ``dict(cls=[...], reg=[...], iou=[...] or None)`` holding d(sum_j g_j * loss_j)/d(head)
        for ``grad_losses = g`` (ones when omitted).  Same kernels as ``__call__`` + ``backward``; every
        call on it is a plain stream-ordered launch, so a whole step can be captured in a CUDA graph."""
        iou = list(iou_pred) if iou_pred is not None else None
        losses, grads, call = self._run(list(box_cls), list(box_regression), iou, targets, anchors, True)
        if grad_losses is not None:
            self._rescale(call, grad_losses.contiguous().float())
        return losses, grads

    # -- the reference's interface --------------------------------------------------------------
    def __call__(self, box_cls, box_regression, iou_pred, targets, anchors, locations=None):
        n_levels = len(box_cls)
        has_iou = iou_pred is not None
        heads = list(box_cls) + list(box_regression) + (list(iou_pred) if has_iou else [])
        losses = _PAALossFunction.apply(self, targets, anchors, n_levels, has_iou, *heads)
        return list(losses) if has_iou else [losses[0], losses[1]]


def make_paa_loss_evaluator(cfg, box_coder):
    return PAALossComputation(cfg, box_coder)      # loss.py:362-364


class ATSSLossComputation(PAALossComputation):
    """Drop-in for paa_core.modeling.rpn.atss.loss.ATSSLossComputation (atss/loss.py:27-279) with
    all three POSITIVE_TYPEs: 'ATSS' (the default rule), 'SSC' (FCOS's rule on anchor centres) and 'IoU' (Matcher
    labels, ignored anchors).  Anchors are assigned on the device (`paa_atss_assign`), the losses come from the
    same streaming pass as PAA's with the centerness targets as regression weights / BCE targets.  Returns
    ``(cls_loss, reg_loss * REG_LOSS_WEIGHT, centerness_loss)`` like the reference."""

    def __init__(self, cfg, box_coder):
        atss = cfg.MODEL.ATSS
        self.positive_type = getattr(atss, "POSITIVE_TYPE", "ATSS")
        if self.positive_type not in _lib.ATSS_POSITIVE_TYPES:
            raise NotImplementedError                                  # atss/loss.py:227-228
        if coder_regression_type(box_coder) != "BOX":
            raise NotImplementedError("only the 'BOX' BoxCoder regression type is supported")
        self.cfg = cfg
        self.gamma = scalar(atss.LOSS_GAMMA)
        self.alpha = scalar(atss.LOSS_ALPHA)
        # Matcher(FG, BG, True) of atss/loss.py:33: read by POSITIVE_TYPE 'IoU' only
        self.iou_threshold = float(getattr(atss, "FG_IOU_THRESHOLD", 0.5))
        self.bg_iou_threshold = float(getattr(atss, "BG_IOU_THRESHOLD", 0.4))
        self.topk = int(atss.TOPK)
        self.box_coder = box_coder
        self.reg_loss_type = "iou"
        self.iou_loss_weight = 1.0                   # centerness loss carries no extra weight (atss/loss.py:273)
        self.reg_loss_weight = float(atss.REG_LOSS_WEIGHT)
        self._lib = _lib.load()
        self.debug = False
        self.last_debug = None
        self.teacher_combined_loss = None
        self._workspace = None
        self._ones = None
        self._flavour = _lib.LOSS_ATSS

    def __call__(self, box_cls, box_regression, centerness, targets, anchors):
        n_levels = len(box_cls)
        heads = list(box_cls) + list(box_regression) + list(centerness)
        losses = _PAALossFunction.apply(self, targets, anchors, n_levels, True, *heads)
        return losses[0], losses[1], losses[2]


def make_atss_loss_evaluator(cfg, box_coder):
    return ATSSLossComputation(cfg, box_coder)       # atss/loss.py:279-281


def generate_retinanet_labels(matched_targets):
    """retinanet/loss.py:84-86 (kept for the constructor's signature; the kernel applies exactly this rule)."""
    return matched_targets.get_field("labels")


class RetinaNetLossComputation(PAALossComputation):
    """Drop-in for paa_core.modeling.rpn.retinanet.loss.RetinaNetLossComputation (retinanet/loss.py:19-81 on
    rpn/loss.py:41-88) with the reference's constructor: a Matcher-like object (``high_threshold``,
    ``low_threshold``, ``allow_low_quality_matches``), the RPN ``BoxCoder`` (``weights``), the label function and a
    SigmoidFocalLoss-like object (``gamma``, ``alpha``).  IoU matching is the PAA kernel's, labelling is
    `paa_retinanet_assign`, the losses come from the same streaming pass as PAA's with smooth-L1 in place of GIoU.
    ``__call__(anchors, box_cls, box_regression, targets)`` returns ``(cls_loss, regression_loss)``."""

    def __init__(self, proposal_matcher, box_coder, generate_labels_func=generate_retinanet_labels,
                 sigmoid_focal_loss=None, bbox_reg_beta=0.11, regress_norm=1.0):
        if not getattr(proposal_matcher, "allow_low_quality_matches", True):
            raise NotImplementedError("only Matcher(..., allow_low_quality_matches=True) is supported")
        name = getattr(generate_labels_func, "__name__", "")
        if name != "generate_retinanet_labels":
            raise NotImplementedError("only generate_retinanet_labels is supported, got %r" % (name,))
        self.proposal_matcher = proposal_matcher
        self.box_coder = box_coder
        self.box_cls_loss_func = sigmoid_focal_loss
        self.generate_labels_func = generate_labels_func
        self.copied_fields = ["labels"]
        self.discard_cases = ["between_thresholds"]
        self.bbox_reg_beta = float(bbox_reg_beta)
        self.regress_norm = float(regress_norm)
        self.iou_threshold = float(proposal_matcher.high_threshold)
        self.bg_iou_threshold = float(proposal_matcher.low_threshold)
        self.box_code_weights = tuple(float(w) for w in getattr(box_coder, "weights", (10.0, 10.0, 5.0, 5.0)))
        self.gamma = scalar(getattr(sigmoid_focal_loss, "gamma", 2.0))
        self.alpha = scalar(getattr(sigmoid_focal_loss, "alpha", 0.25))
        self.topk = 1                                # unused by this flavour
        self.iou_loss_weight = 0.0
        self.reg_loss_weight = 1.0
        self._lib = _lib.load()
        self.debug = False
        self.last_debug = None
        self.teacher_combined_loss = None
        self._workspace = None
        self._ones = None
        self._flavour = _lib.LOSS_RETINANET

    def forward_backward(self, anchors, box_cls, box_regression, targets, grad_losses=None):
        return super(RetinaNetLossComputation, self).forward_backward(box_cls, box_regression, None, targets, anchors,
                                                                      grad_losses)

    def __call__(self, anchors, box_cls, box_regression, targets):
        heads = list(box_cls) + list(box_regression)
        losses = _PAALossFunction.apply(self, targets, anchors, len(box_cls), False, *heads)
        return losses[0], losses[1]


def make_retinanet_loss_evaluator(cfg, box_coder):
    """retinanet/loss.py:89-107."""
    from types import SimpleNamespace
    rn = cfg.MODEL.RETINANET
    matcher = SimpleNamespace(high_threshold=rn.FG_IOU_THRESHOLD, low_threshold=rn.BG_IOU_THRESHOLD,
                              allow_low_quality_matches=True)
    focal = SimpleNamespace(gamma=rn.LOSS_GAMMA, alpha=rn.LOSS_ALPHA)
    return RetinaNetLossComputation(matcher, box_coder, generate_retinanet_labels, focal,
                                    bbox_reg_beta=rn.BBOX_REG_BETA, regress_norm=rn.BBOX_REG_WEIGHT)


class _Points(object):
    """The FCOS locations of one level as degenerate boxes (x, y, x, y): what the kernels read as `anchors`."""

    def __init__(self, bbox, size):
        self.bbox = bbox
        self.size = size


class FCOSLossComputation(PAALossComputation):
    """Drop-in for paa_core.modeling.rpn.fcos.loss.FCOSLossComputation (fcos/loss.py:36-281):
    ``FCOSLossComputation(cfg)``, ``__call__(locations, box_cls, box_regression, centerness, targets)`` ->
    ``(cls_loss, reg_loss, centerness_loss)``.  Target assignment is `paa_fcos_assign`, the losses come from the
    same streaming pass as PAA's with IOULoss on the (l, t, r, b) maps weighted by the centerness targets."""

    def __init__(self, cfg):
        fcos = cfg.MODEL.FCOS
        self.cfg = cfg
        self.gamma = scalar(fcos.LOSS_GAMMA)
        self.alpha = scalar(fcos.LOSS_ALPHA)
        self.fpn_strides = list(fcos.FPN_STRIDES)
        self.center_sampling_radius = float(fcos.CENTER_SAMPLING_RADIUS)
        self.iou_loss_type = fcos.IOU_LOSS_TYPE
        self.norm_reg_targets = bool(fcos.NORM_REG_TARGETS)
        if self.iou_loss_type not in _lib.IOU_LOSS_TYPES:
            raise NotImplementedError(self.iou_loss_type)              # layers/iou_loss.py:43-44
        self.iou_threshold = 0.0
        self.topk = 1
        self.iou_loss_weight = 1.0
        self.reg_loss_weight = 1.0
        self._lib = _lib.load()
        self.debug = False
        self.last_debug = None
        self.teacher_combined_loss = None
        self._workspace = None
        self._ones = None
        self._flavour = _lib.LOSS_FCOS
        self._points = {}

    def _as_points(self, locations, targets):
        if len(locations) > len(self.fpn_strides) or len(locations) > 5:
            raise IndexError("list index out of range")               # object_sizes_of_interest[l], fcos/loss.py:116
        points = []
        for loc in locations:
            key = (loc.data_ptr(), tuple(loc.shape))
            if key not in self._points:
                self._points[key] = torch.cat([loc, loc], dim=1).to(torch.float32).contiguous()
            points.append(self._points[key])
        return [[_Points(p, t.size) for p in points] for t in targets]

    def forward_backward(self, locations, box_cls, box_regression, centerness, targets, grad_losses=None):
        return super(FCOSLossComputation, self).forward_backward(box_cls, box_regression, centerness, targets,
                                                                 self._as_points(locations, targets), grad_losses)

    def __call__(self, locations, box_cls, box_regression, centerness, targets):
        heads = list(box_cls) + list(box_regression) + list(centerness)
        losses = _PAALossFunction.apply(self, targets, self._as_points(locations, targets), len(box_cls), True, *heads)
        return losses[0], losses[1], losses[2]


def make_fcos_loss_evaluator(cfg):
    return FCOSLossComputation(cfg)                  # fcos/loss.py:284-286
