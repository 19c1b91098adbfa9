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
    available (``PAA_NORM_EXCHANGE=nccl`` forces the fallback).

    Like the all-reduce it replaces, the wait for the peers has no deadline by default: a rank that is busy
    saving a checkpoint (engine/trainer.py:110-111) or waiting for its data loader is simply waited for.
    ``PAA_PEER_TIMEOUT_S`` (seconds, default 0 = none) sets one; when it expires the waiting kernel records the
    rank it was waiting for in `status` (pinned host memory) and traps -- an error, never NaN data -- and the
    next call of an evaluator raises with that record (`raise_if_timed_out`)."""
    _by_device = {}

    def __init__(self, device, buffer):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.buffer = buffer
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, group=dist.group.WORLD.group_name)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.status = torch.zeros(4, dtype=torch.int32).pin_memory()
        self._status_np = self.status.numpy()
        self.timeout_s = float(os.environ.get("PAA_PEER_TIMEOUT_S", "0") or 0.0)
        torch.cuda.synchronize(device)      # zeroed before the collective in get() lets anyone write into it

    def raise_if_timed_out(self):
        st = self._status_np
        if st[0]:
            raise RuntimeError("paa_b200: the peer exchange of the loss normalisers timed out after %g s waiting for "
                               "rank %d (exchange step %d); the CUDA context of this process is no longer usable"
                               % (self.timeout_s, int(st[1]), int(st[2])))

    @staticmethod
    def _allocate(device):
        """The local, non-collective half of the set-up: anything that can fail on one rank alone (import,
        unsupported device, allocation) fails here, BEFORE the collective rendezvous."""
        import torch.distributed._symmetric_memory as symm_mem
        return symm_mem.empty(_lib.PEER_BUFFER_DOUBLES, dtype=torch.float64, device=device)

    @classmethod
    def get(cls, device):
        """The exchange for `device`, or None when it cannot be used (then the caller all-reduces).  The
        decision is collective and taken before any rank enters the rendezvous: every rank first reports whether
        its local allocation worked (all-reduce MIN), and only if all did do the ranks rendezvous -- a rank that
        cannot allocate symmetric memory makes everybody fall back instead of leaving the others blocked."""
        key = (device.type, device.index)
        if key not in cls._by_device:
            state = None
            import torch.distributed as dist
            usable = (os.environ.get("PAA_NORM_EXCHANGE", "peer") != "nccl" and dist.is_available()
                      and dist.is_initialized() and dist.get_world_size() == get_num_gpus()
                      and dist.get_world_size() <= _lib.MAX_PEERS)
            if usable:
                why, buffer = "", None
                try:
                    buffer = cls._allocate(device)
                except Exception as e:  # noqa: BLE001 - symmetric memory unavailable on this rank
                    why = str(e)
                ok = torch.tensor([1 if buffer is not None else 0], dtype=torch.int32, device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 1:
                    state = cls(device, buffer)
                elif buffer is None:
                    import warnings
                    warnings.warn("paa_b200: peer-memory normaliser exchange unavailable (%s); "
                                  "using all_reduce" % (why,))
            cls._by_device[key] = state
        return cls._by_device[key]


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("paa_b200 has no CPU path: %s is on %s" % (name, t.device))


_warned_layout = set()


def _dense_both_ways(t):
    """One channel or one location: the NCHW and the channels-last arrangement are the same bytes."""
    return t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def call_layout(box_cls):
    """The memory layout of a call: channels-last (NHWC) when the classification logits -- the tensors that carry
    the bytes -- are dense `torch.channels_last` tensors, else NCHW (what PAAHead.forward produces, paa.py:90-108).
    Both are consumed in place (include/paa_b200.h, `head_layout`); tensors of the call in any other arrangement are
    copied to the call's layout by `_head`, with a warning."""
    telling = [t for t in box_cls if t.dim() == 4 and not _dense_both_ways(t)]
    if telling and all(t.is_contiguous(memory_format=torch.channels_last) for t in telling):
        return _lib.LAYOUT_NHWC
    return _lib.LAYOUT_NCHW


def _head(t, name, layout=0):
    """A head tensor as the kernels read it: float32, dense in the call's layout (NCHW-contiguous or channels-last).
    Anything else -- a sliced view, a tensor in the other layout than the call's logits -- is copied, which for the
    classification logits of a 16-image batch is a 115 MB round trip per call: say so once instead of hiding it."""
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    fmt = torch.channels_last if layout == _lib.LAYOUT_NHWC else torch.contiguous_format
    if t.is_contiguous(memory_format=fmt):
        return t
    if name not in _warned_layout:
        _warned_layout.add(name)
        import warnings
        warnings.warn("paa_b200: %s is not dense in the call's layout (%s; strides %s for shape %s); it is copied on "
                      "every call -- %.1f MB here.  Keep all head outputs of a call either in torch.contiguous_format "
                      "or in torch.channels_last."
                      % (name, "channels-last" if layout == _lib.LAYOUT_NHWC else "NCHW", tuple(t.stride()),
                         tuple(t.shape), t.numel() * 4 / 1e6), stacklevel=3)
    return t.contiguous(memory_format=fmt)


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
    layout = call_layout(box_cls)
    cls = [_head(t, "box_cls", layout) for t in box_cls]
    reg = [_head(t, "box_regression", layout) for t in box_regression]
    iou = None if iou_pred is None else [_head(t, "iou_pred", layout) for t in iou_pred]
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
                anchor_ptrs=level_ptrs, anchor_stride=stride, keep=keep, layout=layout)


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


class _GraphedStep(object):
    """One assign+loss step of an evaluator on a fixed set of head / anchor tensors, captured in a CUDA graph that is
    replayed for every new set of targets (PaaLossArgs.gt_offsets_dev: no kernel parameter or grid size depends on
    the ground-truth counts).  Per call the host only checks the targets, writes their ranges into a pinned buffer,
    concatenates the boxes / labels into the graph's static buffers and launches the graph."""
    RING = 4

    def __init__(self, owner, box_cls, box_reg, iou_pred, anchors, need_grad):
        lv = gather_levels(list(box_cls), list(box_reg), None if iou_pred is None else list(iou_pred), anchors)
        self.owner, self.lv = owner, lv
        N, L, A = lv["N"], lv["L"], lv["A"]
        device = lv["cls"][0].device
        self.device, self.N = device, N
        self.cap_img = int(owner.gt_per_image_capacity)
        self.cap = N * self.cap_img
        self.has_iou = iou_pred is not None
        self.world = owner._world()
        self.grads = owner._alloc_grads(lv, self.has_iou) if need_grad else None
        args = owner._make_args(lv, self.has_iou, self.world, self.grads)
        self.gt_boxes = torch.zeros((self.cap, 4), dtype=torch.float32, device=device)
        self.gt_labels = torch.ones(self.cap, dtype=torch.int64, device=device)
        self.offsets_dev = torch.zeros(N + 1, dtype=torch.int32, device=device)
        self.offsets_host = [torch.zeros(N + 1, dtype=torch.int32).pin_memory() for _ in range(self.RING)]
        self.offsets_np = [t.numpy() for t in self.offsets_host]
        self.copied = [None] * self.RING
        self.calls = 0
        args.gt_boxes, args.gt_labels = self.gt_boxes.data_ptr(), self.gt_labels.data_ptr()
        args.gt_offsets_dev = self.offsets_dev.data_ptr()
        args.gt_capacity, args.gt_per_image_capacity = self.cap, self.cap_img
        nbytes = owner._lib.paa_loss_workspace_bytes(N, A, self.cap, L, owner.topk)
        self.ws = torch.empty(nbytes + 512, dtype=torch.uint8, device=device)       # pinned by the graph: its own
        base = (self.ws.data_ptr() + 255) // 256 * 256
        args.workspace, args.workspace_bytes = base, self.ws.numel() - (base - self.ws.data_ptr())
        self.normalisers = torch.empty(2, dtype=torch.float64, device=device)
        self.losses = torch.empty(3, dtype=torch.float32, device=device)
        args.normalisers, args.losses, args.grad_losses = self.normalisers.data_ptr(), self.losses.data_ptr(), None
        self.args = args
        self.graph = None

    def _capture(self):
        owner, dev = self.owner, self.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # loads the kernels, sets up the peer exchange: none of it capturable
            owner._launch(self.args, dev, self.world, self.normalisers, None)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            owner._launch(self.args, dev, self.world, self.normalisers, None)
        self.graph = graph

    def run(self, targets, anchors):
        owner, N = self.owner, self.N
        offsets, boxes, labels, sum_g = owner._collect_targets(targets, anchors, N)
        gmax = 0
        for i in range(N):
            g = offsets[i + 1] - offsets[i]
            if g <= 0:                                               # matcher.py:53-58
                raise ValueError("No ground-truth boxes available for one of the images during training "
                                 "(image %d)" % i)
            gmax = g if g > gmax else gmax
        if sum_g > self.cap or gmax > self.cap_img:
            return None                                              # does not fit the plan: eager path
        slot = self.calls % self.RING
        self.calls += 1
        if self.copied[slot] is not None:
            self.copied[slot].synchronize()                          # the copy that last read this buffer is done
        else:
            self.copied[slot] = torch.cuda.Event()
        self.offsets_np[slot][:] = offsets
        self.offsets_dev.copy_(self.offsets_host[slot], non_blocking=True)
        self.copied[slot].record()
        b0, l0 = boxes[0], labels[0]
        if b0.dtype == torch.float32 and b0.device == self.device and l0.dtype == torch.int64 and \
                l0.device == self.device:
            torch.cat(boxes, dim=0, out=self.gt_boxes[:sum_g])
            torch.cat(labels, dim=0, out=self.gt_labels[:sum_g])
        else:
            self.gt_boxes[:sum_g].copy_(torch.cat(boxes, dim=0))
            self.gt_labels[:sum_g].copy_(torch.cat(labels, dim=0))
        if self.graph is None:
            self.args.gt_offsets[:N + 1] = offsets                   # the plan's host-side argument checks
            self._capture()
        peer = PeerNormExchange.get(self.device) if self.world > 1 else None
        if peer is not None:
            peer.raise_if_timed_out()
        self.graph.replay()
        # the three losses leave as a copy (a caller may keep them across steps); the gradients alias the graph's
        # buffers until the next call on the same head tensors
        return self.losses.clone(), self.grads, dict(args=self.args, keep=(), device=self.device)


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
        self._init_runtime()
        self._flavour = _lib.LOSS_PAA       # which assignment / weighting the kernels apply (ATSS subclass)

    # -- plumbing -------------------------------------------------------------------------------
    def _init_runtime(self):
        """State shared by every flavour of evaluator (the subclasses have their own constructors)."""
        self._lib = _lib.load()          # raises if the CUDA library is missing
        self.debug = False               # True: keep per-stage parity outputs of the last call
        self.last_debug = None
        self.teacher_combined_loss = None   # [N, A] float32 cuda tensor: stage-wise parity protocol
        self._workspace = None
        self._ones = None
        # Graph mode (opt-in, `use_graph = True` or PAA_B200_GRAPH=1): the step is captured once per set of head /
        # anchor tensors and REPLAYED for every later call on the same tensors, whatever the targets -- the GT ranges
        # live in device memory (PaaLossArgs.gt_offsets_dev), so the captured grids do not depend on them.  Constraints:
        # the returned gradients (and nothing else) alias buffers that the next call on the same tensors overwrites;
        # one evaluator serves one stream; images hold at most `gt_per_image_capacity` ground-truth boxes (a batch
        # that does not fit simply takes the eager path).
        self.use_graph = os.environ.get("PAA_B200_GRAPH", "0") not in ("", "0")
        self.gt_per_image_capacity = 128
        self._graphs = {}

    def _workspace_for(self, device, nbytes):
        ws = self._workspace
        if ws is None or ws.device != device or ws.numel() < nbytes:
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    def _collect_targets(self, targets, anchors, N):
        """Per-image GT ranges and the tensors to concatenate, with the reference's checks."""
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
        return offsets, boxes, labels, sum_g

    def _world(self):
        # the RetinaNet loss normalises by this rank's own counts (retinanet/loss.py:70,79): no exchange
        return 1 if self._flavour == _lib.LOSS_RETINANET else get_num_gpus()

    def _make_args(self, lv, has_iou, world, grads):
        """PaaLossArgs with everything that does not depend on the targets."""
        N, L = lv["N"], lv["L"]
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
        args.head_layout = lv["layout"]
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
        return args

    @staticmethod
    def _alloc_grads(lv, has_iou):
        """Gradient tensors shaped (and laid out in memory) like the heads, cut from ONE allocation (15 allocator
        calls -> 1)."""
        nhwc = lv["layout"] == _lib.LAYOUT_NHWC
        groups = [lv["cls"], lv["reg"]] + ([lv["iou"]] if has_iou else [])
        # every piece starts on a 16-byte boundary: the kernels write float4
        sizes = [[(t.numel() + 3) // 4 * 4 for t in g] for g in groups]
        flat = torch.empty(sum(sum(g) for g in sizes), dtype=torch.float32, device=lv["cls"][0].device)
        out, o = [], 0
        for g, sz in zip(groups, sizes):
            views = []
            for t, n in zip(g, sz):
                piece = flat[o:o + t.numel()]
                if nhwc:       # the channels-last arrangement of the same logical [N, ch, H, W] tensor
                    b, ch, h, w = t.shape
                    views.append(piece.view(b, h, w, ch).permute(0, 3, 1, 2))
                else:
                    views.append(piece.view(t.shape))
                o += n
            out.append(views)
        return dict(cls=out[0], reg=out[1], iou=out[2] if has_iou else None, flat=flat)

    def _assign_fn(self):
        name = {_lib.LOSS_PAA: "paa_assign", _lib.LOSS_ATSS: "paa_atss_assign",
                _lib.LOSS_RETINANET: "paa_retinanet_assign", _lib.LOSS_FCOS: "paa_fcos_assign"}[self._flavour]
        return name, getattr(self._lib, name)

    def _launch(self, args, device, world, normalisers, dbg):
        """Enqueues the step described by `args` on torch's current stream."""
        stream = _lib.stream_handle(device)
        assign_name, assign = self._assign_fn()
        with _lib.device_guard(device):
            peer = PeerNormExchange.get(device) if world > 1 else None
            if peer is not None:
                peer.raise_if_timed_out()
                args.rank = peer.rank
                for r, ptr in enumerate(peer.ptrs):
                    args.peer_norm[r] = ptr
                args.peer_timeout_s = peer.timeout_s
                args.peer_status = peer.status.data_ptr()
            if world == 1 and self._flavour == _lib.LOSS_PAA:
                # one rank: nothing crosses ranks between the two halves -- the single fused entry point
                _lib.check(self._lib.paa_assign_loss(C.byref(args), stream), "paa_assign_loss")
                if dbg is not None:
                    dbg["local_normalisers"] = normalisers
            else:
                _lib.check(assign(C.byref(args), stream), assign_name)
                if dbg is not None:
                    dbg["local_normalisers"] = normalisers.clone()  # this rank's own pair, before the exchange
                if world > 1 and peer is None:
                    reduce_normalisers(normalisers)                     # loss.py:321,338 in one message
                _lib.check(self._lib.paa_loss(C.byref(args), stream), "paa_loss")

    def _run(self, box_cls, box_reg, iou_pred, targets, anchors, need_grad):
        if self.use_graph and not self.debug and self.teacher_combined_loss is None:
            out = self._run_graphed(box_cls, box_reg, iou_pred, targets, anchors, need_grad)
            if out is not None:
                return out
        lv = gather_levels(box_cls, box_reg, iou_pred, anchors)
        N, L, A = lv["N"], lv["L"], lv["A"]
        device = lv["cls"][0].device
        offsets, boxes, labels, sum_g = self._collect_targets(targets, anchors, N)
        gt_boxes = torch.cat(boxes, dim=0)
        if gt_boxes.dtype != torch.float32 or gt_boxes.device != device:
            gt_boxes = gt_boxes.to(device=device, dtype=torch.float32)
        gt_labels = torch.cat(labels, dim=0)
        if gt_labels.dtype != torch.int64 or gt_labels.device != device:
            gt_labels = gt_labels.to(device=device, dtype=torch.int64)
        has_iou = iou_pred is not None
        world = self._world()
        grads = self._alloc_grads(lv, has_iou) if need_grad else None
        args = self._make_args(lv, has_iou, world, grads)
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
        self._launch(args, device, world, normalisers, dbg)
        if dbg is not None:
            dbg["normalisers"] = normalisers
            dbg["gt_offsets"] = offsets
            self.last_debug = dbg
        # keep every tensor whose pointer the kernels use alive until the stream work is enqueued
        call = dict(args=args, keep=(lv, gt_boxes, gt_labels, ws, normalisers, teacher), device=device)
        return losses, grads, call

    # -- graph mode -----------------------------------------------------------------------------
    def _run_graphed(self, box_cls, box_reg, iou_pred, targets, anchors, need_grad):
        """The step as ONE graph launch.  Returns None when this batch cannot take the captured path (too many
        ground-truth boxes for the capacities the graph was planned with): the caller then runs eagerly."""
        has_iou = iou_pred is not None
        heads = list(box_cls) + list(box_reg) + (list(iou_pred) if has_iou else [])
        if not heads or not _anchors_shared(anchors, len(anchors), len(box_cls)):
            return None                  # per-image anchor tensors are stacked afresh on every call: nothing to pin
        key = (tuple(t.data_ptr() for t in heads), heads[0].shape[0], tuple(b.bbox.data_ptr() for b in anchors[0]),
               bool(need_grad), self._world(), tuple(heads[0].stride()))
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 8:                       # bounded: a training loop cycles through few shapes
                self._graphs.pop(next(iter(self._graphs)))
            entry = _GraphedStep(self, box_cls, box_reg, iou_pred, anchors, need_grad)
            self._graphs[key] = entry
        return entry.run(targets, anchors)

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
        self._init_runtime()
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
        self._init_runtime()
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


_POINTS_CACHE_ENTRIES = 16


def points_of(cache, locations):
    """The FCOS locations of every level as degenerate boxes (x, y, x, y).  The reference rebuilds its location
    tensors on every forward (fcos/fcos.py:185-209) and frees them, so a cache keyed by address would return stale
    points once the allocator hands the address to a different grid: an entry is only reused for the very same
    tensor object at the same version (`entry[0] is loc`; holding `loc` also keeps its address from being recycled),
    and the cache is bounded."""
    out = []
    for loc in locations:
        key = id(loc)
        hit = cache.get(key)
        if hit is None or hit[0] is not loc or hit[1] != loc._version:
            if len(cache) >= _POINTS_CACHE_ENTRIES:
                cache.pop(next(iter(cache)))
            hit = (loc, loc._version, torch.cat([loc, loc], dim=1).to(torch.float32).contiguous())
            cache[key] = hit
        out.append(hit[2])
    return out


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
        self._init_runtime()
        self._flavour = _lib.LOSS_FCOS
        self._points = {}

    def _as_points(self, locations, targets):
        if len(locations) > len(self.fpn_strides) or len(locations) > 5:
            raise IndexError("list index out of range")               # object_sizes_of_interest[l], fcos/loss.py:116
        return [[_Points(p, t.size) for p in points_of(self._points, locations)] for t in targets]

    def forward_backward(self, locations, box_cls, box_regression, centerness, targets, grad_losses=None):
        return super(FCOSLossComputation, self).forward_backward(box_cls, box_regression, centerness, targets,
                                                                 self._as_points(locations, targets), grad_losses)

    def __call__(self, locations, box_cls, box_regression, centerness, targets):
        heads = list(box_cls) + list(box_regression) + list(centerness)
        losses = _PAALossFunction.apply(self, targets, self._as_points(locations, targets), len(box_cls), True, *heads)
        return losses[0], losses[1], losses[2]


def make_fcos_loss_evaluator(cfg):
    return FCOSLossComputation(cfg)                  # fcos/loss.py:284-286
