"""PAAPostProcessor -- the reference's inference-side entry point for the PAA hot path, backed by
libpaa_b200.so.

Mirror of paa_core/modeling/rpn/paa/inference.py: same constructor arguments
(inference.py:11-34), ``forward(box_cls, box_regression, iou_pred, anchors) -> list[BoxList]`` with
xyxy boxes, ``.size = (w, h)``, fields ``labels`` (int64, 1-based) and ``scores`` (float32), and the
factory ``make_paa_postprocessor(config, box_coder)`` (inference.py:162-177).  Per-level candidate
selection, decoding, clipping, label-aware NMS, the detections-per-image cut and score voting all
run on the device in one stream-ordered sequence of kernels; the only host synchronisation is the
single read of the per-image detection counts needed to size the returned BoxLists.

Row order: the reference returns detections in ascending pre-NMS index (ml_nms.cu:132-135) where
the order inside a level is whatever ``topk(sorted=False)`` produced (inference.py:64, unspecified).
Here the order inside a level is ascending (location, class) index, so results are deterministic.
"""
import ctypes as C

import torch

from paa_b200 import _lib
from paa_b200.box_coder import coder_regression_type
from paa_b200.loss import gather_levels, points_of
from paa_b200.structures import BoxList


class PAAPostProcessor(torch.nn.Module):
    def __init__(self, pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n, min_size, num_classes,
                 box_coder, bbox_aug_enabled=False, bbox_aug_vote=False, score_voting=False):
        super(PAAPostProcessor, self).__init__()
        self.pre_nms_thresh = pre_nms_thresh
        self.pre_nms_top_n = pre_nms_top_n
        self.nms_thresh = nms_thresh
        self.fpn_post_nms_top_n = fpn_post_nms_top_n
        self.min_size = min_size
        self.num_classes = num_classes
        self.bbox_aug_enabled = bbox_aug_enabled
        self.box_coder = box_coder
        self.bbox_aug_vote = bbox_aug_vote
        self.score_voting = score_voting
        if min_size != 0:
            raise NotImplementedError("min_size is 0 in the reference's factory (inference.py:169)")
        if coder_regression_type(box_coder) != "BOX":
            raise NotImplementedError("only the 'BOX' BoxCoder regression type is supported")
        self._lib = _lib.load()
        self.debug = False
        self.last_debug = None
        self._workspace = None
        # how (anchor, regression) becomes a box: the ATSS 'BOX' coder unless a subclass says otherwise
        self._decode = (_lib.DECODE_ATSS_BOX, (0.0, 0.0, 0.0, 0.0), 0.0)

    def _workspace_for(self, device, nbytes):
        ws = self._workspace
        if ws is None or ws.device != device or ws.numel() < nbytes + 256:
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    def run_device(self, box_cls, box_regression, iou_pred, anchors):
        """Enqueues the whole post-processing path and returns device tensors
        ``(boxes [N,cap,4], scores [N,cap], labels [N,cap] int64, count [N] int32)`` without
        synchronising -- the capturable core of ``forward``."""
        lv = gather_levels(list(box_cls), list(box_regression), None if iou_pred is None else list(iou_pred),
                           anchors)
        N, L, A, Cn = lv["N"], lv["L"], lv["A"], lv["C"]
        if Cn != self.num_classes - 1:
            raise RuntimeError("box_cls has %d classes per anchor, NUM_CLASSES-1 is %d" % (Cn, self.num_classes - 1))
        if N > _lib.MAX_IMAGES:
            raise RuntimeError("at most %d images per call" % _lib.MAX_IMAGES)
        device = lv["cls"][0].device
        cap = L * int(self.pre_nms_top_n)
        args = _lib.PaaPostArgs()
        args.num_images, args.num_levels, args.num_classes = N, L, Cn
        args.anchors_per_loc, args.pre_nms_top_n = lv["apl"], int(self.pre_nms_top_n)
        args.detections_per_img = int(self.fpn_post_nms_top_n)
        args.score_voting = int(bool(self.score_voting))
        args.skip_nms = int(bool(self.bbox_aug_enabled and not self.bbox_aug_vote))      # inference.py:96
        args.pre_nms_thresh, args.nms_thresh = float(self.pre_nms_thresh), float(self.nms_thresh)
        args.anchor_image_stride = lv["anchor_stride"]
        args.head_layout = lv["layout"]          # NCHW or channels-last heads, consumed in place
        args.box_decode = self._decode[0]
        for k in range(4):
            args.decode_weights[k] = float(self._decode[1][k])
        args.decode_clip = float(self._decode[2])
        cls_l, reg_l, iou_l = lv["cls"], lv["reg"], lv["iou"]
        anchor_ptrs, hw_l, grid_w = lv["anchor_ptrs"], lv["hw"], lv["grid_w"]
        levels = args.levels
        for l in range(L):
            s = levels[l]
            s.box_cls, s.box_regression = cls_l[l].data_ptr(), reg_l[l].data_ptr()
            if iou_l is not None:
                s.iou_pred = iou_l[l].data_ptr()
            s.anchors, s.hw, s.grid_w = anchor_ptrs[l], hw_l[l], grid_w[l]
        image_wh = args.image_wh
        for i in range(N):
            w, h = anchors[i][0].size
            wh = image_wh[i]
            wh[0], wh[1] = w, h
        nbytes = self._lib.paa_postprocess_workspace_bytes(N, A, Cn, L, int(self.pre_nms_top_n))
        ws = self._workspace_for(device, nbytes)
        base = (ws.data_ptr() + 255) // 256 * 256
        args.workspace, args.workspace_bytes = base, ws.numel() - (base - ws.data_ptr())
        boxes = torch.empty((N, cap, 4), dtype=torch.float32, device=device)
        scores = torch.empty((N, cap), dtype=torch.float32, device=device)
        labels = torch.empty((N, cap), dtype=torch.int64, device=device)
        count = torch.empty(N, dtype=torch.int32, device=device)
        args.out_boxes, args.out_scores = boxes.data_ptr(), scores.data_ptr()
        args.out_labels, args.out_count = labels.data_ptr(), count.data_ptr()
        dbg = None
        if self.debug:
            dbg = dict(pre_boxes=torch.zeros((N, cap, 4), dtype=torch.float32, device=device),
                       pre_scores=torch.zeros((N, cap), dtype=torch.float32, device=device),
                       pre_labels=torch.zeros((N, cap), dtype=torch.int32, device=device),
                       pre_count=torch.zeros((N, L), dtype=torch.int32, device=device),
                       nms_keep=torch.zeros((N, cap), dtype=torch.uint8, device=device))
            args.dbg_pre_boxes, args.dbg_pre_scores = dbg["pre_boxes"].data_ptr(), dbg["pre_scores"].data_ptr()
            args.dbg_pre_labels, args.dbg_pre_count = dbg["pre_labels"].data_ptr(), dbg["pre_count"].data_ptr()
            args.dbg_nms_keep = dbg["nms_keep"].data_ptr()
        stream = _lib.stream_handle(device)
        with _lib.device_guard(device):
            _lib.check(self._lib.paa_postprocess(C.byref(args), stream), "paa_postprocess")
        self.last_debug = dbg
        self._keep = (lv, ws)
        return boxes, scores, labels, count

    def forward(self, box_cls, box_regression, iou_pred, anchors):
        boxes, scores, labels, count = self.run_device(box_cls, box_regression, iou_pred, anchors)
        # per-image views are cut while the kernels run; only the row counts wait for the device
        per_image = list(zip(boxes.unbind(0), labels.unbind(0), scores.unbind(0)))
        counts = count.tolist()                       # the one host sync: result sizes
        results = []
        for i, c in enumerate(counts):
            b, l, s = per_image[i]
            bl = BoxList(b.narrow(0, 0, c), anchors[i][0].size, mode="xyxy")
            bl.add_field("labels", l.narrow(0, 0, c))
            bl.add_field("scores", s.narrow(0, 0, c))
            results.append(bl)
        return results


def make_paa_postprocessor(config, box_coder):
    return PAAPostProcessor(
        pre_nms_thresh=config.MODEL.PAA.INFERENCE_TH,
        pre_nms_top_n=config.MODEL.PAA.PRE_NMS_TOP_N,
        nms_thresh=config.MODEL.PAA.NMS_TH,
        fpn_post_nms_top_n=config.TEST.DETECTIONS_PER_IMG,
        min_size=0,
        num_classes=config.MODEL.PAA.NUM_CLASSES,
        bbox_aug_enabled=config.TEST.BBOX_AUG.ENABLED,
        box_coder=box_coder,
        bbox_aug_vote=config.TEST.BBOX_AUG.VOTE,
        score_voting=config.MODEL.PAA.INFERENCE_SCORE_VOTING,
    )


class ATSSPostProcessor(PAAPostProcessor):
    """``paa_core/modeling/rpn/atss/inference.py:9-119`` on the same kernels: the ATSS head's
    third output is a centerness map where PAA's is an IoU prediction, both enter the score as
    ``sqrt(sigmoid(cls) * sigmoid(third))``, the threshold is applied to ``sigmoid(cls)`` alone, and there is no
    score voting.  (ATSS ranks the per-level top-k by the product and takes the square root afterwards,
    atss/inference.py:53-75; the square root is monotone, so the selected set is the same unless two distinct
    products at the k-th place round to the same float32 root.)"""

    def __init__(self, pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n, min_size, num_classes,
                 box_coder, bbox_aug_enabled=False, bbox_aug_vote=False):
        super(ATSSPostProcessor, self).__init__(pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n,
                                                min_size, num_classes, box_coder, bbox_aug_enabled, bbox_aug_vote,
                                                score_voting=False)

    def forward(self, box_cls, box_regression, centerness, anchors):
        if centerness is None:
            raise RuntimeError("ATSSPostProcessor needs the centerness maps")
        return super(ATSSPostProcessor, self).forward(box_cls, box_regression, centerness, anchors)


def make_atss_postprocessor(config, box_coder):
    """atss/inference.py:122-137."""
    return ATSSPostProcessor(
        pre_nms_thresh=config.MODEL.ATSS.INFERENCE_TH,
        pre_nms_top_n=config.MODEL.ATSS.PRE_NMS_TOP_N,
        nms_thresh=config.MODEL.ATSS.NMS_TH,
        fpn_post_nms_top_n=config.TEST.DETECTIONS_PER_IMG,
        min_size=0,
        num_classes=config.MODEL.ATSS.NUM_CLASSES,
        bbox_aug_enabled=config.TEST.BBOX_AUG.ENABLED,
        box_coder=box_coder,
        bbox_aug_vote=config.TEST.BBOX_AUG.VOTE,
    )


class RetinaNetPostProcessor(PAAPostProcessor):
    """``paa_core/modeling/rpn/retinanet/inference.py:14-173`` on the same kernels: no third head
    output, several anchors per location, boxes decoded with the RPN ``BoxCoder`` (modeling/box_coder.py:51-95,
    weights (10, 10, 5, 5)).  The reference runs one single-class NMS per class and concatenates the classes
    (rows class-major, score-descending); the label-aware NMS here keeps the same set and returns it in
    ascending pre-NMS index like the PAA post-processor -- consumers (coco_eval) do not depend on row order."""

    def __init__(self, pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n, min_size, num_classes,
                 box_coder=None):
        import math
        weights = tuple(getattr(box_coder, "weights", (10.0, 10.0, 5.0, 5.0)))
        clip = float(getattr(box_coder, "bbox_xform_clip", math.log(1000.0 / 16)))
        super(RetinaNetPostProcessor, self).__init__(pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n,
                                                     min_size, num_classes, _BoxCoderTag(), False, False,
                                                     score_voting=False)
        self.box_coder = box_coder
        self._decode = (_lib.DECODE_LEGACY, weights, clip)

    def forward(self, anchors, box_cls, box_regression, targets=None):
        """Argument order of RPNPostProcessor.forward (rpn/inference.py:123): anchors first."""
        return super(RetinaNetPostProcessor, self).forward(box_cls, box_regression, None, anchors)


class _BoxCoderTag(object):
    """Stands in for a coder object where the decode is selected by `_decode` instead."""
    regression_type = "BOX"


def make_retinanet_postprocessor(config, rpn_box_coder, is_train=False):
    """retinanet/inference.py:176-194."""
    return RetinaNetPostProcessor(
        pre_nms_thresh=config.MODEL.RETINANET.INFERENCE_TH,
        pre_nms_top_n=config.MODEL.RETINANET.PRE_NMS_TOP_N,
        nms_thresh=config.MODEL.RETINANET.NMS_TH,
        fpn_post_nms_top_n=config.TEST.DETECTIONS_PER_IMG,
        min_size=0,
        num_classes=config.MODEL.RETINANET.NUM_CLASSES,
        box_coder=rpn_box_coder,
    )


class FCOSPostProcessor(PAAPostProcessor):
    """``paa_core/modeling/rpn/fcos/inference.py:12-147`` on the same kernels: anchor-free --
    every location (x, y) regresses its distances to the four box edges (decoded as
    ``(x - l, y - t, x + r, y + b)``, :93-98) and carries a centerness logit; score = sqrt(sigmoid(cls) *
    sigmoid(centerness)), threshold on sigmoid(cls) alone."""

    def __init__(self, pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n, min_size, num_classes,
                 bbox_aug_enabled=False):
        super(FCOSPostProcessor, self).__init__(pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n,
                                                min_size, num_classes, _BoxCoderTag(), bbox_aug_enabled, False,
                                                score_voting=False)
        self._decode = (_lib.DECODE_LTRB, (0.0, 0.0, 0.0, 0.0), 0.0)
        self._points = {}

    def forward(self, locations, box_cls, box_regression, centerness, image_sizes):
        """locations: list[L] of [H*W, 2] tensors (fcos.py compute_locations); image_sizes: [(h, w), ...]."""
        points = points_of(self._points, locations)        # (x, y, x, y); validated, bounded cache
        anchors = [[BoxList(p, (int(w), int(h)), mode="xyxy") for p in points] for (h, w) in image_sizes]
        return super(FCOSPostProcessor, self).forward(box_cls, box_regression, centerness, anchors)


def make_fcos_postprocessor(config):
    """fcos/inference.py:169-184."""
    return FCOSPostProcessor(
        pre_nms_thresh=config.MODEL.FCOS.INFERENCE_TH,
        pre_nms_top_n=config.MODEL.FCOS.PRE_NMS_TOP_N,
        nms_thresh=config.MODEL.FCOS.NMS_TH,
        fpn_post_nms_top_n=config.TEST.DETECTIONS_PER_IMG,
        min_size=0,
        num_classes=config.MODEL.FCOS.NUM_CLASSES,
        bbox_aug_enabled=config.TEST.BBOX_AUG.ENABLED,
    )


def ml_nms(boxes, scores, labels, nms_thresh):
    """Device replacement for ``paa_core._C.ml_nms`` (csrc/ml_nms.h:10-27): indices of the kept boxes,
    ascending.  boxes [n,4] float32 cuda, scores [n], labels [n] (any float/int dtype, integer valued)."""
    lib = _lib.load()
    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device="cpu")      # ml_nms.h:18-19
    if not boxes.is_cuda:
        raise RuntimeError("CPU version not implemented")             # ml_nms.h:26
    device = boxes.device
    b = boxes.contiguous().float()
    s = scores.contiguous().float()
    lab = labels.contiguous().float()
    keep = torch.empty(n, dtype=torch.uint8, device=device)
    num = torch.empty(1, dtype=torch.int32, device=device)
    nbytes = lib.paa_ml_nms_workspace_bytes(n)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    base = (ws.data_ptr() + 255) // 256 * 256
    stream = _lib.stream_handle(device)
    with _lib.device_guard(device):
        _lib.check(lib.paa_ml_nms(b.data_ptr(), s.data_ptr(), lab.data_ptr(), n, float(nms_thresh), keep.data_ptr(),
                                  num.data_ptr(), base, ws.numel() - (base - ws.data_ptr()), stream), "paa_ml_nms")
    return torch.nonzero(keep).squeeze(1)


def boxlist_ml_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores", label_field="labels"):
    """boxlist_ops.py:35-59 on top of the device kernel."""
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = ml_nms(boxlist.bbox, boxlist.get_field(score_field), boxlist.get_field(label_field).float(), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep.to(boxlist.bbox.device)].convert(mode)
