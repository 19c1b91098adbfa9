"""Merging of test-time-augmentation detections on the device.

Mirror of ``merge_result_from_multi_scales`` / ``boxlist_nms`` / ``bbox_vote`` / ``soft_bbox_vote`` in
``paa_core/engine/bbox_aug_vote.py:140-310``: the detections pooled from all scales and flips of an image are
merged class by class ('nms', 'vote' or 'soft-vote'), concatenated in class order and cut to the best
``cfg.MODEL.ATSS.PRE_NMS_TOP_N``.  The reference does this with numpy loops on the CPU (one device->host copy
and one host->device copy per class); here one call of ``paa_box_vote`` per image does all classes.
Configuration keys are the ones the reference reads from its global cfg: ``MODEL.RETINANET.NUM_CLASSES``,
``MODEL.RETINANET.INFERENCE_TH``, ``MODEL.ATSS.NMS_TH``, ``MODEL.ATSS.PRE_NMS_TOP_N``.
"""
import ctypes as C

import torch

from paa_b200 import _lib
from paa_b200.structures import BoxList

MODES = {"nms": 0, "vote": 1}


def merge_result_from_multi_scales(boxlists, cfg, nms_type="nms", vote_thresh=0.65):
    lib = _lib.load()
    mode = MODES.get(nms_type, 2)                      # anything else is soft voting (bbox_aug_vote.py:188-191)
    num_classes = int(cfg.MODEL.RETINANET.NUM_CLASSES)
    nms_thresh = float(cfg.MODEL.ATSS.NMS_TH)
    max_det = int(cfg.MODEL.ATSS.PRE_NMS_TOP_N)
    soft_thresh = float(cfg.MODEL.RETINANET.INFERENCE_TH)
    results = []
    for boxlist in boxlists:
        boxlist = boxlist.convert("xyxy")
        boxes = boxlist.bbox
        if not boxes.is_cuda:
            raise RuntimeError("paa_b200 has no CPU path: detections are on %s" % boxes.device)
        scores = boxlist.get_field("scores").to(torch.float32)
        labels = boxlist.get_field("labels")
        valid = (labels >= 1) & (labels < num_classes)          # the class loop of :150 skips everything else
        if not bool(valid.all()):
            boxes, scores, labels = boxes[valid], scores[valid], labels[valid]
        boxes = boxes.to(torch.float32).contiguous()
        scores = scores.contiguous()
        flabels = labels.to(torch.float32).contiguous()
        n = int(boxes.shape[0])
        device = boxes.device
        if mode != 0 and nms_thresh <= 0:                        # boxlist_nms :179-180 returns its input
            mode_i, cap = 0, n
            nms_t = 2.0                                           # IoU never exceeds 1: nothing is suppressed
        else:
            mode_i, cap, nms_t = mode, (2 * n if mode == 2 else n), nms_thresh
        cap = max(cap, 1)
        out_boxes = torch.empty((cap, 4), dtype=torch.float32, device=device)
        out_scores = torch.empty(cap, dtype=torch.float32, device=device)
        out_labels = torch.empty(cap, dtype=torch.int64, device=device)
        out_count = torch.empty(1, dtype=torch.int32, device=device)
        ws = torch.empty(lib.paa_box_vote_workspace_bytes(n) + 256, dtype=torch.uint8, device=device)
        base = (ws.data_ptr() + 255) // 256 * 256
        with _lib.device_guard(device):
            _lib.check(lib.paa_box_vote(boxes.data_ptr(), scores.data_ptr(), flabels.data_ptr(), n, mode_i,
                                        float(vote_thresh), nms_t, soft_thresh, max_det, out_boxes.data_ptr(),
                                        out_scores.data_ptr(), out_labels.data_ptr(), out_count.data_ptr(), base,
                                        ws.numel() - (base - ws.data_ptr()),
                                        _lib.stream_handle(device)), "paa_box_vote")
        c = int(out_count.item())                                 # the one host sync: result size
        result = BoxList(out_boxes[:c], boxlist.size, mode="xyxy")
        result.add_field("scores", out_scores[:c])
        result.add_field("labels", out_labels[:c])
        results.append(result)
    return results
