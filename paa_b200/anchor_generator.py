"""Anchor generation for the PAA head on the device (the step right before the hot path).

Mirror of ``paa_core/modeling/rpn/anchor_generator.py``: ``AnchorGenerator(sizes, aspect_ratios,
anchor_strides, straddle_thresh)`` with ``forward(image_list, feature_maps) -> list[N] of list[L] BoxList``
(xyxy, ``.size = (w, h)``, field ``visibility``; anchor_generator.py:112-125) and the factory
``make_anchor_generator_paa(config)`` (:192-212).  The cell anchors -- a handful of numbers per level -- are
computed on the host exactly like ``generate_anchors`` (:252-330); the grid (``grid_anchors``, :73-95) and the
visibility field (:97-110) are CUDA kernels behind ``paa_grid_anchors`` / ``paa_anchor_visibility``.

The reference regenerates every grid on every forward (paa.py:128); here a grid is generated once per
(grid size, device) and the per-level tensors are shared by all images, which is also what lets the
loss / post-processing kernels read one anchor tensor per level for the whole batch.
"""
import math

import numpy as np
import torch

from paa_b200 import _lib
from paa_b200.structures import BoxList


def _whctrs(anchor):
    w = anchor[2] - anchor[0] + 1.0
    h = anchor[3] - anchor[1] + 1.0
    return w, h, anchor[0] + 0.5 * (w - 1.0), anchor[1] + 0.5 * (h - 1.0)


def _boxes_around(ws, hs, cx, cy):
    ws = np.asarray(ws, np.float64).reshape(-1, 1)
    hs = np.asarray(hs, np.float64).reshape(-1, 1)
    return np.hstack((cx - 0.5 * (ws - 1.0), cy - 0.5 * (hs - 1.0), cx + 0.5 * (ws - 1.0), cy + 0.5 * (hs - 1.0)))


def generate_cell_anchors(stride, sizes, aspect_ratios):
    """[len(ratios) * len(sizes), 4] float64 anchors of one cell: the window (0, 0, stride-1, stride-1) is
    reshaped to every aspect ratio (widths / heights rounded to integers, area kept), then scaled so that
    sqrt(area) ~ size (anchor_generator.py:252-330; ratios outer, sizes inner)."""
    scales = np.asarray(sizes, np.float64) / float(stride)
    ratios = np.asarray(aspect_ratios, np.float64)
    base = np.array([1.0, 1.0, stride, stride], np.float64) - 0.5
    w, h, cx, cy = _whctrs(base)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    by_ratio = _boxes_around(ws, hs, cx, cy)
    out = []
    for a in by_ratio:
        w, h, cx, cy = _whctrs(a)
        out.append(_boxes_around(w * scales, h * scales, cx, cy))
    return np.vstack(out)


class AnchorGenerator(torch.nn.Module):
    """FPN flavour of the reference's generator (one size tuple per stride, anchor_generator.py:46-68)."""

    def __init__(self, sizes=(128, 256, 512), aspect_ratios=(0.5, 1.0, 2.0), anchor_strides=(8, 16, 32),
                 straddle_thresh=0):
        super(AnchorGenerator, self).__init__()
        if len(anchor_strides) != len(sizes):
            raise RuntimeError("FPN should have #anchor_strides == #sizes")
        self.strides = tuple(anchor_strides)
        self.straddle_thresh = straddle_thresh
        self.cell_anchors = [
            torch.from_numpy(generate_cell_anchors(s, z if isinstance(z, (tuple, list)) else (z,),
                                                   aspect_ratios)).float()
            for s, z in zip(anchor_strides, sizes)]
        self._lib = _lib.load()
        self._grid_cache = {}
        self._vis_cache = {}

    def num_anchors_per_location(self):
        return [int(c.shape[0]) for c in self.cell_anchors]

    def _cells_on(self, device):
        key = ("cells", str(device))
        if key not in self._grid_cache:
            self._grid_cache[key] = [c.to(device).contiguous() for c in self.cell_anchors]
        return self._grid_cache[key]

    def grid_anchors(self, grid_sizes, device=None):
        """list[L] of [H*W*a, 4] float32 tensors on `device` (anchor_generator.py:73-95)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("paa_b200 has no CPU path: anchors are generated on a CUDA device")
        key = (tuple((int(h), int(w)) for h, w in grid_sizes), str(device))
        if key in self._grid_cache:
            return self._grid_cache[key]
        cells = self._cells_on(device)
        out = []
        stream = _lib.stream_handle(device)
        with _lib.device_guard(device):
            for (h, w), stride, cell in zip(key[0], self.strides, cells):
                a = int(cell.shape[0])
                t = torch.empty((h * w * a, 4), dtype=torch.float32, device=device)
                _lib.check(self._lib.paa_grid_anchors(cell.data_ptr(), a, h, w, float(stride), t.data_ptr(), stream),
                           "paa_grid_anchors")
                out.append(t)
        self._grid_cache[key] = out
        return out

    def _visibility(self, anchors, grid_key, level, image_wh):
        key = (grid_key, level, (float(image_wh[0]), float(image_wh[1])))
        if key not in self._vis_cache:
            vis = torch.empty(anchors.shape[0], dtype=torch.uint8, device=anchors.device)
            stream = _lib.stream_handle(anchors.device)
            with _lib.device_guard(anchors.device):
                _lib.check(self._lib.paa_anchor_visibility(anchors.data_ptr(), anchors.shape[0], float(image_wh[0]),
                                                           float(image_wh[1]), float(self.straddle_thresh),
                                                           vis.data_ptr(), stream), "paa_anchor_visibility")
            self._vis_cache[key] = vis.to(torch.bool) if self.straddle_thresh >= 0 else vis
        return self._vis_cache[key]

    def forward(self, image_list, feature_maps):
        """image_list: the reference's ImageList (``.image_sizes`` = [(height, width), ...]) or that list."""
        image_sizes = image_list.image_sizes if hasattr(image_list, "image_sizes") else image_list
        grid_sizes = [tuple(int(v) for v in f.shape[-2:]) for f in feature_maps]
        device = feature_maps[0].device
        per_level = self.grid_anchors(grid_sizes, device)
        grid_key = (tuple(grid_sizes), str(device))
        anchors = []
        for (image_height, image_width) in image_sizes:
            in_image = []
            for l, t in enumerate(per_level):
                boxlist = BoxList(t, (image_width, image_height), mode="xyxy")
                boxlist.add_field("visibility", self._visibility(t, grid_key, l, (image_width, image_height)))
                in_image.append(boxlist)
            anchors.append(in_image)
        return anchors


def make_anchor_generator_paa(config):
    """anchor_generator.py:192-212."""
    paa = config.MODEL.PAA
    if len(paa.ANCHOR_STRIDES) != len(paa.ANCHOR_SIZES):
        raise AssertionError("Only support FPN now")
    new_sizes = []
    for size in paa.ANCHOR_SIZES:
        per_layer = []
        for k in range(paa.SCALES_PER_OCTAVE):
            per_layer.append(math.pow(paa.OCTAVE, k / float(paa.SCALES_PER_OCTAVE)) * size)
        new_sizes.append(tuple(per_layer))
    return AnchorGenerator(tuple(new_sizes), paa.ASPECT_RATIOS, paa.ANCHOR_STRIDES, paa.STRADDLE_THRESH)
