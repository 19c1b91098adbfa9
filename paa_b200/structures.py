"""BoxList container with the part of the reference's interface the PAA path touches
(paa_core/structures/bounding_box.py:9-255, boxlist_ops.py:130-156).

The path only needs xyxy boxes, an image size ``(width, height)`` and named per-box fields; the
data-augmentation methods of the reference's class (resize / transpose / crop) are outside the path
and are not provided.  Any object with ``.bbox``, ``.size``, ``.mode`` and ``get_field`` -- in
particular the reference's own ``BoxList`` -- is accepted wherever this one is.
"""
import torch


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2:
            raise ValueError("bbox should have 2 dimensions, got {}".format(bbox.ndimension()))
        if bbox.size(-1) != 4:
            raise ValueError("last dimension of bbox should have a size of 4, got {}".format(bbox.size(-1)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size
        self.mode = mode
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        x1, y1, c, d = self.bbox.unbind(dim=-1)
        if mode == "xywh":
            new = torch.stack((x1, y1, c - x1 + 1, d - y1 + 1), dim=-1)
        else:
            new = torch.stack((x1, y1, x1 + (c - 1).clamp(min=0), y1 + (d - 1).clamp(min=0)), dim=-1)
        out = BoxList(new, self.size, mode=mode)
        out.extra_fields.update(self.extra_fields)
        return out

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    def copy_with_fields(self, fields, skip_missing=False):
        out = BoxList(self.bbox, self.size, self.mode)
        for f in fields if isinstance(fields, (list, tuple)) else [fields]:
            if self.has_field(f):
                out.add_field(f, self.get_field(f))
            elif not skip_missing:
                raise KeyError("Field '{}' not found in {}".format(f, self))
        return out

    def __repr__(self):
        return "BoxList(num_boxes={}, image_width={}, image_height={}, mode={})".format(
            len(self), self.size[0], self.size[1], self.mode)


def cat_boxlist(bboxes):
    """boxlist_ops.py:130-156."""
    assert isinstance(bboxes, (list, tuple)) and len(bboxes) > 0
    size, mode = bboxes[0].size, bboxes[0].mode
    assert all(b.size == size and b.mode == mode for b in bboxes)
    fields = set(bboxes[0].fields())
    assert all(set(b.fields()) == fields for b in bboxes)
    out = BoxList(torch.cat([b.bbox for b in bboxes], dim=0) if len(bboxes) > 1 else bboxes[0].bbox, size, mode)
    for f in fields:
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], dim=0))
    return out


def boxlist_iou(boxlist1, boxlist2):
    """Pairwise IoU of two BoxLists on the device, ``[N, M]`` (boxlist_ops.py:81-116: "+1" convention,
    the reference's float32 operation order, so equality tests against it are bit-exact).  The consumer of
    this in the reference is the COCO evaluation and the test-time-augmentation voting."""
    if boxlist1.size != boxlist2.size:
        raise RuntimeError("boxlists should have same image size, got {}, {}".format(boxlist1, boxlist2))
    from paa_b200 import _lib
    lib = _lib.load()
    b1 = boxlist1.convert("xyxy").bbox
    b2 = boxlist2.convert("xyxy").bbox
    if not b1.is_cuda or not b2.is_cuda:
        raise RuntimeError("paa_b200 has no CPU path: boxlist_iou needs CUDA tensors")
    b1 = b1.to(torch.float32).contiguous()
    b2 = b2.to(device=b1.device, dtype=torch.float32).contiguous()
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    with _lib.device_guard(b1.device):
        _lib.check(lib.paa_boxlist_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0], out.data_ptr(),
                                       _lib.stream_handle(b1.device)), "paa_boxlist_iou")
    return out
