"""GPU: out-of-bounds writes around the buffers the kernels are handed.  compute-sanitizer is closed on this GPU pool
(it answers "closed on this pool and stays closed"), so the memcheck it would have done on the workspace and the
gradient buffers is approximated here: the evaluators' allocations are placed inside larger buffers filled with a
canary pattern, and after a step (eager, crowded GT-list split, graph mode, post-processing) every guard byte must be
untouched.  Run-to-run bit-identity of all outputs (tests/test_gpu_loss.py::test_full_size_properties_*) stands in for
racecheck."""
import numpy as np
import pytest
import torch

from paa_b200 import synthetic
from tests.helpers import to_device_inputs

pytestmark = pytest.mark.gpu

GUARD = 1 << 16          # bytes on either side
PATTERN = 0xA5


class Guarded(object):
    """Allocations carved out of canary-filled buffers; `check()` verifies the guards."""

    def __init__(self):
        self.buffers = []

    def alloc(self, nbytes, device):
        nbytes = (int(nbytes) + 255) // 256 * 256
        raw = torch.full((nbytes + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=device)
        self.buffers.append((raw, nbytes))
        return raw[GUARD:GUARD + nbytes]

    def check(self):
        torch.cuda.synchronize()
        for raw, nbytes in self.buffers:
            lo, hi = raw[:GUARD], raw[GUARD + nbytes:]
            assert bool((lo == PATTERN).all()), "write below a buffer of %d bytes" % nbytes
            assert bool((hi == PATTERN).all()), "write past a buffer of %d bytes" % nbytes


def _guard_evaluator(ev, guard):
    def workspace_for(device, nbytes):
        ws = guard.alloc(nbytes + 256, device)
        ev._workspace = ws
        return ws

    def alloc_grads(lv, has_iou):
        groups = [lv["cls"], lv["reg"]] + ([lv["iou"]] if has_iou else [])
        sizes = [[(t.numel() + 3) // 4 * 4 for t in g] for g in groups]
        flat = guard.alloc(4 * sum(sum(g) for g in sizes), lv["cls"][0].device).view(torch.float32)
        out, o = [], 0
        for g, sz in zip(groups, sizes):
            views = []
            for t, n in zip(g, sz):
                views.append(flat[o:o + t.numel()].view(t.shape))
                o += n
            out.append(views)
        return dict(cls=out[0], reg=out[1], iou=out[2] if has_iou else None, flat=flat)

    ev._workspace_for = workspace_for
    ev._alloc_grads = alloc_grads


@pytest.mark.parametrize("case", ["small", "ragged_levels", "crowded"])
def test_loss_step_stays_inside_its_buffers(case):
    import paa_b200
    kw = dict(small=dict(seed=901, num_images=2, image_hw=(256, 320), gt_per_image=(2, 8)),
              ragged_levels=dict(seed=902, num_images=3, image_hw=(200, 360), gt_per_image=(1, 30)),
              crowded=dict(seed=903, num_images=1, image_hw=(320, 320), gt_per_image=300))[case]
    b = synthetic.make_batch(**kw)
    cls, reg, iou, targets, anchors = to_device_inputs(b)
    cfg = paa_b200.default_cfg()
    guard = Guarded()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    _guard_evaluator(ev, guard)
    losses, grads = ev.forward_backward(cls, reg, iou, targets, anchors)
    guard.check()
    want, want_grads = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg)).forward_backward(
        cls, reg, iou, targets, anchors)
    assert torch.equal(losses, want)
    for a, c in zip(grads["cls"] + grads["reg"] + grads["iou"], want_grads["cls"] + want_grads["reg"] + want_grads["iou"]):
        assert torch.equal(a, c)
    assert np.isfinite(losses.cpu().numpy()).all()


def test_postprocessing_stays_inside_its_workspace():
    import paa_b200
    ib = synthetic.make_inference_batch(seed=911, num_images=2, image_hw=(256, 320))
    cls, reg, iou, _, anchors = to_device_inputs(ib)
    cfg = paa_b200.default_cfg()
    pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
    guard = Guarded()

    def workspace_for(device, nbytes):
        ws = guard.alloc(nbytes + 256, device)
        pp._workspace = ws
        return ws

    pp._workspace_for = workspace_for
    out = pp(cls, reg, iou, anchors)
    guard.check()
    want = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))(cls, reg, iou, anchors)
    for a, c in zip(out, want):
        assert torch.equal(a.bbox, c.bbox) and torch.equal(a.get_field("scores"), c.get_field("scores"))
