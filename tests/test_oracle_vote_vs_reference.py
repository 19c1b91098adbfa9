"""CPU, only where the reference tree is mounted: oracle/vote_oracle.py against the reference's own bbox_vote /
soft_bbox_vote (paa_core/engine/bbox_aug_vote.py:198-310), bit for bit.  The reference functions end with
``.cuda()``; that call is patched to the identity for this test, and the modules bbox_aug_vote.py imports
but the two functions do not use (yacs config, data transforms, image lists) are stubbed."""
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_shim, vote_oracle

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def _load_reference_vote(monkeypatch):
    ref = ref_shim.load_reference()
    ns = types.SimpleNamespace
    cfgmod = types.ModuleType("paa_core.config")
    cfgmod.cfg = ns(MODEL=ns(RETINANET=ns(INFERENCE_TH=0.05, NUM_CLASSES=81), ATSS=ns(NMS_TH=0.6, PRE_NMS_TOP_N=1000)))
    monkeypatch.setitem(sys.modules, "paa_core.config", cfgmod)
    data = types.ModuleType("paa_core.data")
    data.transforms = types.ModuleType("paa_core.data.transforms")
    monkeypatch.setitem(sys.modules, "paa_core.data", data)
    monkeypatch.setitem(sys.modules, "paa_core.data.transforms", data.transforms)
    il = types.ModuleType("paa_core.structures.image_list")
    il.to_image_list = lambda *a, **k: None
    monkeypatch.setitem(sys.modules, "paa_core.structures.image_list", il)
    layers = types.ModuleType("paa_core.layers")
    layers.nms = ref._C.nms
    monkeypatch.setitem(sys.modules, "paa_core.layers", layers)
    monkeypatch.delitem(sys.modules, "paa_core.engine.bbox_aug_vote", raising=False)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    from paa_core.engine import bbox_aug_vote
    return bbox_aug_vote


def test_vote_oracle_matches_reference_functions(monkeypatch):
    rv = _load_reference_vote(monkeypatch)
    g = torch.Generator().manual_seed(5)
    for trial in range(30):
        n = int(torch.randint(2, 200, (1,), generator=g))
        ctr = torch.rand((max(2, n // 6), 2), generator=g) * 300
        which = torch.randint(0, ctr.shape[0], (n,), generator=g)
        xy = ctr[which] + torch.randn((n, 2), generator=g) * 6
        wh = 60 + torch.rand((n, 2), generator=g) * 30
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.randperm(n, generator=g).float() / n * 0.9 + 0.05
        for soft in (False, True):
            fn = rv.soft_bbox_vote if soft else rv.bbox_vote
            rb, rs = fn(boxes, scores, 0.66)
            ob, os_ = vote_oracle.vote_class(boxes.numpy(), scores.numpy(), 0.66, soft=soft)
            assert np.array_equal(rs.numpy(), os_), (trial, soft)
            assert np.array_equal(rb.numpy(), ob), (trial, soft)
