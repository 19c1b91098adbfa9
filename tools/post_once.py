"""Measurement aid: a few eager post-processing steps on the C4 shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paa_b200
from paa_b200 import synthetic
from paa_b200.structures import BoxList
dev = torch.device("cuda", 0)
images = int(sys.argv[1]) if len(sys.argv) > 1 else 8
batch = synthetic.make_inference_batch(seed=4000, num_images=images, image_hw=(800, 1333), candidates_per_level=4000)
cfg = paa_b200.default_cfg()
pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
cls = [t.to(dev) for t in batch.box_cls]; reg = [t.to(dev) for t in batch.box_regression]
iou = [t.to(dev) for t in batch.iou_pred]; anc = [a.to(dev) for a in batch.anchors]
anchors = [[BoxList(a, batch.image_sizes[i]) for a in anc] for i in range(images)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    out = pp.run_device(cls, reg, iou, anchors)
torch.cuda.synchronize()
print("ok", [int(c) for c in out[3].tolist()])
