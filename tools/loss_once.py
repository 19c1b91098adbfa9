"""Measurement aid: a few eager assign+loss steps on the C2 shape (for ncu captures).
    python tools/loss_once.py [images] [nchw|nhwc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paa_b200
from paa_b200 import synthetic
dev = torch.device("cuda", 0)
images = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nhwc = len(sys.argv) > 2 and sys.argv[2] == "nhwc"
b = synthetic.make_batch(seed=2000, num_images=images, image_hw=(800, 1333), gt_per_image=(1, 100))
cfg = paa_b200.default_cfg()
ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
cls, reg, iou, targets, anchors = synthetic.to_device_inputs(b, requires_grad=True, channels_last=nhwc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    losses, grads = ev.forward_backward(cls, reg, iou, targets, anchors)
torch.cuda.synchronize()
print("ok", [float(x) for x in losses], "nhwc" if nhwc else "nchw")
