"""Measurement aid: fused assign+loss step (CUDA-graph replay, L2 flushed) on the other BASELINE configs.
python tools/bench_config.py c3|c5|c2"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paa_b200
from paa_b200 import synthetic
from paa_b200.synthetic import to_device_inputs

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":      # dense crowd: 1333x1333, 500 GT/img, 4 images per GPU (batch 32 over 8)
    b = synthetic.make_batch(seed=3000, num_images=4, image_hw=(1333, 1333), gt_per_image=500)
elif which == "c5":    # multi-scale: 8 images per GPU (batch 64 over 8)
    b = synthetic.make_batch(seed=5000, num_images=8, image_hw=(0, 0), gt_per_image=(1, 100),
                             per_image_hw=synthetic.multiscale_hw(5000, 8))
else:
    b = synthetic.make_batch(seed=2000, num_images=16, image_hw=(800, 1333), gt_per_image=(1, 100))
dev = torch.device("cuda", 0)
cfg = paa_b200.default_cfg()
ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
cls, reg, iou, targets, anchors = to_device_inputs(b, device=dev)
step = lambda: ev.forward_backward(cls, reg, iou, targets, anchors)
for _ in range(3):
    step()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for k in range(25):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    if k >= 5:
        ms.append(s.elapsed_time(e))
med = statistics.median(ms)
import ctypes
from paa_b200 import _lib
lib = _lib.load()
per = {}
for name in ("pass1", "match_score", "select_gmm", "final_loss"):
    lib.paa_kernel_timing_begin(_lib.KERNEL_IDS[name])
    for _ in range(3):
        flush.zero_(); step()
    t, n = ctypes.c_float(0), ctypes.c_int32(0)
    lib.paa_kernel_timing_end(ctypes.byref(t), ctypes.byref(n))
    per[name] = round(1000.0 * t.value / max(1, n.value), 1)
n_gt = sum(int(x.shape[0]) for x in b.gt_boxes)
print(which, "images", b.num_images, "anchors", b.num_anchors, "GT", n_gt, "ms/step %.4f" % med,
      "images/s %.0f" % (b.num_images / med * 1e3), per)
