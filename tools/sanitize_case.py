"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): one PAA assign+loss step with gradients (eager
and through the graph-mode path's device-resident GT ranges), one post-processing call, the stand-alone NMS and focal
layer.  `compute-sanitizer --tool memcheck python tools/sanitize_case.py`"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paa_b200
from paa_b200 import synthetic
from paa_b200.synthetic import to_device_inputs

torch.cuda.set_device(0)
cfg = paa_b200.default_cfg()
coder = paa_b200.BoxCoder(cfg)
b = synthetic.make_batch(seed=77, num_images=2, image_hw=(256, 320), gt_per_image=(2, 8))
cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=True)
ev = paa_b200.make_paa_loss_evaluator(cfg, coder)
losses = ev(cls, reg, iou, targets, anchors, None)
sum(losses).backward()
torch.cuda.synchronize()
print("loss", [float(x.detach()) for x in losses])
big = synthetic.make_batch(seed=78, num_images=1, image_hw=(256, 320), gt_per_image=140)      # the GT-list split
c2, r2, i2, t2, a2 = to_device_inputs(big)
print("crowded", ev.forward_backward(c2, r2, i2, t2, a2)[0].tolist())
evg = paa_b200.make_paa_loss_evaluator(cfg, coder)
evg.use_graph = True
cls_n, reg_n, iou_n, _, _ = to_device_inputs(b)
for _ in range(2):
    lg, _ = evg.forward_backward(cls_n, reg_n, iou_n, targets, anchors)
torch.cuda.synchronize()
print("graph", lg.tolist())
ib = synthetic.make_inference_batch(seed=79, num_images=2, image_hw=(256, 320))
pc, pr, pi, _, pa = to_device_inputs(ib)
pp = paa_b200.make_paa_postprocessor(cfg, coder)
out = pp(pc, pr, pi, pa)
print("post", [len(o) for o in out])
from paa_b200.inference import ml_nms
g = torch.Generator().manual_seed(1)
boxes = torch.rand(300, 4, generator=g) * 100
boxes[:, 2:] += boxes[:, :2]
keep = ml_nms(boxes.cuda(), torch.rand(300, generator=g).cuda(), torch.randint(1, 4, (300,), generator=g).float().cuda(), 0.5)
print("nms", int(keep.numel()))
from paa_b200.layers import sigmoid_focal_loss_cuda
x = torch.randn(100, 80, device="cuda", requires_grad=True)
t = torch.randint(0, 81, (100,), dtype=torch.int32, device="cuda")
sigmoid_focal_loss_cuda(x, t, 2.0, 0.25).sum().backward()
torch.cuda.synchronize()
print("SANITIZE_CASE_DONE")
