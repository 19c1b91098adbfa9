"""Measurement of the operators next to the hot path: device anchor generation, boxlist_iou,
and the ATSS / RetinaNet / FCOS post-processors on the PAA kernels.  CUDA events, L2 flushed before every
timed call, median of 10.   python tools/bench_aux.py"""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
import paa_b200
from paa_b200 import synthetic
from paa_b200.anchor_generator import make_anchor_generator_paa
from paa_b200.inference import make_atss_postprocessor, make_retinanet_postprocessor, make_fcos_postprocessor

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    return statistics.median(ms)


out = {}
# anchors: 800x1333 grid, generated (cache bypassed) vs what the reference does on the host every forward
cfg = paa_b200.default_cfg()
gen = make_anchor_generator_paa(cfg)
grids = synthetic.level_grids(*synthetic.padded_size(800, 1333))
def gen_once():
    gen._grid_cache = {k: v for k, v in gen._grid_cache.items() if k[0] == "cells"}
    return gen.grid_anchors(grids, dev)
ms = timed(gen_once)
out["grid_anchors_800x1333"] = {"ms": ms, "anchors": 22400, "note": "5 launches (one per level), cache bypassed"}
t0 = time.perf_counter()
for _ in range(20):
    [synthetic.level_anchors(g, s) for g, s in zip(grids, synthetic.STRIDES)]
out["grid_anchors_800x1333"]["torch_cpu_ms"] = (time.perf_counter() - t0) / 20 * 1e3

# boxlist_iou: 100 GT x 22400 anchors and 5000 x 5000
for n1, n2 in ((100, 22400), (5000, 5000)):
    g = torch.Generator().manual_seed(n1)
    def boxes(n):
        xy = torch.rand((n, 2), generator=g) * 1000
        return torch.cat([xy, xy + torch.rand((n, 2), generator=g) * 300], 1)
    a, b = paa_b200.BoxList(boxes(n1).to(dev), (1333, 800)), paa_b200.BoxList(boxes(n2).to(dev), (1333, 800))
    ms = timed(lambda: paa_b200.boxlist_iou(a, b))
    out["boxlist_iou_%dx%d" % (n1, n2)] = {"ms": ms, "pairs_per_s": n1 * n2 / (ms / 1e3),
                                           "write_GBps": n1 * n2 * 4 / (ms / 1e3) / 1e9}

# post-processor flavours, 8 images of 800x1333
b = synthetic.make_inference_batch(seed=4000, num_images=8, image_hw=(800, 1333), candidates_per_level=4000)
cls = [t.to(dev) for t in b.box_cls]; reg = [t.to(dev) for t in b.box_regression]; ctr = [t.to(dev) for t in b.iou_pred]
anc = [a.to(dev) for a in b.anchors]
anchors = [[paa_b200.BoxList(a, b.image_sizes[i]) for a in anc] for i in range(8)]
atss_cfg = NS(MODEL=NS(ATSS=NS(INFERENCE_TH=0.05, PRE_NMS_TOP_N=1000, NMS_TH=0.6, NUM_CLASSES=81, REGRESSION_TYPE="BOX")),
              TEST=NS(DETECTIONS_PER_IMG=100, BBOX_AUG=NS(ENABLED=False, VOTE=False)))
pp = make_atss_postprocessor(atss_cfg, paa_b200.BoxCoder(atss_cfg))
ms = timed(lambda: pp.run_device(cls, reg, ctr, anchors))
out["atss_post_8img"] = {"ms": ms, "images_per_s": 8 / (ms / 1e3)}
fc_cfg = NS(MODEL=NS(FCOS=NS(INFERENCE_TH=0.05, PRE_NMS_TOP_N=1000, NMS_TH=0.6, NUM_CLASSES=81)),
            TEST=NS(DETECTIONS_PER_IMG=100, BBOX_AUG=NS(ENABLED=False)))
fp = make_fcos_postprocessor(fc_cfg)
locs = [l.to(dev) for l in synthetic.fcos_locations(b.grids)]
pts = [torch.cat([l, l], 1).contiguous() for l in locs]
fanchors = [[paa_b200.BoxList(p, b.image_sizes[i]) for p in pts] for i in range(8)]
dist = [torch.exp(t * 0.5) * 8.0 * s for t, s in zip(reg, synthetic.STRIDES)]
ms = timed(lambda: fp.run_device(cls, dist, ctr, fanchors))
out["fcos_post_8img"] = {"ms": ms, "images_per_s": 8 / (ms / 1e3)}
rb = synthetic.make_retinanet_batch(seed=4300, num_images=8, image_hw=(800, 1333), cls_mean=-5.0, cls_std=1.2)
rcls = [t.to(dev) for t in rb.box_cls]; rreg = [t.to(dev) for t in rb.box_regression]
ranc = [a.to(dev) for a in rb.anchors]
ranchors = [[paa_b200.BoxList(a, rb.image_sizes[i]) for a in ranc] for i in range(8)]
rn_cfg = NS(MODEL=NS(RETINANET=NS(INFERENCE_TH=0.05, PRE_NMS_TOP_N=1000, NMS_TH=0.4, NUM_CLASSES=81)),
            TEST=NS(DETECTIONS_PER_IMG=100))
rp = make_retinanet_postprocessor(rn_cfg, None)
ms = timed(lambda: rp.run_device(rcls, rreg, None, ranchors))
logit_bytes = sum(t.numel() * 4 for t in rcls)
out["retinanet_post_8img_9anchors"] = {"ms": ms, "images_per_s": 8 / (ms / 1e3), "logit_MB": logit_bytes / 1e6}
# ATSS training step (assignment + losses + gradients) on the C2 shape, CUDA-graph replay like bench.py
from paa_b200.synthetic import to_device_inputs
tb_ = synthetic.make_batch(seed=2000, num_images=16, image_hw=(800, 1333), gt_per_image=(1, 100))
acfg = NS(MODEL=NS(ATSS=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, TOPK=9, REG_LOSS_WEIGHT=2.0, POSITIVE_TYPE="ATSS",
                           REGRESSION_TYPE="BOX")))
aev = paa_b200.make_atss_loss_evaluator(acfg, paa_b200.BoxCoder(acfg))
acls, areg, actr, atargets, aanchors = to_device_inputs(tb_, device=dev)
astep = lambda: aev.forward_backward(acls, areg, actr, atargets, aanchors)
for _ in range(3):
    astep()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    astep()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    astep()
ms = timed(lambda: graph.replay())
out["atss_loss_step_16img_800x1333"] = {"ms": ms, "images_per_s": 16 / (ms / 1e3),
                                        "note": "paa_atss_assign + paa_loss (forward + gradients), graph replay"}

# RetinaNet training step (IoU matching + Matcher labels + focal / smooth-L1 + gradients), 9 anchors per location
rb = synthetic.make_retinanet_batch(seed=2100, num_images=16, image_hw=(800, 1333), gt_per_image=(1, 100))
rcfg = NS(MODEL=NS(RETINANET=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4,
                                BBOX_REG_BETA=0.11, BBOX_REG_WEIGHT=4.0)))
rev = paa_b200.make_retinanet_loss_evaluator(rcfg, NS(weights=(10.0, 10.0, 5.0, 5.0)))
rcls, rreg, _, rtargets, ranchors = to_device_inputs(rb, device=dev)
rstep = lambda: rev.forward_backward(ranchors, rcls, rreg, rtargets)
for _ in range(3):
    rstep()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    rstep()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
rgraph = torch.cuda.CUDAGraph()
with torch.cuda.graph(rgraph):
    rstep()
ms = timed(lambda: rgraph.replay())
rbytes = sum(t.numel() * 4 for t in rcls)
out["retinanet_loss_step_16img_800x1333_9anchors"] = {
    "ms": ms, "images_per_s": 16 / (ms / 1e3), "logit_MB": rbytes / 1e6,
    "logit_read_plus_grad_write_GBps": 2 * rbytes / (ms / 1e3) / 1e9,
    "note": "paa_retinanet_assign + paa_loss (forward + gradients), graph replay; 201600 anchors x 80 classes per image"}

# FCOS training step (anchor-free assignment + focal / IOULoss / centerness + gradients) on the C2 shape
fb = synthetic.make_batch(seed=2200, num_images=16, image_hw=(800, 1333), gt_per_image=(1, 100), trained_like=False)
fb.box_regression = [(t.abs() * 40.0 + 1.0) for t in fb.box_regression]
fcfg = NS(MODEL=NS(FCOS=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FPN_STRIDES=[8, 16, 32, 64, 128],
                           CENTER_SAMPLING_RADIUS=1.5, IOU_LOSS_TYPE="giou", NORM_REG_TARGETS=True)))
fev = paa_b200.make_fcos_loss_evaluator(fcfg)
fcls, freg, fctr, ftargets, _ = to_device_inputs(fb, device=dev)
flocs = [p.to(dev) for p in synthetic.fcos_locations(fb.grids)]
fstep = lambda: fev.forward_backward(flocs, fcls, freg, fctr, ftargets)
for _ in range(3):
    fstep()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    fstep()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
fgraph = torch.cuda.CUDAGraph()
with torch.cuda.graph(fgraph):
    fstep()
ms = timed(lambda: fgraph.replay())
out["fcos_loss_step_16img_800x1333"] = {"ms": ms, "images_per_s": 16 / (ms / 1e3),
                                        "note": "paa_fcos_assign + paa_loss (forward + gradients), graph replay; "
                                                "centre sampling 1.5, GIoU loss, normalised targets"}

# TTA merging: 14 augmentations x ~100 detections of one image, 20 classes present
from paa_b200.bbox_aug_vote import merge_result_from_multi_scales
g = torch.Generator().manual_seed(9)
n_obj, n_aug = 100, 14
ctr = torch.rand((n_obj, 2), generator=g) * 900
size = 40 + torch.rand((n_obj, 2), generator=g) * 200
lab = torch.randint(1, 21, (n_obj,), generator=g)
which = torch.arange(n_obj).repeat(n_aug)
xy = ctr[which] + torch.randn((n_obj * n_aug, 2), generator=g) * 4
wh = size[which] * (1 + 0.05 * torch.randn((n_obj * n_aug, 2), generator=g))
tb = torch.cat([xy, xy + wh], 1)
ts = torch.randperm(n_obj * n_aug, generator=g).float() / (n_obj * n_aug) * 0.9 + 0.06
tl = lab[which]
bl = paa_b200.BoxList(tb.to(dev), (1333, 800)); bl.add_field("scores", ts.to(dev)); bl.add_field("labels", tl.to(dev))
tcfg = NS(MODEL=NS(RETINANET=NS(NUM_CLASSES=81, INFERENCE_TH=0.05), ATSS=NS(NMS_TH=0.6, PRE_NMS_TOP_N=1000)))
for kind in ("vote", "soft-vote"):
    ms = timed(lambda: merge_result_from_multi_scales([bl], tcfg, kind, 0.66))
    out["tta_merge_%s_1400_boxes" % kind.replace("-", "_")] = {"ms": ms,
                                                              "note": "includes the host read of the result count; "
                                                                      "CPU numbers: tests/tta_cpu_baseline.py"}
print(json.dumps(out))
