"""Per-kernel share of the RetinaNet training step (IoU matching / final loss) through the library's per-launch
CUDA-event timers, L2 flushed before every step.   python tools/retina_profile.py [num_images]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
import paa_b200
from paa_b200 import _lib, synthetic

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
b = synthetic.make_retinanet_batch(seed=2100, num_images=n_img, image_hw=(800, 1333), gt_per_image=(1, 100))
cfg = NS(MODEL=NS(RETINANET=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4,
                               BBOX_REG_BETA=0.11, BBOX_REG_WEIGHT=4.0)))
ev = paa_b200.make_retinanet_loss_evaluator(cfg, NS(weights=(10.0, 10.0, 5.0, 5.0)))
cls, reg, _, targets, anchors = synthetic.to_device_inputs(b, device=dev)
step = lambda: ev.forward_backward(anchors, cls, reg, targets)
for _ in range(3):
    step()
torch.cuda.synchronize()
lib = _lib.load()
for name in ("pass1", "final_loss"):
    lib.paa_kernel_timing_begin(_lib.KERNEL_IDS[name])
    for _ in range(5):
        flush.zero_()
        step()
    ms, n = ctypes.c_float(0), ctypes.c_int32(0)
    lib.paa_kernel_timing_end(ctypes.byref(ms), ctypes.byref(n))
    print("%-12s %.1f us per launch (%d launches)" % (name, 1000.0 * ms.value / max(1, n.value), n.value))
ts = []
for _ in range(10):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); step(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
print("eager step median %.1f us" % (1000 * ts[len(ts) // 2]))
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
ts = []
for _ in range(10):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); graph.replay(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
print("graph replay median %.1f us" % (1000 * ts[len(ts) // 2]))
