"""Measurement aid: per-GT cycle counts of select_gmm_kernel (needs probes/libprof_gmm.so built with
-DPAA_PROFILE_GMM and PAA_B200_LIB pointing at it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paa_b200
from paa_b200 import synthetic
from paa_b200.synthetic import to_device_inputs
if os.environ.get("GMM_PROFILE_SHAPE", "C2") == "C3":        # the dense-crowd shape, throughput-bound
    b = synthetic.make_batch(seed=3000, num_images=int(os.environ.get("GMM_PROFILE_IMAGES", "32")),
                             image_hw=(1333, 1333), gt_per_image=500)
else:
    b = synthetic.make_batch(seed=2000, num_images=16, image_hw=(800, 1333), gt_per_image=(1, 100))
cfg = paa_b200.default_cfg()
ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
ev.debug = True
cls, reg, iou, targets, anchors = to_device_inputs(b)
for _ in range(3):
    ev(cls, reg, iou, targets, anchors, None)
torch.cuda.synchronize()
ev.debug = False
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    ev(cls, reg, iou, targets, anchors, None)
t1.record(); torch.cuda.synchronize()
print("eager step (no debug outputs): %.1f us" % (t0.elapsed_time(t1) * 200))
g = ev.last_debug["gmm"].cpu().numpy()
cnt = ev.last_debug["cand_cnt"].cpu().numpy()
scan, em, it = g[:, 0], g[:, 1], g[:, 6]
mhz = 1965.0
print("GTs", len(scan), "cand mean/max", cnt.mean(), cnt.max())
print("scan+sort us: mean %.1f p50 %.1f p95 %.1f max %.1f" % tuple(np.array([scan.mean(), np.percentile(scan, 50), np.percentile(scan, 95), scan.max()]) / mhz))
print("EM us:        mean %.1f p50 %.1f p95 %.1f max %.1f" % tuple(np.array([em.mean(), np.percentile(em, 50), np.percentile(em, 95), em.max()]) / mhz))
print("n_iter: mean %.1f p50 %d p95 %d max %d" % (it.mean(), np.percentile(it, 50), np.percentile(it, 95), it.max()))
sel = it > 0
k, c = np.polyfit(it[sel], em[sel], 1)
print("EM cycles ~= %.0f * n_iter + %.0f  (%.2f us per iteration)" % (k, c, k / mhz))
print("slowest GT: total %.1f us (scan %.1f, em %.1f, n_iter %d)" % ((scan + em).max() / mhz, scan[(scan + em).argmax()] / mhz, em[(scan + em).argmax()] / mhz, it[(scan + em).argmax()]))
