"""Measurement aid: cycles per EM iteration by section (needs a -DPAA_PROFILE_GMM build, PAA_B200_LIB pointing at it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paa_b200
from paa_b200 import synthetic
from paa_b200.synthetic import to_device_inputs
b = synthetic.make_batch(seed=2000, num_images=16, image_hw=(800, 1333), gt_per_image=(1, 100))
cfg = paa_b200.default_cfg()
ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
ev.debug = True
cls, reg, iou, targets, anchors = to_device_inputs(b)
for _ in range(3):
    ev(cls, reg, iou, targets, anchors, None)
torch.cuda.synchronize()
g = ev.last_debug["gmm"].cpu().numpy()
cnt = ev.last_debug["cand_cnt"].cpu().numpy()
it = g[:, 6]
sel = it > 3
names = ["E-step (exp, rcp, products)", "reductions (prod + sum7)", "M-step to var", "log + sqrt/div to loop end", "next log-probs"]
cols = [2, 3, 4, 5, 7]
tot = 0
for nm, c in zip(names, cols):
    per = g[sel, c] / np.maximum(it[sel] - (1 if c == 7 else 0), 1)
    print("%-32s %7.0f cycles / iteration (median %.0f)" % (nm, per.mean(), np.median(per)))
    tot += per.mean()
print("sum %.0f; n_iter max %d; candidates max %d" % (tot, it.max(), cnt.max()))
for lo, hi in ((2, 16), (17, 32), (33, 64)):
    m = sel & (cnt >= lo) & (cnt <= hi)
    if m.any():
        print("n in [%d,%d]: %d GTs," % (lo, hi, m.sum()), " ".join("%.0f" % (g[m, c] / np.maximum(it[m], 1)).mean() for c in cols))
