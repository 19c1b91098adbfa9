"""Debug aid: run one synthetic batch on the GPU and through the oracle, print every GT whose positive set
differs (candidate lists, n_iter, mixture parameters, score gaps).  python tests/debug_gt.py c3|c1|c5 (test infrastructure: it runs the oracle)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paa_b200
from paa_b200 import synthetic
from oracle import paa_oracle, gmm_oracle
from tests.helpers import to_device_inputs, topk_tie_exempt, gmm_tie_exempt

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    b = synthetic.make_batch(seed=3000, num_images=1, image_hw=(1333, 1333), gt_per_image=500)
elif which == "c5":
    b = synthetic.make_batch(seed=5000, num_images=3, image_hw=(0, 0), gt_per_image=(1, 30),
                             per_image_hw=synthetic.multiscale_hw(5000, 3))
else:
    b = synthetic.make_batch(seed=1000, num_images=2, image_hw=(800, 1333), gt_per_image=20)
ref_losses, ref_grads, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes,
                                                        b.gt_labels, b.anchors, with_grad=False)
cfg = paa_b200.default_cfg()
ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
ev.debug = True
cls, reg, iou, targets, anchors = to_device_inputs(b)
ev(cls, reg, iou, targets, anchors, None)
torch.cuda.synchronize()
d = ev.last_debug
got = d["paa_labels"].cpu().numpy()
exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
diff = sorted({(int(i), int(asg.matched_idx[i][a])) for i, a in zip(*np.nonzero(got != asg.paa_labels.numpy()))})
print("diff", diff, "not exempt", [x for x in diff if x not in exempt])
off = d["gt_offsets"]
gm = d["gmm"].cpu().numpy(); cand = d["cand_idx"].cpu().numpy(); cnt = d["cand_cnt"].cpu().numpy()
npos = d["num_pos"].cpu().numpy(); cl = d["combined_loss"].cpu().numpy()
n_it_diff = 0
for i, recs in enumerate(asg.gmm_records):
    for r in recs:
        gi = off[i] + r["gt"]
        if r.get("fit") is not None and int(gm[gi, 6]) != int(r["fit"]["n_iter"]):
            n_it_diff += 1
print("GTs with different n_iter:", n_it_diff)
for (i, g) in diff:
    r = asg.gmm_records[i][g]
    gi = off[i] + g
    print("=== image %d gt %d exempt=%s  n=%d  oracle n_pos=%d  gpu n_cand=%d n_pos=%d" % (
        i, g, (i, g) in exempt, r["n"], r["n_pos"], cnt[gi], npos[gi]))
    same_c = np.array_equal(r["sorted_idx"], cand[gi, :cnt[gi]])
    print("  candidate lists equal:", same_c)
    if not same_c:
        print("  oracle idx", r["sorted_idx"].tolist())
        print("  gpu    idx", cand[gi, :cnt[gi]].tolist())
    fit = r["fit"]
    if fit is None:
        continue
    print("  oracle n_iter %d  w %s mu %s var %s" % (fit["n_iter"], fit["weights"], fit["means"], fit["variances"]))
    print("  gpu    n_iter %d  w %s mu %s var %s" % (gm[gi, 6], gm[gi, 0:2], gm[gi, 2:4], gm[gi, 4:6]))
    fg = fit["components"] == 0
    s = fit["scores"]
    print("  oracle comps", fit["components"].tolist())
    print("  oracle scores", np.array2string(s, precision=6))
    xs = r["sorted_loss"]; xg = cl[i][cand[gi, :cnt[gi]]]
    print("  loss oracle", np.array2string(xs, precision=6)); print("  loss gpu   ", np.array2string(xg, precision=6))
    f2 = gmm_oracle.fit_two_component(xg.reshape(-1, 1).astype(np.float32), impl="numpy")
    print("  numpy-oracle on GPU losses: n_iter %d n_pos %d" % (f2["n_iter"], gmm_oracle.positive_prefix_length(f2)))
