#!/usr/bin/env python
"""Host-side cost of the reference-facing calls with device-resident inputs (what a training / inference
loop of the reference pays per call once it has switched to paa_b200).

    python tools/host_overhead.py [--images 16] [--calls 50] [--profile]

For the C2 batch it reports, per call of ``PAALossComputation.__call__`` + ``torch.autograd.grad``:
  host_us : wall-clock time the Python thread spends inside the call (no synchronisation inside)
  step_us : wall-clock per call of a back-to-back loop, synchronised at both ends (max(host, device))
and the same for ``PAAPostProcessor.forward`` (which returns BoxLists and therefore synchronises).
With --profile the 25 most expensive host functions (cProfile, cumulative) are listed.
"""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--calls", type=int, default=50)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    import paa_b200
    from paa_b200 import synthetic
    from paa_b200.structures import BoxList
    dev = torch.device("cuda", 0)
    batch = synthetic.make_batch(seed=2000, num_images=args.images, image_hw=(800, 1333), gt_per_image=(1, 100))
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    post = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
    d_anchor = [a.to(dev) for a in batch.anchors]
    anchors = [[BoxList(a, batch.image_sizes[i]) for a in d_anchor] for i in range(args.images)]
    targets = []
    for i in range(args.images):
        t = BoxList(batch.gt_boxes[i].to(dev), batch.image_sizes[i])
        t.add_field("labels", batch.gt_labels[i].to(dev))
        targets.append(t)
    cls = [t.to(dev).requires_grad_(True) for t in batch.box_cls]
    reg = [t.to(dev).requires_grad_(True) for t in batch.box_regression]
    iou = [t.to(dev).requires_grad_(True) for t in batch.iou_pred]

    def train_call():
        losses = ev(cls, reg, iou, targets, anchors, None)
        return torch.autograd.grad(losses[0] + losses[1] + losses[2], cls + reg + iou)

    def test_call():
        with torch.no_grad():
            return post(cls, reg, iou, anchors)

    out = {"images": args.images, "calls": args.calls}
    for name, fn in (("loss", train_call), ("post", test_call)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        host = []
        t_all = time.perf_counter()
        for _ in range(args.calls):
            t0 = time.perf_counter()
            fn()
            host.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t_all
        host.sort()
        out[name] = {"host_us_median": 1e6 * host[len(host) // 2], "host_us_min": 1e6 * host[0],
                     "step_us": 1e6 * t_all / args.calls}
        if args.profile:
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(args.calls):
                fn()
            pr.disable()
            torch.cuda.synchronize()
            s = io.StringIO()
            pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25)
            sys.stderr.write("==== %s ====\n%s\n" % (name, s.getvalue()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
