"""Measurement aid: where each kernel of the loss step sits on the GPU's %globaltimer inside a CUDA-graph replay.

Event pairs serialise the programmatic dependent launches they are meant to observe; this reads timestamps that the
kernels of a -DPAA_TRACE build leave themselves (common.cuh: TraceScope).  Usage (GPU box):

    python tools/step_trace.py [--images 16] [--config C2] [--replays 20]

builds paa_b200/libpaa_b200_trace.so if needed, re-executes itself with PAA_B200_LIB pointing at it, and prints per
kernel: first block start, first block past its dependency wait, last block start, last block end (us, relative to the
start of the step's first kernel; median over the replays).
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "paa_b200", "libpaa_b200_trace.so")
NAMES = ["prep_step", "iou_match", "match_score", "select_gmm", "bulk_focal", "positive_list", "gmm_fit_only",
         "pos_at_wait"]       # slot 7: positive_list_kernel's blocks arriving at their dependency wait (first / last)
POST_FIRST = 8
POST_NAMES = ["post_candidates", "post_threshold", "post_filter", "post_select", "post_rank", "post_group",
              "post_class_rank", "post_segments", "post_nms_runs", "post_nms_scan", "post_finish", "post_vote",
              "post_nms_mask"]
SLOTS = 24


def build_trace_lib():
    from paa_b200 import build as b
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-DPAA_TRACE"] + [os.path.join(b.CSRC, s) for s in b.SOURCES] + ["-o", TRACE_LIB]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return TRACE_LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--gt", type=int, nargs=2, default=[1, 100])
    ap.add_argument("--hw", type=int, nargs=2, default=[800, 1333], help="image height and width of the loss batch")
    ap.add_argument("--replays", type=int, default=20)
    ap.add_argument("--build-only", action="store_true")
    ap.add_argument("--post", action="store_true", help="trace the NMS + voting step (C4 shapes) instead of the loss step")
    args = ap.parse_args()
    if args.build_only:
        print(build_trace_lib())
        return
    if os.environ.get("PAA_B200_LIB") != TRACE_LIB:
        if not os.path.exists(TRACE_LIB):
            build_trace_lib()
        env = dict(os.environ, PAA_B200_LIB=TRACE_LIB)
        sys.exit(subprocess.call([sys.executable] + sys.argv, env=env))

    import numpy as np
    import torch
    import paa_b200
    from paa_b200 import _lib, synthetic
    from paa_b200.synthetic import to_device_inputs
    lib = _lib.load()
    raw = ctypes.CDLL(TRACE_LIB)
    raw.paa_trace_set.argtypes = [ctypes.c_void_p]
    raw.paa_trace_set.restype = ctypes.c_int
    trace = torch.zeros(SLOTS * 4, dtype=torch.int64, device="cuda")
    assert raw.paa_trace_set(ctypes.c_void_p(trace.data_ptr())) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cfg = paa_b200.default_cfg()
    if args.post:
        from paa_b200.structures import BoxList
        b = synthetic.make_inference_batch(seed=4000, num_images=args.images, image_hw=(800, 1333), candidates_per_level=4000)
        pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
        cls = [t.cuda() for t in b.box_cls]
        reg = [t.cuda() for t in b.box_regression]
        iou = [t.cuda() for t in b.iou_pred]
        anc = [a.cuda() for a in b.anchors]
        anchors = [[BoxList(a, b.image_sizes[i]) for a in anc] for i in range(args.images)]
        step = lambda: pp.run_device(cls, reg, iou, anchors)
        names, first = POST_NAMES, POST_FIRST
        what = "%d images (C4 shapes)" % args.images
    else:
        b = synthetic.make_batch(seed=2000, num_images=args.images, image_hw=tuple(args.hw), gt_per_image=tuple(args.gt))
        ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
        cls, reg, iou, targets, anchors = to_device_inputs(b)
        step = lambda: ev.forward_backward(cls, reg, iou, targets, anchors)
        names, first = NAMES, 0
        what = "%d images, %d GT" % (args.images, sum(len(x) for x in b.gt_boxes))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    init = torch.zeros(SLOTS, 4, dtype=torch.int64)
    init[:, 0] = init[:, 1] = (1 << 62)
    init = init.reshape(-1).cuda()
    rows = []
    for _ in range(args.replays):
        flush.fill_(1)
        trace.copy_(init)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        t = trace.cpu().numpy().reshape(SLOTS, 4)[first:first + len(names)].astype(np.float64)
        ran = t[:, 2] > 0
        t0 = t[ran, 0].min()
        rows.append(np.where(ran[:, None], (t - t0) / 1000.0, np.nan))
    med = np.nanmedian(np.stack(rows), axis=0)
    print("%s; us relative to the first block of the step's first kernel (median of %d graph replays, L2 flushed)"
          % (what, args.replays))
    print("%-16s %10s %12s %12s %10s %8s" % ("kernel", "first start", "first waited", "last start", "last end", "span"))
    end = 0.0
    for i, nm in enumerate(names):
        s0, w, e, ls = med[i]
        if not np.isfinite(e):
            continue
        w = w if w < 1e9 else float("nan")
        print("%-16s %10.1f %12.1f %12.1f %10.1f %8.1f" % (nm, s0, w, ls, e, e - (w if np.isfinite(w) else s0)))
        end = max(end, e)
    print("step: %.1f us" % end)


if __name__ == "__main__":
    main()
