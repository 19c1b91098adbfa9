"""NMS + score-voting images/sec (second half of BASELINE.json's metric), config C4 of SURVEY.md 8d:
a batch of 64 images of 800x1333, >= 1000 candidates per level above 0.05, 80 classes, NMS 0.6, score
voting on, 100 detections per image.  The headline number runs the whole 64-image batch on one GPU (it
fits); the 8-image figure (one rank's share when the batch is spread over 8 GPUs) and the dense stress
variant are reported beside it.  Imported by bench.py (rank 0) and runnable on its own:
python bench_post.py
"""
import ctypes
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

POST_KERNELS = ("post_candidates", "post_filter", "post_select", "post_rank", "post_nms_mask", "post_nms_scan",
                "post_finish", "post_vote")


def measure_post(dev, steps=10, warmup=3, images=64, with_cpu=True, candidates_per_level=4000, with_dense=True,
                 with_share=True):
    """candidates_per_level: expected (location, class) pairs above 0.05 per level (detector-like
    sparsity, every level still has > 1000 so the per-level cap of 1000 is active everywhere);
    None = the dense variant where ~49 % of ALL logits are candidates (stress case, reported too)."""
    import paa_b200
    from paa_b200 import _lib, synthetic
    from paa_b200.structures import BoxList
    lib = _lib.load()
    batch = synthetic.make_inference_batch(seed=4000, num_images=images, image_hw=(800, 1333),
                                           candidates_per_level=candidates_per_level)
    cfg = paa_b200.default_cfg()
    pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
    cls = [t.to(dev) for t in batch.box_cls]
    reg = [t.to(dev) for t in batch.box_regression]
    iou = [t.to(dev) for t in batch.iou_pred]
    anc = [a.to(dev) for a in batch.anchors]
    anchors = [[BoxList(a, batch.image_sizes[i]) for a in anc] for i in range(images)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        return pp.run_device(cls, reg, iou, anchors)

    for _ in range(max(3, warmup)):
        out = step()
    torch.cuda.synchronize()
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("post: CUDA graph capture failed (%s); timing eager launches\n" % (e,))
        graph = None
        torch.cuda.synchronize()
    ms = []
    for k in range(warmup + steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if graph is not None:
            graph.replay()
        else:
            out = step()
        e.record()
        torch.cuda.synchronize()
        if k >= warmup:
            ms.append(s.elapsed_time(e))
    per_kernel = {}
    for name in POST_KERNELS:
        lib.paa_kernel_timing_begin(_lib.KERNEL_IDS[name])
        for _ in range(3):
            flush.zero_()
            step()
        t, n = ctypes.c_float(0), ctypes.c_int32(0)
        lib.paa_kernel_timing_end(ctypes.byref(t), ctypes.byref(n))
        per_kernel[name + "_us"] = 1000.0 * t.value / max(1, n.value)
    med = statistics.median(ms)
    A = batch.num_anchors
    res = {"metric": "PAA NMS+voting images/sec", "value": images / (med / 1000.0), "unit": "images/s",
           "images_per_gpu": images, "ms_per_step": med, "per_kernel_us": per_kernel,
           "detections": [int(c) for c in out[3].tolist()],
           "config": "C4: 800x1333, 1000 pre-NMS candidates/level, 80 classes, NMS 0.6, voting, 100 dets/img; "
                     + ("~%d candidates above 0.05 per level" % candidates_per_level if candidates_per_level
                        else "dense: ~49% of all logits above 0.05"),
           "launch": "CUDA graph replay" if graph is not None else "eager",
           # candidates kernel: one read of logits + regression + iou_pred per anchor
           "candidates_kernel_GBps": (A * (4 * 80 + 4) * images / 1e9) /
                                     (per_kernel["post_candidates_us"] / 1e6) if per_kernel["post_candidates_us"] else None}
    if with_cpu:
        from oracle import post_oracle
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        post_oracle.postprocess([t[:1] for t in batch.box_cls], [t[:1] for t in batch.box_regression],
                                [t[:1] for t in batch.iou_pred], batch.anchors, batch.image_sizes[:1])
        res["cpu_baseline"] = {"value": 1.0 / (time.perf_counter() - t0), "unit": "images/s",
                               "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "1 image of the batch, one run"}
    del cls, reg, iou, anc, anchors, batch, graph, out
    torch.cuda.empty_cache()
    keep = ("value", "unit", "images_per_gpu", "ms_per_step", "per_kernel_us", "config")
    if with_share and images > 8:
        d = measure_post(dev, steps=steps, warmup=warmup, images=8, with_cpu=False,
                         candidates_per_level=candidates_per_level, with_dense=False, with_share=False)
        res["one_rank_share_of_8"] = {k: d[k] for k in keep}
    if with_dense and candidates_per_level is not None:
        d = measure_post(dev, steps=max(3, steps // 2), warmup=warmup, images=8, with_cpu=False,
                         candidates_per_level=None, with_dense=False, with_share=False)
        res["dense_variant"] = {k: d[k] for k in keep}
    return res


if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(json.dumps(measure_post(torch.device("cuda", 0))))
