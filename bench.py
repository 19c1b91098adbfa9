#!/usr/bin/env python
"""bench.py -- PAA assign+loss images/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config C2 of SURVEY.md 8d, BASELINE.json configs[1]): the paa_R_50_FPN_1x training shape --
a batch of 16 images of 800x1333 (padded 800x1344 => 22 400 anchors on P3..P7), 80 classes, 1..100
ground-truth boxes per image, TOPK 9 -- synthetic "trained-like" head outputs (paa_b200/synthetic.py).
The whole 16-image batch runs on every GPU (weak scaling: N GPUs process 16*N images per step), and
each rank assigns its own images; with N > 1 the two loss normalisers cross ranks in one 2-element
NCCL all-reduce per step.

A step is the training-step semantics of the path: forward (assignment + three losses) AND the
gradients w.r.t. box_cls / box_regression / iou_pred.  `value` calls the evaluator's fused
`forward_backward` (capturable); `e2e` calls the reference-facing `PAALossComputation.__call__` +
`torch.autograd.grad`.  Both run the same kernels.

  value : images/s with inputs resident in HBM; the step is captured once into a CUDA graph and
          replayed; per-step CUDA events, L2 flushed between steps (not timed), max over ranks.
  e2e   : the same step called eagerly, with each step's inputs copied from pinned host memory and
          the three losses read back to the host inside the timed region.
  roofline : final_loss_kernel (reads every logit once, writes its gradient once), timed per launch
          with CUDA events inside the library during an eager pass of the same K steps.
  cold_clean_l2, eager_api_resident : side measurements -- the graph replay from an L2 whose flush left no
          dirty lines, and the reference-facing eager call on resident inputs (host-bound).
  cpu_baseline : the CPU port of the reference path (oracle/, torch CPU ops + scikit-learn) on a
          bounded sample of the same workload, rank 0 at N=1 only.

`--impl reference` times that CPU port alone (the reference's own Python cannot travel to the GPU box;
see DESIGN.md) and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "PAA assign+loss images/sec"
UNIT = "images/s"
IMAGE_HW = (800, 1333)
GT_RANGE = (1, 100)
SEED_BASE = 2000            # 1000 * config index (C2) + rank, SURVEY.md 8d
CPU_SAMPLE_IMAGES = 2       # bounded sample for the CPU arm


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-gpu", type=int, default=16)
    ap.add_argument("--no-graph", action="store_true", help="time the eager path instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="skip the NMS+voting side measurement")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local_rank, world


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (bench.py is one of the places allowed to run it)
# ------------------------------------------------------------------------------------------------
def cpu_port_images_per_sec(batch, n_images, repeats, warmup=1, time_budget_s=None):
    """Median time of `repeats` steps of the CPU port after `warmup` untimed ones; stops early (after at least one
    timed step) once `time_budget_s` of wall clock is spent, so that the arm stays bounded on a slow host."""
    from oracle import paa_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    sl = slice(0, n_images)
    args = ([t[sl] for t in batch.box_cls], [t[sl] for t in batch.box_regression],
            [t[sl] for t in batch.iou_pred], batch.gt_boxes[sl], batch.gt_labels[sl], batch.anchors)
    times = []
    begin = time.perf_counter()
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        paa_oracle.assign_and_loss(*args, with_grad=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time_budget_s is not None and times and time.perf_counter() - begin > time_budget_s:
            break
    return n_images / statistics.median(times), times


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (port), rank 0 only."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from paa_b200 import synthetic
    n = min(CPU_SAMPLE_IMAGES, args.images_per_gpu)
    batch = synthetic.make_batch(seed=SEED_BASE, num_images=n, image_hw=IMAGE_HW, gt_per_image=GT_RANGE)
    # K steps and W warm-up steps as asked, each a 2-image sample (~0.5-1.5 s of CPU work); a wall-clock budget
    # keeps the whole arm within a couple of minutes on any host, and the line reports the steps actually timed
    warm = max(1, min(args.warmup, 3))
    ips, times = cpu_port_images_per_sec(batch, n, max(1, args.steps), warm, time_budget_s=90.0)
    steps = len(times)
    cores = torch.get_num_threads()
    sample = "%d of the %d images/GPU of the workload per step (seed %d), fwd+bwd, %d timed steps (median)" % (
        n, args.images_per_gpu, SEED_BASE, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, note="CPU port of the reference path on a bounded sample"),
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, note=None):
    cfg = {"workload": "C2 paa_R_50_FPN_1x training shape: %d images/GPU of 800x1333 (22400 anchors, P3-P7), "
                       "80 classes, 1-100 GT/img, topk 9, forward+gradients" % args.images_per_gpu,
           "images_per_gpu": args.images_per_gpu, "global_batch": args.images_per_gpu * args.gpus,
           "anchors_per_image": 22400, "gt_per_image": list(GT_RANGE), "parallelism": "dp%d" % args.gpus,
           "l2": "flushed between timed steps (256 MiB write, untimed)"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in a background thread
    (the same counters nvidia-smi prints, without forking a process next to the timed launches);
    falls back to an `nvidia-smi -lms` child if pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.002):
        self.gpu_index = gpu_index
        self.period_s = period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: resolve the torch device through its UUID
            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:  # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _poll_nvml(self):
        nv = self.nvml
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(self.period_s)

    def _poll_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            p = [q.strip() for q in line.split(",")]
            if len(p) < 6:
                continue
            try:
                self.sm.append(float(p[0]))
                self.max_mhz = float(p[1])
            except ValueError:
                continue
            for name, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._poll_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def pin(ts):
    return [t.contiguous().pin_memory() for t in ts]


def run_ours(args):
    import paa_b200
    from paa_b200 import _lib, synthetic
    from paa_b200.structures import BoxList
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU path for the product")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    n_img = args.images_per_gpu
    batch = synthetic.make_batch(seed=SEED_BASE + rank, num_images=n_img, image_hw=IMAGE_HW,
                                 gt_per_image=GT_RANGE)
    L = len(batch.box_cls)
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    lib = _lib.load()

    # host (pinned) and device copies
    h_cls, h_reg, h_iou = pin(batch.box_cls), pin(batch.box_regression), pin(batch.iou_pred)
    h_gtb = [t.pin_memory() for t in batch.gt_boxes]
    h_gtl = [t.pin_memory() for t in batch.gt_labels]
    d_anchor = [a.to(dev) for a in batch.anchors]
    anchors = [[BoxList(a, batch.image_sizes[i]) for a in d_anchor] for i in range(n_img)]

    def targets_from(boxes, labels):
        out = []
        for i in range(n_img):
            t = BoxList(boxes[i], batch.image_sizes[i])
            t.add_field("labels", labels[i])
            out.append(t)
        return out

    d_cls = [t.to(dev).requires_grad_(True) for t in h_cls]
    d_reg = [t.to(dev).requires_grad_(True) for t in h_reg]
    d_iou = [t.to(dev).requires_grad_(True) for t in h_iou]
    d_targets = targets_from([t.to(dev) for t in h_gtb], [t.to(dev) for t in h_gtl])
    heads = d_cls + d_reg + d_iou

    def step_resident():
        # the fused entry point of the same evaluator: losses + gradients, no autograd bookkeeping, so the
        # whole step is capturable; identical kernels to `ev(...)` + backward (which the e2e leg times)
        return ev.forward_backward(d_cls, d_reg, d_iou, d_targets, anchors)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (eager) and graph capture ------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        out = step_resident()
    torch.cuda.synchronize()
    graph = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step_resident()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_out = step_resident()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 - report and fall back to eager timing
            sys.stderr.write("CUDA graph capture failed (%s); timing the eager path\n" % (e,))
            graph = None
            torch.cuda.synchronize()

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            step_resident()

    for _ in range(args.warmup):
        flush_buf.zero_()
        one_step()
    barrier()

    # ---- timed: K steps, per-step events, L2 flush between ----------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush_buf.zero_()
        starts[k].record()
        one_step()
        ends[k].record()
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = sum(step_ms)

    # ---- side measurement: the same K steps from a cold but CLEAN L2 ---------------------------------
    # The 256 MiB flush write leaves the L2 full of dirty lines whose write-back the first kernel of the step could
    # be paying for.  Here the write is followed by a 256 MiB read of a second buffer: the workload's data is gone
    # from L2 just the same, but the lines it evicts are clean.  Reported next to `value`, not instead of it
    # (measured on B200: 0.1551 against 0.1562 ms per step -- the headline does not hinge on the flush style).
    clean_ms = None
    if world == 1:                       # single-GPU runs only: the scaling runs carry nothing but the contract
        try:
            flush_rd = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
            c_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
            c_ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
            for k in range(args.steps):
                flush_buf.zero_()
                flush_rd.sum()
                c_starts[k].record()
                one_step()
                c_ends[k].record()
            torch.cuda.synchronize()
            clean_ms = sum(s.elapsed_time(e) for s, e in zip(c_starts, c_ends))
            del flush_rd
        except Exception as e:  # noqa: BLE001 - a side measurement must not take the bench line down
            sys.stderr.write("cold_clean_l2 side measurement failed: %s\n" % (e,))
            clean_ms = None

    # ---- roofline: the dominant kernel, per-launch CUDA events inside the library, eager pass -----
    lib.paa_kernel_timing_begin(_lib.KERNEL_IDS["final_loss"])
    for k in range(args.steps):
        flush_buf.zero_()
        step_resident()
    k_ms, k_n = ctypes.c_float(0), ctypes.c_int32(0)
    _lib.check(lib.paa_kernel_timing_end(ctypes.byref(k_ms), ctypes.byref(k_n)), "paa_kernel_timing_end")
    clocks = sampler.stop()

    # other kernels' share (one eager pass each; reported, not part of `value`)
    shares = {}
    for name in ("pass1", "match_score", "select_gmm"):
        lib.paa_kernel_timing_begin(_lib.KERNEL_IDS[name])
        for k in range(3):
            flush_buf.zero_()
            step_resident()
        ms, n = ctypes.c_float(0), ctypes.c_int32(0)
        lib.paa_kernel_timing_end(ctypes.byref(ms), ctypes.byref(n))
        shares[name + "_us"] = 1000.0 * ms.value / max(1, n.value)
    shares["final_loss_us"] = 1000.0 * k_ms.value / max(1, k_n.value)

    # ---- side measurement: the reference-facing call with RESIDENT inputs, eager ----------------------
    # `PAALossComputation.__call__` + autograd on the device tensors, back to back, wall clock between two
    # synchronisations: what a caller pays per step when nothing else keeps the GPU busy (host-bound).
    def step_eager_api():
        losses = ev(d_cls, d_reg, d_iou, d_targets, anchors, None)
        return torch.autograd.grad(losses[0] + losses[1] + losses[2], heads)

    eager_ms = None
    if world == 1:
        try:
            for _ in range(3):
                step_eager_api()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(args.steps):
                step_eager_api()
            torch.cuda.synchronize()
            eager_ms = 1000.0 * (time.perf_counter() - t0)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("eager_api_resident side measurement failed: %s\n" % (e,))
            eager_ms = None

    # ---- e2e: host buffers in, losses out, eager, public API -------------------------------------
    # The step's inputs live in pinned host memory, packed the way a collate function would leave them (one
    # block for the head outputs, one each for the boxes / labels of all images).  Every step copies them to
    # the device on a copy stream into one of two device buffers (so the copy of step k+1 overlaps the
    # kernels of step k), calls the reference-facing evaluator + autograd on views of that buffer, and
    # reads the three losses back into pinned memory.
    heads_h = h_cls + h_reg + h_iou
    sizes = [t.numel() for t in heads_h]
    h_pack = torch.empty(sum(sizes), dtype=torch.float32).pin_memory()
    o = 0
    for t, sz in zip(heads_h, sizes):
        h_pack[o:o + sz].copy_(t.reshape(-1))
        o += sz
    gt_counts = [int(t.shape[0]) for t in h_gtb]
    h_boxes = torch.cat(h_gtb, 0).contiguous().pin_memory()
    h_labels = torch.cat(h_gtl, 0).contiguous().pin_memory()
    h_losses = torch.empty(3, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in (h_pack, h_boxes, h_labels))
    copy_stream = torch.cuda.Stream()
    slots = [dict(pack=torch.empty_like(h_pack, device=dev), boxes=torch.empty_like(h_boxes, device=dev),
                  labels=torch.empty_like(h_labels, device=dev), copied=torch.cuda.Event(),
                  free=torch.cuda.Event()) for _ in range(2)]
    state = {"k": 0}

    def enqueue_copy(k):
        sl = slots[k % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])          # the step that last used this slot has finished
            sl["pack"].copy_(h_pack, non_blocking=True)
            sl["boxes"].copy_(h_boxes, non_blocking=True)
            sl["labels"].copy_(h_labels, non_blocking=True)
            sl["copied"].record(copy_stream)

    def step_e2e():
        k = state["k"]
        state["k"] = k + 1
        sl = slots[k % 2]
        enqueue_copy(k + 1)                              # next step's inputs travel while this step computes
        main = torch.cuda.current_stream()
        main.wait_event(sl["copied"])
        views, o = [], 0
        for t, sz in zip(heads_h, sizes):
            views.append(sl["pack"][o:o + sz].view(t.shape).requires_grad_(True))
            o += sz
        cls, reg, iou = views[:L], views[L:2 * L], views[2 * L:]
        tg = targets_from(list(sl["boxes"].split(gt_counts)), list(sl["labels"].split(gt_counts)))
        losses = ev(cls, reg, iou, tg, anchors, None)
        grads = torch.autograd.grad(losses[0] + losses[1] + losses[2], cls + reg + iou)
        h_losses.copy_(torch.stack([l.detach() for l in losses]), non_blocking=True)
        sl["free"].record(main)
        return grads

    for sl in slots:
        sl["free"].record(torch.cuda.current_stream())
    enqueue_copy(0)
    for _ in range(max(3, args.warmup)):
        step_e2e()
    barrier()
    e_start = torch.cuda.Event(enable_timing=True)
    e_end = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e_start.record()
    for k in range(args.steps):
        step_e2e()
    e_end.record()
    barrier()
    e2e_wall_ms = 1000.0 * (time.perf_counter() - t0)
    e2e_ms = max(e_start.elapsed_time(e_end), 0.0)

    # ---- max over ranks ---------------------------------------------------------------------------
    vals = torch.tensor([total_ms, e2e_ms, e2e_wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(vals, op=torch.distributed.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_wall_ms = [float(v) for v in vals]
    images = n_img * world * args.steps
    value = images / (total_ms / 1000.0)
    e2e_value = images / (max(e2e_ms, e2e_wall_ms) / 1000.0)

    # ---- roofline numbers -------------------------------------------------------------------------
    A = batch.num_anchors
    # algorithmic bytes of bulk_focal_kernel per image: every logit read once and its gradient
    # written once (A*4C each); the regression / IoU-prediction gradients belong to positive_terms_kernel
    kernel_bytes_per_image = A * (4 * 80 + 4 * 80)
    kernel_ms = k_ms.value / max(1, k_n.value)
    achieved = kernel_bytes_per_image * n_img / (kernel_ms / 1000.0) / 1e9 if kernel_ms > 0 else 0.0
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    # DRAM bytes of one launch of the same kernel from the committed ncu --set full capture (profiles/)
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(traffic_path) and n_img == 16:
        try:
            k = json.load(open(traffic_path))["kernels"]["bulk_focal_kernel<1, 1>"]
            traffic = k["dram_read_bytes"] + k["dram_write_bytes"]
        except (KeyError, ValueError):
            traffic = None
    roofline = {"bound": "hbm", "kernel": "bulk_focal_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_us": 1000.0 * kernel_ms, "algorithmic_bytes_per_launch": kernel_bytes_per_image * n_img,
                "whole_step_frac_of_hbm_roofline": (A * 696 * n_img / ((total_ms / args.steps) / 1000.0) / 1e9) / peak,
                "per_kernel_us": shares}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, note=("CUDA graph replay" if graph is not None else "eager launches")),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 12,
                "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps},
        # our kernels per step: assign_pass1, match_score, select_gmm, bulk_focal, positive_terms, finish_loss
        # (+ norm_wait_kernel with more than one rank)
        "gpu_launches": (6 + (1 if world > 1 else 0)) * args.steps,
        "roofline": roofline,
        "step_ms_min_med_max": [min(step_ms), statistics.median(step_ms), max(step_ms)],
    }
    # side measurements (single-GPU runs; not the headline): see the comments where they are taken
    if clean_ms is not None:
        line["cold_clean_l2"] = {"value": images / (clean_ms / 1000.0), "unit": UNIT,
                                 "ms_per_step": clean_ms / args.steps,
                                 "l2": "256 MiB write then 256 MiB read of another buffer before every step (untimed)"}
    if eager_ms is not None:
        line["eager_api_resident"] = {"value": images / (eager_ms / 1000.0), "unit": UNIT,
                                      "ms_per_step": eager_ms / args.steps,
                                      "what": "PAALossComputation.__call__ + torch.autograd.grad, inputs resident, "
                                              "no graph, wall clock between two synchronisations"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = min(CPU_SAMPLE_IMAGES, n_img)
        ips, times = cpu_port_images_per_sec(batch, n, repeats=3, warmup=1)
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "first %d images of the step's batch, fwd+bwd, median of 3 after 1 warm-up"
                                          % n}
    if rank == 0 and not args.no_post:
        try:
            from bench_post import measure_post
            line["post"] = measure_post(dev, steps=min(args.steps, 10), warmup=3)
        except Exception as e:  # noqa: BLE001
            line["post"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroying the process group while a captured CUDA graph still
        # references its communicator can block forever, and there is nothing left to clean up.
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
