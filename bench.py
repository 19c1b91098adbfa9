#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200, ONE JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--metric loss|post] [--config C1|C2|C3|C5 (loss) | C4 (post)] [--scaling strong|weak]

Metric: "PAA assign+loss images/sec at 1/2/4/8 B200; NMS+voting images/sec".  The default line is the first half
on config C2 (BASELINE.json configs[1], the paa_R_50_FPN_1x training shape) and carries the second half (config C4)
as its `post` object; `--metric post` prints the second half as a line of its own.

Workloads (synthetic, seed = 1000 * config):
  C1  2 images of 800x1333 (22 400 anchors on P3..P7), 20 GT each           -- the reference's CPU-runnable case
  C2  16 images of 800x1333, 1..100 GT each, TOPK 9                         -- headline
  C3  32 images of 1333x1333 (37 606 anchors), 500 GT each                  -- dense crowd
  C5  64 images, short side 640..800, padded to the batch maximum           -- multi-scale
  C4  64 images of 800x1333, ~4000 candidates above 0.05 per level, 1000 kept per level, NMS 0.6, voting
A config names a GLOBAL batch.  Default `--scaling strong` is BASELINE's own split: the global batch is sharded
over the N ranks exactly like the reference's DistributedSampler (data/build.py:110-115; C2: 16 images on one GPU,
2 per GPU on eight).  The weak-scaling measurement (every rank runs the whole N=1 batch on its own data) rides
along as `weak_scaling`; `--scaling weak` swaps the two.

A step of the loss metric is the training-step semantics of the path: forward (assignment + three losses) AND the
gradients w.r.t. box_cls / box_regression / iou_pred.
  value    : images/s with inputs resident in HBM, the step captured once in a CUDA graph and replayed; per-step CUDA
             events, L2 flushed between steps (256 MiB write, untimed), barrier + synchronize on both sides, max over
             ranks.
  e2e      : the same step through the reference-facing call (`PAALossComputation.__call__` + `torch.autograd.grad`,
             `PAAPostProcessor.forward` + BoxLists moved to the CPU) with every step's inputs copied from pinned host
             memory and the result read back inside the timed region.
  roofline : the WHOLE step against the HBM roofline (algorithmic bytes of SURVEY.md 8d / step time), `kernel` = the
             kernel with the largest event time, `per_kernel` = every kernel of the step timed per launch by CUDA
             events inside the library during an eager pass of the same steps, with its algorithmic bytes, fraction
             and what bounds it.
  cpu_baseline : the CPU port of the reference path (oracle/) on a bounded sample, rank 0 at N=1 only.
`--impl reference` times that CPU port alone (the reference's own Python cannot travel to the GPU box)
and prints the same line with "impl": "reference"; with N > 1 rank 0 alone runs it.
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

LOSS_METRIC = "PAA assign+loss images/sec"
POST_METRIC = "PAA NMS+voting images/sec"
UNIT = "images/s"
NUM_CLASSES = 80

# name -> global batch and how paa_b200.synthetic draws it
CONFIGS = {
    "C1": dict(kind="loss", images=2, make=dict(seed=1000, image_hw=(800, 1333), gt_per_image=20),
               what="C1 reference-runnable case: %d images of 800x1333 (22400 anchors, P3-P7), 80 classes, 20 GT/img"),
    "C2": dict(kind="loss", images=16, make=dict(seed=2000, image_hw=(800, 1333), gt_per_image=(1, 100)),
               what="C2 paa_R_50_FPN_1x training shape: %d images of 800x1333 (22400 anchors, P3-P7), 80 classes, "
                    "1-100 GT/img"),
    "C3": dict(kind="loss", images=32, make=dict(seed=3000, image_hw=(1333, 1333), gt_per_image=500),
               what="C3 dense crowd: %d images of 1333x1333 (37606 anchors), 80 classes, 500 GT/img"),
    "C5": dict(kind="loss", images=64, make=dict(seed=5000, image_hw=(0, 0), gt_per_image=(1, 100)), multiscale=True,
               what="C5 multi-scale R-101 2x shape: %d images, short side 640-800, padded to the batch maximum, "
                    "80 classes, 1-100 GT/img"),
    "C4": dict(kind="post", images=64, make=dict(seed=4000, image_hw=(800, 1333), candidates_per_level=4000),
               what="C4 PAAPostProcessor: %d images of 800x1333, ~4000 candidates above 0.05 per level, 1000 kept per "
                    "level, 80 classes, NMS 0.6, score voting, 100 detections/img"),
}
C2_BATCH_KW = dict(num_images=CONFIGS["C2"]["images"], **CONFIGS["C2"]["make"])     # tests compare this batch with the oracle
CPU_SAMPLE_IMAGES = 2       # bounded sample for the CPU arm of the loss metric
FLUSH_BYTES = 256 << 20

# The kernels of one assign+loss step: (report name, timer id in paa_b200._lib.KERNEL_IDS, what bounds it,
# algorithmic bytes per image as a function of (A anchors, C classes), evidence for kernels that no HBM fraction
# describes -- static, from the committed ncu --set full capture named there).
LOSS_KERNELS = [
    ("prep_step_kernel", "prep_step", "latency", lambda A, C: 0,
     "clears the workspace prefix, leaves the GT ranges in device memory; hidden behind the next launch"),
    ("iou_match_kernel", "pass1", "latency", lambda A, C: A * (16 + 8),
     "anchors x GT IoU with warp-level culling, no head tensor touched: issue / latency"),
    ("match_score_kernel", "match_score", "hbm", lambda A, C: A * (4 * C + 8 + 4 + 16 + 16 + 12),
     "Matcher + anchor scores; the class sums read the logits of IoU-positive anchors only (~45 % of the anchors on "
     "C2; measured DRAM traffic 92 MB of the 115 MB of logits)"),
    ("select_gmm_kernel", "select_gmm", "latency", lambda A, C: A * 12,
     "one warp per GT runs a serial f64 EM chain; the slowest GT of the batch decides (ncu: warps active ~14 %, "
     "top stalls wait / short_scoreboard)"),
    ("bulk_focal_early_kernel", "final_loss", "hbm", lambda A, C: A * (4 * C + 4 * C + 4 + 20),
     "every logit read once, its gradient written once (chunks handed out dynamically; in the graph some are processed "
     "unscaled in the shadow of the last EM fits and scaled in place afterwards -- an event-timed launch has no shadow), "
     "plus the zero fill of the non-positive anchors' regression / IoU gradients; calls with few chunks, several ranks "
     "or the two-launch fit take the static bulk_focal_kernel"),
    ("positive_list_kernel", "positive_terms", "latency", lambda A, C: 0,
     "the positives from per-GT lists (dependent loads per positive) -- in the graph in front of its dependency wait, in "
     "the shadow of the bulk pass; behind it one store per positive and the loss fold in the last block"),
]
POST_KERNELS = [
    ("post_candidates_kernel", "post_candidates", "hbm", lambda A, C: A * (4 * C + 4 + 16 + 16),
     "flat stream over the logits; gated elements drained through a block queue"),
    ("post_threshold_kernel", "post_threshold", "latency", lambda A, C: 0, "histogram walk"),
    ("post_filter_kernel", "post_filter", "latency", lambda A, C: 0, "8 B per candidate"),
    ("post_select_kernel", "post_select", "latency", lambda A, C: 0,
     "rank the boundary bin, counting sort by candidate index, decode per (image, level)"),
    ("post_group_kernel+post_class_rank_kernel", "post_rank", "latency", lambda A, C: 0, "order by (label, score)"),
    ("post_nms_runs_kernel", "post_nms_runs", "alu", lambda A, C: 0,
     "greedy NMS of every label run (<= 256 boxes) by one warp: candidates against the run's kept boxes, no mask in memory"),
    ("post_nms_mask_kernel", "post_nms_mask", "latency", lambda A, C: 0,
     "longer runs only (64x64 bit tiles of pair IoUs); images without one return at once"),
    ("post_nms_scan_kernel", "post_nms_scan", "latency", lambda A, C: 0,
     "longer runs only (one warp per label run scans the mask greedily); images without one return at once"),
    ("post_finish_kernel", "post_finish", "latency", lambda A, C: 0, "top-100 cut by radix select, ordered compaction"),
    ("post_vote_kernel", "post_vote", "alu", lambda A, C: 0, "one warp per detection votes"),
]


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--metric", default="loss", choices=["loss", "post"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--images-per-gpu", type=int, default=None, help="override the config's per-GPU share")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"],
                    help="memory layout of the head tensors: nchw = what the reference's heads produce (default); "
                         "nhwc = torch.channels_last heads, consumed in place")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="loss metric: skip the NMS+voting object")
    ap.add_argument("--no-side", action="store_true", help="skip the side measurements (other scaling mode, eager API)")
    args = ap.parse_args(argv)
    if args.config is None:
        args.config = "C4" if args.metric == "post" else "C2"
    if CONFIGS[args.config]["kind"] != args.metric:
        ap.error("--config %s belongs to --metric %s" % (args.config, CONFIGS[args.config]["kind"]))
    args.warmup = max(3, args.warmup)          # timing rule: at least three warm-up steps
    return args


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_global_batch(name, num_images=None, seed_offset=0):
    from paa_b200 import synthetic
    c = CONFIGS[name]
    n = num_images or c["images"]
    kw = dict(c["make"])
    kw["seed"] = kw["seed"] + seed_offset
    if c["kind"] == "post":
        return synthetic.make_inference_batch(num_images=n, **kw)
    if c.get("multiscale"):
        kw["per_image_hw"] = synthetic.multiscale_hw(kw["seed"], n)
    return synthetic.make_batch(num_images=n, **kw)


def rank_batch(args, scaling, rank, world):
    """This rank's images under `scaling`, and (images per GPU, global images)."""
    from paa_b200 import synthetic
    c = CONFIGS[args.config]
    if args.images_per_gpu:
        return make_global_batch(args.config, args.images_per_gpu, seed_offset=rank), args.images_per_gpu, \
            args.images_per_gpu * world
    if scaling == "strong":
        if c["images"] % world:
            raise SystemExit("config %s: %d images do not split over %d GPUs" % (args.config, c["images"], world))
        per = c["images"] // world
        g = make_global_batch(args.config)
        return (g if world == 1 else synthetic.slice_batch(g, rank * per, (rank + 1) * per)), per, c["images"]
    return make_global_batch(args.config, seed_offset=rank), c["images"], c["images"] * world


def workload_config(args, scaling, per_gpu, total, world):
    c = CONFIGS[args.config]
    cfg = {"workload": (c["what"] % total) + (", topk 9, forward+gradients" if c["kind"] == "loss" else ""),
           "config": args.config, "images_per_gpu": per_gpu, "global_batch": total,
           "parallelism": "dp%d" % world,
           "split": ("BASELINE's split: the global batch of %d sharded over the ranks" % c["images"]) if scaling == "strong"
           else "weak scaling: every rank runs a whole %d-image batch of its own" % per_gpu,
           "l2": "flushed between timed steps (256 MiB write, untimed)",
           "head_layout": "NCHW (contiguous, what the reference's heads produce)" if LAYOUT == "nchw"
           else "NHWC (torch.channels_last heads, consumed in place)",
           "collective": ("none (one rank)" if world == 1 else
                          "nvlink-peer: 16 bytes per rank stored into every peer's symmetric-memory buffer by the last "
                          "block of select_gmm_kernel, read by norm_wait_kernel (no NCCL call on the data path)")
           if c["kind"] == "loss" else "none (post-processing has no exchange step)"}
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
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._poll_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

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
# CPU arm: the oracle port of the reference path (bench.py is one of the places allowed to run it)
# ------------------------------------------------------------------------------------------------
def host_threads():
    n = os.cpu_count() or 1
    torch.set_num_threads(n)      # explicit: torchrun exports OMP_NUM_THREADS=1 to its children
    return torch.get_num_threads()


def _timed_repeats(fn, repeats, warmup, time_budget_s):
    times, begin = [], time.perf_counter()
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time_budget_s is not None and times and time.perf_counter() - begin > time_budget_s:
            break
    return times


def cpu_port_loss(batch, n_images, repeats, warmup=1, time_budget_s=None):
    """(images/s, step times) of the CPU port of PAALossComputation.__call__ + backward on the first `n_images`."""
    from oracle import paa_oracle
    sl = slice(0, n_images)
    a = ([t[sl] for t in batch.box_cls], [t[sl] for t in batch.box_regression], [t[sl] for t in batch.iou_pred],
         batch.gt_boxes[sl], batch.gt_labels[sl], batch.anchors)
    times = _timed_repeats(lambda: paa_oracle.assign_and_loss(*a, with_grad=True), repeats, warmup, time_budget_s)
    return n_images / statistics.median(times), times


def cpu_port_post(batch, n_images, repeats, warmup=1, time_budget_s=None):
    """(images/s, step times) of the CPU port of PAAPostProcessor.forward on the first `n_images`."""
    from oracle import post_oracle
    sl = slice(0, n_images)
    a = ([t[sl] for t in batch.box_cls], [t[sl] for t in batch.box_regression], [t[sl] for t in batch.iou_pred],
         batch.anchors, batch.image_sizes[sl])
    times = _timed_repeats(lambda: post_oracle.postprocess(*a), repeats, warmup, time_budget_s)
    return n_images / statistics.median(times), times


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (its port, oracle/), rank 0 only; each step
    a bounded sample of the workload."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cores = host_threads()
    c = CONFIGS[args.config]
    per_gpu = args.images_per_gpu or (c["images"] // max(1, args.gpus) if args.scaling == "strong" else c["images"])
    total = per_gpu * max(1, args.gpus)
    # K steps after W warm-up steps as asked, each a bounded sample (1-2 images, ~0.5-1.5 s of CPU work); the
    # wall-clock budget only cuts a run short on a very slow host, and the line reports the steps actually timed
    warm = args.warmup
    if args.metric == "post":
        n = 1
        batch = make_global_batch(args.config, n)
        ips, times = cpu_port_post(batch, n, max(1, args.steps), warm, time_budget_s=150.0)
        what = "PAAPostProcessor.forward"
    else:
        n = min(CPU_SAMPLE_IMAGES, per_gpu)
        batch = make_global_batch(args.config, n)
        ips, times = cpu_port_loss(batch, n, max(1, args.steps), warm, time_budget_s=150.0)
        what = "PAALossComputation.__call__ + backward"
    sample = "%s on %d image(s) drawn like the workload's (seed %d) per step, %d timed steps (median), %d threads" % (
        what, n, c["make"]["seed"], len(times), cores)
    line = {
        "impl": "reference", "metric": POST_METRIC if args.metric == "post" else LOSS_METRIC, "value": ips,
        "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
        "ms_per_step": 1000.0 * statistics.median(times), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.scaling, per_gpu, total, max(1, args.gpus)),
        "launch": "CPU port of the reference path (oracle/), a bounded sample of the workload per step",
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm: shared plumbing
# ------------------------------------------------------------------------------------------------
class Ctx(object):
    def __init__(self):
        self.rank, self.local_rank, self.world = dist_env()
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: there is no CPU path for the product")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return [float(v) for v in t]

    def min_over_ranks(self, values):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        return [float(v) for v in t]


def capture(step, no_graph):
    """Warm-up on a side stream, then one capture of `step` (a callable that only enqueues stream work)."""
    if no_graph:
        return None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        torch.cuda.synchronize()
        return graph
    except Exception as e:  # noqa: BLE001 - report and fall back to eager timing
        sys.stderr.write("CUDA graph capture failed (%s); timing eager launches\n" % (e,))
        torch.cuda.synchronize()
        return None


def timed_steps(ctx, one_step, steps, warmup):
    """W untimed then exactly K timed steps, per-step CUDA events on the launching stream, L2 flushed before every
    step, barrier + synchronize on both sides.  Returns the per-step times of this rank (ms)."""
    for _ in range(warmup):
        ctx.flush.zero_()
        one_step()
    ctx.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for k in range(steps):
        ctx.flush.zero_()
        starts[k].record()
        one_step()
        ends[k].record()
    ctx.barrier()
    return [s.elapsed_time(e) for s, e in zip(starts, ends)]


def per_kernel_events(ctx, lib, kernel_table, step, reps):
    """Average device time per launch of every kernel in `kernel_table`, one kernel at a time: the library brackets
    each launch of the chosen kernel with CUDA events on its own stream during `reps` eager steps (L2 flushed before
    each).  The event pair serialises the launch behind its predecessor, so these are stand-alone kernel times: under
    the graph the dependent launches overlap and the step is shorter than their sum."""
    from paa_b200 import _lib
    out = {}
    for name, key, _, _, _ in kernel_table:
        lib.paa_kernel_timing_begin(_lib.KERNEL_IDS[key])
        for _ in range(reps):
            ctx.flush.zero_()
            step()
        ms, n = ctypes.c_float(0), ctypes.c_int32(0)
        _lib.check(lib.paa_kernel_timing_end(ctypes.byref(ms), ctypes.byref(n)), "paa_kernel_timing_end")
        out[name] = (1000.0 * ms.value / n.value if n.value else 0.0, n.value // max(1, reps))
    return out


def roofline_report(kernel_table, per_kernel, A, C, n_img, step_ms, step_bytes_per_image, traffic_file):
    peak, peak_src = hbm_peak()
    traffic = {}
    path = os.path.join(ROOT, "profiles", traffic_file)
    if os.path.exists(path):
        try:
            traffic = json.load(open(path))
        except ValueError:
            traffic = {}
    same_shape = traffic.get("images") == n_img and traffic.get("anchors_per_image") == A
    rows, longest = [], None
    total_us = sum(us for us, _ in per_kernel.values()) or 1.0
    for name, _, bound, bytes_fn, why in kernel_table:
        us, launches = per_kernel.get(name, (0.0, 0))
        if launches == 0:
            continue
        b = bytes_fn(A, C) * n_img
        row = {"kernel": name, "us": round(us, 2), "share_of_kernel_time": round(us / total_us, 4), "bound": bound,
               "algorithmic_bytes": b, "why": why}
        if b and us > 0:
            row["achieved_GBps"] = round(b / (us * 1e-6) / 1e9, 1)
            row["frac_of_hbm_peak"] = round(b / (us * 1e-6) / 1e9 / peak, 4)
        k = traffic.get("kernels", {}).get(name) if same_shape else None
        if k:
            row["traffic"] = k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0)
            for key in ("issue_active_pct", "warps_active_pct", "dram_throughput_pct"):
                if key in k:
                    row["ncu_" + key] = k[key]
        rows.append(row)
        if longest is None or us > longest["us"]:
            longest = row
    achieved = step_bytes_per_image * n_img / (step_ms * 1e-3) / 1e9
    return {"bound": "hbm", "scope": "whole step (all kernels of the path, CUDA-graph replay)",
            "kernel": longest["kernel"] if longest else None,
            "kernel_us": longest["us"] if longest else None, "kernel_bound": longest["bound"] if longest else None,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "algorithmic_bytes_per_step": step_bytes_per_image * n_img,
            "traffic": (sum(r.get("traffic", 0) for r in rows) or None) if same_shape else None,
            "traffic_source": ("static: dram__bytes_read.sum + dram__bytes_write.sum summed over the step's kernels, "
                               "from the committed ncu --set full capture of the same kernels on the same batch "
                               "(tools/loss_once.py / tools/post_once.py, tools/capture_r2z.sh), profiles/%s"
                               % traffic_file) if same_shape and traffic
            else "no capture of this shape committed",
            "peak_source": peak_src, "per_kernel": rows}


# ------------------------------------------------------------------------------------------------
# head layout of the run (--layout)
# ------------------------------------------------------------------------------------------------
LAYOUT = "nchw"


def head_format():
    return torch.channels_last if LAYOUT == "nhwc" else torch.contiguous_format


def head_bytes_in_order(t):
    """The values of a logical [N, C, H, W] head tensor in the order the run's layout keeps them in memory."""
    return t.permute(0, 2, 3, 1).reshape(-1) if LAYOUT == "nhwc" else t.reshape(-1)


def head_view(flat, shape):
    """A logical [N, C, H, W] view of a flat buffer holding the tensor in the run's layout."""
    if LAYOUT == "nhwc":
        n, c, h, w = shape
        return flat.view(n, h, w, c).permute(0, 3, 1, 2)
    return flat.view(shape)


# ------------------------------------------------------------------------------------------------
# assign + loss
# ------------------------------------------------------------------------------------------------
class LossRun(object):
    """One rank's share of a loss workload: resident inputs, the evaluator, the captured step."""

    def __init__(self, ctx, batch, no_graph):
        import paa_b200
        from paa_b200 import _lib
        from paa_b200.structures import BoxList
        self.ctx, self.batch = ctx, batch
        dev = ctx.dev
        self.n_img = batch.num_images
        self.L = len(batch.box_cls)
        cfg = paa_b200.default_cfg()
        self.ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
        self.lib = _lib.load()
        pin = lambda ts: [t.contiguous(memory_format=head_format()).pin_memory() for t in ts]       # noqa: E731
        self.h_cls, self.h_reg, self.h_iou = pin(batch.box_cls), pin(batch.box_regression), pin(batch.iou_pred)
        self.h_gtb = [t.pin_memory() for t in batch.gt_boxes]
        self.h_gtl = [t.pin_memory() for t in batch.gt_labels]
        d_anchor = [a.to(dev) for a in batch.anchors]
        self.anchors = [[BoxList(a, batch.image_sizes[i]) for a in d_anchor] for i in range(self.n_img)]
        self.d_cls = [t.to(dev).requires_grad_(True) for t in self.h_cls]
        self.d_reg = [t.to(dev).requires_grad_(True) for t in self.h_reg]
        self.d_iou = [t.to(dev).requires_grad_(True) for t in self.h_iou]
        self.d_targets = self.targets_from([t.to(dev) for t in self.h_gtb], [t.to(dev) for t in self.h_gtl])
        self.heads = self.d_cls + self.d_reg + self.d_iou
        for _ in range(3):
            self.step_resident()
        torch.cuda.synchronize()
        self.graph = capture(self.step_resident, no_graph)

    def targets_from(self, boxes, labels):
        from paa_b200.structures import BoxList
        out = []
        for i in range(self.n_img):
            t = BoxList(boxes[i], self.batch.image_sizes[i])
            t.add_field("labels", labels[i])
            out.append(t)
        return out

    def step_resident(self):
        # the fused entry point of the same evaluator: losses + gradients, no autograd bookkeeping, so the whole
        # step is capturable; identical kernels to `ev(...)` + backward (which the e2e leg times)
        return self.ev.forward_backward(self.d_cls, self.d_reg, self.d_iou, self.d_targets, self.anchors)

    def one_step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.step_resident()

    def step_eager_api(self):
        losses = self.ev(self.d_cls, self.d_reg, self.d_iou, self.d_targets, self.anchors, None)
        return torch.autograd.grad(losses[0] + losses[1] + losses[2], self.heads)

    def eager_api_ms(self, steps):
        """`PAALossComputation.__call__` + autograd on resident inputs, back to back, wall clock between two
        synchronisations: what a caller pays per step when nothing else keeps the GPU busy (host-bound)."""
        for _ in range(3):
            self.step_eager_api()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step_eager_api()
        torch.cuda.synchronize()
        return 1000.0 * (time.perf_counter() - t0) / steps

    def graph_mode_ms(self, steps, autograd):
        """The evaluator in graph mode (`use_graph`): ONE captured step replayed for every call, with a different
        set of targets each time (the second set is the first with the images' ground truth rotated by one image) --
        the way a training loop feeds it.  Wall clock between two synchronisations, back to back.  `autograd`: through
        `__call__` + `torch.autograd.grad` (the drop-in call); else through `forward_backward`."""
        import paa_b200
        cfg = paa_b200.default_cfg()
        ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
        ev.use_graph = True
        ev.gt_per_image_capacity = 128
        rotated = self.targets_from([t.bbox for t in self.d_targets[1:] + self.d_targets[:1]],
                                    [t.get_field("labels") for t in self.d_targets[1:] + self.d_targets[:1]])
        sets = [self.d_targets, rotated]

        def step(k):
            if autograd:
                losses = ev(self.d_cls, self.d_reg, self.d_iou, sets[k & 1], self.anchors, None)
                return torch.autograd.grad(losses[0] + losses[1] + losses[2], self.heads)
            return ev.forward_backward(self.d_cls, self.d_reg, self.d_iou, sets[k & 1], self.anchors)

        for k in range(4):
            step(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            step(k)
        torch.cuda.synchronize()
        ms = 1000.0 * (time.perf_counter() - t0) / steps
        graphs = len(ev._graphs)
        calls = sum(g.calls for g in ev._graphs.values())
        return ms, graphs, calls

    def exchange_check(self):
        """N > 1: the normalisers the peer-memory exchange delivered against an NCCL all-reduce of the ranks' own
        pairs, on every rank.  Returns (max relative difference, 1.0 if the peer path was in use else 0.0)."""
        from paa_b200 import loss as paa_loss
        ev = self.ev
        ev.debug = True
        try:
            ev.forward_backward(self.d_cls, self.d_reg, self.d_iou, self.d_targets, self.anchors)
            got = ev.last_debug["normalisers"].clone()
            want = ev.last_debug["local_normalisers"].clone()
        finally:
            ev.debug = False
        torch.distributed.all_reduce(want, op=torch.distributed.ReduceOp.SUM)
        torch.cuda.synchronize()
        rel = float(((got - want).abs() / want.abs().clamp(min=1e-300)).max())
        used_peer = paa_loss.PeerNormExchange.get(self.ctx.dev) is not None
        return rel, (1.0 if used_peer else 0.0)

    def e2e(self, steps, warmup, graph_mode=True):
        """Host buffers in, losses out, public API.  The step's inputs live in pinned host memory, packed the
        way a collate function would leave them (one block for the head outputs, one each for the boxes / labels of all
        images).  Every step copies them to the device on a copy stream into one of two device buffers (the copy of
        step k+1 overlaps the kernels of step k), calls the reference-facing evaluator + autograd on views of that
        buffer, and reads the three losses back into pinned memory.  Returns (device ms, wall ms, h2d bytes, d2h)."""
        import paa_b200
        ctx, dev, L = self.ctx, self.ctx.dev, self.L
        # the evaluator in graph mode (a public switch of PAALossComputation): the two device input buffers alternate,
        # so two captured steps serve all calls; targets change through the device-resident GT ranges
        cfg = paa_b200.default_cfg()
        ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
        ev.use_graph = graph_mode
        heads_h = self.h_cls + self.h_reg + self.h_iou
        sizes = [t.numel() for t in heads_h]
        h_pack = torch.empty(sum(sizes), dtype=torch.float32).pin_memory()
        o = 0
        for t, sz in zip(heads_h, sizes):
            h_pack[o:o + sz].copy_(head_bytes_in_order(t))
            o += sz
        gt_counts = [int(t.shape[0]) for t in self.h_gtb]
        h_boxes = torch.cat(self.h_gtb, 0).contiguous().pin_memory()
        h_labels = torch.cat(self.h_gtl, 0).contiguous().pin_memory()
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

        def step():
            k = state["k"]
            state["k"] = k + 1
            sl = slots[k % 2]
            enqueue_copy(k + 1)                              # next step's inputs travel while this step computes
            main = torch.cuda.current_stream()
            main.wait_event(sl["copied"])
            views, o = [], 0
            for t, sz in zip(heads_h, sizes):
                views.append(head_view(sl["pack"][o:o + sz], t.shape).requires_grad_(True))
                o += sz
            cls, reg, iou = views[:L], views[L:2 * L], views[2 * L:]
            tg = self.targets_from(list(sl["boxes"].split(gt_counts)), list(sl["labels"].split(gt_counts)))
            losses = ev(cls, reg, iou, tg, self.anchors, None)
            grads = torch.autograd.grad(losses[0] + losses[1] + losses[2], cls + reg + iou)
            h_losses.copy_(torch.stack([l.detach() for l in losses]), non_blocking=True)
            sl["free"].record(main)
            return grads

        for sl in slots:
            sl["free"].record(torch.cuda.current_stream())
        enqueue_copy(0)
        for _ in range(warmup):
            step()
        ctx.barrier()
        if os.environ.get("PAA_BENCH_PROFILE"):          # measurement aid: where the host time of an e2e step goes
            import cProfile
            import pstats
            prof = cProfile.Profile()
            prof.enable()
            for _ in range(50):
                step()
            prof.disable()
            torch.cuda.synchronize()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(35)
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e_start.record()
        for _ in range(steps):
            step()
        e_end.record()
        ctx.barrier()
        wall_ms = 1000.0 * (time.perf_counter() - t0)
        return max(e_start.elapsed_time(e_end), 0.0), wall_ms, h2d_bytes, 12


def measure_loss(ctx, args, scaling, full):
    """One scaling mode of the loss metric on this rank; everything that enters the JSON line is already the
    max over ranks.  `full`: also the per-kernel pass, e2e and the side measurements."""
    batch, per_gpu, total = rank_batch(args, scaling, ctx.rank, ctx.world)
    run = LossRun(ctx, batch, args.no_graph)
    A = batch.num_anchors
    sampler = ClockSampler(ctx.local_rank).start() if full else None
    step_ms = timed_steps(ctx, run.one_step, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    total_ms, = ctx.max_over_ranks([sum(step_ms)])
    res = {"per_gpu": per_gpu, "total": total, "A": A, "graph": run.graph is not None,
           "ms_per_step": total_ms / args.steps, "value": total * args.steps / (total_ms / 1000.0),
           "step_ms_min_med_max": [min(step_ms), statistics.median(step_ms), max(step_ms)], "clocks": clocks,
           "num_gt": int(sum(int(t.shape[0]) for t in batch.gt_boxes))}
    if ctx.world > 1:
        rel, peer = run.exchange_check()
        rel, = ctx.max_over_ranks([rel])
        peer, = ctx.min_over_ranks([peer])
        res["exchange_check"] = {"what": "normalisers delivered by the exchange vs an NCCL all-reduce of the ranks' own "
                                         "pairs, checked on every rank (max over ranks)",
                                 "max_rel_diff": rel, "ok": bool(rel <= 1e-12), "ranks": ctx.world,
                                 "path": "nvlink-peer" if peer else "nccl all-reduce (peer memory unavailable)"}
    if full:
        per_kernel = per_kernel_events(ctx, run.lib, LOSS_KERNELS, run.step_resident, 3)
        res["roofline"] = roofline_report(LOSS_KERNELS, per_kernel, A, NUM_CLASSES, per_gpu, res["ms_per_step"],
                                          A * 696, "r2_traffic_loss.json")
        res["kernel_launches_per_step"] = sum(n for _, n in per_kernel.values()) + (1 if ctx.world > 1 else 0)
        e_ms, e_wall, h2d, d2h = run.e2e(args.steps, args.warmup)
        e_ms, e_wall = ctx.max_over_ranks([e_ms, e_wall])
        res["e2e"] = {"value": total * args.steps / (max(e_ms, e_wall) / 1000.0), "unit": UNIT,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "ms_per_step": max(e_ms, e_wall) / args.steps,
                      "what": "PAALossComputation.__call__ + torch.autograd.grad (evaluator in graph mode, use_graph), "
                              "inputs copied from pinned host memory and losses read back every step (per-rank bytes)"}
        if not args.no_side:
            g_ms, g_wall, _, _ = run.e2e(args.steps, args.warmup, graph_mode=False)
            g_ms, g_wall = ctx.max_over_ranks([g_ms, g_wall])
            res["e2e"]["eager_launches"] = {"value": total * args.steps / (max(g_ms, g_wall) / 1000.0), "unit": UNIT,
                                            "ms_per_step": max(g_ms, g_wall) / args.steps,
                                            "what": "the same with use_graph off (every kernel launched by the host)"}
        if ctx.world == 1 and not args.no_side:
            try:
                ms = run.eager_api_ms(args.steps)
                res["eager_api_resident"] = {"value": total / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms,
                                             "what": "PAALossComputation.__call__ + torch.autograd.grad, inputs "
                                                     "resident, no graph, wall clock between two synchronisations"}
            except Exception as e:  # noqa: BLE001 - a side measurement must not take the bench line down
                sys.stderr.write("eager_api_resident side measurement failed: %s\n" % (e,))
            try:
                for key, autograd in (("graph_forward_backward_new_targets", False), ("graph_api_new_targets", True)):
                    ms, graphs, calls = run.graph_mode_ms(max(args.steps, 20), autograd)
                    res[key] = {"value": total / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms,
                                "captured_graphs": graphs, "replays_and_capture": calls,
                                "what": ("PAALossComputation.__call__ + torch.autograd.grad" if autograd else
                                         "PAALossComputation.forward_backward") +
                                        " in graph mode (use_graph): one captured step replayed with a different target "
                                        "set every call, inputs resident, wall clock between two synchronisations"}
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("graph-mode side measurement failed: %s\n" % (e,))
        if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            n = min(CPU_SAMPLE_IMAGES, per_gpu)
            ips, times = cpu_port_loss(batch, n, repeats=3, warmup=1, time_budget_s=60.0)
            res["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "first %d images of the step's batch, fwd+bwd, median of %d after 1 "
                                             "warm-up, %d threads" % (n, len(times), cores)}
    del run
    torch.cuda.empty_cache()
    return res


def loss_line(ctx, args):
    other = "weak" if args.scaling == "strong" else "strong"
    main = measure_loss(ctx, args, args.scaling, full=True)
    side = None
    if ctx.world > 1 and not args.no_side and not args.images_per_gpu:
        side = measure_loss(ctx, args, other, full=False)
    line = {
        "metric": LOSS_METRIC, "value": main["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.scaling, main["per_gpu"], main["total"], ctx.world),
        "launch": "CUDA graph replay" if main["graph"] else "eager launches",
        "clocks": main["clocks"], "e2e": main["e2e"],
        "gpu_launches": main["kernel_launches_per_step"] * args.steps,
        "roofline": main["roofline"], "step_ms_min_med_max": main["step_ms_min_med_max"],
        "num_gt_this_rank": main["num_gt"],
    }
    for key in ("exchange_check", "eager_api_resident", "graph_forward_backward_new_targets", "graph_api_new_targets",
                "cpu_baseline"):
        if key in main:
            line[key] = main[key]
    if side is not None:
        line[other + "_scaling"] = {"value": side["value"], "unit": UNIT, "ms_per_step": side["ms_per_step"],
                                    "images_per_gpu": side["per_gpu"], "global_batch": side["total"],
                                    "exchange_check": side.get("exchange_check")}
    return line


# ------------------------------------------------------------------------------------------------
# NMS + score voting
# ------------------------------------------------------------------------------------------------
class PostRun(object):
    def __init__(self, ctx, batch, no_graph):
        import paa_b200
        from paa_b200 import _lib
        from paa_b200.structures import BoxList
        self.ctx, self.batch = ctx, batch
        dev = ctx.dev
        self.n_img = batch.num_images
        cfg = paa_b200.default_cfg()
        self.pp = paa_b200.make_paa_postprocessor(cfg, paa_b200.BoxCoder(cfg))
        self.lib = _lib.load()
        self.cls = [t.to(dev).contiguous(memory_format=head_format()) for t in batch.box_cls]
        self.reg = [t.to(dev).contiguous(memory_format=head_format()) for t in batch.box_regression]
        self.iou = [t.to(dev).contiguous(memory_format=head_format()) for t in batch.iou_pred]
        anc = [a.to(dev) for a in batch.anchors]
        self.anchors = [[BoxList(a, batch.image_sizes[i]) for a in anc] for i in range(self.n_img)]
        for _ in range(3):
            self.out = self.step_resident()
        torch.cuda.synchronize()
        self.graph = capture(self.step_resident, no_graph)

    def step_resident(self):
        return self.pp.run_device(self.cls, self.reg, self.iou, self.anchors)

    def one_step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.step_resident()

    def e2e(self, steps, warmup):
        """`PAAPostProcessor.forward` with host logits in and BoxLists out on the CPU: every step copies the head
        outputs from pinned host memory (double-buffered on a copy stream), calls forward() and moves every image's
        BoxList to the host the way engine/inference.py:38-39 does (`[o.to(cpu_device) for o in output]`)."""
        ctx, dev = self.ctx, self.ctx.dev
        heads_h = [t.contiguous() for t in self.batch.box_cls + self.batch.box_regression + self.batch.iou_pred]
        sizes = [t.numel() for t in heads_h]
        L = len(self.batch.box_cls)
        h_pack = torch.empty(sum(sizes), dtype=torch.float32).pin_memory()
        o = 0
        for t, sz in zip(heads_h, sizes):
            h_pack[o:o + sz].copy_(head_bytes_in_order(t))
            o += sz
        copy_stream = torch.cuda.Stream()
        slots = [dict(pack=torch.empty_like(h_pack, device=dev), copied=torch.cuda.Event(), free=torch.cuda.Event())
                 for _ in range(2)]
        state = {"k": 0, "d2h": 0}
        cpu = torch.device("cpu")

        def enqueue_copy(k):
            sl = slots[k % 2]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(sl["free"])
                sl["pack"].copy_(h_pack, non_blocking=True)
                sl["copied"].record(copy_stream)

        def step():
            k = state["k"]
            state["k"] = k + 1
            sl = slots[k % 2]
            enqueue_copy(k + 1)
            main = torch.cuda.current_stream()
            main.wait_event(sl["copied"])
            views, o = [], 0
            for t, sz in zip(heads_h, sizes):
                views.append(head_view(sl["pack"][o:o + sz], t.shape))
                o += sz
            out = self.pp(views[:L], views[L:2 * L], views[2 * L:], self.anchors)
            host = [b.to(cpu) for b in out]
            sl["free"].record(main)
            state["d2h"] = sum(len(b) * (16 + 4 + 8) for b in host) + 4 * len(host)
            return host

        for sl in slots:
            sl["free"].record(torch.cuda.current_stream())
        enqueue_copy(0)
        for _ in range(warmup):
            step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        ctx.barrier()
        wall_ms = 1000.0 * (time.perf_counter() - t0)
        return wall_ms, h_pack.numel() * 4, state["d2h"]


def measure_post(ctx, args, scaling, full, cfg_name="C4"):
    pargs = argparse.Namespace(**vars(args))
    pargs.config = cfg_name
    if args.metric != "post":
        pargs.images_per_gpu = None
    batch, per_gpu, total = rank_batch(pargs, scaling, ctx.rank, ctx.world)
    run = PostRun(ctx, batch, args.no_graph)
    A = batch.num_anchors
    steps = args.steps if args.metric == "post" else min(args.steps, 10)
    sampler = ClockSampler(ctx.local_rank).start() if full else None
    step_ms = timed_steps(ctx, run.one_step, steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    total_ms, = ctx.max_over_ranks([sum(step_ms)])
    res = {"metric": POST_METRIC, "value": total * steps / (total_ms / 1000.0), "unit": UNIT, "n_gpus": ctx.world,
           "steps": steps, "warmup": args.warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(pargs, scaling, per_gpu, total, ctx.world),
           "launch": "CUDA graph replay" if run.graph is not None else "eager launches",
           "clocks": clocks, "detections_first_images": [int(c) for c in run.out[3][:4].tolist()]}
    if full:
        per_kernel = per_kernel_events(ctx, run.lib, POST_KERNELS, run.step_resident, 3)
        res["roofline"] = roofline_report(POST_KERNELS, per_kernel, A, NUM_CLASSES, per_gpu, res["ms_per_step"],
                                          A * (4 * NUM_CLASSES + 16 + 4 + 16), "r2_traffic_post.json")
        res["gpu_launches"] = (sum(n for _, n in per_kernel.values()) + 1) * steps      # + post_class_rank_kernel
        wall, h2d, d2h = run.e2e(steps, args.warmup)
        wall, = ctx.max_over_ranks([wall])
        res["e2e"] = {"value": total * steps / (wall / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": wall / steps,
                      "what": "PAAPostProcessor.forward with the head outputs copied from pinned host memory and every "
                              "BoxList moved to the CPU each step (per-rank bytes, wall clock)"}
        if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            ips, times = cpu_port_post(batch, 1, repeats=3, warmup=1, time_budget_s=60.0)
            res["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "first image of the step's batch, median of %d after 1 warm-up, %d threads"
                                             % (len(times), cores)}
    del run
    torch.cuda.empty_cache()
    return res


def post_line(ctx, args):
    other = "weak" if args.scaling == "strong" else "strong"
    line = measure_post(ctx, args, args.scaling, full=True, cfg_name=args.config if args.metric == "post" else "C4")
    if ctx.world > 1 and not args.no_side and not args.images_per_gpu:
        s = measure_post(ctx, args, other, full=False, cfg_name=args.config if args.metric == "post" else "C4")
        line[other + "_scaling"] = {k: s[k] for k in ("value", "unit", "ms_per_step")}
        line[other + "_scaling"].update(images_per_gpu=s["config"]["images_per_gpu"],
                                        global_batch=s["config"]["global_batch"])
    return line


def run_ours(args):
    ctx = Ctx()
    if args.metric == "post":
        line = post_line(ctx, args)
    else:
        line = loss_line(ctx, args)
        if not args.no_post:
            try:
                line["post"] = post_line(ctx, args)
            except Exception as e:  # noqa: BLE001
                line["post"] = {"unavailable": str(e)[:300]}
    if ctx.rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        # Leave without tearing NCCL down: destroying the process group while a captured CUDA graph still
        # references its communicator can block forever, and there is nothing left to clean up.
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    global LAYOUT
    args = parse_args()
    LAYOUT = args.layout
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
