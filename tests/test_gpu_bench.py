"""GPU: bench.py prints ONE JSON line with the contract's keys for both halves of the metric (a short run on a
small per-GPU share), and its roofline names the kernel with the largest event time."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}


def _bench(*extra):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3",
                           "--no-cpu-baseline", "--no-side"] + list(extra), cwd=ROOT, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert proc.returncode == 0, proc.stderr[-3000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def _check_roofline(r):
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    rows = r["per_kernel"]
    assert len(rows) >= 5 and all(row["us"] > 0 for row in rows)
    longest = max(rows, key=lambda row: row["us"])
    assert r["kernel"] == longest["kernel"]            # the dominant kernel is the longest one, not the best one
    assert all(row["bound"] in ("hbm", "latency", "alu") for row in rows)


def test_loss_line_small_share():
    line = _bench("--no-post", "--images-per-gpu", "2")
    assert KEYS <= set(line)
    assert line["metric"] == "PAA assign+loss images/sec" and line["value"] > 0 and line["steps"] == 3
    assert line["config"]["images_per_gpu"] == 2 and line["dtype"] == "f32" and line["vs_baseline"] is None
    assert line["e2e"]["h2d_bytes_per_step"] > 2 * 22400 * 85 * 4 and line["e2e"]["value"] > 0
    assert line["gpu_launches"] >= 6 * 3
    _check_roofline(line["roofline"])


def test_post_line_small_share():
    line = _bench("--metric", "post", "--images-per-gpu", "2")
    assert KEYS <= set(line)
    assert line["metric"] == "PAA NMS+voting images/sec" and line["value"] > 0
    assert line["config"]["config"] == "C4" and line["config"]["images_per_gpu"] == 2
    assert line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] > 0
    _check_roofline(line["roofline"])
