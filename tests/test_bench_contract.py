"""CPU: the reference arm of bench.py (`--impl reference`, the CPU port of the reference path timed on the host
cores) prints exactly one JSON line with the contract's keys; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONTRACT_KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                 "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(extra_env=None, extra_args=()):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1"] + list(extra_args), cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    return [ln for ln in proc.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert CONTRACT_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "PAA assign+loss images/sec"
    assert line["unit"] == "images/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] >= 1 and line["dtype"] == "f32" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_reference_arm_of_the_post_metric():
    lines = _run(extra_args=("--metric", "post"))
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert CONTRACT_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "PAA NMS+voting images/sec"
    assert line["config"]["config"] == "C4" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and "PAAPostProcessor.forward" in line["cpu_baseline"]["sample"]


def test_configs_name_baselines_global_batches():
    """BASELINE.json configs: C2 = 16 images (2/GPU x 8), C3 = 32, C4 / C5 = 64; the strong split is the default."""
    sys.path.insert(0, ROOT)
    import bench
    assert {k: v["images"] for k, v in bench.CONFIGS.items()} == {"C1": 2, "C2": 16, "C3": 32, "C4": 64, "C5": 64}
    a = bench.parse_args([])
    assert (a.metric, a.config, a.scaling, a.gpus) == ("loss", "C2", "strong", 1) and a.warmup >= 3
    assert bench.parse_args(["--metric", "post"]).config == "C4"
    assert bench.C2_BATCH_KW == dict(num_images=16, seed=2000, image_hw=(800, 1333), gt_per_image=(1, 100))
