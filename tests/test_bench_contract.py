"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the reference's loop on the host
cores) prints exactly ONE JSON line on stdout with the keys the driver reads, and the GPU arm refuses to run without a
device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                          env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pairs/sec @1024 pts" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("VCR-Net partial-to-partial")          # BASELINE.json configs[1]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return                                          # the GPU arm itself is exercised by the driver / gpu_round.sh
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in r.stderr
