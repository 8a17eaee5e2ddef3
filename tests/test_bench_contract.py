"""bench.py's contract on a machine without a GPU: the reference arm (`--impl reference`, the reference's op sequence on the host
cores through the oracle) prints ONE JSON line with the keys the driver reads, and the CUDA arm refuses to run without a device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env,
                          cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--budget", "5")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ppft_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == pytest.approx(d["value"]) and "PPFT step" in cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == pytest.approx(d["value"]) and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a CUDA device")
def test_cuda_arm_refuses_to_run_without_a_device():
    res = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert res.returncode != 0
    assert "no CUDA device" in (res.stderr + res.stdout)
    assert not [l for l in res.stdout.splitlines() if l.strip().startswith("{")]      # no number from a fallback
