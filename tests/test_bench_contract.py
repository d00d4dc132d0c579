"""bench.py contract on a machine without a GPU: the reference arm prints ONE JSON line with the agreed keys
(a bounded CPU run of the oracle port), non-zero ranks of a torchrun launch stay silent, and the CUDA arm refuses
to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.pop("RANK", None), e.pop("WORLD_SIZE", None), e.pop("LOCAL_RANK", None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    res = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("GripNet fwd+bwd edges/sec") and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 2 and d["gpu_launches"] == 0
    assert d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "epochs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "2"], env={"RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = _run(["--steps", "2"], timeout=300)
    assert res.returncode != 0 and res.stdout.strip() == ""
    assert "no CUDA device" in res.stderr and "no CPU fallback" in res.stderr
