"""bench.py's CPU legs (no GPU): the reference arm prints one JSON line with the contract's keys, on rank 0 only
under a multi-rank launch, and the product arm fails loudly without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--nside", "16", "--steps", "2", "--warmup", "1", "--gpus", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "it/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c2: nside=16" in d["config"]["workload"] and d["config"]["n_cg_iterations"]["min"] >= 2


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--nside", "16", "--steps", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--nside", "16", "--steps", "1", "--warmup", "1", "--no-cpu"])
    assert r.returncode != 0
    assert "dang_gpu" in (r.stderr + r.stdout) or "CUDA" in (r.stderr + r.stdout)
