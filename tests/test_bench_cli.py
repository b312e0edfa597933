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


def test_roofline_block_reports_the_bounding_pipe(monkeypatch):
    """HBM-bound kernels report bytes / time against the measured copy peak; a kernel listed in
    profiles/ncu_pipes.json (per-pixel Metropolis: an execution pipe bounds it) reports that pipe, with the
    HBM figures kept beside it."""
    sys.path.insert(0, ROOT)
    import bench
    stats = {"cg_pass_kernel": {"launches": 2, "ms": 2.0, "bytes": 10.0e9},
             "mh_perpixel_kernel": {"launches": 2, "ms": 20.0, "bytes": 4.0e9},
             "scalar kernels": {"launches": 4, "ms": 0.1, "bytes": 0.0}}
    monkeypatch.setattr(bench, "ncu_pipe", lambda k, c: None)
    monkeypatch.setattr(bench, "ncu_traffic", lambda k: 4_000_000_000 if k == "cg_pass_kernel" else None)
    r = bench.roofline_block("cg_pass_kernel", stats, "c2", 6388.0, "measured", capture_share=1.0)
    assert r["bound"] == "hbm" and r["achieved"] == 5000.0 and abs(r["frac"] - 5000.0 / 6388.0) < 1e-4
    assert r["traffic"] == 4_000_000_000
    assert bench.roofline_block("cg_pass_kernel", stats, "c2", 6388.0, "measured", capture_share=0.125)["traffic"] == 500_000_000
    assert bench.roofline_block("cg_pass_kernel", stats, "c2", 6388.0, "measured")["traffic"] is None   # another workload
    assert r["per_kernel"]["scalar kernels"]["GBps"] is None and r["avg_launch_us"] == 1000.0
    monkeypatch.setattr(bench, "ncu_pipe", lambda k, c: {"bound": "fp64", "busy_pct": 41.5, "issue_active_pct": 60.0,
                                                         "capture": "profiles/x.csv"} if (k, c) == ("mh_perpixel_kernel", "c3") else None)
    r = bench.roofline_block("mh_perpixel_kernel", stats, "c3", 6388.0, "measured")
    assert r["bound"] == "fp64" and r["achieved"] == 41.5 and r["peak"] == 100.0 and r["frac"] == 0.415
    assert r["hbm"]["achieved"] == 200.0 and r["avg_launch_us"] == 10000.0
    r = bench.roofline_block("mh_perpixel_kernel", stats, "c4", 6388.0, "measured")
    assert r["bound"] == "hbm"
