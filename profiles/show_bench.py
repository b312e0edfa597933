"""Pretty-print the bench.py JSON line(s) found in a log (per-kernel CUDA-event timings)."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        e = d.get("e2e", {})
        print(path, "|", d["config"]["workload"][:40], "| gpus", d["n_gpus"], "| it/s", d["value"], "| ms/step", d["ms_per_step"],
              "| n_cg", d["config"].get("n_cg_iterations"), "| e2e", e.get("value"), e.get("ms_per_step_device_rank0"),
              e.get("ms_per_step_wall_rank0"))
        for k, v in (d.get("roofline") or {}).get("per_kernel", {}).items():
            print(f"    {k:28s} {v}")
