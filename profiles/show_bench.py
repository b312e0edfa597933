"""Pretty-print one bench.py JSON line (per-kernel CUDA-event timings)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read())
print(d["config"]["workload"][:60], "| it/s", d["value"], "| ms/step", d["ms_per_step"], "| n_cg", d["config"]["n_cg_iterations"],
      "| e2e", d["e2e"]["value"])
for k, v in d["roofline"]["per_kernel"].items():
    print(f"    {k:28s} {v}")
