"""Bare pinned-memory copy bandwidth on this box, all ranks copying at the same time:

    python scripts/pcie_probe.py                                   (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29555 scripts/pcie_probe.py                  (N GPUs, one process each)

Answers VERDICT r1 "weak" #8: is the end-to-end curve of bench.py (host buffers every step) limited by each
GPU's own host link, by the box's aggregate host bandwidth, or by something in the staging code?  Prints one
JSON line: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions together, at the sizes the
c2 bench moves per step and rank (eta up: 50 MB / N, amplitude planes down: 100 MB / N) and at 256 MB.
"""
import json
import os
import sys

import torch


def main():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"n_gpus": world, "cases": {}}
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    for label, up_mb, dn_mb in (("bench_step", 50.3 / world, 100.7 / world), ("256MB", 256.0, 256.0)):
        n_up, n_dn = int(up_mb * 1e6 / 8), int(dn_mb * 1e6 / 8)
        h_up = torch.empty(n_up, dtype=torch.float64).pin_memory()
        h_dn = torch.empty(n_dn, dtype=torch.float64).pin_memory()
        d_up = torch.empty(n_up, dtype=torch.float64, device="cuda")
        d_dn = torch.zeros(n_dn, dtype=torch.float64, device="cuda")
        res = {}
        for mode in ("h2d", "d2h", "both"):
            reps = 20
            for timed in (False, True):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s_up.wait_event(e0)
                s_dn.wait_event(e0)
                for _ in range(reps):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s_up):
                            d_up.copy_(h_up, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s_dn):
                            h_dn.copy_(d_dn, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s_up)
                torch.cuda.current_stream().wait_stream(s_dn)
                e1.record()
                barrier()
            ms = allmax(e0.elapsed_time(e1)) / reps
            by = (n_up * 8 if mode != "d2h" else 0) + (n_dn * 8 if mode != "h2d" else 0)
            res[mode] = {"ms_per_rep_max_over_ranks": round(ms, 4), "GBps_per_rank": round(by / ms / 1e6, 2),
                         "GBps_aggregate": round(world * by / ms / 1e6, 2)}
        out["cases"][label] = {"h2d_MB_per_rank": round(up_mb, 2), "d2h_MB_per_rank": round(dn_mb, 2), **res}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
