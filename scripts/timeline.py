"""Per-stream event log of a few Gibbs iterations (no nsys in this image): every kernel launch on the compute stream
and every staged copy on the copy streams, with start / end from CUDA events (DANG_OPT_PROFILE, dang_gpu_timeline).

    python scripts/timeline.py [--mode value|e2e] [--steps 4] [--out profiles/xyz.md]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29577 \
        scripts/timeline.py ...                         (rank 0 prints; every rank runs the same loop)

`value`: the device-resident loop of bench.py; `e2e`: its host-buffer loop (deviates up, maps down, every step)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="value")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--nside", type=int, default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--defer", type=int, default=0, help="1: DANG_OPT_DEFER_SCALARS (value mode)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench import pinned_array
    from dang_b200.engine import OPT_PROFILE, Engine, setup_torch_comm
    from dang_b200.healpix import ring_partition
    from dang_b200.synth import make_config, make_sky
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = make_config(args.config, nside=args.nside)
    sky = make_sky(cfg)
    bounds = ring_partition(cfg.nside, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    eng = Engine(cfg, sky, device=local, pix_range=(lo, hi))
    if world > 1:
        setup_torch_comm(eng, mailboxes=os.environ.get("DANG_GPU_MAILBOX", "1") != "0")
    if world > 1 and os.environ.get("DANG_GPU_MAILBOX", "1") != "0":
        us = C.c_double()
        for cnt in (4, 32):
            eng._ck(eng.lib.dang_gpu_comm_probe(eng.h, 2000, cnt, C.byref(us)))
            if rank == 0:
                print(json.dumps({"comm_probe": f"{cnt} doubles over NVLink mailboxes, {world} ranks", "us_per_exchange": round(us.value, 3)}), flush=True)
    npix = cfg.npix
    eta_hs = [pinned_array(eng.lib, (2 * npix,)) for _ in range(3)]
    for i, e in enumerate(eta_hs):
        e[:] = np.random.default_rng(1 + i).standard_normal(2 * npix)
    amp_h = [pinned_array(eng.lib, (cfg.nmaps, npix)) for _ in cfg.comps]
    z_h, u_h = [pinned_array(eng.lib, (cfg.nsample,))], [pinned_array(eng.lib, (cfg.nsample,))]
    z_h[0][:] = np.random.default_rng(2).standard_normal(cfg.nsample)
    u_h[0][:] = np.random.default_rng(3).random(cfg.nsample)

    pending = []

    def step(it):
        if args.mode == "value":
            r1, _ = eng.gibbs_iteration(2 + it, seed=it)
            if r1[0][0] == -1:  # deferred scalars: read the previous iteration's numbers
                pending.append(eng.iteration_mark())
                if len(pending) > 1:
                    eng.iteration_scalars(pending.pop(0))
        else:
            eng.stage_eta(eta_hs[(it + 1) % 3])
            eng.sample_cg_groups(eta=None)
            for ic in range(len(cfg.comps)):
                eng.amplitude_async(ic, amp_h[ic])
            eng.sample_spectral_parameters(z=z_h, u=u_h, seed=7 + 2 * it)
            eng.index_fullsky(1, 0, 2)
            eng.index_fullsky(1, 0, 3)

    if args.mode != "value":
        eng.stage_eta(eta_hs[0])
    step(0)
    if args.defer and args.mode == "value":
        from dang_b200.engine import OPT_DEFER_SCALARS
        eng.set_option(OPT_DEFER_SCALARS, 1)
    for w in range(1, 4):
        step(w)
    eng.download_wait()
    eng.sync()
    if world > 1:
        dist.barrier()
    eng.set_option(OPT_PROFILE, 1)
    import time
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(10 + k)
    while pending:
        eng.iteration_scalars(pending.pop(0))
    eng.download_wait()
    eng.sync()
    wall_us = (time.perf_counter() - t0) * 1e6
    n = C.c_int()
    kind = (C.c_int * 4096)()
    a0 = np.zeros(4096)
    a1 = np.zeros(4096)
    eng._ck(eng.lib.dang_gpu_timeline(eng.h, 4096, C.byref(n), kind, a0.ctypes.data_as(C.POINTER(C.c_double)),
                                      a1.ctypes.data_as(C.POINTER(C.c_double))))
    eng.set_option(OPT_PROFILE, 0)
    if rank == 0:
        rows = [(eng.lib.dang_gpu_kernel_name(kind[i]).decode(), a0[i], a1[i]) for i in range(n.value)]
        span = max(r[2] for r in rows) - min(r[1] for r in rows)
        lines = [f"# Event log: config {cfg.name} nside {cfg.nside}, {world} GPU(s), mode `{args.mode}`, {args.steps} Gibbs iterations (rank 0)\n",
                 f"host wall time of the loop {wall_us:.0f} us ({wall_us / args.steps:.0f} us / iteration); device span of the log {span:.0f} us "
                 f"(profiling adds two event records per launch, so absolute times are a few per cent above bench.py's)\n",
                 "| # | stream | what | start us | end us | duration us | gap before (same stream) us |", "|---:|---|---|---:|---:|---:|---:|"]
        last_end = {}
        busy = {}
        for i, (name, s, e) in enumerate(rows):
            stream = "copy h2d" if name.startswith("h2d") else "copy d2h" if name.startswith("d2h") else "compute"
            gap = s - last_end[stream] if stream in last_end else 0.0
            last_end[stream] = e
            busy[stream] = busy.get(stream, 0.0) + (e - s)
            lines.append(f"| {i} | {stream} | {name} | {s:.1f} | {e:.1f} | {e - s:.1f} | {gap:.1f} |")
        lines.append("")
        for st, b in busy.items():
            lines.append(f"* `{st}` busy {b:.0f} us of {span:.0f} us ({100 * b / span:.1f} %)")
        text = "\n".join(lines)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            open(args.out, "w").write(text + "\n")
        print(text if len(rows) < 80 else "\n".join(lines[:3] + lines[-8:]))
        print(json.dumps({"n_gpus": world, "mode": args.mode, "us_per_iteration_wall": round(wall_us / args.steps, 1),
                          "busy_us": {k: round(v, 1) for k, v in busy.items()}, "span_us": round(span, 1)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
