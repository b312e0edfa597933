"""Same-box A/B of kernel variants on the headline config (boxes differ by a few per cent, so variants are only
ever compared inside one gpurun call):   python scripts/ab_probe.py [--config c2] [--nside 512] [--steps 10]

Prints, per variant, ms per Gibbs iteration (CUDA events over `steps` iterations, device-resident) and the average
launch time of every kernel family (DANG_OPT_PROFILE)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--nside", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variants", default="old,persistent,ring,ring+persistent,ring+persistent+eta,ring+persistent+optimize")
    args = ap.parse_args()
    from dang_b200.engine import OPT_CG_PERSISTENT, OPT_PROFILE, OPT_STREAM_RING, Engine
    from dang_b200.synth import make_config, make_sky
    cfg = make_config(args.config, nside=args.nside)
    sky = make_sky(cfg)
    rng = np.random.default_rng(1)
    eta_h = rng.standard_normal(2 * cfg.npix)
    for name in args.variants.split(","):
        eng = Engine(cfg, sky)
        eng.set_option(OPT_STREAM_RING, int("ring" in name))
        eng.set_option(OPT_CG_PERSISTENT, int("persistent" in name))
        for kv in [t for t in name.split("+") if "=" in t]:
            k, v = kv.split("=")
            eng.set_option(int(k), float(v))
        eta = eta_h if "eta" in name else None
        mode = "optimize" if "optimize" in name else "sample"

        def step(it):
            r1 = eng.sample_cg_groups(ml_mode=mode, eta=eta, seed=2 * it)
            eng.sample_spectral_parameters(seed=2 * it + 1)
            return r1[0][0]

        for w in range(3):
            step(w)
        eng.sync()
        eng.event_record(0)
        for k in range(args.steps):
            n_cg = step(3 + k)
        eng.event_record(1)
        eng.sync()
        ms = eng.event_elapsed_ms(0, 1) / args.steps
        eng.kernel_stats(reset=True)
        eng.set_option(OPT_PROFILE, 1)
        for k in range(5):
            step(100 + k)
        st = eng.kernel_stats(reset=True)
        eng.set_option(OPT_PROFILE, 0)
        per = {k: {"n": v["launches"], "avg_us": round(1e3 * v["ms"] / max(v["launches"], 1), 2),
                   "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] > 0 else None}
               for k, v in st.items() if v["launches"]}
        print(json.dumps({"variant": name, "ms_per_step": round(ms, 4), "it_per_s": round(1e3 / ms, 2), "n_cg": n_cg,
                          "chisq": eng.compute_chisq(), "kernels": per}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
