"""Generates tests/golden/c1_nside4.npz from the CPU oracle (NOT from the reference, which cannot
be built or run here): three Gibbs iterations of config c1 at nside 4 with seeded injected
deviates.  Re-run with `python tests/golden/make_golden.py` only when the oracle is deliberately
changed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def run_case():
    from dang_b200.synth import make_config, make_sky
    from oracle.binding import Oracle
    cfg = make_config("c1", nside=4)
    sky = make_sky(cfg)
    ora = Oracle(cfg, sky)
    rng = np.random.default_rng(20260103)
    nsample = 10
    out = {}
    for it in (1, 2, 3):
        eta = rng.standard_normal(2 * cfg.npix)
        its, delta = ora.sample_cg_group(0, 1, eta)
        out[f"it{it}_n_cg"] = np.array(its)
        out[f"it{it}_chisq_cg"] = np.array(ora.compute_chisq()[0])
        if it > 1:
            z, u = rng.standard_normal(nsample * cfg.npix), rng.random(nsample * cfg.npix)
            ora.sample_spectral_parameters(nsample, 1, z, u)
            out[f"it{it}_chisq_mh"] = np.array(ora.compute_chisq()[0])
        out[f"it{it}_amp_synch"] = ora.amplitude(0).copy()
        out[f"it{it}_amp_dust"] = ora.amplitude(1).copy()
        out[f"it{it}_beta_s"] = ora.indices(0).copy()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_nside4.npz"), **run_case())
    print("wrote c1_nside4.npz")
