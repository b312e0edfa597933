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


def run_intensity_case():
    """Round-2 oracle features: dust + HI template (`hi_fit`) + `monopole` in one Stokes-I CG group (two amplitude
    draws, band monopoles, a full-sky T_d draw of the HI component) and the three resolution operators."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import intensity_case
    from oracle.binding import Oracle, load, udgrade
    load().ora_set_T_CMB(2.7255)
    cfg, sky, _, _ = intensity_case(4, with_hi=True)
    cfg.cg_groups[0].converge, cfg.cg_groups[0].max_iter = 1e-22, 500
    spec = cfg.comps[1].indices[0]
    spec.sample, spec.region, spec.prior = True, "fullsky", "gaussian"
    sky.indices["hi"][0][:] = 20.0
    ora = Oracle(cfg, sky)
    rng = np.random.default_rng(20260105)
    out = {}
    for it in (1, 2):
        ora.sample_cg_group(0, 1, rng.standard_normal(cfg.npix))
        out[f"i{it}_chisq"] = np.array(ora.compute_chisq()[0])
        out[f"i{it}_hi_amp"] = ora.template_amplitudes(1).copy()
        out[f"i{it}_monopole"] = ora.template_amplitudes(len(cfg.comps) - 1).copy()
        ora.update_sky_model()
        out[f"i{it}_res"] = ora.res_map().copy()
        z, u = rng.standard_normal(8), rng.random(8)
        acc, dec, _ = ora.sample_index_mh(1, 0, 1, 8, 1, z, u, want_trace=True)
        out[f"i{it}_dec"] = np.array(dec[:8])
        out[f"i{it}_T"] = ora.indices(1)[0][0][:1].copy()
    m = rng.standard_normal((2, 12 * 8 * 8))               # [nmaps][npix]
    m[0, 5] = -1.6375e30                                   # a bad pixel
    out["ud_ring_8_2"] = udgrade("ring", m, 8, 2)
    out["ud_ring_2_4"] = udgrade("ring", out["ud_ring_8_2"], 2, 4)
    out["ud_rms_8_4"] = udgrade("rms", np.abs(rng.standard_normal((1, 12 * 64))) + 0.1, 8, 4)
    out["ud_mask_8_2"] = udgrade("mask", (rng.random((1, 12 * 64)) > 0.3).astype(float), 8, 2, threshold=0.5)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "intensity_nside4.npz"), **run_intensity_case())
    print("wrote intensity_nside4.npz")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_nside4.npz"), **run_case())
    print("wrote c1_nside4.npz")
