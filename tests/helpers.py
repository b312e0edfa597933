"""Shared builders for the parity tests: a small synthetic sky + matching Oracle / Engine."""
import copy

import numpy as np

from dang_b200.synth import make_config, make_sky


def small_case(name="c1", nside=16, seed=20260101, perturb=True):
    cfg = make_config(name, nside=nside)
    sky = make_sky(cfg, seed=seed, noise_seed=seed + 1)
    if perturb:
        # start from non-trivial amplitudes / index maps so every code path sees real numbers
        rng = np.random.default_rng(seed + 2)
        for c in cfg.comps:
            sky.amplitude[c.label][1:3] = sky.truth[c.label][1:3] * (1.0 + 0.05 * rng.standard_normal((2, cfg.npix)))
            for k, s in enumerate(c.indices):
                if s.region == "per-pixel":
                    v = s.init + 0.02 * rng.standard_normal(cfg.npix)
                    sky.indices[c.label][k][:] = v[None, :]
    return cfg, sky


def deviates(cfg, nsample, ncalls=1, seed=20260103):
    rng = np.random.default_rng(seed)
    n = nsample * cfg.npix * ncalls
    return rng.standard_normal(n), rng.random(n)


def clone_sky(sky):
    return copy.deepcopy(sky)


def template_case(nside=8, seed=7, noise=True, unfitted_band=2):
    """Synchrotron (diffuse, power law) + a dust TEMPLATE component fitted per band (the configuration of
    arXiv:2201.03530, src/dang_cg_mod.f90 template branches): sig = a_s sed_s + tamp_j T(p) + noise.
    Returns (cfg, sky, tamp_true [nbands])."""
    from dang_b200.config import Component
    from dang_b200.synth import band_sed, band_sigma
    cfg = make_config("c2", nside=nside)
    synch = cfg.comps[0]
    nb, npix = cfg.nbands, cfg.npix
    corr = [j != unfitted_band for j in range(nb)]
    cfg.comps = [synch, Component(label="dust", type="template", nu_ref_ghz=353.0, cg_group=1, amp_sample=True,
                                  indices=[], corr=corr)]
    sky = make_sky(cfg, seed=seed, noise_seed=seed + 1)   # synch only (dust has no entry in the simulated sky)
    rng = np.random.default_rng(seed + 5)
    T = np.ones((3, npix))
    T[1:3] = np.abs(rng.normal(0.0, 1.0, size=(2, npix))) + 0.05
    T[1] /= T[1].max()
    T[2] /= T[2].max()
    tamp_true = np.array([40.0 * (cfg.bands[j].nu_ghz / 353.0) ** 1.6 if corr[j] else 0.0 for j in range(nb)])
    sig = np.zeros_like(sky.sig)
    a_s = sky.truth["synch"]
    for j, b in enumerate(cfg.bands):
        sig[j, 1:3] = a_s[1:3] * band_sed(b, synch, -3.1) + tamp_true[j] * T[1:3]
        if noise:
            sig[j, 1:3] += sky.rms[j, 1:3] * rng.standard_normal((2, npix))
    sky.sig = sig
    sky.indices["synch"][0][:] = -3.1
    sky.template = {"dust": T}
    sky.template_amplitudes = {"dust": np.zeros((3, nb))}
    sky.amplitude["dust"] = np.zeros((3, npix))
    sky.indices["dust"] = np.zeros((0, 3, npix))
    return cfg, sky, tamp_true
