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
