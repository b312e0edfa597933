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


def intensity_case(nside=8, seed=11, noise=True, with_hi=False, unfitted_band=None):
    """A Stokes-I run (TQU = 'T', CG_POLTYPE = 'T'): thermal dust (mbb, diffuse) + a `monopole` component (one offset
    per fitted band, src/dang_component_mod.f90:579-597) and optionally an `hi_fit` component
    (s = A_nu * HI(p) * B_nu(T_d(p)) in RJ units, :599-700), the Stokes-I border rows of compute_Ax
    (src/dang_cg_mod.f90:717-744, :833-866).  Returns (cfg, sky, mono_true [nbands], hi_true [nbands] or None)."""
    from dang_b200.config import Band, CGGroup, Component, IndexSpec, RunConfig
    from dang_b200.synth import Sky, band_sigma, sed_mbb
    bands = [Band(nu) for nu in (100.0, 143.0, 217.0, 353.0, 545.0, 857.0)]
    nb = len(bands)
    dust = Component(label="dust", type="mbb", nu_ref_ghz=353.0, cg_group=1, amp_sample=True,
                     indices=[IndexSpec("BETA", init=1.55, sample=False, poltype="T"),
                              IndexSpec("T", init=19.6, sample=False, poltype="T")])
    corr = [j != unfitted_band for j in range(nb)]
    mono = Component(label="monopole", type="monopole", nu_ref_ghz=353.0, cg_group=1, amp_sample=True, indices=[], corr=corr)
    comps = [dust, mono]
    if with_hi:
        comps = [dust, Component(label="hi", type="hi_fit", nu_ref_ghz=353.0, cg_group=1, amp_sample=True,
                                 indices=[IndexSpec("T", init=20.0, sample=False, region="per-pixel", prior="gaussian",
                                                    gauss=(20.0, 2.0), uni=(10.0, 40.0), step=0.5, poltype="T")],
                                 corr=[True] * nb), mono]
        dust.amp_sample = False  # the HI template stands for the dust in the fit (as in the reference's HI runs)
    cfg = RunConfig("intensity", nside, bands, comps, [CGGroup(sample=True, max_iter=400, converge=1e-18, poltype="T")],
                    nsample=10, ngibbs=3, tqu="T")
    npix = cfg.npix
    rng = np.random.default_rng(seed)
    a_d = np.zeros((3, npix))
    a_d[0] = np.abs(rng.normal(0.0, 50.0, npix)) + 5.0
    mono_true = np.array([3.0 + 0.7 * j if corr[j] else 0.0 for j in range(nb)])
    hi_map = np.zeros((3, npix))
    hi_map[0] = np.abs(rng.normal(1.0, 0.5, npix)) + 0.1
    t_hi = 20.0 + 1.5 * rng.standard_normal(npix)
    hi_true = np.array([2.0e-2 * (b.nu_ghz / 353.0) ** 1.5 for b in bands]) if with_hi else None
    sig = np.zeros((nb, 3, npix))
    rms = np.ones((nb, 3, npix))
    for j, b in enumerate(bands):
        if not with_hi:
            sig[j, 0] = a_d[0] * sed_mbb(b.nu_ghz, 353.0, 1.55, 19.6)
        else:
            from oracle.binding import planck_rj
            sig[j, 0] = hi_true[j] * hi_map[0] * planck_rj(b.nu_ghz * 1e9, t_hi)
        sig[j, 0] += mono_true[j]
        rms[j, 0] = band_sigma(b.nu_ghz) * (1.0 + 0.3 * rng.random(npix))
        if noise:
            sig[j, 0] += rms[j, 0] * rng.standard_normal(npix)
    from dang_b200.healpix import pix2z_ring
    z = pix2z_ring(nside, np.arange(npix))
    mask = np.where(np.abs(z) < np.sin(np.deg2rad(5.0)), 0.0, 1.0)
    amp0 = {"dust": a_d * (1.0 + 0.05 * rng.standard_normal((3, npix))) if not with_hi else np.zeros((3, npix)),
            "monopole": np.zeros((3, npix)), "hi": np.zeros((3, npix))}
    idx0 = {"dust": np.stack([np.full((3, npix), 1.55), np.full((3, npix), 19.6)]), "monopole": np.zeros((0, 3, npix)),
            "hi": np.stack([np.tile(t_hi, (3, 1))])}
    sky = Sky(sig=sig, rms=rms, mask=mask, gain=np.ones(nb), offset=np.zeros(nb), amplitude=amp0, indices=idx0,
              truth={"dust": a_d})
    sky.template = {"hi": hi_map, "monopole": None}
    sky.template_amplitudes = {"hi": np.zeros((3, nb)), "monopole": np.zeros((3, nb))}
    return cfg, sky, mono_true, hi_true
