"""Seeded synthetic Q/U skies and the BASELINE.json configurations.

There is no network and the reference ships no data, so every run uses synthetic maps:
RING order, nmaps = 3 with an empty I plane, rms_nu(pix) = sigma_nu (1 + 0.3 U[0,1)),
Galactic-plane mask |b| < 5 deg, sky = power-law synchrotron + modified-blackbody dust.
Arrays come back in the reference's Fortran layout `A(0:npix-1, nmaps, nbands)`, i.e.
C order `[band][stokes][pix]`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np

from .config import Band, CGGroup, Component, IndexSpec, RunConfig
from .healpix import pix2z_ring

# src/dang_util_mod.f90:12-13
K_B = 1.3806503e-23
H_PLANCK = 1.0545726691251021e-34 * 2.0 * np.pi

SEED_SKY, SEED_NOISE, SEED_DEVIATES = 20260101, 20260102, 20260103


def _tabulated_band(nu_ghz: float, n_bp: int) -> Band:
    """Top-hat with Gaussian edges over +-15 % of nu_c, n_bp samples (config c3)."""
    nu = np.linspace(0.85 * nu_ghz, 1.15 * nu_ghz, n_bp)
    x = (nu - nu_ghz) / (0.15 * nu_ghz)
    tau = np.where(np.abs(x) < 0.7, 1.0, np.exp(-0.5 * ((np.abs(x) - 0.7) / 0.1) ** 2))
    return Band(nu_ghz=nu_ghz, label=f"{nu_ghz:g}", bp_nu_ghz=nu, bp_tau=tau)


def _synch(beta_sample: bool, region: str = "per-pixel") -> Component:
    return Component(
        label="synch", type="power-law", nu_ref_ghz=30.0, cg_group=1, amp_sample=True,
        indices=[IndexSpec("BETA", init=-3.0, sample=beta_sample, region=region, prior="gaussian",
                           gauss=(-3.1, 0.1), uni=(-4.0, -2.0), step=0.05, poltype="Q+U")])


def _dust(beta_sample: bool, t_sample: bool, region: str = "per-pixel", step_beta: float = 0.05) -> Component:
    return Component(
        label="dust", type="mbb", nu_ref_ghz=353.0, cg_group=1, amp_sample=True,
        indices=[IndexSpec("BETA", init=1.5, sample=beta_sample, region=region, prior="gaussian",
                           gauss=(1.55, 0.1), uni=(1.0, 2.2), step=step_beta, poltype="Q+U"),
                 IndexSpec("T", init=19.0, sample=t_sample, region=region, prior="gaussian",
                           gauss=(19.6, 1.0), uni=(10.0, 35.0), step=0.5, poltype="Q+U")])


def make_config(name: str, nside: int | None = None) -> RunConfig:
    """The five BASELINE.json configurations; `nside` overrides the map size for small tests."""
    cg = [CGGroup(sample=True, max_iter=100, converge=1e-12, poltype="Q+U")]
    if name == "c1":
        bands = [Band(nu) for nu in (23.0, 30.0, 44.0, 70.0, 353.0)]
        return RunConfig("c1", nside or 64, bands, [_synch(True), _dust(False, False)], cg,
                         nsample=50, ngibbs=10)
    c2_nu = (22.8, 28.4, 33.0, 40.6, 44.1, 60.8, 70.4, 353.0)
    if name in ("c2", "c5"):
        bands = [Band(nu) for nu in c2_nu]
        return RunConfig(name, nside or 512, bands,
                         [_synch(False), _dust(True, False, region="fullsky", step_beta=0.01)], cg,
                         nsample=20, ngibbs=10)
    if name == "c3":
        bands = [_tabulated_band(nu, 128) for nu in c2_nu + (93.5, 100.0, 143.0, 217.0)]
        return RunConfig("c3", nside or 1024, bands, [_synch(True), _dust(True, False)], cg,
                         nsample=20, ngibbs=10)
    if name == "c4":
        nus = np.exp(np.linspace(np.log(20.0), np.log(857.0), 20))
        bands = [Band(float(nu)) for nu in nus]
        return RunConfig("c4", nside or 2048, bands, [_synch(False), _dust(True, True)], cg,
                         nsample=20, ngibbs=10)
    raise ValueError(f"unknown config {name!r}")


_SIGMA_C1 = {23.0: 2.0, 30.0: 2.5, 44.0: 3.0, 70.0: 3.5, 353.0: 1.0}


def band_sigma(nu_ghz: float) -> float:
    if nu_ghz in _SIGMA_C1:
        return _SIGMA_C1[nu_ghz]
    # smooth WMAP/Planck-like noise levels (uK_RJ): best near 30-70 GHz and at 353 for dust
    return float(1.0 + 2.5 * np.exp(-0.5 * (np.log(nu_ghz / 60.0) / 0.8) ** 2))


def sed_powerlaw(nu_ghz, nu_ref_ghz, beta):
    return (np.asarray(nu_ghz) / nu_ref_ghz) ** beta


def sed_mbb(nu_ghz, nu_ref_ghz, beta, td):
    z = H_PLANCK / (K_B * td)
    nu, nu0 = np.asarray(nu_ghz) * 1e9, nu_ref_ghz * 1e9
    return (np.exp(z * nu0) - 1.0) / (np.exp(z * nu) - 1.0) * (nu / nu0) ** (beta + 1.0)


def band_sed(band: Band, comp: Component, *theta) -> float:
    f = sed_powerlaw if comp.type == "power-law" else sed_mbb
    if band.is_delta:
        return float(f(band.nu_ghz, comp.nu_ref_ghz, *theta))
    tau = band.bp_tau / band.bp_tau.sum()
    return float(np.sum(tau * f(band.bp_nu_ghz, comp.nu_ref_ghz, *theta)))


@dataclass
class Sky:
    sig: np.ndarray       # [nbands][nmaps][npix]
    rms: np.ndarray       # [nbands][nmaps][npix]
    mask: np.ndarray      # [npix]
    gain: np.ndarray      # [nbands]
    offset: np.ndarray    # [nbands]
    amplitude: Dict[str, np.ndarray]   # initial component amplitudes [nmaps][npix]
    indices: Dict[str, np.ndarray]     # initial index maps [nindices][nmaps][npix]
    truth: Dict[str, np.ndarray]       # true amplitudes used to build the sky
    template: Dict[str, np.ndarray] = None             # type 'template': c%template [nmaps][npix] (max = 1 per plane)
    template_amplitudes: Dict[str, np.ndarray] = None  # c%template_amplitudes [nmaps][nbands]


TRUE_THETA = {"synch": (-3.1,), "dust": (1.55, 19.6)}
TRUE_AMP_SIGMA = {"synch": 10.0, "dust": 50.0}


def make_sky_slice(cfg: RunConfig, lo: int, hi: int, seed: int = SEED_SKY, noise_seed: int = SEED_NOISE) -> Sky:
    """The same sky model for ONE rank of a large sharded run: full-sky-shaped arrays (the C ABI takes
    the reference's full-size allocatables) of which only pixels [lo, hi) are ever written, so the
    untouched pages of the calloc'ed arrays never become resident (nside 2048 x 20 bands is 48 GB of
    maps; eight ranks each building the full sky would not fit the host).  Draws are seeded per
    (array, band, plane, lo), so the slices of different ranks are independent; they are NOT the
    numbers make_sky() draws -- parity tests use make_sky(), benches of big configs use this."""
    npix, nb, nm = cfg.npix, cfg.nbands, cfg.nmaps
    n = hi - lo
    truth, amp0, idx0 = {}, {}, {}
    sig = np.zeros((nb, nm, npix))
    rms = np.zeros((nb, nm, npix))
    for ic, c in enumerate(cfg.comps):
        amp0[c.label] = np.zeros((nm, npix))
        idx0[c.label] = np.zeros((len(c.indices), nm, npix))
        for k, s in enumerate(c.indices):
            idx0[c.label][k][:, lo:hi] = s.init
        if c.label in TRUE_THETA:
            rng = np.random.default_rng([seed, 1, ic, lo])
            a = rng.normal(0.0, TRUE_AMP_SIGMA[c.label], size=(2, n))
            for j, b in enumerate(cfg.bands):
                sig[j, 1:3, lo:hi] += a * band_sed(b, c, *TRUE_THETA[c.label])
        truth[c.label] = None
    # one noise realisation per plane, rolled by a band-dependent lag (cheap on the host, and every
    # (pixel, band) still sees an independent-looking draw)
    nrng = np.random.default_rng([noise_seed, 2, lo])
    q0, g0 = nrng.random(size=(2, n)), nrng.standard_normal(size=(2, n))
    for j, b in enumerate(cfg.bands):
        r = band_sigma(b.nu_ghz) * (1.0 + 0.3 * np.roll(q0, 7919 * (j + 1), axis=1))
        rms[j, :, lo:hi] = 1.0
        rms[j, 1:3, lo:hi] = r
        sig[j, 1:3, lo:hi] += r * np.roll(g0, 104729 * (j + 1), axis=1)
    mask = np.zeros(npix)
    z = pix2z_ring(cfg.nside, np.arange(lo, hi))
    mask[lo:hi] = np.where(np.abs(z) < np.sin(np.deg2rad(5.0)), 0.0, 1.0)
    return Sky(sig=sig, rms=rms, mask=mask, gain=np.ones(nb), offset=np.zeros(nb),
               amplitude=amp0, indices=idx0, truth=truth)


def make_sky(cfg: RunConfig, seed: int = SEED_SKY, noise_seed: int = SEED_NOISE) -> Sky:
    npix, nb, nm = cfg.npix, cfg.nbands, cfg.nmaps
    rng = np.random.default_rng(seed)
    nrng = np.random.default_rng(noise_seed)
    truth, amp0, idx0 = {}, {}, {}
    sig = np.zeros((nb, nm, npix))
    for c in cfg.comps:
        a = np.zeros((nm, npix))
        amp0[c.label] = np.zeros((nm, npix))
        idx0[c.label] = (np.stack([np.full((nm, npix), s.init) for s in c.indices]) if c.indices
                         else np.zeros((0, nm, npix)))
        if c.label in TRUE_THETA:  # synch / dust carry the simulated sky; other components start empty
            a[1:3] = rng.normal(0.0, TRUE_AMP_SIGMA[c.label], size=(2, npix))
            for j, b in enumerate(cfg.bands):
                sig[j, 1:3] += a[1:3] * band_sed(b, c, *TRUE_THETA[c.label])
        truth[c.label] = a
    rms = np.ones((nb, nm, npix))
    for j, b in enumerate(cfg.bands):
        rms[j, 1:3] = band_sigma(b.nu_ghz) * (1.0 + 0.3 * nrng.random(size=(2, npix)))
        sig[j, 1:3] += rms[j, 1:3] * nrng.standard_normal(size=(2, npix))
    z = pix2z_ring(cfg.nside, np.arange(npix))
    mask = np.where(np.abs(z) < np.sin(np.deg2rad(5.0)), 0.0, 1.0)
    return Sky(sig=sig, rms=rms, mask=mask, gain=np.ones(nb), offset=np.zeros(nb),
               amplitude=amp0, indices=idx0, truth=truth)
