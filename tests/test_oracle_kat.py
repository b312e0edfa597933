"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden vectors,
so these analytic cases plus line-by-line correspondence are all that anchors it)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from dang_b200.config import MISSVAL
from dang_b200.synth import H_PLANCK, K_B, TRUE_THETA, band_sed, make_config, make_sky
from helpers import small_case
from oracle.binding import Oracle, load, philox_normals, philox_uniforms


def truth_case(nside=4, noiseless=True):
    cfg = make_config("c1", nside=nside)
    sky = make_sky(cfg)
    for c in cfg.comps:
        for k, v in enumerate(TRUE_THETA[c.label]):
            sky.indices[c.label][k][:] = v
    if noiseless:
        sky.sig[:] = 0.0
        for c in cfg.comps:
            for j, b in enumerate(cfg.bands):
                sky.sig[j, 1:3] += sky.truth[c.label][1:3] * band_sed(b, c, *TRUE_THETA[c.label])
    return cfg, sky


def test_sed_identities():
    cfg, sky = truth_case()
    ora = Oracle(cfg, sky)
    # delta band at nu_ref: SED == 1 for both laws (bands 1 = 30 GHz synch ref, 4 = 353 GHz dust ref)
    assert ora.eval_sed(0, 1, 0, 2) == 1.0
    assert abs(ora.eval_sed(1, 4, 0, 2) - 1.0) < 1e-15
    # power law: (nu/nu_ref)**beta with beta from the index map or from theta
    assert abs(ora.eval_sed(0, 0, 0, 2) - (23.0 / 30.0) ** -3.1) < 1e-15
    assert abs(ora.eval_sed(0, 0, 0, 2, theta=[-2.5]) - (23.0 / 30.0) ** -2.5) < 1e-15
    # modified blackbody against an independent numpy evaluation
    z = H_PLANCK / (K_B * 19.6)
    ref = (np.exp(z * 353e9) - 1) / (np.exp(z * 44e9) - 1) * (44.0 / 353.0) ** 2.55
    assert abs(ora.eval_sed(1, 2, 5, 3) / ref - 1.0) < 1e-14


def test_bandpass_normalisation_and_delta_limit():
    from dang_b200.config import Band
    cfg, sky = truth_case()
    nu = np.linspace(43.999, 44.001, 9)
    cfg.bands[2] = Band(44.0, bp_nu_ghz=nu, bp_tau=np.full(9, 7.0))  # un-normalised weights
    ora = Oracle(cfg, sky)
    # tau0 is normalised to unit sum, so a very narrow top-hat reproduces the delta band
    assert abs(ora.eval_sed(0, 2, 0, 2) / (44.0 / 30.0) ** -3.1 - 1.0) < 1e-8
    assert abs(ora.eval_sed(1, 2, 0, 2) / band_sed(Band(44.0), cfg.comps[1], 1.55, 19.6) - 1.0) < 1e-8


def test_normal_prior_and_box_muller():
    lib = load()
    assert abs(lib.ora_eval_normal_prior(1.5, 1.5, 0.2) - 1.0 / (0.2 * np.sqrt(2 * np.pi))) < 1e-15
    assert abs(lib.ora_eval_normal_prior(1.7, 1.5, 0.2) / lib.ora_eval_normal_prior(1.5, 1.5, 0.2) - np.exp(-0.5)) < 1e-15
    # rand_normal: r = sqrt(-2 ln u1) = 1 at u1 = exp(-1/2); theta = 2 pi u2 = pi/2 at u2 = 1/4
    assert abs(lib.ora_rand_normal_from_uniform(3.0, 2.0, np.exp(-0.5), 0.25) - 5.0) < 1e-14
    assert abs(lib.ora_rand_normal_from_uniform(0.0, 1.0, 0.3, 0.5)) < 1e-15  # sine branch: sin(pi) = 0


def test_philox_known_answers_and_moments():
    lib = load()
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kats:
        c = (C.c_uint * 4)(*ctr)
        lib.ora_philox_raw(c, key[0], key[1])
        assert tuple(c) == out
    u = philox_uniforms(42, 3, 0, 200000)
    assert 0.0 < u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    z = philox_normals(42, 2, 0, 200000)
    assert abs(z.mean()) < 1e-2 and abs(z.std() - 1.0) < 1e-2
    assert abs(np.mean(z ** 4) - 3.0) < 0.1


def test_rhs_and_Ax_are_consistent_blocks():
    cfg, sky = truth_case(noiseless=True)
    ora = Oracle(cfg, sky)
    npix = cfg.npix
    m = sky.mask != 0
    # noiseless data: b == A a_true, where a_true is packed [synch Q, synch U, dust Q, dust U]
    a_true = np.concatenate([sky.truth["synch"][1], sky.truth["synch"][2], sky.truth["dust"][1], sky.truth["dust"][2]])
    a_true = a_true * np.tile(m, 4)
    b = ora.compute_rhs()
    assert rel_err(ora.compute_Ax(a_true), b) < 1e-13
    # A is symmetric positive definite on the unmasked pixels and zero on the masked ones
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(4 * npix), rng.standard_normal(4 * npix)
    Ax, Ay = ora.compute_Ax(x), ora.compute_Ax(y)
    assert abs(y @ Ax - x @ Ay) < 1e-10 * abs(y @ Ax)
    assert x @ Ax > 0
    assert np.all(Ax.reshape(4, npix)[:, ~m] == 0.0)
    # explicit 2x2 block at one pixel / Stokes
    p = int(np.flatnonzero(m)[5])
    s1 = np.array([ora.eval_sed(0, j, p, 2) for j in range(cfg.nbands)])
    s2 = np.array([ora.eval_sed(1, j, p, 2) for j in range(cfg.nbands)])
    w = 1.0 / sky.rms[:, 1, p] ** 2
    e = np.zeros(4 * npix)
    e[p] = 1.0
    col = ora.compute_Ax(e)
    assert abs(col[p] - np.sum(s1 * s1 * w)) < 1e-13 * col[p]
    assert abs(col[2 * npix + p] - np.sum(s1 * s2 * w)) < 1e-13 * abs(col[2 * npix + p])


def test_cg_recovers_noiseless_sky_and_zero_chisq():
    cfg, sky = truth_case(nside=4)
    ora = Oracle(cfg, sky)
    its, delta = ora.sample_cg_group(ml_mode=0)
    assert 1 < its[0] < cfg.cg_groups[0].max_iter and delta[0] < cfg.cg_groups[0].converge
    m = sky.mask != 0
    for ic, c in enumerate(cfg.comps):
        assert np.max(np.abs(ora.amplitude(ic)[1:3][:, m] - sky.truth[c.label][1:3][:, m])) < 1e-6
        assert np.all(ora.amplitude(ic)[:, ~m] == 0.0)  # masked pixels keep the initial amplitude (0)
    chisq, planes = ora.compute_chisq()
    assert chisq < 1e-12 and planes[0] == 0.0
    # Q3: the loop counter starts at 1, so i_max = 1 means no iteration at all
    cfg.cg_groups[0].max_iter = 1
    ora2 = Oracle(cfg, sky)
    its2, _ = ora2.sample_cg_group(ml_mode=0)
    assert its2[0] == 1 and np.all(ora2.amplitude(0) == 0.0)


def test_sample_vector_quirk_q1():
    cfg, sky = truth_case()
    ora = Oracle(cfg, sky)
    npix = cfg.npix
    eta = np.random.default_rng(1).standard_normal(2 * npix)
    ref = ora.compute_sample_vector(eta).reshape(4, npix)
    fix = ora.compute_sample_vector(eta, fix_q1=True).reshape(4, npix)
    # reference indexing: the LAST diffuse component's term lands in slot 1, slot 2 gets nothing
    assert np.all(ref[2:] == 0.0)
    assert np.array_equal(ref[:2], fix[2:])
    assert not np.allclose(fix[:2], fix[2:])


def test_lnl_matches_chisq_and_is_zero_for_exact_model():
    cfg, sky = small_case("c1", 4)
    ora = Oracle(cfg, sky)
    ora.update_sky_model()
    chisq, planes = ora.compute_chisq()
    dp = C.POINTER(C.c_double)
    mi = (C.c_int * 2)(2, 3)
    mask = np.where((sky.mask == 0) | (sky.mask == MISSVAL), 0.0, sky.mask)
    lnl = ora.lib.ora_evaluate_lnL.__class__  # noqa: F841 (symbol exists)
    ora.lib.ora_evaluate_lnL.restype = C.c_double
    ora.lib.ora_evaluate_lnL.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int), C.c_int, dp]
    sm = np.ascontiguousarray(ora.sky_model())
    v = ora.lib.ora_evaluate_lnL(ora.st, sky.sig.ctypes.data_as(dp), sky.rms.ctypes.data_as(dp),
                                 sm.ctypes.data_as(dp), mi, -1, mask.ctypes.data_as(dp))
    # lnL = -1/2 sum ((d-m)/sigma)^2 = -1/2 * nbands * sum(chi_map)
    assert abs(v + 0.5 * cfg.nbands * (planes[1] + planes[2])) < 1e-10 * abs(v)
    v0 = ora.lib.ora_evaluate_lnL(ora.st, sm.ctypes.data_as(dp), sky.rms.ctypes.data_as(dp),
                                  sm.ctypes.data_as(dp), mi, -1, mask.ctypes.data_as(dp))
    assert v0 == 0.0


def test_metropolis_semantics():
    cfg, sky = small_case("c1", 4)
    cfg.comps[0].indices[0].prior = "uniform"
    cfg.comps[0].indices[0].uni = (-3.05, -2.95)
    nsample = 16
    rng = np.random.default_rng(2)
    z, u = rng.standard_normal(nsample * cfg.npix), rng.random(nsample * cfg.npix)
    ora = Oracle(cfg, sky)
    before = ora.indices(0).copy()
    acc, dec, lnl = ora.sample_index_mh(0, 0, -1, nsample, 1, z, u, want_trace=True)
    dec = dec.reshape(nsample, cfg.npix)
    m = sky.mask != 0
    assert np.all(dec[:, ~m] == 3) and np.all(dec[:, m] != 3)
    assert (dec == 2).sum() > 0 and acc == (dec == 1).sum()
    assert np.all(np.isnan(lnl.reshape(nsample, -1)[dec >= 2]))      # Q5: no evaluation, no uniform
    idx = ora.indices(0)
    assert np.all(idx[0][1:3][:, ~m] == 0.0)                         # masked pixels end up at 0
    assert np.array_equal(idx[0][1], idx[0][2])                      # Q+U draw writes both planes
    assert np.array_equal(idx[0][0], before[0][0])                   # I plane untouched
    assert np.all((idx[0][1][m] >= -3.05 - 0.2) & (idx[0][1][m] <= -2.95 + 0.2))
    # optimize mode only ever accepts improvements: the final lnL is >= every accepted one before it
    ora2 = Oracle(cfg, sky)
    acc2, dec2, lnl2 = ora2.sample_index_mh(0, 0, -1, nsample, 0, z, u, want_trace=True)
    l2, d2 = lnl2.reshape(nsample, -1), dec2.reshape(nsample, -1)
    for p in np.flatnonzero(m)[:20]:
        accepted = l2[d2[:, p] == 1, p]
        assert np.all(np.diff(accepted) > 0)


def test_tuner_thresholds():
    cfg, sky = small_case("c2", 4, perturb=False)
    spec = cfg.comps[1].indices[0]
    spec.tune, spec.step = True, 1.0
    spec.uni = (1.5 - 1e-9, 1.5 + 1e-9)  # every proposal is out of bounds: accept rate 0 -> halved
    ora = Oracle(cfg, sky)
    rng = np.random.default_rng(3)
    z, u = rng.standard_normal(20 * 3), rng.random(20 * 3)
    nb, step = ora.tune_step(1, 0, -1, 20, 1, z, u, 3)
    assert nb == 3 and step == 1.0 / 8


def test_golden_vectors():
    """Oracle outputs frozen in tests/golden/ (made by tests/golden/make_golden.py): guards the
    checker itself against regressions."""
    import os
    from golden.make_golden import run_case
    path = os.path.join(os.path.dirname(__file__), "golden", "c1_nside4.npz")
    gold = np.load(path)
    now = run_case()
    for k in gold.files:
        assert np.array_equal(gold[k], now[k]) or rel_err(now[k], gold[k]) < 1e-13, k


def test_golden_vectors_intensity_components_and_udgrade():
    """The round-2 oracle features frozen the same way: hi_fit + monopole in a Stokes-I CG group (amplitudes,
    band monopoles, residual, a full-sky T_d chain) and the resolution operators."""
    import os
    from golden.make_golden import run_intensity_case
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "intensity_nside4.npz"))
    now = run_intensity_case()
    assert set(gold.files) == set(now)
    for k in gold.files:
        assert np.array_equal(gold[k], now[k]) or rel_err(now[k], gold[k]) < 1e-12, k
    load().ora_set_T_CMB(2.7255)


def test_template_fit_recovers_a_noiseless_sky():
    """Template branches of compute_rhs / compute_Ax / unpack_amplitudes (src/dang_cg_mod.f90:560-587,
    :745-768, :867-893, :1374-1392): on a noiseless sky = synchrotron + per-band template amplitudes the CG
    returns the input template amplitudes and synchrotron map, and the chi-square of the result vanishes."""
    from helpers import template_case
    from oracle.binding import Oracle
    cfg, sky, tamp_true = template_case(8, noise=False)
    ora = Oracle(cfg, sky)
    its, delta = ora.sample_cg_group(0, 0, np.zeros(2 * cfg.npix))  # optimize mode: no fluctuation
    assert 2 < its[0] < cfg.cg_groups[0].max_iter and delta[0] <= cfg.cg_groups[0].converge
    ta = ora.template_amplitudes(1)
    assert np.allclose(ta[1], tamp_true, rtol=1e-7, atol=1e-9) and np.array_equal(ta[1], ta[2])
    assert np.all(ta[0] == 0.0)  # the I plane is never written by a Q+U solve (:1379-1381)
    m = sky.mask != 0
    assert np.max(np.abs(ora.amplitude(0)[1:3][:, m] - sky.truth["synch"][1:3][:, m])) < 1e-5
    assert ora.compute_chisq()[0] < 1e-12


# ------------------------------------------------------------------ Stokes-I component types
def test_planck_rj_sed_of_t_cmb_and_hi_fit():
    """evaluate_T_cmb / evaluate_hi_fit (src/dang_component_mod.f90:815-884) against the closed form
    B_nu(T) / (2 k nu^2 / c^2) * 1e6 = (h nu / k) / (exp(h nu / k T) - 1) * 1e6 [uK_RJ]."""
    import sys
    sys.path.insert(0, "tests")
    from helpers import intensity_case
    from oracle.binding import Oracle, load
    load().ora_set_T_CMB(2.7255)
    cfg, sky, _, hi_true = intensity_case(4, with_hi=True)
    sky.template_amplitudes["hi"][0] = hi_true
    ora = Oracle(cfg, sky)
    h, k_B = 1.0545726691251021e-34 * 2.0 * np.pi, 1.3806503e-23
    for j, b in enumerate(cfg.bands):
        nu, T = b.nu_ghz * 1e9, sky.indices["hi"][0][0, 5]
        closed = (h * nu / k_B) / np.expm1(h * nu / (k_B * T)) * 1e6
        sed = ora.eval_sed(1, j, 5, 1)                      # hi_fit: template(pix) * planck
        assert abs(sed - sky.template["hi"][0, 5] * closed) <= 1e-12 * abs(sed)
        sig = ora.lib.ora_eval_signal(ora.st, 1, j, 5, 1, None)
        assert abs(sig - hi_true[j] * sed) <= 1e-15 * abs(sig)
        assert ora.eval_sed(2, j, 5, 1) == 1.0              # monopole: template(:,1) = 1
        assert ora.eval_sed(2, j, 5, 2) == 0.0              # ... and 0 in polarisation


def test_monopole_and_hi_fit_recover_a_noiseless_sky():
    """Known answer for the border rows of compute_rhs / compute_Ax (src/dang_cg_mod.f90:522-559, 717-744, 833-866):
    on a noiseless sky the CG recovers the HI band amplitudes and the band monopoles, update_sky_model turns the
    monopoles into the offsets (src/dang_data_mod.f90:357-361), and chi-square -> 0."""
    import sys
    sys.path.insert(0, "tests")
    from helpers import intensity_case
    from oracle.binding import Oracle, load
    load().ora_set_T_CMB(2.7255)
    cfg, sky, mono, hi = intensity_case(8, noise=False, with_hi=True)
    ora = Oracle(cfg, sky)
    its, delta = ora.sample_cg_group(0, 0, None)
    assert 1 < its[0] < cfg.cg_groups[0].max_iter
    assert np.allclose(ora.template_amplitudes(1)[0], hi, rtol=1e-6)
    assert np.allclose(ora.template_amplitudes(2)[0], mono, rtol=1e-6)
    off = np.ctypeslib.as_array(ora.lib.ora_offset(ora.st), shape=(cfg.nbands,))
    assert np.array_equal(off, ora.template_amplitudes(2)[0])
    assert ora.compute_chisq()[0] < 1e-12
    # dust (diffuse) + monopole, one band left unfitted: chi-square -> 0 as well
    cfg, sky, mono, _ = intensity_case(8, noise=False, with_hi=False, unfitted_band=3)
    ora = Oracle(cfg, sky)
    ora.sample_cg_group(0, 0, None)
    assert ora.compute_chisq()[0] < 1e-12
    assert np.allclose(ora.template_amplitudes(1)[0], mono, atol=1e-5)


def test_sample_vector_slots_interleave_with_two_border_components():
    """compute_sample_vector's template counter `l` (src/dang_cg_mod.f90:970) is never reset per
    component, so with two border components its slots interleave (band-major) while compute_Ax lays the
    components out one after the other.  With a single border component both layouts coincide."""
    import sys
    sys.path.insert(0, "tests")
    from helpers import intensity_case
    from oracle.binding import Oracle
    cfg, sky, _, _ = intensity_case(4, with_hi=True)
    ora = Oracle(cfg, sky)
    npix, nb = cfg.npix, cfg.nbands
    eta = np.random.default_rng(2).standard_normal(npix)
    sv = ora.compute_sample_vector(eta)           # no diffuse component in this group: the vector is the tail
    assert sv.shape == (2 * nb,)
    m = sky.mask != 0
    hi_rows = np.array([np.sum((eta / sky.rms[j, 0])[m] * np.array([ora.eval_sed(1, j, i, 1) for i in range(npix)])[m]) for j in range(nb)])
    mono_rows = np.array([np.sum((eta / sky.rms[j, 0])[m]) for j in range(nb)])
    assert np.allclose(sv[0::2], hi_rows, rtol=1e-12) and np.allclose(sv[1::2], mono_rows, rtol=1e-12)


# ------------------------------------------------------------------ udgrade
def test_healpix_index_conversion_and_udgrade_known_answers():
    """nest2ring at nside 2 against the published HEALPix table (healpy.nest2ring(2, range(48))); round trips; and
    the algebra of udgrade_ring / udgrade_rms / udgrade_mask (src/dang_util_mod.f90:341-376)."""
    from oracle.binding import load, udgrade
    lib = load()
    table = [13, 5, 4, 0, 15, 7, 6, 1, 17, 9, 8, 2, 19, 11, 10, 3, 28, 20, 27, 12, 30, 22, 21, 14, 32, 24, 23, 16,
             34, 26, 25, 18, 44, 37, 36, 29, 45, 39, 38, 31, 46, 41, 40, 33, 47, 43, 42, 35]
    assert [lib.ora_nest2ring(2, i) for i in range(48)] == table
    assert [lib.ora_nest2ring(1, i) for i in range(12)] == list(range(12))
    for ns in (4, 32):
        n = 12 * ns * ns
        assert [lib.ora_ring2nest(ns, lib.ora_nest2ring(ns, i)) for i in range(n)] == list(range(n))
    ns = 16
    n = 12 * ns * ns
    rng = np.random.default_rng(3)
    m = rng.standard_normal((2, n))
    d = udgrade("ring", m, ns, 4)
    assert np.isclose(d.sum() * 16, m.sum())                                    # averaging conserves the mean
    assert np.allclose(udgrade("ring", udgrade("ring", d, 4, ns), ns, 4), d, rtol=1e-14)  # upgrade copies parents
    bad = m.copy()
    bad[0, :] = -1.6375e30
    assert np.all(udgrade("ring", bad, ns, 4)[0] == -1.6375e30)                 # all children bad -> bad
    rms = np.abs(m) + 0.5
    assert np.allclose(udgrade("rms", np.full((1, n), 3.0), ns, 4), 3.0 * 4 / 16)  # sqrt(mean sigma^2) * nside_out/nside_in
    assert np.allclose(udgrade("rms", rms, ns, 8) ** 2, udgrade("ring", rms ** 2, ns, 8) / 4.0)
    mask = (rng.random((1, n)) < 0.7).astype(float)
    dm = udgrade("mask", mask, ns, 4)
    assert set(np.unique(dm)) <= {0.0, 1.0} and np.array_equal(dm, (udgrade("ring", mask, ns, 4) >= 0.5).astype(float))
