"""GPU parity for the Stokes-I component types: `monopole`, `hi_fit` and `T_cmb`
(src/dang_component_mod.f90:579-700, 815-884) -- their border rows in the CG (src/dang_cg_mod.f90:522-559,
717-744, 833-866, 1044-1067), update_sky_model's monopole -> offset rule (src/dang_data_mod.f90:357-361), the
Metropolis data that subtracts the monopole on top of the offset (src/dang_sample_mod.f90:173-196) and the global
T_CMB update after a T_cmb draw (:76-78).  Every call goes through the C ABI and is compared with the oracle.

As for `template` (test_gpu_parity.py) the bordered systems are ill-conditioned, so the CG operator is compared on
the first residual norms and the solutions at a tight threshold, not on the iteration count."""
import numpy as np
import pytest

from conftest import TOL, rel_err
from helpers import intensity_case

pytestmark = pytest.mark.gpu


def _pair(cfg, sky):
    from dang_b200.engine import Engine
    from oracle.binding import Oracle, load
    load().ora_set_T_CMB(2.7255)
    return Oracle(cfg, sky), Engine(cfg, sky)


@pytest.mark.parametrize("with_hi", [False, True])
def test_border_cg_operator_matches_oracle(with_hi):
    cfg, sky, _, _ = intensity_case(16, with_hi=with_hi, unfitted_band=None if with_hi else 3)
    cfg.cg_groups[0].converge, cfg.cg_groups[0].max_iter = 1e-14, 500
    ora, eng = _pair(cfg, sky)
    eta = np.random.default_rng(1).standard_normal(cfg.npix)
    it_o, delta_o, tr_o = ora.cg_search_trace(0, 0, 1, eta)
    it_g, delta_g = eng.cg_solve(0, 0, "sample", eta)
    tr_g = eng.cg_trace()
    assert np.max(np.abs(tr_g[:5] - tr_o[:5]) / tr_o[:5]) < 1e-11, (tr_g[:5], tr_o[:5])
    assert delta_g <= cfg.cg_groups[0].converge and delta_o <= cfg.cg_groups[0].converge


@pytest.mark.parametrize("ml_mode", ["optimize", "sample"])
@pytest.mark.parametrize("with_hi", [False, True])
def test_monopole_and_hi_fit_draws(with_hi, ml_mode):
    """Dust + monopole, and HI template + monopole (two border components in one group): amplitudes, band
    monopoles -> offsets, chi-square, sky model / residual, and a full-sky draw whose data follow :173-196."""
    cfg, sky, mono_true, hi_true = intensity_case(16, with_hi=with_hi, unfitted_band=None if with_hi else 3)
    cfg.ml_mode = ml_mode
    cfg.cg_groups[0].converge, cfg.cg_groups[0].max_iter = 1e-22, 2000
    sampled = cfg.comps[1].indices[0] if with_hi else cfg.comps[0].indices[0]   # hi T_d / dust beta
    sampled.sample, sampled.region, sampled.prior = True, "fullsky", "gaussian"
    if with_hi:
        sky.indices["hi"][0][:] = 20.0   # a full-sky index starts from a uniform map
    else:
        sampled.gauss, sampled.uni, sampled.step = (1.55, 0.1), (1.0, 2.2), 0.002
    ora, eng = _pair(cfg, sky)
    ic_m = len(cfg.comps) - 1
    rng = np.random.default_rng(3)
    ml = 1 if ml_mode == "sample" else 0
    tol = 1e-8
    for it in range(1, 3):
        eta = rng.standard_normal(cfg.npix)
        ora.sample_cg_group(0, ml, eta)
        r = eng.sample_cg_groups(ml_mode=ml_mode, eta=eta)
        chisq_o, _ = ora.compute_chisq()
        assert abs(r[1] - chisq_o) <= tol * chisq_o, (it, r[1], chisq_o)
        ta_g, ta_o = eng.template_amplitudes(ic_m), ora.template_amplitudes(ic_m)
        assert rel_err(ta_g, ta_o) < tol, (ta_g[0], ta_o[0])
        assert np.all(ta_g[1:] == 0.0)   # hi_fit / monopole amplitudes live on plane 1 only
        if with_hi:
            assert rel_err(eng.template_amplitudes(1), ora.template_amplitudes(1)) < tol
        else:
            assert rel_err(eng.amplitude(0), ora.amplitude(0)) < tol
        if it == 1 and ml_mode == "optimize" and not with_hi:   # the fit sees the true monopoles through the noise
            fitted = np.array(cfg.comps[ic_m].corr)
            assert np.allclose(ta_g[0][fitted], mono_true[fitted], atol=0.5)
        sky_g, res_g, chi_g = eng.update_sky_model()
        ora.update_sky_model()
        assert rel_err(sky_g, ora.sky_model()) < tol and rel_err(res_g, ora.res_map()) < tol
        nsample = 12
        z, u = rng.standard_normal(nsample), rng.random(nsample)
        ic_s = 1 if with_hi else 0
        acc_o, dec_o, lnl_o = ora.sample_index_mh(ic_s, 0, 1, nsample, ml, z, u, want_trace=True)
        acc_g = eng.sample_index_mh(ic_s, 0, 1, nsample, ml_mode, z, u)
        dec_g, lnl_g = eng.decisions(nsample, fullsky=True)
        assert np.array_equal(dec_g, dec_o[:nsample]) and acc_g == acc_o
        ev = dec_o[:nsample] < 2
        assert rel_err(lnl_g[ev], lnl_o[:nsample][ev]) < 1e-8
        ora.update_sky_model()
        chisq_o, _ = ora.compute_chisq()
        assert abs(eng.compute_chisq() - chisq_o) <= tol * chisq_o
        assert rel_err(eng.indices(ic_s), ora.indices(ic_s)) < 1e-14


def test_hi_fit_per_pixel_temperature_draw():
    """Per-pixel T_d chains of an hi_fit component (one plane): decisions identical to the oracle."""
    from dang_b200.engine import OPT_RECORD
    cfg, sky, _, hi_true = intensity_case(8, with_hi=True)
    spec = cfg.comps[1].indices[0]
    spec.sample, spec.region = True, "per-pixel"
    sky.template_amplitudes["hi"][0] = hi_true * 1.02   # non-zero band amplitudes to sample against
    ora, eng = _pair(cfg, sky)
    eng.set_option(OPT_RECORD, 1)
    nsample = 10
    rng = np.random.default_rng(5)
    z, u = rng.standard_normal(nsample * cfg.npix), rng.random(nsample * cfg.npix)
    acc_o, dec_o, lnl_o = ora.sample_index_mh(1, 0, 1, nsample, 1, z, u, want_trace=True)
    acc_g = eng.sample_index_mh(1, 0, 1, nsample, "sample", z, u)
    dec_g, lnl_g = eng.decisions(nsample, fullsky=False)
    assert np.array_equal(dec_g, dec_o) and acc_g == acc_o
    ev = dec_o < 2
    assert 0 < (dec_o == 1).sum() < ev.sum()
    assert rel_err(lnl_g[ev], lnl_o[ev]) < TOL
    assert rel_err(eng.indices(1)[0][0], ora.indices(1)[0][0]) < 1e-14


def test_t_cmb_component_and_global_temperature():
    """A 'T_cmb' component (eval_signal = B_nu(T) in RJ units, no amplitude) next to dust and a 'cmb' component:
    chi-square / sky model, a full-sky T draw, and the global T_CMB update that the 'cmb' SED (1 / a2t) follows."""
    from dang_b200.config import Band, CGGroup, Component, IndexSpec, RunConfig
    from dang_b200.synth import Sky, band_sigma, sed_mbb
    from oracle.binding import planck_rj
    nside = 8
    bands = [Band(nu) for nu in (30.0, 44.0, 70.0, 100.0, 143.0, 217.0)]
    nb = len(bands)
    dust = Component(label="dust", type="mbb", nu_ref_ghz=143.0, cg_group=1, amp_sample=True,
                     indices=[IndexSpec("BETA", init=1.55, poltype="T"), IndexSpec("T", init=19.6, poltype="T")])
    cmb = Component(label="cmb", type="cmb", nu_ref_ghz=100.0, cg_group=1, amp_sample=True, indices=[])
    tcmb = Component(label="tcmb", type="T_cmb", nu_ref_ghz=100.0, cg_group=2, amp_sample=False,
                     indices=[IndexSpec("T", init=2.7255, sample=True, region="fullsky", prior="gaussian", gauss=(2.7255, 1e-3),
                                        uni=(2.70, 2.75), step=2e-6, poltype="T")])
    cfg = RunConfig("tcmb", nside, bands, [dust, cmb, tcmb], [CGGroup(sample=True, max_iter=300, converge=1e-16, poltype="T")],
                    nsample=16, tqu="T")
    npix = cfg.npix
    rng = np.random.default_rng(9)
    sig = np.zeros((nb, 3, npix))
    rms = np.ones((nb, 3, npix))
    a_d = np.abs(rng.normal(0, 30.0, npix)) + 1.0
    a_c = rng.normal(0, 70.0, npix)
    for j, b in enumerate(bands):
        rms[j, 0] = band_sigma(b.nu_ghz) * (1.0 + 0.3 * rng.random(npix))
        sig[j, 0] = (a_d * sed_mbb(b.nu_ghz, 143.0, 1.55, 19.6) + planck_rj(b.nu_ghz * 1e9, 2.72552)
                     + a_c + rms[j, 0] * rng.standard_normal(npix))
    amp0 = {"dust": np.zeros((3, npix)), "cmb": np.zeros((3, npix)), "tcmb": np.zeros((3, npix))}
    amp0["dust"][0], amp0["cmb"][0] = a_d, a_c
    idx0 = {"dust": np.stack([np.full((3, npix), 1.55), np.full((3, npix), 19.6)]), "cmb": np.zeros((0, 3, npix)),
            "tcmb": np.full((1, 3, npix), 2.7255)}
    sky = Sky(sig=sig, rms=rms, mask=np.ones(npix), gain=np.ones(nb), offset=np.zeros(nb), amplitude=amp0, indices=idx0, truth={})
    ora, eng = _pair(cfg, sky)
    ora.update_sky_model()
    chisq_o, _ = ora.compute_chisq()
    assert abs(eng.compute_chisq() - chisq_o) <= TOL * chisq_o
    sky_g, res_g, _ = eng.update_sky_model()
    assert rel_err(sky_g, ora.sky_model()) < TOL
    nsample = cfg.nsample
    z, u = rng.standard_normal(nsample * npix), rng.random(nsample * npix)
    ora.sample_spectral_parameters(nsample, 1, z, u)           # draws T, then T_CMB = indices(0,1,1) (:76-78)
    acc, chisq_g = eng.sample_spectral_parameters(nsample=nsample, z=z, u=u)
    t_new = ora.indices(2)[0, 0, 0]
    assert t_new != 2.7255 and ora.lib.ora_get_T_CMB() == t_new
    assert eng.indices(2)[0, 0, 0] == t_new
    chisq_o, _ = ora.compute_chisq()
    assert abs(chisq_g - chisq_o) <= TOL * chisq_o             # the 'cmb' SED moved with T_CMB on both sides
    ora.lib.ora_set_T_CMB(2.7255)


def test_border_layout_limits_are_loud():
    from dang_b200.engine import DangGpuError, Engine
    cfg, sky, _, _ = intensity_case(4)
    cfg.cg_groups[0].poltype = "Q+U"   # hi_fit / monopole rows address plane 1 only
    eng = Engine(cfg, sky)
    with pytest.raises(DangGpuError, match="Stokes-I"):
        eng.sample_cg_groups(eta=np.zeros(2 * cfg.npix))
