"""Parity at BASELINE.json's full size (config c2: nside = 512, 3.1 M pixels, 8 bands) through
size-independent properties -- the CPU oracle would need minutes per iteration here:

  * known answer: a noiseless sky is recovered by the optimize-mode CG, chi-square -> 0;
  * the three CG forms (checkpointed recompute, streaming, classic two-pass) and the two
    full-sky likelihood forms (sufficient statistics, per-proposal streaming) agree;
  * determinism / idempotence: repeating a call reproduces every bit;
  * linearity of the amplitude draw in the data;
  * sharding invariance: chi-square summed over two half-sky handles equals the full-sky value.
"""
import numpy as np
import pytest

from conftest import TOL, rel_err

pytestmark = pytest.mark.gpu

NSIDE = 512


@pytest.fixture(scope="module")
def case():
    from dang_b200.synth import TRUE_THETA, band_sed, make_config, make_sky
    cfg = make_config("c2", nside=NSIDE)
    sky = make_sky(cfg)
    clean = np.zeros_like(sky.sig)
    for c in cfg.comps:
        for j, b in enumerate(cfg.bands):
            clean[j, 1:3] += sky.truth[c.label][1:3] * band_sed(b, c, *TRUE_THETA[c.label])
    return cfg, sky, clean


def engine(cfg, sky, opts=None):
    from dang_b200.engine import Engine
    eng = Engine(cfg, sky)
    for k, v in (opts or {}).items():
        eng.set_option(k, v)
    return eng


def test_noiseless_sky_is_recovered(case):
    import copy
    from dang_b200.synth import TRUE_THETA
    cfg, sky, clean = case
    s2 = copy.copy(sky)
    s2.sig = clean
    s2.indices = {c.label: np.stack([np.full((3, cfg.npix), v) for v in TRUE_THETA[c.label]]) for c in cfg.comps}
    eng = engine(cfg, s2)
    it, delta = eng.cg_solve(0, 0, "optimize")
    assert 1 < it < cfg.cg_groups[0].max_iter and delta < cfg.cg_groups[0].converge
    m = sky.mask != 0
    for ic, c in enumerate(cfg.comps):
        a = eng.amplitude(ic)
        assert np.max(np.abs(a[1:3][:, m] - sky.truth[c.label][1:3][:, m])) < 1e-6
        assert np.all(a[:, ~m] == 0.0)
    assert eng.compute_chisq() < 1e-12


def test_cg_forms_agree_and_are_deterministic(case):
    from dang_b200.engine import OPT_CG_CHECKPOINT, OPT_CG_TWO_PASS
    cfg, sky, _ = case
    res = {}
    for name, opts in {"recompute": {}, "again": {}, "streaming": {OPT_CG_CHECKPOINT: 0},
                       "two_pass": {OPT_CG_TWO_PASS: 1}}.items():
        eng = engine(cfg, sky, opts)
        it, delta = eng.cg_solve(0, 0, "sample", seed=99)
        res[name] = (it, delta, eng.amplitude(0)[1:3].copy(), eng.amplitude(1)[1:3].copy(), eng.compute_chisq())
        eng.close()
    assert res["recompute"][0] == res["streaming"][0] == res["two_pass"][0]
    # same build, same inputs: bit-identical
    assert res["recompute"][1] == res["again"][1]
    assert np.array_equal(res["recompute"][2], res["again"][2]) and res["recompute"][4] == res["again"][4]
    # the recompute form performs the streaming form's operations in the same order
    assert np.array_equal(res["recompute"][2], res["streaming"][2])
    assert np.array_equal(res["recompute"][3], res["streaming"][3])
    for k in (2, 3):
        assert rel_err(res["two_pass"][k], res["recompute"][k]) < TOL
    assert abs(res["two_pass"][4] - res["recompute"][4]) < TOL * res["recompute"][4]


def test_fullsky_likelihood_forms_agree(case):
    from dang_b200.engine import OPT_FULLSKY_STREAM
    cfg, sky, _ = case
    rng = np.random.default_rng(3)
    z, u = rng.standard_normal(cfg.nsample), rng.random(cfg.nsample)
    out = {}
    for stream in (0, 1):
        eng = engine(cfg, sky, {OPT_FULLSKY_STREAM: stream})
        eng.cg_solve(0, 0, "sample", seed=5)
        acc = eng.sample_index_mh(1, 0, -1, cfg.nsample, "sample", z, u)
        dec, lnl = eng.decisions(cfg.nsample, fullsky=True)
        out[stream] = (acc, dec, lnl, eng.indices(1)[0, 1, :4].copy())
        eng.close()
    assert np.array_equal(out[0][1], out[1][1]) and out[0][0] == out[1][0]
    ev = out[0][1] < 2
    assert ev.sum() > 0
    assert rel_err(out[0][2][ev], out[1][2][ev]) < TOL
    assert np.array_equal(out[0][3], out[1][3])


def test_amplitude_draw_is_linear_in_the_data(case):
    """optimize mode, x0 = 0: the CG polynomial depends on the spectrum of A and on r0 only up to
    scale, so doubling the data (and quadrupling the stop threshold) doubles the solution."""
    import copy
    cfg, sky, _ = case
    cfg2 = copy.deepcopy(cfg)
    cfg2.cg_groups[0].converge = 4 * cfg.cg_groups[0].converge
    s2 = copy.copy(sky)
    s2.sig = 2.0 * sky.sig
    e1, e2 = engine(cfg, sky), engine(cfg2, s2)
    it1, d1 = e1.cg_solve(0, 0, "optimize")
    it2, d2 = e2.cg_solve(0, 0, "optimize")
    assert it1 == it2 and abs(d2 - 4 * d1) <= 1e-9 * d2
    for ic in range(2):
        assert rel_err(e2.amplitude(ic), 2.0 * e1.amplitude(ic)) < 1e-12


def test_chisq_is_sharding_invariant(case):
    from dang_b200.engine import Engine
    from dang_b200.healpix import ring_partition
    cfg, sky, _ = case
    full = Engine(cfg, sky)
    planes, n = full.chisq_planes()
    b = ring_partition(cfg.nside, 2, weights=(sky.mask != 0).astype(float))
    parts = []
    for g in range(2):
        e = Engine(cfg, sky, pix_range=(int(b[g]), int(b[g + 1])))
        parts.append(e.chisq_planes())
        e.close()
    assert parts[0][1] + parts[1][1] == n
    assert rel_err(parts[0][0] + parts[1][0], planes) < 1e-13


def test_one_statistics_pass_serves_both_chisquares(case):
    """Headline iteration at full size: chi-square from the cached sufficient statistics (before and after the
    full-sky draw) against the streaming chi-square kernel, and the draw itself against the uncached path."""
    from dang_b200.engine import OPT_STAT_CACHE
    cfg, sky, _ = case
    rng = np.random.default_rng(8)
    z, u = rng.standard_normal(cfg.nsample), rng.random(cfg.nsample)
    out = {}
    for cache in (1, 0):
        eng = engine(cfg, sky, {OPT_STAT_CACHE: cache})
        for it in (1, 2):
            eng.cg_solve(0, 0, "sample", seed=40 + it)
            c_before = eng.compute_chisq()
            eng.kernel_stats(reset=True)
            acc = eng.sample_index_mh(1, 0, -1, cfg.nsample, "sample", z, u)
            c_after = eng.compute_chisq()
            st = eng.kernel_stats()
        out[cache] = (c_before, acc, c_after, eng.index_fullsky(1, 0, 2), st["mh_suffstat_kernel"]["launches"], st["chisq_kernel"]["launches"])
        eng.close()
    assert out[1][4] == 0 and out[1][5] == 0      # cached: no pass over the maps in the draw or after it
    assert out[0][4] == 1 and out[0][5] == 1      # uncached: one statistics pass + one chi-square pass
    assert abs(out[1][0] - out[0][0]) <= 1e-12 * out[0][0]
    assert out[1][1] == out[0][1] and out[1][3] == out[0][3]
    assert abs(out[1][2] - out[0][2]) <= 1e-12 * out[0][2]


@pytest.mark.parametrize("form", [1, 3])
def test_screened_perpixel_chains_equal_fp64_chains_at_full_size(form):
    """Config c4's per-pixel beta_d and T_d draws at nside 512 (3.1 M chains x 20 proposals x 20 bands): the
    fp32-screened kernel leaves bit-identical index maps and acceptance counts, with a tiny fallback rate."""
    from dang_b200.engine import OPT_PERPIXEL_FAST
    from dang_b200.synth import make_config, make_sky
    cfg = make_config("c4", nside=NSIDE)
    sky = make_sky(cfg)
    a, b = engine(cfg, sky, {OPT_PERPIXEL_FAST: form}), engine(cfg, sky, {OPT_PERPIXEL_FAST: 0})
    n_unmasked = int((sky.mask != 0).sum())
    for eng in (a, b):
        eng.cg_solve(0, 0, "sample", seed=3)
    for nind in (0, 1):
        acc_a = a.sample_index_mh(1, nind, -1, cfg.nsample, "sample", seed=11 + nind)
        acc_b = b.sample_index_mh(1, nind, -1, cfg.nsample, "sample", seed=11 + nind)
        fallbacks, _ = a.perpixel_stats()
        assert acc_a == acc_b and 0 < acc_a < cfg.nsample * n_unmasked
        assert fallbacks < 2e-3 * cfg.nsample * n_unmasked, fallbacks
        ia, ib = a.indices(1)[nind], b.indices(1)[nind]
        assert np.array_equal(ia, ib)
        assert np.all(ia[1:3][:, sky.mask == 0] == 0.0)
