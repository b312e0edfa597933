"""Parity against the CPU oracle AT THE SIZES BASELINE.json CONFIGURES:

  c1  exactly as configured: nside 64, 5 bands, 10 Gibbs iterations, NUMSAMPLE 50, per-pixel beta_s
      (serial oracle: the strict reference semantics);
  c2  the headline config at nside 512, 8 bands, CG + full-sky beta_d, NUMSAMPLE 20, 3 iterations
      (OpenMP oracle: same source, the reference's own `!$OMP PARALLEL DO` pixel loops);
  c3  12 bands x 128-sample tabulated bandpasses, per-pixel beta_s and beta_d, NUMSAMPLE 20, nside 64;
  c4  20 bands, per-pixel beta_d then T_d, NUMSAMPLE 20, nside 64, 3 iterations.

Every call goes through the C ABI with injected deviates (eta, z, u).  Asserted every iteration:
the CG iteration count and the whole residual trajectory, amplitude maps and chi-square within
1e-10 relative, every Metropolis accept / reject / out-of-bounds decision identical
(`array_equal`), the acceptance counts, lnL of every evaluated proposal within 1e-10, index maps
within 1e-13 (accepted proposals are stored verbatim).
"""
import numpy as np
import pytest

from conftest import TOL, rel_err

pytestmark = pytest.mark.gpu


def _sampled_calls(cfg):
    """(ic, nind, map_n, fullsky) in sample_spectral_parameters' order (src/dang_sample_mod.f90:39-74)."""
    from dang_b200.config import flag_to_map_n, return_poltype_flag
    out = []
    for ic, c in enumerate(cfg.comps):
        for j, s in enumerate(c.indices):
            if s.sample:
                for flag in return_poltype_flag(s.poltype):
                    out.append((ic, j, flag_to_map_n(flag), s.region == "fullsky"))
    return out


def run_gibbs(name, nside, n_iter, omp, first_draw_iter=2, seed=20260103):
    """The loop body of src/dang.f90:87-126 on both sides, compared after every block."""
    from dang_b200.engine import OPT_RECORD, Engine
    from dang_b200.synth import make_config, make_sky
    from oracle.binding import Oracle
    cfg = make_config(name, nside=nside)
    sky = make_sky(cfg)
    ora, eng = Oracle(cfg, sky, omp=omp), Engine(cfg, sky)
    eng.set_option(OPT_RECORD, 1)
    rng = np.random.default_rng(seed)
    nsample = cfg.nsample
    calls = _sampled_calls(cfg)
    n_cg, n_eval, n_acc = [], 0, 0
    unmasked = sky.mask != 0
    for it in range(1, n_iter + 1):
        # ---- sample_cg_groups: rhs -> cg_search -> unpack, update_sky_model, chi-square
        eta = rng.standard_normal(2 * cfg.npix)
        it_o, delta_o, trace_o = ora.cg_search_trace(ml_mode=1, eta=eta)
        ora.update_sky_model()
        r = eng.sample_cg_groups(eta=eta)
        it_g, delta_g = r[0]
        trace_g = eng.cg_trace()
        assert it_g == it_o, (name, it, it_g, it_o, trace_g[-3:], trace_o[-3:])
        assert len(trace_g) == len(trace_o)
        assert np.allclose(trace_g, trace_o, rtol=1e-7, atol=0), (name, it)
        n_cg.append(it_g)
        chisq_o, _ = ora.compute_chisq()
        assert abs(r[1] - chisq_o) <= TOL * chisq_o, (name, it, r[1], chisq_o)
        for ic in range(len(cfg.comps)):
            assert rel_err(eng.amplitude(ic), ora.amplitude(ic)) < TOL, (name, it, ic)
        # ---- sample_spectral_parameters (iter > 1 in the reference's loop)
        if it >= first_draw_iter:
            for ic, j, map_n, fullsky in calls:
                n = nsample if fullsky else nsample * cfg.npix
                z, u = rng.standard_normal(n), rng.random(n)
                acc_o, dec_o, lnl_o = ora.sample_index_mh(ic, j, map_n, nsample, 1, z, u, want_trace=True)
                acc_g = eng.sample_index_mh(ic, j, map_n, nsample, "sample", z, u)
                dec_g, lnl_g = eng.decisions(nsample, fullsky=fullsky)
                if not fullsky:  # compare the chains of unmasked pixels; masked ones are flagged 3 on both sides
                    assert np.array_equal(dec_g.reshape(nsample, -1)[:, ~unmasked], dec_o.reshape(nsample, -1)[:, ~unmasked])
                assert np.array_equal(dec_g, dec_o[:n]), (name, it, ic, j, int((dec_g != dec_o[:n]).sum()))
                assert acc_g == acc_o, (name, it, ic, j, acc_g, acc_o)
                ev = dec_o[:n] < 2
                assert rel_err(lnl_g[ev], lnl_o[:n][ev]) < TOL, (name, it, ic, j)
                n_eval += int(ev.sum())
                n_acc += int((dec_o[:n] == 1).sum())
                if fullsky:
                    assert eng.index_fullsky(ic, j, 2) == ora.indices(ic)[j, 1, 0]
            ora.update_sky_model()
            chisq_o, _ = ora.compute_chisq()
            chisq_g = eng.compute_chisq()
            assert abs(chisq_g - chisq_o) <= TOL * chisq_o, (name, it, chisq_g, chisq_o)
        for ic, c in enumerate(cfg.comps):
            if c.indices:
                assert rel_err(eng.indices(ic), ora.indices(ic)) < 1e-13, (name, it, ic)
    assert 0 < n_acc < n_eval  # non-trivial chains
    eng.close()
    ora.close()
    return n_cg


def test_c1_as_configured():
    """BASELINE configs[0]: nside 64, 5 delta bands, 10 Gibbs iterations, NUMSAMPLE 50, per-pixel beta_s,
    CG_GROUP_MAX_ITER 100, CG_CONVERGE_THRESH 1e-12 -- against the serial oracle."""
    from dang_b200.synth import make_config
    cfg = make_config("c1")
    assert (cfg.nside, cfg.nbands, cfg.ngibbs, cfg.nsample) == (64, 5, 10, 50)
    n_cg = run_gibbs("c1", 64, cfg.ngibbs, omp=False)
    assert all(1 < n < 100 for n in n_cg), n_cg  # every solve converged by the stop rule, not by i_max


def test_c2_headline_config_at_nside_512():
    """BASELINE configs[1] (the bench.py workload) at its real size: same n_cg as the oracle under the
    absolute stop rule, the same delta trajectory, identical full-sky decisions."""
    n_cg = run_gibbs("c2", 512, 3, omp=True)
    assert all(1 < n < 100 for n in n_cg), n_cg


def test_c3_bandpass_config_at_nside_64():
    """BASELINE configs[2] with its real band count (12) and 128-sample bandpasses: one amplitude draw,
    then per-pixel beta_s and beta_d chains (the device uses the 8-node Gauss rule and the moment
    series; the oracle sums all 128 samples per proposal)."""
    run_gibbs("c3", 64, 1, omp=True, first_draw_iter=1)


def test_c4_target_config_at_nside_64():
    """BASELINE configs[3] with its real band count (20): CG + per-pixel beta_d then T_d (fp32-screened
    chains with fp64 fallback), 3 Gibbs iterations."""
    run_gibbs("c4", 64, 3, omp=True)


def test_multi_gpu_parity_world_2():
    """tests/multi_gpu_check.py under torch.distributed.run on 2 GPUs (skipped on a single-GPU box):
    ring-range shards, NVLink mailboxes, every rank's slice against the full-sky oracle (nside 64)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DANG_MGC_NSIDE="64")
    for mailbox in ("1", "0"):
        env["DANG_GPU_MAILBOX"] = mailbox
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29533",
                            os.path.join(root, "tests", "multi_gpu_check.py")],
                           env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
        assert p.stdout.count("multi-GPU parity ok") == 2
