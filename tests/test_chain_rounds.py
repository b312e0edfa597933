"""The full-sky chain kernel (mh_suff_chain_kernel, dang_b200/csrc/kernels_mh.cuh) evaluates proposals in speculative
rounds: 32 lanes evaluate proposals l0 .. l0+31 against the CURRENT state, the first accepted one ends the round.  This
restates both loops in Python -- the reference's serial loop (src/dang_sample_mod.f90:282-324, with Q5: an out-of-bounds
proposal is skipped without consuming its uniform) and the round structure -- on the same lnL / prior functions and
checks that decisions, traces, the final sample and the acceptance count are identical, whatever the acceptance rate."""
import math

import numpy as np
import pytest


def serial(sample, lnl, step, lo, hi, z, u, optimize):
    n = len(z)
    dec = np.full(n, 3, dtype=np.int8)
    trace = np.full(n, np.nan)
    lnl_old, accept = lnl(sample), 0
    for l in range(n):
        th = sample + (0.0 + step * z[l])
        if th < lo or th > hi:
            dec[l] = 2
            continue
        lnl_new = lnl(th)
        ratio = math.exp(min(lnl_new - lnl_old, 700.0))
        acc = ratio > 1.0 if optimize else ratio > u[l]
        trace[l] = lnl_new
        dec[l] = 1 if acc else 0
        if acc:
            sample, lnl_old, accept = th, lnl_new, accept + 1
    return sample, accept, dec, trace


def rounds(sample, lnl, step, lo, hi, z, u, optimize, width=32):
    n = len(z)
    dec = np.full(n, 3, dtype=np.int8)
    trace = np.full(n, np.nan)
    lnl_old, accept, l0, nrounds = lnl(sample), 0, 0, 0
    while l0 < n:
        nrounds += 1
        lanes = []
        for j in range(width):                     # every lane against the same (sample, lnl_old)
            l = l0 + j
            if l >= n:
                lanes.append(None)
                continue
            th = sample + (0.0 + step * z[l])
            oob = th < lo or th > hi
            lnl_new, acc = 0.0, False
            if not oob:
                lnl_new = lnl(th)
                ratio = math.exp(min(lnl_new - lnl_old, 700.0))
                acc = ratio > 1.0 if optimize else ratio > u[l]
            lanes.append((th, oob, lnl_new, acc))
        first = next((j for j, v in enumerate(lanes) if v is not None and not v[1] and v[3]), width)
        for j, v in enumerate(lanes):
            if v is None or j > first:
                continue
            dec[l0 + j] = 2 if v[1] else (1 if v[3] else 0)
            if not v[1]:
                trace[l0 + j] = v[2]
        if first < width:
            sample, lnl_old, accept = lanes[first][0], lanes[first][2], accept + 1
            l0 += first + 1
        else:
            l0 += width
    return sample, accept, dec, trace, nrounds


@pytest.mark.parametrize("optimize", [False, True])
@pytest.mark.parametrize("nsample,step,width", [(20, 0.01, 32), (50, 0.05, 32), (200, 0.3, 32), (100, 2.0, 32), (70, 0.05, 8)])
def test_speculative_rounds_equal_the_serial_chain(nsample, step, width, optimize):
    rng = np.random.default_rng(nsample)
    for case in range(40):
        mu, w = 1.5 + 0.2 * rng.standard_normal(), 10.0 ** rng.uniform(-2.0, 0.5)
        lnl = lambda x: -0.5 * ((x - mu) / w) ** 2 - 0.5 * ((x - 1.55) / 0.1) ** 2     # likelihood + Gaussian prior
        z, u = rng.standard_normal(nsample), rng.random(nsample)
        s0 = 1.5 + 0.1 * rng.standard_normal()
        a = serial(s0, lnl, step, 1.0, 2.2, z, u, optimize)
        b = rounds(s0, lnl, step, 1.0, 2.2, z, u, optimize, width)
        assert a[0] == b[0] and a[1] == b[1]
        assert np.array_equal(a[2], b[2])
        assert np.array_equal(a[3], b[3], equal_nan=True)
        # rounds = accepted moves + the windows that ended without one
        assert b[4] <= a[1] + math.ceil(nsample / width) + 1
