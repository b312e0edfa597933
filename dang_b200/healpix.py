"""Minimal HEALPix RING-scheme geometry used by the host side of the GPU path.

Only what the hot path needs: pixel -> z = cos(theta) (to build the synthetic Galactic
mask) and the iso-latitude ring table used to shard pixels across GPUs by contiguous
ring ranges.  The reference gets all of this from HEALPix-F90
(`nside2npix`, `pix_tools`, used at src/dang.f90:49-50); nothing here is on the device path.
"""
from __future__ import annotations

import numpy as np


def nside2npix(nside: int) -> int:
    return 12 * nside * nside


def ring_starts(nside: int) -> np.ndarray:
    """First RING-ordered pixel of every iso-latitude ring, plus npix as sentinel.

    Rings are numbered 1..4*nside-1 from the north pole; entry r-1 is the start of ring r.
    """
    nring = 4 * nside - 1
    npix = nside2npix(nside)
    ncap = 2 * nside * (nside - 1)
    starts = np.empty(nring + 1, dtype=np.int64)
    r = np.arange(1, nring + 1, dtype=np.int64)
    north = r <= nside - 1
    equat = (r >= nside) & (r <= 3 * nside)
    south = r > 3 * nside
    starts[:-1][north] = 2 * r[north] * (r[north] - 1)
    starts[:-1][equat] = ncap + (r[equat] - nside) * 4 * nside
    rs = 4 * nside - r[south]  # ring index counted from the south pole
    starts[:-1][south] = npix - 2 * rs * (rs + 1)
    starts[-1] = npix
    return starts


def pix2z_ring(nside: int, pix: np.ndarray) -> np.ndarray:
    """z = cos(colatitude) of RING-ordered pixels."""
    pix = np.asarray(pix, dtype=np.int64)
    starts = ring_starts(nside)
    ring = np.searchsorted(starts, pix, side="right")  # 1-based ring number
    z = np.empty(pix.shape, dtype=np.float64)
    fact2 = 4.0 / (12.0 * nside * nside)
    fact1 = 2.0 * nside * fact2
    north = ring < nside
    south = ring > 3 * nside
    equat = ~(north | south)
    z[north] = 1.0 - ring[north].astype(np.float64) ** 2 * fact2
    z[equat] = (2 * nside - ring[equat]).astype(np.float64) * fact1
    rs = (4 * nside - ring[south]).astype(np.float64)
    z[south] = -1.0 + rs * rs * fact2
    return z


def ring_partition(nside: int, nranks: int, weights: np.ndarray | None = None) -> np.ndarray:
    """Split RING-ordered pixels into `nranks` contiguous ring ranges.

    Returns `bounds` of length nranks+1 with bounds[g] = first pixel owned by rank g.
    Boundaries always fall on ring starts; ranges are chosen to balance `weights`
    (e.g. the unmasked-pixel indicator) or plain pixel counts.
    """
    starts = ring_starts(nside)
    npix = nside2npix(nside)
    if weights is None:
        csum_at_start = starts.astype(np.float64)
    else:
        w = np.asarray(weights, dtype=np.float64)
        assert w.shape == (npix,)
        csum = np.concatenate([[0.0], np.cumsum(w)])
        csum_at_start = csum[starts]
    total = csum_at_start[-1]
    bounds = np.empty(nranks + 1, dtype=np.int64)
    bounds[0] = 0
    bounds[-1] = npix
    for g in range(1, nranks):
        target = total * g / nranks
        k = int(np.argmin(np.abs(csum_at_start - target)))
        bounds[g] = starts[k]
    # keep the ranges non-decreasing (tiny maps with many ranks)
    for g in range(1, nranks + 1):
        bounds[g] = max(bounds[g], bounds[g - 1])
    return bounds
