"""Host-side logic of the multi-rank path, exercised on CPU with world_size-2 gloo:
ring-range partitioning, rank-ordered (deterministic) combination of per-rank partial sums, and
the broadcast of the NCCL unique id through torch.distributed."""
import os
import socket

import numpy as np
import pytest

from dang_b200.healpix import nside2npix, pix2z_ring, ring_partition, ring_starts


def test_ring_starts_and_z():
    for nside in (1, 2, 4, 16):
        st = ring_starts(nside)
        npix = nside2npix(nside)
        assert st[0] == 0 and st[-1] == npix and len(st) == 4 * nside
        sizes = np.diff(st)
        assert sizes.max() == 4 * nside and sizes.min() == 4
        assert np.array_equal(sizes, sizes[::-1])  # north/south symmetry
        z = pix2z_ring(nside, np.arange(npix))
        assert np.all(np.diff(z) <= 1e-15)  # RING order runs from north to south
        assert abs(z.sum()) < 1e-9
        for r in range(4 * nside - 1):  # constant latitude within a ring
            assert np.ptp(z[st[r]:st[r + 1]]) == 0.0


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_ring_partition_is_contiguous_ring_aligned_and_balanced(nranks):
    nside = 32
    npix = nside2npix(nside)
    z = pix2z_ring(nside, np.arange(npix))
    w = (np.abs(z) >= np.sin(np.deg2rad(5.0))).astype(float)
    b = ring_partition(nside, nranks, weights=w)
    assert b[0] == 0 and b[-1] == npix and np.all(np.diff(b) > 0)
    assert set(b.tolist()) <= set(ring_starts(nside).tolist())
    loads = np.array([w[b[g]:b[g + 1]].sum() for g in range(nranks)])
    assert loads.max() / loads.mean() < 1.15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nside = 16
    npix = nside2npix(nside)
    rng = np.random.default_rng(5)
    mask = (rng.random(npix) > 0.1).astype(float)
    per_pixel = rng.standard_normal(npix) ** 2
    b = ring_partition(nside, world, weights=mask)
    lo, hi = int(b[rank]), int(b[rank + 1])
    # what every scalar exchange in libdang_gpu does: local partial -> all-gather -> rank-ordered sum
    local = torch.tensor([float(np.sum(per_pixel[lo:hi] * mask[lo:hi])), float(mask[lo:hi].sum())], dtype=torch.float64)
    rows = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(rows, local)
    total = 0.0
    for g in range(world):
        total = total + float(rows[g][0])
    # the NCCL unique id travels as 128 raw bytes from rank 0
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    q.put((rank, lo, hi, total, float(sum(r[1] for r in rows)), bytes(uid.numpy().tobytes())))
    dist.destroy_process_group()


def test_rank_ordered_scalar_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, n0, u0), (r1, lo1, hi1, t1, n1, u1) = out
    npix = nside2npix(16)
    assert lo0 == 0 and hi0 == lo1 and hi1 == npix
    assert t0 == t1 and n0 == n1  # identical bits on every rank
    rng = np.random.default_rng(5)
    mask = (rng.random(npix) > 0.1).astype(float)
    per_pixel = rng.standard_normal(npix) ** 2
    assert abs(t0 - np.sum(per_pixel * mask)) < 1e-9 * t0 and n0 == mask.sum()
    assert u0 == u1 == bytes(range(128))
