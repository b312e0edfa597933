"""Config c5: an ensemble of independent Gibbs chains on ONE GPU sharing one device copy of the maps
(dang_gpu_share_maps).  Prints aggregate iterations/s over all chains (chains advance round-robin)."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from dang_b200.engine import Engine
from dang_b200.synth import make_config, make_sky

nchain = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nside = int(sys.argv[2]) if len(sys.argv) > 2 else 512
nit = 4
cfg = make_config("c5", nside=nside)
sky = make_sky(cfg)
owner = Engine(cfg, sky)
chains = [owner] + [Engine(cfg, sky, share_maps_with=owner) for _ in range(nchain - 1)]
for it in (1, 2):  # warm-up (first iteration has no spectral draw)
    for k, e in enumerate(chains):
        e.gibbs_iteration(it, seed=1000 * k)
for e in chains:
    e.sync()
t0 = time.perf_counter()
for it in range(3, 3 + nit):
    for k, e in enumerate(chains):
        e.gibbs_iteration(it, seed=1000 * k)
for e in chains:
    e.sync()
dt = time.perf_counter() - t0
import torch
print(f"c5: {nchain} chains x nside {nside}: {nchain * nit / dt:.1f} Gibbs iterations/s aggregate "
      f"({1e3 * dt / (nchain * nit):.3f} ms per chain-iteration), device memory in use "
      f"{(torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30:.1f} GiB")
chis = [e.compute_chisq() for e in chains[:4]]
print("chi-square of the first chains (different deviate seeds):", [round(c, 6) for c in chis])
