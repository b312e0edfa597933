import sys, numpy as np
sys.path.insert(0, '.')
from dang_b200.engine import Engine, OPT_CG_CHECKPOINT
from dang_b200.synth import make_config, make_sky
cfg = make_config("c2", nside=int(sys.argv[1]) if len(sys.argv) > 1 else 64); sky = make_sky(cfg)
a = Engine(cfg, sky); b = Engine(cfg, sky); b.set_option(OPT_CG_CHECKPOINT, 0)
for it in range(1):
    ra = a.cg_solve(0, 0, "sample", seed=99 + it); rb = b.cg_solve(0, 0, "sample", seed=99 + it)
    xa = a.cg_x().reshape(2, 2, -1); xb = b.cg_x().reshape(2, 2, -1)
    print(ra, rb, "x equal", np.array_equal(xa, xb))
    for ic in range(2):
        aa, ab = a.amplitude(ic)[1:3], b.amplitude(ic)[1:3]
        print(ic, "amp==x (recompute)", np.array_equal(aa, xa[ic]), "amp==x (streaming)", np.array_equal(ab, xb[ic]), "amps equal", np.array_equal(aa, ab),
              "max diff", np.max(np.abs(aa - xa[ic])), np.argmax(np.abs(aa - xa[ic]).ravel()))

d = np.abs(a.amplitude(0)[1:3] - b.amplitude(0)[1:3])
idx = np.argwhere(d > 0)
print("n differing", len(idx), idx[:5], d.max())
print("x differing", np.sum(a.cg_x() != b.cg_x()))
