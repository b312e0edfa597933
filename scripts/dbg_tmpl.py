import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import template_case
from oracle.binding import Oracle
from dang_b200.engine import Engine
cfg, sky, tamp_true = template_case(16)
cfg.ml_mode = "optimize"
cfg.cg_groups[0].converge, cfg.cg_groups[0].max_iter = 1e-20, 300
ora, eng = Oracle(cfg, sky), Engine(cfg, sky)
eta = np.zeros(2 * cfg.npix)
print(ora.sample_cg_group(0, 0, eta), eng.sample_cg_groups(ml_mode="optimize", eta=eta))
sky_g, res_g, chi_g = eng.update_sky_model()
so, ro = ora.sky_model(), ora.res_map()
for k in range(3):
    d = np.abs(sky_g[:, k] - so[:, k])
    print("plane", k, "max diff", d.max(), "max ref", np.abs(so[:, k]).max(), "argmax", np.unravel_index(d.argmax(), d.shape))
    d = np.abs(res_g[:, k] - ro[:, k])
    print("   res max diff", d.max(), "max ref", np.abs(ro[:, k]).max())
print("tamp", eng.template_amplitudes(1), ora.template_amplitudes(1))
