"""One call of the simultaneous fit at its full draw count (5 runs x 2e5 draws) for ncu captures."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmctoffitting_b200 as M
warnings.simplefilter("ignore")
rs = np.random.RandomState(0)
cfg = M.config.simult(ode_mode=M.config.ODE_RANGE)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
th = np.tile([1878.4, 850, 170, 0.5, 3e4, 2e4, 2e4, 4e4, 4e4], (n, 1)) * (1 + 0.01 * rs.standard_normal((n, 9)))
th[:, 0] = np.clip(th[:, 0], 1826, 1924)
draws = [rs.standard_normal(cfg.n_draws) for _ in range(5)]
extra = [rs.standard_normal(20000) for _ in range(5)]
fn = M.make_lnprob(cfg, [np.ones(t) * 100 for t in cfg.tof_bins], draws, extra_draws=extra)
for _ in range(2):
    out = fn.batch(th)
print(fn.model.stats(), np.isfinite(out).mean())
