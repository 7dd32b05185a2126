import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mcmctoffitting_b200 as M
warnings.simplefilter("ignore")
rs = np.random.RandomState(0)
cfg = M.config.adv(0, ode_mode=M.config.ODE_RANGE)
n = 296
th = np.column_stack([rs.uniform(1020, 1100, n), rs.uniform(0.08, 0.12, n)])
fn = M.make_lnprob(cfg, np.ones(50) * 100, rs.standard_normal(cfg.n_draws))
for _ in range(3):
    out = fn.batch(th)
print(fn.model.stats())
