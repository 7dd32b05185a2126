"""Tiny end-to-end run of every kernel (for compute-sanitizer memcheck / racecheck on the GPU box)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmctoffitting_b200 as M
warnings.simplefilter("ignore")
rs = np.random.RandomState(0)
which = sys.argv[1:] or ["range", "rk4", "simple", "simult", "simult_range", "onebd"]
if "range" in which or "rk4" in which:
    for mode in ([M.config.ODE_RANGE] if "range" in which else []) + ([M.config.ODE_RK4] if "rk4" in which else []):
        cfg = M.config.sweep(ode_mode=mode, n_samples=2048, n_ev_per_loop=1024)
        fn = M.make_lnprob(cfg, np.ones(2048), rs.standard_normal(2048))
        th = np.array([[1050, .1], [1500, .3], [1200, .45], [999, .1]])
        print("adv mode", mode, fn.batch(th), fn.model.cell_counts(th[:1]).sum(), fn.model.model_batch(th[:2], stage="spread").shape)
        fn.model.close()
if "simple" in which:
    cfg = M.config.simple(20000)
    fn = M.make_lnprob(cfg, np.ones(25), (rs.random_sample(20000), rs.standard_normal(20000)))
    print("simple", fn.batch([[1100, -100, 50], [1000, -50, 20]]))
    fn.model.close()
for nm, mode in (("simult", M.config.ODE_RK4), ("simult_range", M.config.ODE_RANGE)):
    if nm in which:
        cfg = M.config.simult(n_samples=3000, n_ev_per_loop=1000, ode_mode=mode)
        fn = M.make_lnprob(cfg, [np.ones(n) for n in cfg.tof_bins], [rs.standard_normal(cfg.n_draws) for _ in range(5)],
                           extra_draws=[rs.standard_normal(3000) for _ in range(5)])
        th = np.array([[1878.4, 850, 170, 0.5, 3e4, 2e4, 2e4, 4e4, 4e4], [1825.0, 1000, 300, 1.2, 3e4, 2e4, 2e4, 4e4, 4e4]])
        print(nm, fn.batch(th), fn.model.cell_counts(th[:1], run=2).sum())
        fn.model.close()
if "onebd" in which:
    cfg = M.config.onebd(n_samples=3000, n_ev_per_loop=1000)
    fn = M.make_lnprob(cfg, [np.ones(25)] * 3, [rs.standard_normal(cfg.n_draws) for _ in range(3)],
                       extra_draws=[rs.random_sample(2000) for _ in range(3)])
    print("onebd", fn.batch([[900.0, 170.0, 0.5, 3e4, 2e4, 4e4, 5.0, 12.0, 0.0]]))
    fn.model.close()
