"""Time the BASELINE.json configs C1-C4 (and oneBD) at their full sizes and BASELINE walker counts on one GPU:
device-resident thetas, CUDA-event kernel time of one ensemble-sized call.  Observables are a model realisation at the
script's guess values (made with the library itself: this is a timing tool, parity at these sizes is in
tests/test_gpu_parity.py), so the finite fraction of the log-likelihoods is meaningful."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mcmctoffitting_b200 as M
warnings.simplefilter("ignore")
which = sys.argv[1:] or ["simple", "intermediate", "adv", "simult", "onebd"]
dev = torch.device("cuda", 0)
rs = np.random.RandomState(0)


def run(name, cfg, thetas, draws, extra=None, reps=2, guess=None, scale=1e4):
    obs = [np.ones(n) * 100 for n in cfg.tof_bins]
    obs = obs[0] if cfg.n_runs == 1 else obs
    t0 = time.time()
    fn = M.make_lnprob(cfg, obs, draws, extra_draws=extra)
    setup = time.time() - t0
    m = fn.model
    if guess is not None:                                  # observables = a realisation of the model at the guess values
        g = np.asarray(guess, dtype=np.float64)[None, :]
        if cfg.kind == M.config.KIND_SIMPLE:
            real = [np.rint(m.model_batch(g, run=0, stage="counts")[0] * (scale / cfg.n_draws))]
        elif cfg.n_runs == 1:
            real = [np.rint(scale * m.model_batch(g, run=0, stage="spread")[0])]
        else:
            real = [np.rint(m.model_batch(g, run=r, stage="spread")[0]) for r in range(cfg.n_runs)]   # scale factors are parameters
        fn.bind_observables(real[0] if cfg.n_runs == 1 else real)
    th = torch.from_numpy(np.ascontiguousarray(thetas)).to(dev)
    out = torch.empty(len(thetas), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    m.lnprob_batch_device(th.data_ptr(), len(thetas), out.data_ptr(), st)
    torch.cuda.synchronize()
    m.set_timing(True)
    best = 1e30
    for _ in range(reps):
        m.lnprob_batch_device(th.data_ptr(), len(thetas), out.data_ptr(), st)
        torch.cuda.synchronize()
        best = min(best, m.last_kernel_ms())
    fin = float(torch.isfinite(out).double().mean())
    print("%-14s walkers=%6d draws/run=%8d  %10.2f ms  %12.1f evals/s  finite=%.2f  setup=%.1fs ctas/sm=%d" % (
        name, len(thetas), cfg.n_draws, best, len(thetas) / (best * 1e-3), fin, setup, m.stats()["ctas_per_sm"]), flush=True)
    m.close()


if "simple" in which:
    cfg = M.config.simple(1000000)
    th = np.array([1100.0, -100.0, 50.0]) + np.array([5.0, 5.0, 2.0]) * rs.standard_normal((32, 3))
    run("C1 simple", cfg, th, (rs.random_sample(cfg.n_draws), rs.standard_normal(cfg.n_draws)), guess=[1100.0, -100.0, 50.0])
if "intermediate" in which:
    for mode, nm in ((M.config.ODE_RK4, "rk4"), (M.config.ODE_RANGE, "range")):
        cfg = M.config.intermediate(0, ode_mode=mode)
        th = np.array([900.0, 0.10]) + np.array([5.0, 5e-3]) * rs.standard_normal((256, 2))
        run("C2 interm/" + nm, cfg, th, rs.standard_normal(cfg.n_draws), reps=1, guess=[900.0, 0.10])
if "adv" in which:
    for mode, nm in ((M.config.ODE_RK4, "rk4"), (M.config.ODE_RANGE, "range")):
        cfg = M.config.adv(0, ode_mode=mode)
        n = 4096                                            # BASELINE.json: 4096 walkers
        th = np.array([1050.0, 0.10]) + np.array([5.0, 5e-3]) * rs.standard_normal((n, 2))
        run("C3 adv/" + nm, cfg, th, rs.standard_normal(cfg.n_draws), reps=1, guess=[1050.0, 0.10])
if "simult" in which:
  for mode, nm in ((M.config.ODE_RK4, "rk4"), (M.config.ODE_RANGE, "range")):
    cfg = M.config.simult(ode_mode=mode)
    n = 16384 if mode == M.config.ODE_RANGE else int(os.environ.get("SIMULT_RK4_N", "1184"))   # BASELINE.json: 16384 walkers
    th = np.tile([1878.4, 850, 170, 0.5, 3e4, 2e4, 2e4, 4e4, 4e4], (n, 1)) * (1 + 0.01 * rs.standard_normal((n, 9)))
    th[:, 0] = np.clip(th[:, 0], 1826, 1924)
    draws = [rs.standard_normal(cfg.n_draws) for _ in range(5)]
    extra = [rs.standard_normal(20000) for _ in range(5)]
    run("C4 simult/" + nm, cfg, th, draws, extra, reps=1, guess=[1878.4, 850, 170, 0.5, 3e4, 2e4, 2e4, 4e4, 4e4])
if "onebd" in which:
    cfg = M.config.onebd()
    n = 4096
    th = np.tile([900.0, 170.0, 0.5, 3e4, 2e4, 4e4, 5.0, 12.0, 0.5], (n, 1)) * (1 + 0.01 * rs.standard_normal((n, 9)))
    draws = [rs.standard_normal(cfg.n_draws) for _ in range(3)]
    extra = [rs.random_sample(4000) for _ in range(3)]
    run("oneBD", cfg, th, draws, extra, reps=2, guess=[900.0, 170.0, 0.5, 3e4, 2e4, 4e4, 5.0, 12.0, 0.5])
if "few" in which:
    # the reference's own run sizes: a few dozen walkers per half-step, 1e5 draws
    cfg = M.config.adv(0, ode_mode=M.config.ODE_RANGE)
    for n in (25, 50, 128):
        th = np.column_stack([rs.uniform(1020, 1100, n), rs.uniform(0.08, 0.12, n)])
        run("adv few n=%d" % n, cfg, th, rs.standard_normal(cfg.n_draws), reps=2)
