"""Cell shapes away from the tested ones -- more than ZR_TPITCH (128) rows, odd row counts (an odd x_bins used to leave the
interval breaks 4-byte aligned in shared memory: "misaligned address"; fixed in range_layout / zrank_layout) -- the
single-launch kernel against the banded pair of round 1, bit for bit; then smoke().  Crash / alignment check: the
observables are random, so most log-likelihoods are -inf (value parity is what tests/test_gpu_parity.py is for)."""
import os, sys, numpy as np, warnings
sys.path.insert(0, os.getcwd())
warnings.simplefilter("ignore")
import mcmctoffitting_b200 as M
rs = np.random.RandomState(5)
for xb, eb in ((160, 120), (129, 150), (33, 240)):
    z = rs.standard_normal(1024)
    obs = np.rint(rs.uniform(0, 50, 2048))
    th = np.vstack([np.array([1050.0, 0.10]) + np.array([10, 1e-2]) * rs.standard_normal((1500, 2)),
                    np.column_stack([rs.uniform(1000, 2600, 500), rs.uniform(0.02, 0.5, 500)])])
    res = {}
    for label, env in (("single", "1"), ("pair", "0")):
        os.environ["TOFGPU_RANGE_ZRANK"] = env
        cfg = M.config.sweep(ode_mode=M.config.ODE_RANGE, x_bins=xb, e_bins=eb)
        fn = M.make_lnprob(cfg, obs, z)
        if os.environ.get("REALISTIC_OBS"):                 # observables = a model realisation: finite log-likelihoods
            real = np.rint(1e4 * fn.model.model_batch(np.array([[1050.0, 0.10]]), run=0, stage="spread")[0])
            fn.bind_observables(real)
        try:
            res[label] = fn.batch(th)
        except Exception as e:
            print(xb, label, 'FAILED', e); raise
        st = fn.model.stats()
        fn.model.close()
        print(xb, label, "launches/call", st["model_launches_per_call"], "finite", int(np.isfinite(res[label]).sum()))
    a, b = res["single"], res["pair"]
    print(xb, "identical:", bool(np.array_equal(a, b, equal_nan=True)))
import __graft_entry__ as g
g.smoke(); print("smoke ok")
