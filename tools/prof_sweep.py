"""One model call of the sweep shape for ncu captures and quick kernel timing -- no sampler, no CPU arm.

    python tools/prof_sweep.py [--n 131072] [--reps 5] [--prior-box] [--check 0]

Prints the kernel time of each repetition (CUDA events inside the library) and evaluations/s.  With --check K the
first K walkers are compared with the oracle (lnprob 1e-9)."""
import argparse, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmctoffitting_b200 as M   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--prior-box", action="store_true", help="walkers uniform over the adv prior (adv:81-82) instead of the bench ensemble")
ap.add_argument("--check", type=int, default=0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--stages", action="store_true", help="per-stage cycle shares (tof_set_stage_timing)")
args = ap.parse_args()
warnings.simplefilter("ignore")

import torch  # noqa: E402
cfg = M.config.sweep(ode_mode=M.config.ODE_RANGE)
z = np.random.RandomState(20260101).standard_normal(1024)
rs = np.random.RandomState(args.seed)
if args.prior_box:
    thetas = np.column_stack([rs.uniform(1000, 2600, args.n), rs.uniform(0.02, 0.5, args.n)])
else:
    thetas = np.array([1050.0, 0.10]) + np.array([10, 1e-2]) * rs.standard_normal((args.n, 2))
    thetas[:, 0] = np.clip(thetas[:, 0], 1000.0 + 1e-6, 2600.0 - 1e-6)
    thetas[:, 1] = np.clip(thetas[:, 1], 0.02 + 1e-9, 0.5 - 1e-9)
obs = np.zeros(2048)
obs[900:1150] = 100.0          # any observables: the kernel does full work regardless
if args.check:
    from oracle import tof_oracle as O
    om = O.sweep_model()
    obs = np.rint(1e5 * om.model_pdf([1050, 0.08], np.random.RandomState(7).standard_normal(1024)))
fn = M.make_lnprob(cfg, obs, z, device=0)
m = fn.model
dev = torch.device("cuda", 0)
th = torch.from_numpy(thetas).to(dev)
out = torch.empty(args.n, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
m.set_timing(True)
ms = []
for _ in range(args.reps):
    m.lnprob_batch_device(th.data_ptr(), args.n, out.data_ptr(), stream)
    torch.cuda.synchronize()
    ms.append(m.last_kernel_ms())
best = min(ms)
print("kernel ms per call:", " ".join("%.3f" % v for v in ms))
print("evals/s (best): %.4g   finite fraction %.3f   stats %s" % (args.n / (best * 1e-3), float(torch.isfinite(out).double().mean()), m.stats()))
if args.stages:
    m.set_stage_timing(True)
    m.lnprob_batch_device(th.data_ptr(), args.n, out.data_ptr(), stream)
    torch.cuda.synchronize()
    print("stage profile:", m.stage_profile())
    m.set_stage_timing(False)
if args.check:
    from oracle import tof_oracle as O
    xs = O.DDNXS()
    got = out[:args.check].cpu().numpy()
    bad = 0
    for k in range(args.check):
        want = om.lnprob(thetas[k], obs, z, xs)
        ok = (got[k] == want) or (np.isfinite(want) and abs(got[k] - want) <= 1e-9 * abs(want))
        bad += int(not ok)
    print("oracle check: %d walkers, %d outside 1e-9, %d finite" % (args.check, bad, int(np.isfinite(got).sum())))
m.close()
