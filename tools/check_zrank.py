"""A/B of the shipped adv_zrank_kernel against adv_planned_kernel + overflow launch (TOFGPU_RANGE_ZRANK=0) and, on a
sample, the oracle: same lnprob bit for bit is expected (same cells; only the order of the normalisation sum differs).

    python tools/check_zrank.py [--n 131072] [--oracle 64]"""
import argparse, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmctoffitting_b200 as M   # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--oracle", type=int, default=0)
args = ap.parse_args()
warnings.simplefilter("ignore")
from oracle import tof_oracle as O   # noqa: E402  (checker)
om = O.sweep_model()
z = np.random.RandomState(20260101).standard_normal(1024)
obs = np.rint(1e5 * om.model_pdf([1050, 0.08], np.random.RandomState(7).standard_normal(1024)))
rs = np.random.RandomState(1)
ens = np.array([1050.0, 0.10]) + np.array([10, 1e-2]) * rs.standard_normal((args.n, 2))
box = np.column_stack([rs.uniform(1000, 2600, args.n), rs.uniform(0.02, 0.5, args.n)])
box[:8] = [[1000.5, 0.0201], [2599.0, 0.499], [1000.5, 0.499], [2599.0, 0.0201], [1800, 0.02001], [1050, 0.3], [1500, 0.25], [2000, 0.1]]
cfg = M.config.sweep(ode_mode=M.config.ODE_RANGE)
dev = torch.device("cuda", 0)
res = {}
for label, env in (("zrank", "1"), ("planned", "0"))[:1 if os.environ.get("ZR_ONLY") else 2]:
    os.environ["TOFGPU_RANGE_ZRANK"] = env
    fn = M.make_lnprob(cfg, obs, z, device=0)
    m = fn.model
    m.set_timing(True)
    for name, th in (("ensemble", ens), ("prior_box", box)):
        t = torch.from_numpy(th).to(dev)
        out = torch.empty(args.n, dtype=torch.float64, device=dev)
        ms = []
        for _ in range(3):
            m.lnprob_batch_device(t.data_ptr(), args.n, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.synchronize()
            ms.append(m.last_kernel_ms())
        res[(label, name)] = out.cpu().numpy()
        print("%-8s %-10s kernel ms %s  -> %.3f M evals/s  finite %.3f  stats %s" % (
            label, name, " ".join("%.2f" % v for v in ms), args.n / min(ms) / 1e3, np.isfinite(res[(label, name)]).mean(),
            {k: v for k, v in m.stats().items() if k in ("band_queued_last", "kernel_launches")}))
    m.close()
for name, th in (("ensemble", ens), ("prior_box", box)):
    if ("planned", name) not in res:
        continue
    a, b = res[("zrank", name)], res[("planned", name)]
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    both = np.isfinite(a) & np.isfinite(b)
    rel = np.abs(a[both] - b[both]) / np.abs(b[both])
    print("%-10s identical %d / %d   max rel (finite both) %.3g   finiteness mismatches %d" % (
        name, same.sum(), len(a), rel.max() if rel.size else 0.0, int((np.isfinite(a) != np.isfinite(b)).sum())))
    if not same.all():
        bad = np.flatnonzero(~same)[:5]
        for k in bad:
            print("   walker %d theta %s: zrank %r planned %r" % (k, th[k], a[k], b[k]))
if args.oracle:
    xs = O.DDNXS()
    for name, th in (("ensemble", ens), ("prior_box", box)):
        a = res[("zrank", name)]
        bad = 0
        worst = 0.0
        for k in range(args.oracle):
            want = om.lnprob(th[k], obs, z, xs)
            if a[k] == want or (np.isnan(a[k]) and np.isnan(want)):
                continue
            r = abs(a[k] - want) / abs(want) if np.isfinite(want) and np.isfinite(a[k]) else np.inf
            worst = max(worst, r)
            bad += int(not r <= 1e-9)
        print("oracle %-10s %d walkers: %d outside 1e-9 (worst %.3g)" % (name, args.oracle, bad, worst))
