"""Experiment: does 2 CTAs/SM x 512 threads beat 1 CTA/SM x 1024 when the cell histogram is small enough?"""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mcmctoffitting_b200 as M
warnings.simplefilter("ignore")
n = 148 * 64
rs = np.random.RandomState(1)
thetas = np.array([1050, 0.10]) + np.array([10, 1e-2]) * rs.standard_normal((n, 2))
z = np.random.RandomState(20260101).standard_normal(1024)
dev = torch.device("cuda", 0)
th = torch.from_numpy(thetas).to(dev)
out = torch.empty(n, dtype=torch.float64, device=dev)
for label, kw in (("full 240 bins", {}), ("100 bins 550-1550", dict(e_bins=100, e_range=(550.0, 1550.0)))):
    for nt in (1024, 512):
        os.environ["TOFGPU_RANGE_THREADS"] = str(nt)
        cfg = M.config.sweep(ode_mode=M.config.ODE_RANGE, **kw)
        with M.TofModel(cfg) as m:
            m.set_observables(np.ones(2048))
            m.set_draws(z)
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(2):
                m.lnprob_batch_device(th.data_ptr(), n, out.data_ptr(), st)
            torch.cuda.synchronize()
            m.set_timing(True)
            ms = []
            for _ in range(3):
                m.lnprob_batch_device(th.data_ptr(), n, out.data_ptr(), st)
                torch.cuda.synchronize()
                ms.append(m.last_kernel_ms())
            s = m.stats()
            print("%-20s NT=%4d  %8.3f ms  %10.0f evals/s  ctas/sm=%d smem=%d" % (label, nt, min(ms), n / (min(ms) * 1e-3), s["ctas_per_sm"], s["smem_bytes"]), flush=True)
