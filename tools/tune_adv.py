"""Time the adv model kernel variants on the sweep shape (run on the GPU box)."""
import os, sys, json, warnings, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mcmctoffitting_b200 as M
from oracle import tof_oracle as O
warnings.simplefilter("ignore")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 32
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["128x8", "256x4", "256x2", "512x2", "512x1", "1024x1"]
om = O.sweep_model()
z = np.random.RandomState(20260101).standard_normal(1024)
obs = np.rint(1e5 * om.model_pdf([1050, .10], np.random.RandomState(7).standard_normal(1024)))
thetas = np.array([1050, 0.10]) + np.array([10, 1e-2]) * np.random.RandomState(1).standard_normal((n, 2))
dev = torch.device("cuda", 0)
th = torch.from_numpy(thetas).to(dev)
out = torch.empty(n, dtype=torch.float64, device=dev)
res = {}
for v in variants:
    os.environ["TOFGPU_ADV_VARIANT"] = v
    for sort in (False, True):
        cfg = M.config.sweep(ode_mode=int(os.environ.get("TOF_ODE_MODE", "0")))
        with M.TofModel(cfg) as m:
            m.set_observables(obs)
            m.set_draws(z, sort=sort)
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
            res[(v, sort)] = min(ms)
            print("variant %-7s sorted=%-5s  %8.3f ms  %10.0f evals/s  ctas/sm=%d smem=%d band=%d/%d queued=%d" % (
                v, sort, min(ms), n / (min(ms) * 1e-3), s["ctas_per_sm"], s["smem_bytes"], s["band_ctas_per_sm"],
                s["band_cells"], s["band_queued_last"]), flush=True)
    if v == variants[0]:
        print("fp64 peak TFLOP/s", M.TofModel(M.config.sweep()).measure_fp64_peak())
