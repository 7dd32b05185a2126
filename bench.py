#!/usr/bin/env python
"""Benchmark of the lnprob hot path: walker lnprob evaluations / second, adv TOF model.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]    # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[4] / SURVEY.md 8d, fits one GPU): adv model, 262144 walkers x 2048 TOF
bins x 1024 Monte-Carlo draws, synthetic observables.  One *step* is one ensemble MCMC step = two
red/blue half-steps = 262144 lnprob evaluations (propose -> forward model + likelihood -> accept
-> all-gather of the updated half when N > 1).  Walkers are sharded over ranks (strong scaling:
the ensemble size is fixed by the named config).

Prints ONE JSON line (rank 0).  ``value`` is device-resident throughput; ``e2e`` is the same metric
through the reference-facing host API (`TofLnProb.batch` = C-ABI `tof_lnprob_batch` with HOST buffers,
host<->device copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_WALKERS = 262144
N_DRAWS = 1024
N_TOF_BINS = 2048
THETA_STAR = (1050.0, 0.10)
DRAW_SEED = 20260101

# Algorithmic FP64 work per evaluation ("Roofline"; convention of SURVEY.md 8d: add/mul/compare = 1,
# FMA = 2, div = 8, log = 24).
#  * RK4 formulation:   F_rk4   = 174*D*X + 14*X*E + 26*E + (31+2K)*T
#  * range-table formulation (shipped): per (draw, x) sample  v = u0 + delta (1), interval compare (1),
#    dt = v - break (1), degree-7 Horner (14), accumulate (1) = 18;  per draw the T1 lookup (7 FMA + 4) = 18
#                                     F_range = 18*D*X + 18*D + 14*X*E + 26*E + (31+2K)*T
X_BINS, E_BINS, N_TAPS = 100, 240, 16
FLOP_PER_EVAL_RK4 = 174 * N_DRAWS * X_BINS + 14 * X_BINS * E_BINS + 26 * E_BINS + (31 + 2 * N_TAPS) * N_TOF_BINS
FLOP_PER_EVAL_RANGE = (18 * N_DRAWS * X_BINS + 18 * N_DRAWS + 14 * X_BINS * E_BINS + 26 * E_BINS
                       + (31 + 2 * N_TAPS) * N_TOF_BINS)
FLOP_PER_EVAL_EXECUTED = 18 * N_DRAWS * X_BINS + 18 * N_DRAWS
BYTES_PER_EVAL = 8 * (2 + 1)   # theta in, lnprob out
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12


def workload(om_mod):
    """Synthetic inputs of the named shape, produced with the oracle (checker side)."""
    om = om_mod.sweep_model()
    z = np.random.RandomState(DRAW_SEED).standard_normal(N_DRAWS)
    zstar = np.random.RandomState(7).standard_normal(N_DRAWS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        obs = np.rint(1e5 * om.model_pdf(list(THETA_STAR), zstar))
    thetas = np.array(THETA_STAR) + np.array([10, 1e-2]) * np.random.RandomState(1).standard_normal((N_WALKERS, 2))
    thetas[:, 0] = np.clip(thetas[:, 0], 1000.0 + 1e-6, 2600.0 - 1e-6)   # into the prior (adv:81-82)
    thetas[:, 1] = np.clip(thetas[:, 1], 0.02 + 1e-9, 0.5 - 1e-9)
    return om, z, obs, thetas


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(z, obs):
    from oracle import tof_oracle as O
    warnings.simplefilter("ignore")
    _W["m"] = O.sweep_model()
    _W["xs"] = O.DDNXS()
    _W["z"], _W["obs"] = z, obs


def _cpu_eval(theta):
    return float(_W["m"].lnprob(theta, _W["obs"], _W["z"], _W["xs"]))


def _ref_available():
    """True where the reference tree itself can be executed (the build container, or TOF_REFERENCE_ROOT set)."""
    try:
        from oracle import ref_loader
        return ref_loader.available()
    except Exception:
        return False


def _ref_init(z, obs):
    """The reference's OWN lnprob (tests/advIntermediateTOFmodel.py:115-199, executed unmodified through
    oracle/ref_loader.py) at the sweep shape: 1024 draws per loop, one loop, 2048 TOF bins on [128, 256) ns, physical
    mean excitation energy.  Each call draws fresh normals from numpy's global stream, as the script does."""
    from oracle import ref_loader
    warnings.simplefilter("ignore")
    ns = ref_loader.load("advIntermediateTOFmodel")
    ref = ref_loader.load_utilities()
    ns["nEvPerLoop"] = N_DRAWS
    ns["data_x"] = np.repeat(ns["x_binCenters"], N_DRAWS)
    ns["tof_nBins"] = N_TOF_BINS
    ns["tof_range"] = (128.0, 256.0)
    ns["stoppingModel"] = ref.ionStopping.ionStopping.simpleBethe([1, 2, 8.565e-5, 1, 19.2e-3])
    _W["ns"], _W["obs"] = ns, obs
    np.random.seed(os.getpid() & 0xFFFF)


def _ref_eval(theta):
    ns = _W["ns"]
    prior = ns["lnprior"](theta)
    if not np.isfinite(prior):
        return float(-np.inf)
    return float(prior + ns["lnlike"](theta, _W["obs"], nDraws=N_DRAWS))


def cpu_rate(z, obs, thetas, procs, per_proc, chunksize=None, kind="port"):
    """Evaluations/s of the CPU arm through a process pool (emcee `threads=P`, adv:300-302).  kind "port": the numpy
    oracle; kind "loader": the reference's own function bodies (only where the reference tree exists)."""
    import multiprocessing as mp
    init, ev = (_ref_init, _ref_eval) if kind == "loader" else (_cpu_init, _cpu_eval)
    n = procs * per_proc
    sample = [list(t) for t in thetas[:n]]
    if procs == 1:
        init(z, obs)
        ev(sample[0])
        t0 = time.perf_counter()
        for t in sample:
            ev(t)
        return n / (time.perf_counter() - t0)
    with mp.get_context("fork").Pool(procs, initializer=init, initargs=(z, obs)) as pool:
        pool.map(ev, sample[:procs])                              # warm-up: imports, first call
        t0 = time.perf_counter()
        if chunksize:
            list(pool.imap(ev, sample, chunksize=chunksize))
        else:
            pool.map(ev, sample)
        return n / (time.perf_counter() - t0)


def oracle_check(z, obs, thetas, got, procs):
    """Self-check of the timed ensemble: the oracle's lnprob for `thetas` against the values the GPU sampler holds."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs, initializer=_cpu_init, initargs=(z, obs)) as pool:
        want = np.array(pool.map(_cpu_eval, [list(t) for t in thetas]))
    both = np.isfinite(want) & np.isfinite(got)
    same_class = (np.isfinite(want) == np.isfinite(got)) & ~(np.isnan(want) ^ np.isnan(got))
    rel = np.abs(got[both] - want[both]) / np.abs(want[both]) if both.any() else np.zeros(0)
    return {"n": int(len(thetas)), "n_finite": int(both.sum()), "max_rel": float(rel.max()) if rel.size else 0.0,
            "n_outside_1e-9": int((rel > 1e-9).sum()), "n_flips": int((~same_class).sum()), "tolerance": 1e-9,
            "what": "lnprob the sampler holds for %d randomly chosen walkers of the timed ensemble (after the timed steps) "
                    "against the numpy oracle on the same positions and draws; n_flips = finite/-inf disagreements" % len(thetas)}


def run_cpu_baseline():
    """The `cpu_baseline` object of the GPU arm: serial, process pool (emcee threads=P) and an
    MPI-pool-shaped task farm, each on a bounded sample (~10-30 s of CPU work in total)."""
    from oracle import tof_oracle as O
    om, z, obs, thetas = workload(O)
    procs = os.cpu_count() or 1
    serial = cpu_rate(z, obs, thetas, 1, 48)
    pooled = cpu_rate(z, obs, thetas, procs, 24)
    mpi_like = cpu_rate(z, obs, thetas, max(procs - 1, 1), 24, chunksize=1) if procs > 1 else serial
    line = {
        "value": pooled, "unit": "evals/s", "cores": procs, "kind": "port",
        "sample": "%d evaluations (24 per process) of the same workload, numpy oracle" % (24 * procs),
        "serial_evals_per_s": serial, "mpi_pool_emulation_evals_per_s": mpi_like,
        "mpi_pool_emulation": "1 master + %d workers, one task per message (multiprocessing imap, chunksize=1); "
                              "mpi4py is not installed, so mpiTOFmodel.py's MPIPool is emulated" % max(procs - 1, 1)}
    if _ref_available():
        # the reference's own code (loader shim), where its tree exists: SURVEY.md 8d's CPU arm
        line["reference_loader_evals_per_s"] = cpu_rate(z, obs, thetas, procs, 4, kind="loader")
        line["reference_loader_serial_evals_per_s"] = cpu_rate(z, obs, thetas, 1, 6, kind="loader")
        line["reference_loader"] = ("tests/advIntermediateTOFmodel.py lnprior + lnlike executed unmodified through "
                                    "oracle/ref_loader.py, %d processes x 4 evaluations" % procs)
    print(json.dumps(line))
    return 0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import tof_oracle as O
    om, z, obs, thetas = workload(O)
    procs = os.cpu_count() or 1
    # where the reference tree exists its own function bodies are timed (kind "loader"); on the GPU box, which has no
    # /root/reference, the numpy port of the same algorithm (kind "port", ~4x faster per core than the reference)
    kind = "loader" if _ref_available() else "port"
    per_proc = 4 if kind == "loader" else 24
    rates = []
    for i in range(args.warmup + args.steps):
        r = cpu_rate(z, obs, thetas[i * procs * per_proc:], procs, per_proc, kind=kind)
        if i >= args.warmup:
            rates.append(r)
    value = statistics.mean(rates)
    sample = "%d evaluations per step (%d per process x %d processes) of the %d-walker workload, %s" % (
        procs * per_proc, per_proc, procs, N_WALKERS,
        "the reference's own lnprob through oracle/ref_loader.py" if kind == "loader" else "numpy oracle (the reference tree is not on this box)")
    line = {
        "impl": "reference", "metric": "walker lnprob evals/sec (adv TOF model)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * procs * per_proc / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "adv TOF model sweep: %d walkers x %d TOF bins x %d MC draws" % (N_WALKERS, N_TOF_BINS, N_DRAWS),
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": procs, "kind": "reference" if kind == "loader" else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import mcmctoffitting_b200 as M
    from mcmctoffitting_b200.ensemble import EnsembleSampler
    from oracle import tof_oracle as O     # checker side only: builds the synthetic observables / CPU baseline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    om, z, obs, thetas = workload(O)
    ode = {"rk4": M.config.ODE_RK4, "range": M.config.ODE_RANGE}[args.ode]
    f32 = args.precision == "fp32"
    if f32 and ode != M.config.ODE_RANGE:
        raise SystemExit("--precision fp32 needs --ode range")
    cfg = M.config.sweep(ode_mode=ode, precision=M.config.PRECISION_FP32 if f32 else M.config.PRECISION_FP64)
    flop_per_eval = FLOP_PER_EVAL_RANGE if ode == M.config.ODE_RANGE else FLOP_PER_EVAL_RK4
    fn = M.make_lnprob(cfg, obs, z, device=local_rank)
    model = fn.model
    sampler = EnsembleSampler(N_WALKERS, 2, fn, seed=1234, store_chain=False)

    pos = torch.from_numpy(thetas).to(device)
    lp = sampler.initial_lnprob(pos)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = model.stats()["kernel_launches"]
    for _ in range(args.warmup):
        sampler.run_device(pos, lp, 1)
    barrier()
    launches_warm = model.stats()["kernel_launches"]

    clocks = ClockSampler(local_rank)
    if rank == 0:                          # NVML polling takes driver locks: one poller per box is enough
        clocks.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    model.set_timing(True)
    kernel_ms = []
    barrier()
    for i in range(args.steps):
        flush.zero_()                      # L2 flush between timed steps, outside the event bracket
        starts[i].record()
        sampler.run_device(pos, lp, 1)
        stops[i].record()
        stops[i].synchronize()
        kernel_ms.append(model.last_kernel_ms())   # the second half-step's model kernel
    barrier()
    model.set_timing(False)
    clock_info = clocks.finish()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    launches_timed = model.stats()["kernel_launches"] - launches_warm
    value = N_WALKERS * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host API (HOST buffers in, HOST lnprob out) -------------------------------
    per_rank = N_WALKERS // world
    # pinned host buffers for the inputs (positions) and the result (lnprob)
    pin_in = torch.from_numpy(np.ascontiguousarray(thetas[rank * per_rank:(rank + 1) * per_rank])).pin_memory()
    pin_out = torch.empty(per_rank // 2, dtype=torch.float64).pin_memory()
    host_thetas, host_out = pin_in.numpy(), pin_out.numpy()
    half = per_rank // 2
    fn.batch(host_thetas[:half], host_out)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        fn.batch(host_thetas[:half], host_out)   # one call per half-ensemble, as emcee's _get_lnprob does
        fn.batch(host_thetas[half:], host_out)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = N_WALKERS * e2e_steps / float(e2e_s.item())

    finite_frac = float(torch.isfinite(lp).double().mean().item())

    # ---- self-check (outside the timed region): a random sample of the timed ensemble against the oracle -----------
    parity_sample = None
    if rank == 0 and not args.no_parity_sample:
        pick = np.random.RandomState(2026).choice(N_WALKERS, size=256, replace=False)
        pos_h, lp_h = pos.cpu().numpy(), lp.cpu().numpy()
        if f32:
            parity_sample = {"skipped": "FP32 arm: tolerance 1e-4, see fp32_mode"}
        else:
            # in a fresh interpreter: no fork of a process that holds a CUDA context
            import subprocess
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "sample.npz")
                np.savez(path, thetas=pos_h[pick], got=lp_h[pick])
                res = subprocess.run([sys.executable, os.path.abspath(__file__), "--oracle-check", path], stdout=subprocess.PIPE,
                                     text=True, timeout=900)
            parity_sample = (json.loads(res.stdout.strip().splitlines()[-1]) if res.returncode == 0 and res.stdout.strip()
                             else {"error": "rc=%d" % res.returncode})

    # ---- burn-in regime: walkers uniform over the adv prior box (adv:81-82), one half-ensemble call ----------------
    prior_box = None
    if rank == 0 and not f32 and ode == M.config.ODE_RANGE and not args.no_extras:
        rs_box = np.random.RandomState(3)
        nb = N_WALKERS // 2 // world
        box = np.column_stack([rs_box.uniform(1000, 2600, nb), rs_box.uniform(0.02, 0.5, nb)])
        th_box = torch.from_numpy(box).to(device)
        out_box = torch.empty(nb, dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        model.set_timing(True)
        ms_box = []
        for i in range(4):
            flush.zero_()
            model.lnprob_batch_device(th_box.data_ptr(), nb, out_box.data_ptr(), stream)
            torch.cuda.synchronize()
            if i:
                ms_box.append(model.last_kernel_ms())
        model.set_timing(False)
        st_box = model.stats()
        prior_box = {"value": nb / (statistics.mean(ms_box) * 1e-3), "unit": "evals/s", "kernel_ms": statistics.mean(ms_box),
                     "walkers": nb, "wide_walkers": st_box.get("wide_last"), "launches_per_call": st_box.get("model_launches_per_call"),
                     "what": "walkers uniform over the adv prior box 1000<e0<2600, 0.02<sigma0<0.5 (a burn-in ensemble): "
                             "most E-bands exceed the shared-memory histogram and are kept in the L2 scratch slice"}

    # ---- per-stage breakdown from the library's own stage timing (separate instrumented kernels, outside the
    # timed region): share of CTA cycles per stage over one ensemble step --------------------------------------
    stage_profile = None
    if rank == 0 and not f32 and ode == M.config.ODE_RANGE and not args.no_stage_profile:
        model.set_stage_timing(True)
        th_prof = torch.from_numpy(thetas).to(device)
        out_prof = torch.empty(N_WALKERS // 2, dtype=torch.float64, device=device)
        per = N_WALKERS // 2 // world
        model.lnprob_batch_device(th_prof.data_ptr(), per, out_prof.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        torch.cuda.synchronize()
        stage_profile = model.stage_profile()
        stage_profile["what"] = ("tof_set_stage_timing: SM clock cycles per stage summed over CTAs, one call of %d walkers, "
                                 "instrumented instantiation of the same kernel" % per)
        model.set_stage_timing(False)

    # ---- per-evaluation draws (the reference's own behaviour: every lnlike call draws its own normals, adv:128) ----
    fresh_draws = None
    if rank == 0 and not f32 and ode == M.config.ODE_RANGE and not args.no_extras:
        fnf = M.make_lnprob(cfg, obs, None, device=local_rank, fresh_seed=20260101)
        mf = fnf.model
        nb = N_WALKERS // 2 // world
        th_f = torch.from_numpy(thetas[:nb].copy()).to(device)
        out_f = torch.empty(nb, dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        mf.set_timing(True)
        ms_f = []
        for i in range(4):
            flush.zero_()
            mf.lnprob_batch_device(th_f.data_ptr(), nb, out_f.data_ptr(), stream)
            torch.cuda.synchronize()
            if i:
                ms_f.append(mf.last_kernel_ms())
        fresh_draws = {"value": nb / (statistics.mean(ms_f) * 1e-3), "unit": "evals/s", "kernel_ms": statistics.mean(ms_f),
                       "walkers": nb, "finite_lnprob_fraction": float(torch.isfinite(out_f).double().mean().item()),
                       "what": "tof_set_draw_mode(TOF_DRAWS_PER_EVALUATION): every (call, walker) draws its own 1024 sorted "
                               "normals on the device (Philox + exponential spacings + inverse normal CDF, no sort) and is run by "
                               "the per-walker-lookup kernels (adv_planned_kernel + overflow launch); the headline uses the bound "
                               "draw set (common random numbers), the mode the parity goldens pin"}
        mf.close()

    # ---- the reference's own adv configuration (BASELINE.json configs[3]): 4096 walkers x 1e5 draws, 50 TOF bins ------
    adv_c3 = None
    if rank == 0 and world == 1 and not f32 and ode == M.config.ODE_RANGE and not args.no_extras:
        cfg3 = M.config.adv(0, ode_mode=M.config.ODE_RANGE, mean_excitation=19.2e-3)
        rs3 = np.random.RandomState(5)
        fn3 = M.make_lnprob(cfg3, np.ones(cfg3.tof_bins[0]), rs3.standard_normal(cfg3.n_draws), device=local_rank)
        m3 = fn3.model
        # observables: a realisation of the model at the script's region of interest (made with the library: a timing figure)
        real = np.rint(1e4 * m3.model_batch(np.array([[1050.0, 0.10]]), run=0, stage="spread")[0])
        fn3.bind_observables(real)
        n3 = 4096
        th3 = torch.from_numpy(np.array([1050.0, 0.10]) + np.array([5.0, 5e-3]) * rs3.standard_normal((n3, 2))).to(device)
        out3 = torch.empty(n3, dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        m3.set_timing(True)
        ms3 = []
        for i in range(3):
            m3.lnprob_batch_device(th3.data_ptr(), n3, out3.data_ptr(), stream)
            torch.cuda.synchronize()
            if i:
                ms3.append(m3.last_kernel_ms())
        adv_c3 = {"value": n3 / (statistics.mean(ms3) * 1e-3), "unit": "evals/s", "kernel_ms": statistics.mean(ms3), "walkers": n3,
                  "draws_per_walker": cfg3.n_draws, "finite_lnprob_fraction": float(torch.isfinite(out3).double().mean().item()),
                  "launches_per_call": m3.stats().get("model_launches_per_call"),
                  "what": "tests/advIntermediateTOFmodel.py at its own sizes (BASELINE.json configs[3]: 4096 walkers, nDraws 1e5, "
                          "100 x 240 grid, 50 TOF bins), physical mean excitation energy; kernel time of one ensemble-sized call "
                          "(adv_zrank_multi_kernel: tiles of sorted draws)"}
        m3.close()

    # ---- optional FP32 sample stage, reported beside the FP64 headline (N = 1, range formulation) ----------
    fp32_mode = None
    if world == 1 and not f32 and ode == M.config.ODE_RANGE and not args.no_fp32:
        fn32 = M.make_lnprob(M.config.sweep(ode_mode=ode, precision=M.config.PRECISION_FP32), obs, z, device=local_rank)
        m32 = fn32.model
        half_n = N_WALKERS // 2
        th_dev = torch.from_numpy(thetas).to(device)
        out32 = torch.empty(N_WALKERS, dtype=torch.float64, device=device)
        out64 = torch.empty(N_WALKERS, dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        for h in (0, 1):
            model.lnprob_batch_device(th_dev[h * half_n:].data_ptr(), half_n, out64[h * half_n:].data_ptr(), stream)
        m32.set_timing(True)
        ms32 = []
        for i in range(3 + max(1, min(args.steps, 5))):
            flush.zero_()
            for h in (0, 1):
                m32.lnprob_batch_device(th_dev[h * half_n:].data_ptr(), half_n, out32[h * half_n:].data_ptr(), stream)
                torch.cuda.synchronize()
                if i >= 3:
                    ms32.append(m32.last_kernel_ms())
        both = torch.isfinite(out32) & torch.isfinite(out64)
        dev32 = ((out32[both] - out64[both]).abs() / out64[both].abs())
        fp32_mode = {
            "value": half_n / (statistics.mean(ms32) * 1e-3), "unit": "evals/s", "kernel_ms": statistics.mean(ms32),
            "what": "precision=PRECISION_FP32: phase-1 sample arithmetic in FP32, reductions and later stages in FP64; "
                    "kernel time of one half-ensemble call (device-resident)",
            "max_rel_dev_vs_fp64": float(dev32.max().item()), "tolerance": 1e-4,
            "finiteness_agreement": float((torch.isfinite(out32) == torch.isfinite(out64)).double().mean().item()),
        }
        m32.close()

    if rank == 0:
        fp64_peak = model.measure_fp64_peak()
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        k_ms = statistics.mean(kernel_ms)
        evals_per_launch = N_WALKERS // 2 // world
        achieved = evals_per_launch * flop_per_eval / (k_ms * 1e-3) / 1e12
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE model call at this size (131072 walkers, one launch of
            # adv_zrank_kernel), from the ncu --set full capture summarised in profiles/r2_zrank_fp64_ncu_summary.txt;
            # algorithmic bytes are evals * 24
            "traffic": (2586368 * evals_per_launch // 131072) if ode == M.config.ODE_RANGE else None,
            "traffic_source": "profiles/r2_zrank_fp64_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum of one "
                              "131072-walker launch, scaled by walkers per launch); algorithmic bytes are evals * 24",
            "kernel": ("adv_zrank_kernel<512,7> (one persistent launch per call, 2 CTAs/SM)"
                       if ode == M.config.ODE_RANGE else "adv_lnprob_kernel"),
            "kernel_ms": k_ms, "evals_per_launch": evals_per_launch,
            "flop_per_eval": flop_per_eval, "flop_per_eval_rk4_formulation": FLOP_PER_EVAL_RK4,
            # executed sample work only (per (draw, x): v, compare, dt, degree-7 Horner, accumulate = 18; T1 lookup 18 per
            # draw), without the per-cell / per-TOF-bin terms the kernel legitimately skips for empty cells and bins
            "flop_per_eval_executed": FLOP_PER_EVAL_EXECUTED,
            "frac_executed": (evals_per_launch * FLOP_PER_EVAL_EXECUTED / (k_ms * 1e-3) / 1e12) / fp64_peak,
            "rk4_equivalent_tflops": evals_per_launch * FLOP_PER_EVAL_RK4 / (k_ms * 1e-3) / 1e12,
            "peak_source": "DFMA microbenchmark run in this process (tof_measure_fp64_peak)",
            "nominal_fp64_tflops": NOMINAL_FP64_TFLOPS,
            "hbm": {"achieved_gbs": evals_per_launch * BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "bytes_per_eval": BYTES_PER_EVAL},
            "note": "path is FP64-pipe bound (SURVEY.md 8d): neither HBM nor tensor cores limit it",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # a fresh interpreter: no fork of a process that holds a CUDA context
            import subprocess
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline"], stdout=subprocess.PIPE,
                                 text=True, timeout=600)
            cpu = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else {"error": "rc=%d" % out.returncode}
        line = {
            "metric": "walker lnprob evals/sec (adv TOF model)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 samples / f64 reductions" if f32 else "f64", "data": "synthetic",
            "config": {"workload": "adv TOF model sweep: %d walkers x %d TOF bins x %d MC draws" % (N_WALKERS, N_TOF_BINS, N_DRAWS),
                       "step": "one ensemble MCMC step = 2 red/blue half-steps = %d lnprob evaluations" % N_WALKERS,
                       "parallelism": "walkers sharded over %d rank(s); all_gather of the updated half per half-step" % world,
                       "ode": ("range-energy tables (exact solution of the autonomous Bethe ODE)" if ode == M.config.ODE_RANGE
                               else "rk4 x%d per x-interval" % cfg.ode_substeps), "threads_per_cta": model.stats()["threads"],
                       "l2": "flushed between timed steps (256 MiB memset outside the event bracket)",
                       "finite_lnprob_fraction": finite_frac},
            "roofline": roofline, "cpu_baseline": cpu, "fp32_mode": fp32_mode, "stage_profile": stage_profile,
            "parity_sample": parity_sample, "prior_box": prior_box, "fresh_draws": fresh_draws, "adv_config3": adv_c3,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": N_WALKERS * 2 * 8,
                    "d2h_bytes_per_step": N_WALKERS * 8,
                    "what": "the reference-facing call: TofLnProb.batch = C-ABI tof_lnprob_batch with pinned HOST buffers, two calls "
                            "of %d walkers per step as emcee's _get_lnprob issues them (H2D copy of theta, model kernel, D2H copy of "
                            "lnprob inside the wall-clock region); %d steps, no L2 flush, without the stretch-move propose/accept "
                            "kernels that `value` includes -- which is why it can exceed `value`" % (N_WALKERS // 2, e2e_steps)},
            "gpu_launches": launches_timed, "clocks": clock_info,
            "kernel_stats": model.stats(),
        }
        print(json.dumps(line))
    barrier()
    model.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--ode", choices=["range", "rk4"], default="range", help="stopping-stage formulation of the CUDA path")
    ap.add_argument("--precision", choices=["fp64", "fp32"], default="fp64",
                    help="fp64 (default, the reference's arithmetic) or the optional FP32 sample stage as the measured arm")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary FP32-mode measurement of the default run")
    ap.add_argument("--no-stage-profile", action="store_true", help="skip the per-stage cycle breakdown (tof_set_stage_timing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-sample", action="store_true", help="skip the oracle check of 256 timed-ensemble walkers")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (prior box, per-evaluation draws)")
    ap.add_argument("--cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--oracle-check", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline:
        return run_cpu_baseline()
    if args.oracle_check:
        from oracle import tof_oracle as O
        _, z, obs, _ = workload(O)
        with np.load(args.oracle_check) as f:
            print(json.dumps(oracle_check(z, obs, f["thetas"], f["got"], min(os.cpu_count() or 1, 32))))
        return 0
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
