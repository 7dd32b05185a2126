"""TofModel: Python owner of one ``tof_ctx`` (include/tofgpu.h) -- tables in, batched lnprob out."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from . import config as cfgmod
from .config import ModelConfig

STAGES = {"counts": 0, "pdf": 1, "spread": 2}


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _as_f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        out = out.reshape(shape)
    return out


class TofModel:
    """A model configuration resident on one GPU.

    ``lnprob_batch(thetas)`` evaluates what the reference evaluates one walker at a time through
    ``lnprob(theta, observables)`` (adv:191-199 and friends); observables and the Monte-Carlo draws
    are bound once with :meth:`set_observables` / :meth:`set_draws`.
    """

    def __init__(self, config: ModelConfig, device: int = 0):
        config.validate()
        self.config = config
        self.device = int(device)
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        self._keep = []  # host tables must outlive tof_create only, kept for introspection
        c = _lib.TofConfig()
        c.abi_version = _lib.ABI_VERSION
        c.model = config.kind
        c.device = self.device
        c.ode_mode = config.ode_mode
        c.precision = config.precision
        c.ode_substeps = config.ode_substeps
        c.ode_from_zero = int(config.ode_from_zero)
        c.prior_strict = int(config.prior_strict)
        c.nan_to_neginf = int(config.nan_to_neginf)
        c.ndim = config.ndim
        c.n_runs = config.n_runs
        c.n_samples = config.n_samples
        c.n_ev_per_loop = config.n_ev_per_loop
        c.n_loops = config.n_loops
        c.speed_of_light = cfgmod.SPEED_OF_LIGHT
        c.mass_deuteron = cfgmod.MASS_DEUTERON
        c.mass_neutron = cfgmod.MASS_NEUTRON
        c.mass_he3 = cfgmod.MASS_HE3
        c.q_ddn = cfgmod.Q_DDN
        c.cell_length = cfgmod.distances.cellLength
        c.simple_neutron_base = cfgmod.distances.cellToZero
        for i, (lo, hi) in enumerate(config.prior):
            c.prior_lo[i] = lo
            c.prior_hi[i] = hi
        for r in range(config.n_runs):
            c.tof_bins[r] = config.tof_bins[r]
            c.tof_min[r], c.tof_max[r] = config.tof_ranges[r]
        if config.kind != cfgmod.KIND_SIMPLE:
            c.x_bins, c.e_bins = config.x_bins, config.e_bins
            c.x_min, c.x_max = config.x_range
            c.e_min, c.e_max = config.e_range
            A, B = cfgmod.bethe_reduced(config.materials)
            c.n_materials = len(A)
            for k in range(len(A)):
                c.bethe_A[k], c.bethe_B[k] = A[k], B[k]
            c.ode_substeps = max(config.ode_substeps, 1)
            tabs = dict(
                x_centers=_as_f64(config.x_centers()),
                e_centers=_as_f64(config.e_centers()),
                neutron_speed=_as_f64(config.neutron_speed()),
                neutron_dist=_as_f64(config.neutron_dist()),
                xs_breaks=_as_f64(cfgmod.DDN_XS_ENERGIES),
                xs_coefs=_as_f64(cfgmod.not_a_knot_cubic(cfgmod.DDN_XS_ENERGIES, cfgmod.DDN_XS_SIGMA0)),
                taps=_as_f64(config.taps),
            )
            c.n_xs = tabs["xs_breaks"].shape[0]
            c.n_taps = tabs["taps"].shape[0]
            c.n_zero_deg = config.n_zero_deg
            if config.n_zero_deg:
                t, w = cfgmod.zero_degree_tables(cfgmod.dd_neutron_energy(config.e_centers()), config.n_zero_deg)
                tabs["zero_deg_times"], tabs["zero_deg_weights"] = _as_f64(t), _as_f64(w)
            if config.kind == cfgmod.KIND_ONEBD:
                tabs["stop_coefs"] = _as_f64(config.stop_coefs())
                tabs["attenuation"] = _as_f64(config.attenuation())
                tabs["taps2"] = _as_f64(config.taps2)
                grid = config.stop_energy_grid()
                c.stop_n, c.n_taps2 = len(grid), len(config.taps2)
                c.stop_lo, c.stop_step = float(grid[0]), float(grid[1] - grid[0])
                c.beam_energy = config.beam_energy
            for name, arr in tabs.items():
                setattr(c, name, _dptr(arr))
            self._keep.append(tabs)
            self.tables = tabs
            self.range_tables = None
            if config.ode_mode == cfgmod.ODE_RANGE:
                from . import range_tables as rt
                rtab = rt.build_cached(config)
                self.range_tables = rtab
                c.t1_q, c.t1_key_lo, c.t1_n = rtab.t1_q, rtab.t1_key_lo, rtab.t1_n
                c.rng_degree, c.rng_n, c.rng_lut_n = rtab.degree, len(rtab.bins), len(rtab.lut)
                c.rng_sign, c.rng_u_max = rtab.sign, rtab.u_max
                c.e_tab_lo, c.e_tab_hi = rtab.e_tab_lo, rtab.e_tab_hi
                rt_arrays = dict(t1=_as_f64(rtab.t1_coefs), br=_as_f64(rtab.breaks), co=_as_f64(rtab.coefs),
                                 bins=np.ascontiguousarray(rtab.bins, dtype=np.int32),
                                 lut=np.ascontiguousarray(rtab.lut, dtype=np.uint16))
                c.t1_coefs, c.rng_breaks, c.rng_coefs = _dptr(rt_arrays["t1"]), _dptr(rt_arrays["br"]), _dptr(rt_arrays["co"])
                c.rng_bins = rt_arrays["bins"].ctypes.data_as(C.POINTER(C.c_int32))
                c.rng_lut = rt_arrays["lut"].ctypes.data_as(C.POINTER(C.c_uint16))
                self._keep.append(rt_arrays)
        rc = self._lib.tof_create(C.byref(c), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.tof_last_error(None)
            self._ctx = C.c_void_p()
            raise _lib.TofError(rc, msg.decode() if msg else "")

    # -- lifetime -------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.tof_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        _lib.check(self._lib, self._ctx, rc)

    @property
    def handle(self) -> C.c_void_p:
        return self._ctx

    # -- inputs ---------------------------------------------------------------------------------------
    def set_observables(self, observables, run: int = 0) -> None:
        obs = _as_f64(observables).ravel()
        self._check(self._lib.tof_set_observables(self._ctx, run, _dptr(obs), obs.shape[0]))

    def set_draws(self, values, run: int = 0, stream: int = 0, sort: bool = False) -> None:
        """Bind explicit draws.  ``sort=True`` sorts stream-0 normals first: the model is a sum over
        draws, so their order only matters for speed (neighbouring threads then hit the same
        histogram bins and merge their updates)."""
        v = _as_f64(values).ravel()
        if sort:
            v = np.sort(v)
        self._check(self._lib.tof_set_draws(self._ctx, run, stream, _dptr(v), v.shape[0]))

    # -- evaluation -----------------------------------------------------------------------------------
    def set_draw_mode(self, per_evaluation: bool, seed: int = 0, epoch0: int = 0) -> None:
        """``per_evaluation=True``: every evaluation draws its own Monte-Carlo numbers on the device, as the reference
        does inside every ``lnlike`` call (adv:128; simple:62-64) -- each walker at each step sees its own noise.
        ``False`` (default): the bound draw set, shared by all walkers and calls (the parity mode).
        Call ``epoch0`` is the key of the next batch call; it advances by one per call (tof_set_draw_mode)."""
        self._check(self._lib.tof_set_draw_mode(self._ctx, 1 if per_evaluation else 0, int(seed), int(epoch0)))

    def generate_draws(self, epoch: int, walker: int, n: int, stream: int = 0, sorted: bool = False, run: int = 0) -> np.ndarray:
        """The draws walker ``walker`` of model call ``epoch`` uses for ``run`` in per-evaluation mode (for parity checks):
        stream 0 normals, 1 uniforms, 3 the simultaneous fit's replacement normals."""
        out = np.empty(int(n), dtype=np.float64)
        self._check(self._lib.tof_generate_draws(self._ctx, int(epoch), int(walker), int(run), int(stream), 1 if sorted else 0,
                                                 _dptr(out), int(n)))
        return out

    def _thetas(self, thetas) -> np.ndarray:
        t = _as_f64(thetas)
        if t.ndim == 1:
            t = t.reshape(1, -1)
        if t.ndim != 2 or t.shape[1] != self.config.ndim:
            raise ValueError("thetas must have shape [n, %d]" % self.config.ndim)
        return t

    def lnprob_batch(self, thetas, out: Optional[np.ndarray] = None) -> np.ndarray:
        """``out``: optional preallocated float64 result buffer (e.g. a view of pinned host memory)."""
        t = self._thetas(thetas)
        if out is None:
            out = np.empty(t.shape[0], dtype=np.float64)
        elif out.dtype != np.float64 or out.shape != (t.shape[0],) or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape [n]")
        self._check(self._lib.tof_lnprob_batch(self._ctx, _dptr(t), t.shape[0], _dptr(out)))
        return out

    def lnprob_batch_device(self, theta_ptr: int, n: int, out_ptr: int, stream: int = 0) -> None:
        """Device pointers (e.g. ``tensor.data_ptr()``), asynchronous on ``stream``."""
        self._check(self._lib.tof_lnprob_batch_device(self._ctx, C.c_void_p(theta_ptr), n, C.c_void_p(out_ptr),
                                                      C.c_void_p(stream)))

    def model_batch(self, thetas, run: int = 0, stage: str = "spread") -> np.ndarray:
        t = self._thetas(thetas)
        out = np.empty((t.shape[0], self.config.tof_bins[run]), dtype=np.float64)
        self._check(self._lib.tof_model_batch(self._ctx, _dptr(t), t.shape[0], run, STAGES[stage], _dptr(out)))
        return out

    def cell_counts(self, thetas, run: int = 0) -> np.ndarray:
        t = self._thetas(thetas)
        out = np.empty((t.shape[0], self.config.x_bins, self.config.e_bins), dtype=np.int64)
        self._check(self._lib.tof_cell_counts_batch(self._ctx, _dptr(t), t.shape[0], run,
                                                    out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def deuteron_counts(self, thetas, run: int = 0) -> np.ndarray:
        """Unweighted per-x histograms of the stopped deuteron energies of the last loop, ``[n, X, E]`` -- the
        ``eD_atEachX`` rows of utilities/ppcTools.py:140-157 (simult model, ODE_RK4) and of
        utilities/ppcTools_oneBD.py:214-224 (oneBD model)."""
        t = self._thetas(thetas)
        out = np.empty((t.shape[0], self.config.x_bins, self.config.e_bins), dtype=np.int64)
        self._check(self._lib.tof_deuteron_counts_batch(self._ctx, _dptr(t), t.shape[0], run,
                                                        out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    # -- sampler kernels (device pointers) ------------------------------------------------------------
    def stretch_propose(self, s_ptr, n, walker0, comp_ptr, n_comp, a, seed, step, half, q_ptr, logzz_ptr, stream=0):
        self._check(self._lib.tof_stretch_propose(self._ctx, C.c_void_p(s_ptr), n, walker0, C.c_void_p(comp_ptr), n_comp,
                                                  a, seed, step, half, C.c_void_p(q_ptr), C.c_void_p(logzz_ptr),
                                                  C.c_void_p(stream)))

    def stretch_accept(self, s_ptr, lnprob_ptr, n, walker0, q_ptr, newlp_ptr, logzz_ptr, seed, step, half,
                       naccept_ptr=0, stream=0):
        self._check(self._lib.tof_stretch_accept(self._ctx, C.c_void_p(s_ptr), C.c_void_p(lnprob_ptr), n, walker0,
                                                 C.c_void_p(q_ptr), C.c_void_p(newlp_ptr), C.c_void_p(logzz_ptr), seed,
                                                 step, half, C.c_void_p(naccept_ptr), C.c_void_p(stream)))

    # -- diagnostics ----------------------------------------------------------------------------------
    def ensemble_step(self, pos_ptr, lnprob_ptr, n_walkers, n_steps, a, seed, step0, naccept_ptr=0, stream=0) -> None:
        """``n_steps`` whole stretch-move steps of an ensemble resident on this GPU, enqueued without host round trips
        (tof_ensemble_step)."""
        self._check(self._lib.tof_ensemble_step(self._ctx, C.c_void_p(pos_ptr), C.c_void_p(lnprob_ptr), n_walkers, n_steps,
                                                a, seed, step0, C.c_void_p(naccept_ptr), C.c_void_p(stream)))

    def ensemble_half_step(self, state_ptr, n_walkers, half, own0, n_own, a, seed, step, naccept_ptr=0, stream=0) -> None:
        """One half-step for the rows this GPU owns of the packed ``[n_walkers, ndim+1]`` state (tof_ensemble_half_step)."""
        self._check(self._lib.tof_ensemble_half_step(self._ctx, C.c_void_p(state_ptr), n_walkers, half, own0, n_own, a, seed,
                                                     step, C.c_void_p(naccept_ptr), C.c_void_p(stream)))

    def stats(self) -> dict:
        s = _lib.TofStats()
        self._check(self._lib.tof_get_stats(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    STAGES = ("setup", "histogram", "normalise", "scatter", "likelihood")

    def set_stage_timing(self, enabled: bool) -> None:
        """Per-stage SM-cycle accounting inside the adv/intermediate range kernel (tof_set_stage_timing)."""
        self._check(self._lib.tof_set_stage_timing(self._ctx, int(enabled)))

    def stage_profile(self) -> dict:
        """Cycles per stage since the last call (and reset): ``{"cycles": {...}, "share": {...}, "walkers": n}``."""
        buf = (C.c_uint64 * (len(self.STAGES) + 1))()
        self._check(self._lib.tof_get_stage_cycles(self._ctx, buf))
        cyc = {k: int(buf[i]) for i, k in enumerate(self.STAGES)}
        tot = max(sum(cyc.values()), 1)
        return {"cycles": cyc, "share": {k: v / tot for k, v in cyc.items()}, "walkers": int(buf[len(self.STAGES)])}

    def set_timing(self, enabled: bool) -> None:
        self._check(self._lib.tof_set_timing(self._ctx, int(enabled)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.tof_last_kernel_ms(self._ctx, C.byref(ms)))
        return float(ms.value)

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._check(self._lib.tof_measure_fp64_peak(self._ctx, C.byref(v)))
        return float(v.value)
