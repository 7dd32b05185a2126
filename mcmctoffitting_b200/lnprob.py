"""Reference-shaped entry points: ``lnprob(theta, observables, ...)`` callables and the pool adapter.

The reference hands emcee a plain function plus fixed kwargs (adv:300-302; simultFit.py:713-718)
and parallelises by passing ``threads=`` or ``pool=`` (mpiTOFmodel.py:199-201; simultFit.py:701-706).
:func:`make_lnprob` returns a callable with the same signature per model; :class:`BatchedPool` is
the object to pass as ``pool=``: emcee calls ``pool.map(fn, positions)`` once per half-step and the
whole half-ensemble is evaluated by one GPU launch.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import config as cfgmod
from .config import ModelConfig
from .model import TofModel


class TofLnProb:
    """Callable drop-in for the reference's ``lnprob``.

    adv / intermediate / simple:  ``lnprob(theta, observables)``            (adv:191, simple:112)
    simult:  ``lnprob(theta, observables, standoffDists, tofRanges, nTOFbins, nDraws=...)`` (simultFit.py:444)

    ``observables`` given at call time are compared with the bound ones and re-uploaded only when
    they differ; the geometry arguments of the simultaneous fit must match the configuration (they
    are baked into device tables) and are checked.
    """

    def __init__(self, model: TofModel):
        self.model = model
        self._bound: List[Optional[np.ndarray]] = [None] * model.config.n_runs

    # observables ------------------------------------------------------------------------------------
    def bind_observables(self, observables) -> None:
        cfg = self.model.config
        runs = [observables] if cfg.n_runs == 1 else list(observables)
        if len(runs) != cfg.n_runs:
            raise ValueError("expected %d observable histograms" % cfg.n_runs)
        for r, obs in enumerate(runs):
            arr = np.array(obs, dtype=np.float64, copy=True).ravel()
            if self._bound[r] is None or not np.array_equal(arr, self._bound[r], equal_nan=True):
                self.model.set_observables(arr, r)
                self._bound[r] = arr

    def _check_geometry(self, standoffDists, tofRanges, nTOFbins, nDraws) -> None:
        cfg = self.model.config
        if standoffDists is not None and tuple(float(v) for v in standoffDists) != tuple(cfg.standoffs):
            raise ValueError("standoffDists differ from the configured geometry")
        if tofRanges is not None and tuple(tuple(float(x) for x in r) for r in tofRanges) != tuple(cfg.tof_ranges):
            raise ValueError("tofRanges differ from the configured TOF windows")
        if nTOFbins is not None and tuple(int(v) for v in nTOFbins) != tuple(cfg.tof_bins):
            raise ValueError("nTOFbins differ from the configured TOF binning")
        if nDraws is not None and int(nDraws) != cfg.n_samples:
            raise ValueError("nDraws=%s differs from the configured n_samples=%d" % (nDraws, cfg.n_samples))

    # evaluation -------------------------------------------------------------------------------------
    def batch(self, thetas, out=None) -> np.ndarray:
        return self.model.lnprob_batch(thetas, out) if out is not None else self.model.lnprob_batch(thetas)

    def __call__(self, theta, observables=None, standoffDists=None, tofRanges=None, nTOFbins=None, nDraws=None):
        if observables is not None:
            self.bind_observables(observables)
        self._check_geometry(standoffDists, tofRanges, nTOFbins, nDraws)
        return float(self.model.lnprob_batch(np.asarray(theta, dtype=np.float64).reshape(1, -1))[0])


def make_lnprob(config: ModelConfig, observables, draws, device: int = 0, extra_draws=None,
                sort_draws: bool = False, fresh_seed=None) -> TofLnProb:
    """Build the GPU context and return the reference-shaped ``lnprob`` callable.

    ``draws``: standard normals, ``[n_draws]`` (one run) or one array per run; for the simple model a
    pair ``(u, z)`` in the order the reference consumes its RNG (uniforms first, simple:62-65).
    ``extra_draws``: per-run replacement normals for the simultaneous fit's rejection loop.
    ``draws=None`` with ``fresh_seed=<int>``: per-evaluation draws generated on the device, the reference's own
    behaviour (every ``lnlike`` call draws from the global stream: adv:128, simple:62-64); simple and adv models."""
    model = TofModel(config, device)
    if draws is None:
        if fresh_seed is None:
            raise ValueError("pass the Monte-Carlo draws, or fresh_seed=<int> for per-evaluation draws")
        model.set_draw_mode(True, seed=fresh_seed)
    elif config.kind == cfgmod.KIND_SIMPLE:
        u, z = draws
        model.set_draws(z, 0, 0)
        model.set_draws(u, 0, 1)
    else:
        per_run = [draws] if config.n_runs == 1 else list(draws)
        for r, z in enumerate(per_run):
            model.set_draws(z, r, 0, sort=sort_draws)
        if extra_draws is not None:
            for r, z in enumerate(extra_draws):
                model.set_draws(z, r, 1)
    fn = TofLnProb(model)
    fn.bind_observables(observables)
    return fn


class BatchedPool:
    """Pool-shaped adapter for emcee's ``pool=`` seam (mpiTOFmodel.py:199-201).

    ``map(fn, positions)`` stacks the positions and issues ONE batched evaluation.  ``fn`` must be
    (or wrap, as emcee's ``_function_wrapper`` does through ``.f``) a :class:`TofLnProb`; anything
    else is refused -- there is deliberately no CPU path to fall back to."""

    def __init__(self, lnprob: TofLnProb):
        self.lnprob = lnprob
        self.n_map_calls = 0

    def _resolve(self, fn) -> TofLnProb:
        inner = getattr(fn, "f", fn)
        if inner is not self.lnprob:
            raise TypeError("BatchedPool only evaluates the TofLnProb it was built for")
        kwargs = getattr(fn, "kwargs", None) or {}
        args = getattr(fn, "args", None) or ()
        if args or kwargs:
            obs = kwargs.get("observables", args[0] if args else None)
            if obs is not None:
                self.lnprob.bind_observables(obs)
            names = ("standoffDists", "tofRanges", "nTOFbins", "nDraws")
            geo = dict(zip(names, args[1:]))
            geo.update({k: v for k, v in kwargs.items() if k in names})
            self.lnprob._check_geometry(geo.get("standoffDists"), geo.get("tofRanges"), geo.get("nTOFbins"),
                                        geo.get("nDraws"))
        return self.lnprob

    def map(self, fn, positions: Iterable[Sequence[float]]) -> List[float]:
        target = self._resolve(fn)
        pos = np.array(list(positions), dtype=np.float64)
        self.n_map_calls += 1
        if pos.size == 0:
            return []
        return [float(v) for v in target.batch(pos)]

    # emcee.utils.MPIPool surface used by the reference (mpiTOFmodel.py:187-201, 238)
    def is_master(self) -> bool:
        return True

    def wait(self) -> None:
        return None

    def close(self) -> None:
        return None
