"""Batched posterior-predictive generation.

``ppcTools.generatePPC`` (utilities/ppcTools.py:283-330) samples parameter vectors from the last 50 steps of
a chain and re-runs ``generateModelData`` for each of them, run by run -- minutes of CPU per hundred samples.
Here the same thing is two batched GPU calls per run: the TOF spectra (``tof_model_batch``) and the integer
(x, E) cell counts (``tof_cell_counts_batch``), whose rows are the reference's ``eN_atEachX`` neutron spectra
(ppcTools.py:170-180).  The unweighted deuteron-energy histograms (``eD_atEachX``) are not produced.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .model import TofModel


def sample_posterior(chain: np.ndarray, n_entries: int, last_steps: int = 50, rng: Optional[np.random.RandomState] = None
                     ) -> np.ndarray:
    """Parameter vectors drawn like ppcTools.py:293-300: uniformly (with repetition) from the flattened last
    ``last_steps`` steps of ``chain[step, walker, param]`` (the layout of ``readChainFromFile``)."""
    rng = rng or np.random.RandomState()
    flat = chain[-last_steps:].reshape(-1, chain.shape[-1])
    return flat[rng.randint(0, flat.shape[0], size=n_entries)]


def generate_ppc(model: TofModel, thetas: np.ndarray) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """For every run: ``(spectra[n, T_run], cell_counts[n, X, E])`` of the posterior samples ``thetas``."""
    thetas = np.ascontiguousarray(thetas, dtype=np.float64)
    spectra, cells = [], []
    for run in range(model.config.n_runs):
        spectra.append(model.model_batch(thetas, run=run, stage="spread"))
        cells.append(model.cell_counts(thetas, run=run))
    return spectra, cells


def ppc_bands(spectra: np.ndarray, quantiles=(0.16, 0.5, 0.84)) -> np.ndarray:
    """Per-bin quantile bands of a stack of PPC spectra ``[n, T]`` (what the PPC plots draw)."""
    return np.quantile(spectra, quantiles, axis=0)
