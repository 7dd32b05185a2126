"""Batched posterior-predictive generation.

``ppcTools.generatePPC`` (utilities/ppcTools.py:283-330) samples parameter vectors from the last 50 steps of
a chain and re-runs ``generateModelData`` for each of them, run by run -- minutes of CPU per hundred samples.
Here the same thing is two batched GPU calls per run: the TOF spectra (``tof_model_batch``) and the integer
(x, E) cell counts (``tof_cell_counts_batch``), whose rows are the reference's ``eN_atEachX`` neutron spectra
(ppcTools.py:170-180); :func:`deuteron_spectra` gives the unweighted deuteron-energy histograms (``eD_atEachX``,
ppcTools.py:151-157) and :func:`sdef_sia_cumulative` the MCNP source card of ppcTools.py:397-422.
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


def deuteron_spectra(model: TofModel, thetas: np.ndarray) -> List[np.ndarray]:
    """Per run ``[n, X, E]`` unweighted histograms of the stopped deuteron energies (last loop only, as the
    reference keeps them: ppcTools.py:141, 151-157; ppcTools_oneBD.py:214, 223-224).  Needs the simult model with
    ``ode_mode=ODE_RK4`` or the oneBD model (``config.onebd()`` / ``config.onebd_ppc()``)."""
    thetas = np.ascontiguousarray(thetas, dtype=np.float64)
    return [model.deuteron_counts(thetas, run=run) for run in range(model.config.n_runs)]


def neutron_spectrum(cell_counts: np.ndarray) -> np.ndarray:
    """ppcTools.py:406-412: the neutron spectra of the FIRST run of every posterior sample, summed along the cell
    and over the samples.  ``cell_counts``: ``[n, X, E]`` of run 0 (the rows are ``eN_atEachX``)."""
    return np.asarray(cell_counts).sum(axis=(0, 1)).astype(np.float64)


def sdef_sia_cumulative(cell_counts: np.ndarray, e_n_centers: np.ndarray, dist_number: int = 100,
                        count_format: str = "%.0f") -> dict:
    """MCNP SDEF ``SI A`` / ``SP`` cards for the neutron distribution marginalised over the cell length
    (ppcTools.makeSDEF_sia_cumulative, ppcTools.py:397-422): energies in MeV with three decimals, counts as
    integers.  ``e_n_centers`` = ``getDDneutronEnergy(eD_binCenters)`` in keV (ppcTools.py:81).
    ``count_format="%.3e"`` gives the card of the oneBD twin (ppcTools_oneBD.py:406-431 writes ``{:.3e}``)."""
    spec = neutron_spectrum(cell_counts)
    if len(spec) != len(e_n_centers):
        raise ValueError("one neutron energy per E-bin is needed")
    energies = "".join(" %.3f" % (e / 1000) for e in e_n_centers)
    weights = "".join((" " + count_format) % c for c in spec)
    return {"si": "si%d a%s" % (dist_number, energies), "sp": "sp%d%s" % (dist_number, weights)}
