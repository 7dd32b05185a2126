"""Data formats either side of the lnprob path.

* measured TOF histograms: tab-separated ``lowEdge  run0 ... runN`` rows (``readMultiStandoffTOFdata``,
  utilities/utilities.py:198-216) and the window selection every script applies to them
  (advIntermediateTOFmodel.py:219-224; simultFit.py:524-532);
* chain files: see :func:`mcmctoffitting_b200.ensemble.write_chain_step` / ``read_chain``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def read_multi_standoff_tof(filename: str, n_runs: int = 4) -> np.ndarray:
    """``[n_bins, 1 + n_runs]`` array: column 0 the lower bin edges, then one column of counts per run."""
    rows = []
    with open(filename, "r") as fh:
        for line in fh:
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 1 + n_runs or not parts[0].strip():
                continue
            rows.append([float(v) for v in parts[:1 + n_runs]])
    return np.array(rows, dtype=np.float64)


def select_window(tof_data: np.ndarray, run: int, tof_range: Tuple[float, float]) -> np.ndarray:
    """Observed counts of ``run`` inside ``[min, max)`` of the lower bin edges (adv:223)."""
    edges = tof_data[:, 0]
    keep = (edges >= tof_range[0]) & (edges < tof_range[1])
    return tof_data[:, run + 1][keep]


def observables_for(config, tof_data: np.ndarray) -> List[np.ndarray]:
    """One observed histogram per run of ``config`` (simultFit.py:528-532)."""
    return [select_window(tof_data, r, config.tof_ranges[r]) for r in range(config.n_runs)]


def write_multi_standoff_tof(filename: str, low_edges: Sequence[float], counts: np.ndarray) -> None:
    """Inverse of :func:`read_multi_standoff_tof` (``counts[n_bins, n_runs]``)."""
    counts = np.asarray(counts, dtype=np.float64)
    with open(filename, "w") as fh:
        for e, row in zip(low_edges, counts):
            fh.write("\t".join([repr(float(e))] + [repr(float(v)) for v in row]) + "\n")
