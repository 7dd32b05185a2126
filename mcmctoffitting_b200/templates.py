"""Template pre-computation for the template-fitting variant.

``tests/devShapeTemplates.py:195-244`` builds, for each of 32 slices ``[e0, e1)`` of initial deuteron energy, a
TOF template with the adv pipeline where ``eZeros = np.random.uniform(e0, e1, n)`` (line 210) and the deuteron
flight time uses ``e0`` (line 231).  ``uniform(e0, e1) == e0 + (e1 - e0) * u``, which is the adv model's
``e0 + (sigma0 * e0) * z`` with ``z = u`` and ``sigma0 = (e1 - e0) / e0``: the templates fall out of one batched
``tof_model_batch`` call with uniform "draws" bound instead of normals.
"""
from __future__ import annotations

import numpy as np

from .model import TofModel


def template_thetas(bounds) -> np.ndarray:
    """``[n, 2]`` adv parameter vectors for the energy slices ``bounds[k] .. bounds[k+1]``
    (``templateEnergyBounds``, devShapeTemplates.py:252-253)."""
    b = np.asarray(bounds, dtype=np.float64)
    lo, hi = b[:-1], b[1:]
    return np.column_stack([lo, (hi - lo) / lo])


def build_templates(model: TofModel, bounds, uniforms, stage: str = "spread") -> np.ndarray:
    """Templates ``[n_slices, T]``.  ``uniforms``: ``U[0,1)`` numbers, ``config.n_draws`` of them, bound as the
    model's draws (they replace the normals of the adv model)."""
    model.set_draws(np.asarray(uniforms, dtype=np.float64))
    return model.model_batch(template_thetas(bounds), stage=stage)


def build_model_tof(coeffs, templates) -> np.ndarray:
    """``buildModelTOF`` (devShapeTemplates.py:256-268): ``scale * sum_k c_k * template_k``."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    return coeffs[0] * (coeffs[1:, None] * np.asarray(templates)[:len(coeffs) - 1]).sum(axis=0)
