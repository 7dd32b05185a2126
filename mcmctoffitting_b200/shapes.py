"""Energy-distribution shapes on explicit draws.

The reference draws initial deuteron energies from one of three shapes through the global ``np.random`` state; the
GPU kernels take the underlying standard normals explicitly and apply the transform on the device.  These host-side
mirrors state the transforms (bit-exact against numpy/scipy, see tests) for users who prepare draw tables:

* normal     ``np.random.normal(e0, sigma0*e0, n)``                 adv:128            -> ``loc + scale*z``
* lognormal  ``beamE - lognorm.rvs(s, loc=eLoss, scale, n)``        simultFit.py:243-244 -> ``beamE - (scale*exp(s*z) + loc)``
* skew-normal ``utilities/pdfs.py:16-28`` (legacy fallback, only imported by simultFit.py:108-117): two normals per draw
"""
from __future__ import annotations

import math

import numpy as np


def normal(loc, scale, z):
    """``np.random.normal(loc, scale)`` on explicit standard normals ``z``."""
    return loc + scale * np.asarray(z, dtype=np.float64)


def lognormal_loss(beam_energy, s, loc, scale, z):
    """``beamE - scipy.stats.lognorm.rvs(s, loc=loc, scale=scale)`` (simultFit.py:243-244; csi_oneBD.py:438-439)."""
    return beam_energy - (np.exp(s * np.asarray(z, dtype=np.float64)) * scale + loc)


def skewnorm_rvs(a, loc, scale, z0, z1):
    """``pdfs.skewnorm.rvs`` (pdfs.py:16-28): ``u0 = scale*z0``, ``v = scale*z1``,
    ``d = a/sqrt(1+a^2)``, ``u1 = d*u0 + v*sqrt(1-d^2)``, ``where(u0 >= 0, u1, -u1) + loc``."""
    u0 = scale * np.asarray(z0, dtype=np.float64)
    v = scale * np.asarray(z1, dtype=np.float64)
    d = a / np.sqrt(1 + a ** 2)
    u1 = d * u0 + v * np.sqrt(1 - d ** 2)
    return np.where(u0 >= 0, u1, -u1) + loc


def skewnorm_pdf(x, loc=0.0, a=0.0, scale=1.0):
    """``pdfs.skewnorm.pdf`` (pdfs.py:12-14): ``2 phi(t) Phi(a t) / scale`` with ``t = (x-loc)/scale``."""
    t = (np.asarray(x, dtype=np.float64) - loc) / scale
    phi = np.exp(-t ** 2 / 2.0) / math.sqrt(2 * math.pi)
    Phi = 0.5 * np.vectorize(math.erfc)(-a * t / math.sqrt(2.0))        # norm.cdf, accurate in the lower tail
    return 2 * phi * Phi / scale
