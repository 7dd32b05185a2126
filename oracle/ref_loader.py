"""Test-only loader that executes the UNMODIFIED reference scripts' function bodies.

TEST INFRASTRUCTURE -- never imported by the product package.

The reference (gcrich/mcmcTOFfitting) model scripts cannot be imported: they import
matplotlib/emcee (absent here), parse argv, read author-machine data files and start MCMC at
import time (tests/advIntermediateTOFmodel.py:23-24,216,300-347).  This loader reads a script
from ``/root/reference`` at run time, ``exec``s only its *prefix* (everything before the first
data-file read) into a fresh namespace with stub plotting/sampler modules, and hands the
namespace back so that ``lnlike/lnprob/generateModelData`` can be called after
``np.random.seed(s)``.  Nothing is copied into this repository; the loader only works where
``/root/reference`` exists (the build container), which is why the values it produces are
committed as fixtures under ``tests/golden/`` by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("TOF_REFERENCE_ROOT", "/root/reference")

# last line (1-based, inclusive) of each script that is safe to execute: everything before the
# first hard-coded data-file read / plotting / sampler construction.
PREFIX_LINES = {
    "simpleTOFmodel": 121,          # model + lnprob end at simpleTOFmodel.py:120
    "intermediateTOFmodel": 204,    # intermediateTOFmodel.py:191-199 is lnprob
    "advIntermediateTOFmodel": 204,  # advIntermediateTOFmodel.py:191-199 is lnprob
    "simultFit": 518,               # simultFit.py:444-469 is lnprob, data read at 521
    "csi_oneBD": 700,
    "devShapeTemplates": 268,       # generateModelData 195-244, template bounds 246-253, buildModelTOF 256-268
}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tests"))


class _LinspaceShim:
    """numpy >= 1.18 rejects float ``num`` in linspace; the reference passes one at
    utilities/utilities.py:249-254.  Coerce with int() exactly as old numpy did."""

    def __init__(self):
        self._orig = np.linspace

    def __call__(self, start, stop, num=50, *a, **k):
        return self._orig(start, stop, int(num), *a, **k)


def load(script: str, argv: list[str] | None = None, prefix_lines: int | None = None) -> dict:
    """Exec the prefix of ``/root/reference/tests/<script>.py``; return its namespace."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, "tests", script + ".py")
    n = prefix_lines or PREFIX_LINES[script]
    with open(path, "r") as f:
        src = "".join(f.readlines()[:n])
    stubs = {}
    for name in ("matplotlib", "matplotlib.pyplot", "emcee", "emcee.utils", "corner"):
        stubs[name] = mock.MagicMock(name=name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns: dict = {"__name__": "ref_" + script, "__file__": path}
    old_argv = sys.argv
    shim = _LinspaceShim()
    try:
        sys.argv = [path] + list(argv or [])
        with mock.patch.dict(sys.modules, stubs):
            np.linspace = shim
            try:
                exec(compile(src, path, "exec"), ns)
            finally:
                np.linspace = shim._orig
    finally:
        sys.argv = old_argv
    return ns


def load_utilities() -> types.ModuleType:
    """Import the reference's importable library modules (utilities, ionStopping, constants)."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    shim = _LinspaceShim()
    np.linspace = shim
    try:
        import utilities.utilities as uu  # noqa
        import utilities.ionStopping as ion  # noqa
        import constants.constants as cc  # noqa
        # beamTimingShape() calls linspace with a float num in its constructor
        uu._beamTiming_instance = uu.beamTimingShape()
    finally:
        np.linspace = shim._orig
    mod = types.SimpleNamespace(utilities=uu, ionStopping=ion, constants=cc)
    return mod
