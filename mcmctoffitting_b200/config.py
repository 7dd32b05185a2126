"""Model configurations: the module-level literals of the reference scripts as frozen dataclasses.

The reference keeps every binning, range, density and prior as a literal at the top of each
``tests/*.py`` script ("Config / flag system").  Here they are one dataclass
per model family with presets:

* :func:`simple`        tests/simpleTOFmodel.py:25-28, 106-110        (config 1; also mpiTOFmodel.py)
* :func:`intermediate`  tests/intermediateTOFmodel.py:44-100, 165      (config 2)
* :func:`adv`           tests/advIntermediateTOFmodel.py:34-100, 165   (config 3)
* :func:`sweep`         adv model at the benchmark shape: 1024 draws x 2048 TOF bins
* :func:`simult`        tests/simultFit.py:121-205, 425-435            (config 4)

Host-side tables (bin centres, neutron speeds, flight paths, cross-section spline coefficients,
timing-response taps) are computed here with numpy in the reference's operation order and handed
to the CUDA library through the C ABI (include/tofgpu.h).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence, Tuple

import numpy as np

# ---- constants/constants.py ------------------------------------------------------------------
SPEED_OF_LIGHT = 29.9792        # cm/ns      constants.py:13
MASS_ELECTRON = 511             # keV/c^2    constants.py:21
MASS_DEUTERON = 1.8756e+06      # constants.py:22
MASS_NEUTRON = 939565.0         # constants.py:23
MASS_HE3 = 2.809414e6           # constants.py:25
Q_DDN = 3268.914                # constants.py:93
AVOGADRO = 6.02214076e23        # scipy.constants.Avogadro, used at ionStopping.py:54
BETHE_FIXED_FACTOR = 1.67489e-14  # ionStopping.py:69


class distances:
    """constants.py:37-57 (tunlSSA_CsI)."""
    cellToZero = 518.055
    cellLength = 2.86
    zeroDegLength = 3.81
    tipToColli = 148.4
    colliToZero = 233.8
    delta1 = 131.09
    delta2 = 52.39
    standoffClose = tipToColli + colliToZero
    standoffMid = standoffClose + delta1
    standoffFar = standoffMid + delta2
    colliToCsI = 59.45
    csiToZero = 355.7
    csiDiameter = 2.341
    standoff_TUNLruns = colliToCsI + csiToZero + csiDiameter + tipToColli


class tofWindows:
    """constants.py:105-107."""
    nBins = {"close": 45, "mid": 50, "far": 70, "production": 65}
    maxRange = {"close": 175.0, "mid": 225.0, "far": 260.0, "production": 260.0}
    minRange = {"close": 130.0, "mid": 175.0, "far": 190.0, "production": 195.0}


# ---- kinematics (utilities.py:48-73) -----------------------------------------------------------
def dd_neutron_energy(e_d):
    """getDDneutronEnergy at 0 degrees, reference operation order (utilities.py:48-62)."""
    e_d = np.asarray(e_d, dtype=np.float64)
    r = np.sqrt(MASS_DEUTERON * MASS_NEUTRON * e_d) / (MASS_NEUTRON + MASS_HE3) * np.cos(0 * np.pi / 180)
    s = (e_d * (MASS_HE3 - MASS_DEUTERON) + Q_DDN * MASS_HE3) / (MASS_NEUTRON + MASS_HE3)
    return np.power(r + np.sqrt(np.power(r, 2) + s), 2)


def speed(mass, energy):
    """The velocity term of getTOF (utilities.py:71)."""
    return SPEED_OF_LIGHT * np.sqrt(2 * np.asarray(energy, dtype=np.float64) / mass)


# ---- Bethe stopping power, reduced (ionStopping.py:38-97) --------------------------------------
def bethe_reduced(materials: Sequence[Sequence[float]], ion_charge: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """``simpleBethe.dEdx`` is algebraically ``-(1/E) * sum_k A_k ln(B_k E)`` with
    ``A_k = 2 pi z^2 F n_e,k m_d / (m_e c^4)`` and ``B_k = 4 m_e / (m_d I_k)``; rows of
    ``materials`` are ``(Z, A, rho, I_keV)`` as in ``addMaterial`` (ionStopping.py:71-76)."""
    A, B = [], []
    for Z, Amass, rho, excitation in materials:
        n_e = AVOGADRO * Z * rho / (Amass * 1)                       # ionStopping.py:54-56
        A.append(4 * math.pi * ion_charge ** 2 * BETHE_FIXED_FACTOR * n_e * MASS_DEUTERON /
                 (2.0 * MASS_ELECTRON * SPEED_OF_LIGHT ** 4))
        B.append(4.0 * MASS_ELECTRON / (MASS_DEUTERON * excitation))
    return np.array(A), np.array(B)


# ---- D(d,n) 0-degree cross section (utilities.py:332-429) ---------------------------------------
DDN_XS_ENERGIES = np.array([float(e) for e in range(20, 101, 10)] + [float(e) for e in range(150, 1001, 50)] +
                           [float(e) for e in range(1100, 3001, 100)] + [float(e) for e in range(3500, 10001, 500)])
DDN_XS_SIGMA0 = np.array([
    0.025, 0.125, 0.31, 0.52, 0.78, 1.06, 1.35, 1.66, 2.00, 3.33, 4.6, 5.9, 7.1, 8.3, 9.4, 10.4, 11.4, 12.4, 13.4,
    14.3, 15.1, 15.8, 16.5, 17.2, 17.8, 18.4, 19.0, 20.0, 21.0, 21.9, 22.7, 23.4, 24.0, 24.6, 25.2, 25.8, 26.4, 26.9,
    27.5, 28.0, 28.4, 28.9, 29.3, 29.8, 30.3, 30.7, 31.2, 33.5, 35.7, 37.8, 40.0, 41.5, 42.9, 43.8, 44.6, 45.2, 45.7,
    46.1, 46.4, 46.5, 46.5])


def not_a_knot_cubic(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Piecewise-cubic coefficients of the interpolating not-a-knot spline -- what
    ``interp1d(kind='cubic')`` (utilities.py:412-413) represents -- as ``[n-1, 4]`` power-basis rows
    (highest order first) about each left breakpoint."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    h = np.diff(x)
    slope = np.diff(y) / h
    # unknowns: first derivatives s_i; C2 continuity rows + not-a-knot end rows
    M = np.zeros((n, n))
    rhs = np.zeros(n)
    for i in range(1, n - 1):
        M[i, i - 1] = h[i]
        M[i, i] = 2.0 * (h[i - 1] + h[i])
        M[i, i + 1] = h[i - 1]
        rhs[i] = 3.0 * (h[i] * slope[i - 1] + h[i - 1] * slope[i])
    d = x[2] - x[0]
    M[0, 0] = h[1]
    M[0, 1] = d
    rhs[0] = ((h[0] + 2.0 * d) * h[1] * slope[0] + h[0] ** 2 * slope[1]) / d
    d = x[-1] - x[-3]
    M[-1, -1] = h[-2]
    M[-1, -2] = d
    rhs[-1] = (h[-1] ** 2 * slope[-2] + (2.0 * d + h[-1]) * h[-2] * slope[-1]) / d
    s = np.linalg.solve(M, rhs)
    t = (s[:-1] + s[1:] - 2.0 * slope) / h
    c = np.empty((n - 1, 4))
    c[:, 0] = t / h
    c[:, 1] = (slope - s[:-1]) / h - t
    c[:, 2] = s[:-1]
    c[:, 3] = y[:-1]
    return c


# ---- timing responses (utilities.py:219-329) -----------------------------------------------------
def beam_timing_taps(sigma: float = 1.1910e+00, tau: float = 1.0110e+00, bin_width: float = 1.0) -> np.ndarray:
    """beamTimingShape: exponential (x) Gaussian sampled at bin centres -4.5 ... 10.5 ns and
    normalised (utilities.py:232-273).  16 taps at the reference's sigma/tau."""
    lo = math.ceil(-1.0 * 5 * sigma)
    hi = math.ceil(10 * tau)
    nb = int(hi - lo)
    centers = np.linspace(lo + bin_width / 2, hi - bin_width / 2, nb)
    exp_arg = sigma ** 2 / (2 * tau ** 2) - centers / tau
    erf_arg = (sigma ** 2 - centers * tau) / (np.sqrt(2) * sigma * tau)
    vals = np.exp(exp_arg) * (1 - np.array([math.erf(v) for v in erf_arg]))
    return vals / np.sum(vals)


def gaussian_timing_taps(sigma: float = 2.7) -> np.ndarray:
    """beamTimingShape.gaussianTiming: 11 taps at -20 ... 20 step 4 (utilities.py:303-306)."""
    centers = np.linspace(-20, 20, 11, True)
    vals = np.exp(-(centers / sigma) ** 2 / 2)
    return vals / np.sum(vals)


def zero_degree_tables(e_n: np.ndarray, n_segments: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """zeroDegreeTimingSpread.getTimesAndWeights for every neutron energy (utilities.py:154-192):
    ``[E, n_segments]`` transit times and normalised interaction weights."""
    seg = distances.zeroDegLength / n_segments
    x_locs = np.linspace(seg / 2, distances.zeroDegLength - seg / 2, n_segments)
    times = np.empty((len(e_n), n_segments))
    weights = np.empty((len(e_n), n_segments))
    for j, en in enumerate(e_n):
        times[j] = x_locs / speed(MASS_NEUTRON, en)
        sigma_np = (4.83 / (np.sqrt(en / 1000)) - 0.578) * 1e-24
        w = np.exp(-1 * sigma_np * 4.82e22 * x_locs)
        weights[j] = w / np.sum(w)
    return times, weights


# ---- model configurations --------------------------------------------------------------------------
KIND_SIMPLE, KIND_ADV, KIND_SIMULT, KIND_ONEBD = 1, 2, 3, 4
ODE_RK4, ODE_RANGE = 0, 1
PRECISION_FP64, PRECISION_FP32 = 0, 1     # tof_precision (include/tofgpu.h)


@dataclass(frozen=True)
class ModelConfig:
    """One model family with every literal the reference script fixes at import time."""
    kind: int
    name: str
    ndim: int
    prior: Tuple[Tuple[float, float], ...]
    prior_strict: bool
    tof_bins: Tuple[int, ...]
    tof_ranges: Tuple[Tuple[float, float], ...]
    standoffs: Tuple[float, ...] = ()
    n_samples: int = 0             # multiplier in np.rint(dataHist*nSamples); = lnlike's nDraws
    n_ev_per_loop: int = 0
    n_loops: int = 1
    x_bins: int = 0
    x_range: Tuple[float, float] = (0.0, distances.cellLength)
    e_bins: int = 0
    e_range: Tuple[float, float] = (0.0, 1.0)
    materials: Tuple[Tuple[float, float, float, float], ...] = ()   # (Z, A, rho, I_keV)
    ode_mode: int = ODE_RK4
    ode_substeps: int = 1
    ode_from_zero: bool = False
    precision: int = PRECISION_FP64   # PRECISION_FP32: optional single-precision sample stage (adv/intermediate + ODE_RANGE)
    zero_deg_half_length_in_path: bool = True   # adv adds zeroDegLength/2 (adv:154), simultFit does not (290-291)
    n_zero_deg: int = 0
    nan_to_neginf: bool = False
    taps: Tuple[float, ...] = field(default_factory=lambda: tuple(beam_timing_taps()))
    # oneBD only (csi_oneBD.py:226-295, 407-411)
    beam_energy: float = 0.0
    stop_grid: Tuple[float, float, float] = ()          # np.arange(lo, hi, step) of betheApprox (csi_oneBD.py:293)
    stop_table: Tuple[Tuple[float, ...], ...] = ()      # [len(grid)][x_bins]; () = integrate like the reference does
    attenuation_length: float = 0.0
    taps2: Tuple[float, ...] = ()

    @property
    def n_runs(self) -> int:
        return len(self.tof_bins)

    @property
    def n_draws(self) -> int:
        """Standard-normal draws consumed per run and evaluation."""
        return self.n_loops * self.n_ev_per_loop

    # tables in the reference's operation order -----------------------------------------------------
    def x_centers(self) -> np.ndarray:
        lo, hi = self.x_range
        s = (hi - lo) / self.x_bins
        return np.linspace(lo + s / 2, hi - s / 2, self.x_bins)               # adv:70-73

    def e_centers(self) -> np.ndarray:
        lo, hi = self.e_range
        s = (hi - lo) / self.e_bins
        return np.linspace(lo + s / 2, hi - s / 2, self.e_bins)               # adv:60-63

    def neutron_speed(self) -> np.ndarray:
        return speed(MASS_NEUTRON, dd_neutron_energy(self.e_centers()))          # adv:100,110

    def neutron_dist(self) -> np.ndarray:
        xc = self.x_centers()
        rows = []
        for so in self.standoffs:
            if self.zero_deg_half_length_in_path:
                rows.append(distances.cellLength - xc + distances.zeroDegLength / 2 + so)   # adv:153-155
            else:
                rows.append(distances.cellLength - xc + so)                                   # simultFit:290-291
        return np.ascontiguousarray(np.array(rows, dtype=np.float64))

    # oneBD tables ---------------------------------------------------------------------------------------
    def stop_energy_grid(self) -> np.ndarray:
        return np.arange(*self.stop_grid, dtype=np.float64)

    def stopping_table(self) -> np.ndarray:
        """betheApprox table z[E0_k, x_i] (ionStopping.py:108-128).  The reference integrates it with scipy's
        dopri5 at default tolerances and then interpolates THAT table, so the same integrator is used here
        (scipy is a dependency of the reference); pass ``stop_table`` to supply one instead."""
        if self.stop_table:
            return np.asarray(self.stop_table, dtype=np.float64)
        try:
            from scipy.integrate import ode
        except ImportError as exc:  # pragma: no cover
            raise RuntimeError("scipy is needed to integrate the betheApprox table; pass stop_table=...") from exc
        A, B = bethe_reduced(self.materials)

        def dedx(x, y):   # simpleBethe.dEdx in its own operation order (ionStopping.py:78-97)
            velocity = np.sqrt(2 * y / MASS_DEUTERON) * SPEED_OF_LIGHT
            lead = 4 * np.pi * 1 ** 2 / (MASS_ELECTRON * SPEED_OF_LIGHT ** 2 * velocity ** 2)
            frac = 0
            for Z, Am, rho, exc_ in self.materials:
                n_e = AVOGADRO * Z * rho / (Am * 1)
                frac = frac + n_e * np.log(2 * MASS_ELECTRON / (SPEED_OF_LIGHT ** 2) * velocity ** 2 / exc_)
            return -1 * lead * BETHE_FIXED_FACTOR * frac

        rows = []
        for e_zero in np.arange(*self.stop_grid):
            solver = ode(dedx).set_integrator("dopri5").set_initial_value(e_zero)
            rows.append(np.array([solver.integrate(x) for x in self.x_centers()]).flatten())
        return np.array(rows)

    def stop_coefs(self) -> np.ndarray:
        """Per x column, the not-a-knot cubic in E0 through the table: what RectBivariateSpline (kx=ky=3, s=0;
        ionStopping.py:130) evaluates to along its own x nodes.  [x_bins][n-1][4]."""
        tab = self.stopping_table()
        grid = self.stop_energy_grid()
        return np.ascontiguousarray(np.array([not_a_knot_cubic(grid, tab[:, i]) for i in range(self.x_bins)]))

    def attenuation(self) -> np.ndarray:
        return np.exp(-self.x_centers() / self.attenuation_length)            # initialization.py:35-40

    def validate(self) -> None:
        if self.kind != KIND_SIMPLE:
            # the reference bins x by value (np.histogram2d, adv:134); the kernels index rows directly,
            # which is identical as long as every centre falls in its own bin -- true for linspace centres
            lo, hi = self.x_range
            edges = np.linspace(lo, hi, self.x_bins + 1)
            idx = np.searchsorted(edges, self.x_centers(), side="right") - 1
            if not np.array_equal(idx, np.arange(self.x_bins)):
                raise ValueError("x bin centres are not aligned with their bins")
            if len(self.standoffs) != self.n_runs or len(self.tof_ranges) != self.n_runs:
                raise ValueError("standoffs / tof_ranges / tof_bins must have one entry per run")
        if self.kind == KIND_ADV and self.ode_mode == ODE_RANGE:
            # tof_create refuses these with TOF_ERR_CAPACITY (the range kernel reuses the cell histogram as the
            # density buffer and keeps E-bin indices in 16 bits); fail here with the same wording
            if self.tof_bins[0] > self.x_bins * self.e_bins:
                raise ValueError("ODE_RANGE needs tof_bins <= x_bins*e_bins; use ode_mode=ODE_RK4")
            if self.e_bins > 65535:
                raise ValueError("ODE_RANGE keeps E-bin indices in 16 bits: e_bins must be <= 65535")
        if len(self.prior) != self.ndim:
            raise ValueError("prior needs one (lo, hi) pair per parameter")
        if self.precision not in (PRECISION_FP64, PRECISION_FP32):
            raise ValueError("precision must be PRECISION_FP64 or PRECISION_FP32")
        if self.precision == PRECISION_FP32 and not (self.kind == KIND_ADV and self.ode_mode == ODE_RANGE):
            raise ValueError("PRECISION_FP32 is built for the adv/intermediate model with ode_mode=ODE_RANGE only")


def simple(n_draws: int = 1000000) -> ModelConfig:
    """tests/simpleTOFmodel.py (config 1)."""
    return ModelConfig(kind=KIND_SIMPLE, name="simple", ndim=3,
                       prior=((800.0, 1200.0), (-200.0, 0.0), (10.0, 100.0)), prior_strict=True,   # simple:108
                       tof_bins=(25,), tof_ranges=((175.0, 200.0),),                                # simple:25-28
                       n_samples=n_draws, n_ev_per_loop=n_draws, n_loops=1)


_RUN_NAME = {0: "mid", 1: "close", 2: "close", 3: "far"}
_RUN_STANDOFF = {0: distances.standoffMid, 1: distances.standoffClose, 2: distances.standoffClose,
                 3: distances.standoffFar}


def _window(name):
    return tofWindows.nBins[name], (tofWindows.minRange[name], tofWindows.maxRange[name])


def adv(run: int = 0, n_samples: int = 100000, n_ev_per_loop: int = 100000, mean_excitation: float = 19.2,
        **overrides) -> ModelConfig:
    """tests/advIntermediateTOFmodel.py (config 3).  ``mean_excitation=19.2`` is the script AS
    WRITTEN (adv:94, keV where eV was meant; dE/dx comes out positive); pass 19.2e-3 for the
    physical value used by simultFit.py:195."""
    nb, rng = _window(_RUN_NAME[run])
    kw = dict(kind=KIND_ADV, name="adv", ndim=2, prior=((1000.0, 2600.0), (0.02, 0.5)), prior_strict=True,  # adv:81-82
              tof_bins=(nb,), tof_ranges=(rng,), standoffs=(_RUN_STANDOFF[run],),
              n_samples=n_samples, n_ev_per_loop=n_ev_per_loop, n_loops=int(n_samples / n_ev_per_loop),    # adv:126
              x_bins=100, x_range=(0.0, distances.cellLength), e_bins=240, e_range=(200.0, 2600.0),       # adv:56-73
              materials=((1, 2, 8.565e-5, mean_excitation),), ode_substeps=1, ode_from_zero=False)        # adv:90-97
    kw.update(overrides)
    return ModelConfig(**kw)


def intermediate(run: int = 0, n_samples: int = 1000000, n_ev_per_loop: int = 10000, mean_excitation: float = 19.2,
                 **overrides) -> ModelConfig:
    """tests/intermediateTOFmodel.py (config 2)."""
    nb, rng = _window(_RUN_NAME[run])
    kw = dict(kind=KIND_ADV, name="intermediate", ndim=2, prior=((750.0, 1200.0), (0.02, 0.17)), prior_strict=True,
              tof_bins=(nb,), tof_ranges=(rng,), standoffs=(_RUN_STANDOFF[run],),
              n_samples=n_samples, n_ev_per_loop=n_ev_per_loop, n_loops=int(n_samples / n_ev_per_loop),
              x_bins=100, x_range=(0.0, distances.cellLength), e_bins=150, e_range=(200.0, 1700.0),
              materials=((1, 2, 8.37e-5, mean_excitation),), ode_substeps=1, ode_from_zero=False)
    kw.update(overrides)
    return ModelConfig(**kw)


def sweep(**overrides) -> ModelConfig:
    """Benchmark shape of BASELINE.json / SURVEY.md 8(d): adv model, physical I, 1024 draws,
    2048 TOF bins on [128, 256) ns."""
    kw = dict(n_samples=1024, n_ev_per_loop=1024, mean_excitation=19.2e-3, name="sweep",
              tof_bins=(2048,), tof_ranges=((128.0, 256.0),))
    kw.update(overrides)
    return adv(0, **kw)


def simult(n_samples: int = 200000, n_ev_per_loop: int = 50000, **overrides) -> ModelConfig:
    """tests/simultFit.py (config 4)."""
    names = ["mid", "close", "close", "far", "production"]
    kw = dict(kind=KIND_SIMULT, name="simult", ndim=9,
              prior=((1825.0, 1925.0), (600.0, 1000.0), (40.0, 300.0), (0.1, 1.2)) + ((0.0, 1.0e6),) * 5,  # simultFit:425-435
              prior_strict=False,
              tof_bins=tuple(tofWindows.nBins[n] for n in names),
              tof_ranges=tuple((tofWindows.minRange[n], tofWindows.maxRange[n]) for n in names),
              standoffs=(distances.standoffMid, distances.standoffClose, distances.standoffClose,
                         distances.standoffFar, distances.standoff_TUNLruns),                                # simultFit:127-131
              n_samples=n_samples, n_ev_per_loop=n_ev_per_loop,
              n_loops=int(math.ceil(n_samples / n_ev_per_loop)),                                             # simultFit:239
              x_bins=10, x_range=(0.0, distances.cellLength), e_bins=50, e_range=(200.0, 1200.0),           # simultFit:158-175
              materials=((1, 2, 8.565e-5, 19.2 * 1e-3),), ode_substeps=4, ode_from_zero=True,               # simultFit:191-201,256
              zero_deg_half_length_in_path=False, n_zero_deg=10, nan_to_neginf=True)                         # simultFit:463-468
    kw.update(overrides)
    return ModelConfig(**kw)


class distances_oneBD:
    """constants.py:59-81 (tunlSSA_CsI_oneBD)."""
    standoffClose = 351.3
    standoffMid = standoffClose + (412.3 - 351.3)
    standoffFar = standoffMid + (444.5 - 412.3)


def onebd(n_samples: int = 200000, n_ev_per_loop: int = 10000, **overrides) -> ModelConfig:
    """tests/csi_oneBD.py with its default flags (the production "one-BD" model)."""
    zc = np.linspace(0, 24, 7, True)
    taps2 = np.exp(-zc / 2.) / np.sum(np.exp(-zc / 2.))                                                     # csi_oneBD:407-408
    kw = dict(kind=KIND_ONEBD, name="onebd", ndim=9,
              prior=((200.0, 2000.0), (10.0, 700.0), (0.05, 3.0)) + ((1e3, 1.0e8),) * 3 + ((0.0, 1e3),) * 3,  # csi_oneBD:595-606
              prior_strict=False,
              tof_bins=(25, 25, 25), tof_ranges=((80.0, 180.0), (100.0, 200.0), (120.0, 220.0)),            # constants.py:114-123
              standoffs=(distances_oneBD.standoffClose, distances_oneBD.standoffMid, distances_oneBD.standoffFar),
              n_samples=n_samples, n_ev_per_loop=n_ev_per_loop, n_loops=int(math.ceil(n_samples / n_ev_per_loop)),
              x_bins=10, x_range=(0.0, distances.cellLength), e_bins=100, e_range=(200.0, 2200.0),          # csi_oneBD:199-212
              materials=((1, 2, 4 * 8.565e-5, 19.2 * 1e-3),),                                                # csi_oneBD:270-288
              zero_deg_half_length_in_path=False, n_zero_deg=0, nan_to_neginf=True,
              taps=tuple(gaussian_timing_taps(2.7)),                                                         # csi_oneBD:266
              beam_energy=2490.0, stop_grid=(100, 2400, 100), attenuation_length=20.0, taps2=tuple(taps2))
    kw.update(overrides)
    return ModelConfig(**kw)


def onebd_ppc(n_samples: int = 50000, n_ev_per_loop: int = 10000, **overrides) -> ModelConfig:
    """utilities/ppcTools_oneBD.py: the posterior-predictive twin of the oneBD model.  Same pipeline as
    :func:`onebd` on the grid ``initialize_oneBD`` sets up today (20 x-bins, 400 E-bins; initialization.py:12-33),
    plus the 10 zero-degree sub-times per cell (ppcTools_oneBD.py:246-248) and a causal transit convolution with
    tau = 4 bins (ppcTools_oneBD.py:87-88; csi_oneBD.py:407-408 uses 2).  nSamples 5e4, nEvPerLoop 1e4 (78-79)."""
    zc = np.linspace(0, 24, 7, True)
    taps2 = np.exp(-zc / 4.) / np.sum(np.exp(-zc / 4.))                                                     # ppcTools_oneBD.py:87-88
    kw = dict(name="onebd_ppc", x_bins=20, e_bins=400, n_zero_deg=10, taps2=tuple(taps2))
    kw.update(overrides)
    return onebd(n_samples=n_samples, n_ev_per_loop=n_ev_per_loop, **kw)
