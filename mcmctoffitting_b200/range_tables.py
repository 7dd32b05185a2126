"""Range-energy tables for the autonomous Bethe ODE (TOF_ODE_RANGE).

``simpleBethe.dEdx`` (ionStopping.py:78-97) does not depend on x: ``dE/dx = f(E)``.  For an autonomous
ODE every deuteron follows the same curve, shifted.  With the path coordinate

    u(E) = integral_{e_min}^{E} dE' / |f(E')|            (the classic CSDA range table, re-based at e_min)

a deuteron that starts with ``E0`` at ``x_start`` satisfies ``u(E(x)) = u(E0) + sgn * (x - x_start)``,
``sgn = sign(f)``.  So "integrate the ODE for every draw and every x" (adv:129, simultFit.py:256-258)
collapses to one table lookup ``u0_d = u(E0_d)`` per draw, and each (draw, x_i) sample is the number
``v = u0_d + sgn*(x_i - x_start)``.  Nothing downstream needs the energy itself:

* the E-bin of a sample (adv:134-137) is the interval of ``v`` between the *bin edges mapped to u*;
* its cross-section weight (adv:131) is ``omega(v) = XS(E(v))``, a fixed 1-D function of ``v``.

This module builds, in extended precision on the host,

* ``T1``: ``u(E)`` as degree-7 polynomials on cells indexed by the exponent/mantissa bits of E;
* ``T2``: ``omega(v)`` as degree-``P`` polynomials on intervals whose breakpoints contain every E-bin edge
  and every cross-section spline knot (so one pointer walk yields both the bin and the weight).

Accuracy target: ``|u_hat - u| <= 2e-12 cm`` and ``|omega_hat - omega| <= 2e-13 * omega`` -- far below the
reference's own LSODA tolerance (rtol 1.5e-8); the integer cell counts agree with the RK4 oracle
except for draws that sit within that distance of a bin edge.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

from . import config as cfgmod

LD = np.longdouble
T1_DEGREE = 7
T1_Q = 4           # 2**Q cells per octave of E


@dataclass
class RangeTables:
    sign: float                 # sign of dE/dx on the table domain
    # T1: u(E) -------------------------------------------------------------------------------------
    t1_q: int
    t1_key_lo: int              # (hi32(E) >> (20-Q)) of the first cell
    t1_n: int
    t1_coefs: np.ndarray        # [n, 8] monomials in t in [-1, 1], lowest order first
    e_tab_lo: float
    e_tab_hi: float
    # T2: omega(v) ---------------------------------------------------------------------------------
    degree: int
    u_max: float                # u(e_max); u(e_min) = 0
    breaks: np.ndarray          # [m+1] interval breakpoints in u, breaks[0] = 0, breaks[m] = u_max
    bins: np.ndarray            # [m] E-bin index of each interval (int32)
    coefs: np.ndarray           # [m, degree+1] monomials in (v - breaks[j]), lowest order first
    lut: np.ndarray             # [lut_n] uint16: interval holding the left edge of each uniform cell
    lut_inv: float              # lut_n / u_max
    edges_u: np.ndarray         # [e_bins+1] E-bin edges mapped to u
    max_err_u: float
    max_err_omega: float


def _gl(n=24):
    x, w = np.polynomial.legendre.leggauss(n)
    return x.astype(LD), w.astype(LD)


_GLX, _GLW = _gl()


class _Stopping:
    def __init__(self, materials):
        A, B = cfgmod.bethe_reduced(materials)
        self.A = A.astype(LD)
        self.B = B.astype(LD)

    def f(self, E):
        E = np.asarray(E, dtype=LD)
        s = np.zeros_like(E)
        for a, b in zip(self.A, self.B):
            s = s + a * np.log(b * E)
        return -s / E

    def inv_absf(self, E):
        return 1.0 / np.abs(self.f(E))

    def integral(self, a, b, pieces=4):
        """integral_a^b dE/|f| by composite 24-point Gauss-Legendre, extended precision."""
        a, b = LD(a), LD(b)
        if a == b:
            return LD(0)
        edges = a + (b - a) * (np.arange(pieces + 1, dtype=LD) / pieces)
        tot = LD(0)
        for lo, hi in zip(edges[:-1], edges[1:]):
            mid, half = (lo + hi) / 2, (hi - lo) / 2
            tot += half * np.sum(_GLW * self.inv_absf(mid + half * _GLX))
        return tot


def _cheb_nodes(n):
    k = np.arange(n, dtype=LD)
    return np.cos(LD(np.pi) * (2 * k + 1) / (2 * n))[::-1]      # ascending in [-1, 1]


def _fit_monomial(vals, nodes):
    """Interpolating polynomial through (nodes in [-1,1], vals); monomial coefficients, lowest first."""
    n = len(nodes)
    V = np.vander(nodes, n, increasing=True).astype(LD)
    # Gaussian elimination in extended precision (numpy.linalg has no longdouble path)
    M = np.concatenate([V, np.asarray(vals, dtype=LD).reshape(-1, 1)], axis=1)
    for c in range(n):
        p = c + int(np.argmax(np.abs(M[c:, c])))
        if p != c:
            M[[c, p]] = M[[p, c]]
        M[c] = M[c] / M[c, c]
        for r in range(n):
            if r != c:
                M[r] = M[r] - M[r, c] * M[c]
    return M[:, n]


def _key(E, q):
    hi = (np.array([E], dtype=np.float64).view(np.uint64)[0] >> np.uint64(32)).astype(np.int64)
    return int(hi >> (20 - q))


def _cell_edges(key, q):
    """[lo, hi) of the T1 cell with the given key."""
    lo_bits = np.uint64(key) << np.uint64(52 - q)
    lo = np.array([lo_bits], dtype=np.uint64).view(np.float64)[0]
    hi = np.array([np.uint64(key + 1) << np.uint64(52 - q)], dtype=np.uint64).view(np.float64)[0]
    return float(lo), float(hi)


def build(config: cfgmod.ModelConfig, degree: int = 7, tol_omega: float = 2e-13, tol_u: float = 2e-12) -> RangeTables:
    st = _Stopping(config.materials)
    e_min, e_max = config.e_range
    # ---- sign and domain --------------------------------------------------------------------------
    probe = np.geomspace(max(e_min, 1e-3), e_max, 4001)
    fv = st.f(probe)
    if not (np.all(fv < 0) or np.all(fv > 0)):
        raise ValueError("dE/dx changes sign inside the histogram range; TOF_ODE_RANGE is not applicable")
    sign = -1.0 if fv[0] < 0 else 1.0
    zero_E = 1.0 / float(np.max(st.B)) if len(st.B) == 1 else None    # single material: f(1/B) = 0
    if sign < 0:
        # energies only fall: draws below e_min can never enter the histogram range
        e_lo = e_min
        e_hi = 16384.0
    else:
        # energies only rise: draws above e_max can never enter the histogram range; stay clear of the
        # zero of f (1/|f| has a pole there)
        e_lo = 0.0625
        e_hi = 2.0 ** np.ceil(np.log2(e_max * 1.0001))
        if zero_E is not None and e_hi > 0.5 * zero_E:
            raise ValueError("histogram range too close to the zero of dE/dx for the range tables")
    if not e_hi > e_max:
        raise ValueError("range-table domain does not cover e_max")
    q = T1_Q
    key_lo = _key(e_lo, q)
    key_hi = _key(np.nextafter(e_hi, 0), q)
    n1 = key_hi - key_lo + 1
    # ---- T1: cumulative u at cell edges, then an interpolant per cell ---------------------------------
    nodes = _cheb_nodes(T1_DEGREE + 1)
    t1 = np.zeros((n1, T1_DEGREE + 1), dtype=np.float64)
    cell_lo_u = np.zeros(n1 + 1, dtype=LD)
    # u at the left edge of the first cell, relative to e_min
    first_lo, _ = _cell_edges(key_lo, q)
    cell_lo_u[0] = -st.integral(first_lo, e_min, pieces=8) if first_lo < e_min else st.integral(e_min, first_lo, pieces=8)
    for c in range(n1):
        lo, hi = _cell_edges(key_lo + c, q)
        cell_lo_u[c + 1] = cell_lo_u[c] + st.integral(lo, hi, pieces=2)
    max_err_u = 0.0
    chk = np.array([-0.93, -0.41, 0.07, 0.55, 0.97], dtype=LD)
    for c in range(n1):
        lo, hi = _cell_edges(key_lo + c, q)
        mid, half = (LD(lo) + LD(hi)) / 2, (LD(hi) - LD(lo)) / 2
        vals = np.array([cell_lo_u[c] + st.integral(lo, mid + half * t, pieces=2) for t in nodes], dtype=LD)
        co = _fit_monomial(vals, nodes)
        t1[c] = co.astype(np.float64)
        for t in chk:
            exact = cell_lo_u[c] + st.integral(lo, mid + half * t, pieces=2)
            approx = np.polynomial.polynomial.polyval(np.float64(t), t1[c])
            max_err_u = max(max_err_u, abs(float(exact - LD(approx))))
    if max_err_u > tol_u:
        raise ValueError("T1 fit error %.3g cm exceeds %.3g" % (max_err_u, tol_u))

    def u_of(E):
        """Extended-precision u(E) using the cell-edge cumulative values (host only)."""
        k = _key(float(E), q) - key_lo
        lo, _ = _cell_edges(key_lo + k, q)
        return cell_lo_u[k] + st.integral(lo, E, pieces=2)

    def e_of(u):
        """Inverse of u(E) by bisection + Newton on the host (extended precision)."""
        lo_c, hi_c = 0, n1
        while hi_c - lo_c > 1:                     # cell whose [u_lo, u_hi) holds u
            m = (lo_c + hi_c) // 2
            if cell_lo_u[m] <= u:
                lo_c = m
            else:
                hi_c = m
        a, b = _cell_edges(key_lo + lo_c, q)
        E = LD(a) + (LD(b) - LD(a)) * (LD(u) - cell_lo_u[lo_c]) / (cell_lo_u[lo_c + 1] - cell_lo_u[lo_c])
        for _ in range(6):
            E = E - (u_of(E) - LD(u)) * np.abs(st.f(E))
            E = min(max(E, LD(a)), LD(np.nextafter(b, 0)))
        return E

    # ---- T2 breakpoints: E-bin edges and cross-section knots inside (e_min, e_max), mapped to u ---------
    eb = config.e_bins
    step = (e_max - e_min) / eb
    e_edges = [e_min + k * step for k in range(eb)] + [e_max]          # np.linspace(e_min, e_max, eb+1)
    edges_u = np.array([u_of(E) for E in e_edges], dtype=LD)
    edges_u[0] = LD(0)
    knots = [E for E in cfgmod.DDN_XS_ENERGIES if e_min < E < e_max]
    pts = [(float(u), b, True) for b, u in enumerate(edges_u)]
    for E in knots:
        pts.append((float(u_of(E)), None, False))
    pts.sort(key=lambda p: p[0])
    # drop knots that coincide with an edge
    merged = []
    for p in pts:
        if merged and abs(p[0] - merged[-1][0]) <= 1e-13 * max(1.0, abs(p[0])):
            if p[2]:
                merged[-1] = p
            continue
        merged.append(p)
    xs_c = cfgmod.not_a_knot_cubic(cfgmod.DDN_XS_ENERGIES, cfgmod.DDN_XS_SIGMA0).astype(LD)
    xs_b = cfgmod.DDN_XS_ENERGIES.astype(LD)

    def xs_ld(E):
        E = min(max(LD(E), xs_b[0]), xs_b[-1])
        i = int(min(max(np.searchsorted(cfgmod.DDN_XS_ENERGIES, float(E), side="right") - 1, 0), len(xs_b) - 2))
        x = E - xs_b[i]
        return ((xs_c[i, 0] * x + xs_c[i, 1]) * x + xs_c[i, 2]) * x + xs_c[i, 3]

    nodes2 = _cheb_nodes(degree + 1)
    chk2 = np.array([-0.9, -0.3, 0.2, 0.8], dtype=LD)

    def fit_interval(a, b):
        mid, half = (LD(a) + LD(b)) / 2, (LD(b) - LD(a)) / 2
        vals = np.array([xs_ld(e_of(mid + half * t)) for t in nodes2], dtype=LD)
        co_t = _fit_monomial(vals, nodes2)                            # in t in [-1, 1]
        # re-expand about the left endpoint in (v - a): t = (v - a)/half - 1
        poly = np.polynomial.polynomial.Polynomial(np.array(co_t, dtype=LD))
        shifted = poly(np.polynomial.polynomial.Polynomial(np.array([LD(-1), LD(1) / half], dtype=LD)))
        co = np.zeros(degree + 1, dtype=LD)
        co[:len(shifted.coef)] = shifted.coef
        err = 0.0
        for t in chk2:
            v = mid + half * t
            exact = xs_ld(e_of(v))
            approx = np.polynomial.polynomial.polyval(np.float64(v - LD(a)), co.astype(np.float64))
            err = max(err, abs(float((LD(approx) - exact) / exact)))
        return co.astype(np.float64), err

    breaks, bins, coefs = [], [], []
    max_err_w = 0.0
    cur_bin = -1
    for k in range(len(merged) - 1):
        a, b_idx, is_edge = merged[k]
        if is_edge:
            cur_bin = b_idx
        b = merged[k + 1][0]
        pieces = 1
        while True:
            sub = np.linspace(a, b, pieces + 1)
            fits = [fit_interval(sub[i], sub[i + 1]) for i in range(pieces)]
            worst = max(f[1] for f in fits)
            if worst <= tol_omega or pieces >= 16:
                break
            pieces *= 2
        for i in range(pieces):
            breaks.append(float(sub[i]))
            bins.append(cur_bin)
            coefs.append(fits[i][0])
        max_err_w = max(max_err_w, worst)
    u_max = float(edges_u[-1])
    breaks.append(u_max)
    breaks = np.array(breaks, dtype=np.float64)
    breaks[0] = 0.0
    m = len(bins)
    lut_n = 1024
    lut = np.zeros(lut_n, dtype=np.uint16)
    j = 0
    for cidx in range(lut_n):
        left = u_max * cidx / lut_n
        while j + 1 < m and left >= breaks[j + 1]:
            j += 1
        lut[cidx] = j
    return RangeTables(sign=sign, t1_q=q, t1_key_lo=key_lo, t1_n=n1, t1_coefs=np.ascontiguousarray(t1),
                       e_tab_lo=float(_cell_edges(key_lo, q)[0]), e_tab_hi=float(_cell_edges(key_hi, q)[1]),
                       degree=degree, u_max=u_max, breaks=breaks, bins=np.array(bins, dtype=np.int32),
                       coefs=np.ascontiguousarray(np.array(coefs, dtype=np.float64)), lut=lut, lut_inv=lut_n / u_max,
                       edges_u=edges_u.astype(np.float64), max_err_u=max_err_u, max_err_omega=max_err_w)


_CACHE = {}


def build_cached(config: cfgmod.ModelConfig, degree: int = 7) -> RangeTables:
    """Tables depend only on the stopping medium and the E binning; cache them per process."""
    key = (tuple(config.materials), tuple(config.e_range), config.e_bins, degree)
    if key not in _CACHE:
        _CACHE[key] = build(config, degree)
    return _CACHE[key]


# ---- numpy emulation of the device arithmetic (host-logic tests; not a product path) -------------------
def t1_eval(tab: RangeTables, E0: np.ndarray) -> np.ndarray:
    """u0 = u(E0) exactly as the kernel evaluates it; -inf for E0 <= 0 / below the table, +inf above."""
    E0 = np.asarray(E0, dtype=np.float64)
    out = np.empty_like(E0)
    bits = E0.view(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.int64)
    q = tab.t1_q
    key = hi >> (20 - q)
    idx = key - tab.t1_key_lo
    below = ~(E0 >= tab.e_tab_lo)           # also NaN and negatives
    above = E0 >= tab.e_tab_hi
    if tab.sign > 0:
        # rising energies: tiny positive E0 are clamped to the first cell (documented bound in DESIGN.md)
        clamp = (E0 > 0) & (E0 < tab.e_tab_lo)
    else:
        clamp = np.zeros_like(below)
    idx = np.clip(idx, 0, tab.t1_n - 1)
    mant = ((bits & np.uint64(0x000FFFFFFFFFFFFF)) | np.uint64(0x3FF0000000000000)).view(np.float64)  # [1, 2)
    c_in_oct = (key & ((1 << q) - 1)).astype(np.float64)
    t = (mant - 1.0) * float(1 << (q + 1)) - (2.0 * c_in_oct + 1.0)
    t = np.where(clamp, -1.0, t)
    co = tab.t1_coefs[idx]
    acc = co[:, T1_DEGREE].copy()
    for k in range(T1_DEGREE - 1, -1, -1):
        acc = acc * t + co[:, k]
    out[:] = acc
    out[below & ~clamp] = -np.inf
    out[above] = np.inf
    return out


def emulate_cell_hist(tab: RangeTables, config: cfgmod.ModelConfig, E0: np.ndarray) -> np.ndarray:
    """Weighted (x, E) histogram H[X, E] through the range tables (float64, draw order)."""
    xc = config.x_centers()
    x_start = 0.0 if config.ode_from_zero else xc[0]
    u0 = t1_eval(tab, E0)
    H = np.zeros((config.x_bins, config.e_bins))
    for i in range(config.x_bins):
        v = u0 + tab.sign * (xc[i] - x_start)
        ok = (v >= 0.0) & (v <= tab.u_max)
        vv = v[ok]
        j = np.clip(np.searchsorted(tab.breaks, vv, side="right") - 1, 0, len(tab.bins) - 1)
        dt = vv - tab.breaks[j]
        co = tab.coefs[j]
        w = co[:, tab.degree].copy()
        for k in range(tab.degree - 1, -1, -1):
            w = w * dt + co[:, k]
        np.add.at(H[i], tab.bins[j], w)
    return H
