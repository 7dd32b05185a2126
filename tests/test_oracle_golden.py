"""Pin the CPU oracle to values produced by the reference's own functions
(tests/golden/reference_golden.json, made by oracle/make_golden.py)."""
import math
import warnings

import numpy as np
import pytest

from oracle import tof_oracle as O
from conftest import parse_floats, unsparse

warnings.simplefilter("ignore")


def rel(a, b):
    if a == b:
        return 0.0
    if not (math.isfinite(a) and math.isfinite(b)):
        return float("inf")
    return abs(a - b) / max(abs(a), abs(b))


def test_kat_kinematics_and_stopping(golden):
    k = golden["kat"]
    E = parse_floats(k["E"])
    En = O.getDDneutronEnergy(E)
    assert np.array_equal(En, parse_floats(k["getDDneutronEnergy"]))
    assert np.array_equal(O.getTOF(O.M_NEUTRON, En, 516.625), parse_floats(k["getTOF_neutron_516.625"]))
    assert np.array_equal(O.getTOF(O.M_DEUTERON, (1500 + E) / 2, 1.43), parse_floats(k["getTOF_deuteron_1.43"]))
    for key, mat in [("dEdx_I19.2e-3", (1, 2, 8.565e-5, 19.2e-3)), ("dEdx_I19.2", (1, 2, 8.565e-5, 19.2)),
                     ("dEdx_oneBD", (1, 2, 4 * 8.565e-5, 19.2e-3))]:
        sb = O.SimpleBethe([mat])
        assert np.array_equal(sb.dEdx(E), parse_floats(k[key])), key
        (A, B), = sb.reduced()
        np.testing.assert_allclose(-(A / E) * np.log(B * E), parse_floats(k[key]), rtol=2e-14)
    assert O.STANDOFF_MID == k["standoffs"]["mid"] and O.STANDOFF_CLOSE == k["standoffs"]["close"]
    assert O.STANDOFF_FAR == k["standoffs"]["far"] and O.STANDOFF_TUNL == k["standoffs"]["tunl"]


def test_kat_xs_and_timing(golden):
    k = golden["kat"]
    xs = O.DDNXS()
    assert np.array_equal(xs.evaluate(parse_floats(k["xs_in"])), parse_floats(k["xs_out"]))
    assert np.array_equal(xs.evaluate(parse_floats(k["xs_dense_in"])), parse_floats(k["xs_dense_out"]))
    assert np.array_equal(O.beam_timing_taps(), parse_floats(k["beamTiming_taps"]))
    assert np.array_equal(O.gaussian_timing_taps(2.7), parse_floats(k["gaussianTiming_2.7_4_taps"]))
    zt, zw = O.ZeroDegreeTimingSpread().getTimesAndWeights(2500.0)
    assert np.array_equal(zt, parse_floats(k["zeroDeg_En2500_times"]))
    assert np.array_equal(zw, parse_floats(k["zeroDeg_En2500_weights"]))


def test_simple_model_goldens(golden, pf):
    g = golden["simple"]
    obs = np.array(g["obs"], dtype=np.float64)
    m = O.SimpleModel()
    for c in g["cases"]:
        if c["nDraws"] > 100000:
            continue  # the 1e6-draw case runs in test_simple_model_full_size
        rs = np.random.RandomState(c["seed"])
        u = rs.random_sample(c["nDraws"])
        z = rs.standard_normal(c["nDraws"])
        got = m.lnlike(c["theta"], obs, u, z)
        want = pf(c["value"])
        assert rel(got, want) <= 1e-13 or (math.isnan(got) and math.isnan(want)), (c, got)


def test_simple_model_full_size(golden, pf):
    g = golden["simple"]
    obs = np.array(g["obs"], dtype=np.float64)
    c = [c for c in g["cases"] if c["nDraws"] == 1000000][0]
    rs = np.random.RandomState(c["seed"])
    u = rs.random_sample(c["nDraws"])
    z = rs.standard_normal(c["nDraws"])
    assert rel(O.SimpleModel().lnprob(c["theta"], obs, u, z), pf(c["value"])) <= 1e-13


@pytest.mark.parametrize("key", ["adv_as_written", "adv_physical"])
def test_adv_model_goldens(golden, pf, key):
    g = golden[key]
    obs = parse_floats(g["obs"])
    xs = O.DDNXS()
    exact = 0
    cases = [c for c in g["cases"] if c["nDraws"] <= 4096]
    for c in cases:
        nd = c["nDraws"]
        m = O.adv_model(0, mean_excitation=g["mean_excitation"], n_ev_per_loop=g["n_ev_per_loop"], n_samples=nd)
        z = np.random.RandomState(c["seed"]).standard_normal(m.n_loops * m.n_ev_per_loop)
        got = m.lnlike(c["theta"], obs, z, xs)
        want = pf(c["value"])
        # the reference's LSODA runs at rtol~1.5e-8, the oracle's RK4 is accurate to ~1e-12: a draw
        # within that distance of an E-bin edge can flip one integer cell count (~1e-5 relative).
        assert rel(got, want) <= 1e-4, (c, got)
        exact += rel(got, want) <= 1e-12
    assert exact >= len(cases) - 1
    for s in g["spectra"]:
        m = O.adv_model(0, mean_excitation=g["mean_excitation"], n_ev_per_loop=1024, n_samples=1024)
        z = np.random.RandomState(s["seed"]).standard_normal(1024)
        assert np.array_equal(m.raw_tof(s["theta"], z, xs, density=False), parse_floats(s["counts"]))
        np.testing.assert_allclose(m.raw_tof(s["theta"], z, xs, density=True), parse_floats(s["pdf"]), rtol=1e-14)


@pytest.mark.parametrize("key", ["intermediate_as_written", "intermediate_physical"])
def test_intermediate_model_goldens(golden, pf, key):
    """BASELINE config 2 through tests/intermediateTOFmodel.py's own functions (-run 3: far standoff, 70 TOF bins;
    150 E-bins on 200-1700 keV, rho = 8.37e-5): log-likelihoods and integer TOF spectra."""
    g = golden[key]
    obs = parse_floats(g["obs"])
    xs = O.DDNXS()
    exact = 0
    for c in g["cases"]:
        nd = c["nDraws"]
        m = O.intermediate_model(g["run"], mean_excitation=g["mean_excitation"], n_ev_per_loop=g["n_ev_per_loop"], n_samples=nd)
        assert (m.tof_bins, [m.tof_min, m.tof_max]) == (g["tof_bins"], g["tof_range"])
        z = np.random.RandomState(c["seed"]).standard_normal(m.n_loops * m.n_ev_per_loop)
        got = m.lnlike(c["theta"], obs, z, xs)
        want = pf(c["value"])
        assert rel(got, want) <= 1e-4, (c["theta"], got, want)      # one LSODA-tolerance count flip at most
        same = rel(got, want) <= 1e-12
        exact += same
        if same:
            assert np.array_equal(m.raw_tof(c["theta"], z, xs, density=False), parse_floats(c["counts"]))
    assert exact >= len(g["cases"]) - 1
    m = O.intermediate_model(g["run"], mean_excitation=g["mean_excitation"], n_ev_per_loop=1000, n_samples=1000)
    assert m.lnprob([700.0, .1], obs, np.zeros(1000), xs) == pf(g["lnprob_outside_prior"]) == -np.inf


@pytest.mark.slow
@pytest.mark.parametrize("key", ["adv_as_written", "adv_physical"])
def test_adv_model_default_ndraws(golden, pf, key):
    """lnprob at the script's default nDraws=1e5 (97 loops x 1024): single-count flips allowed."""
    g = golden[key]
    obs = parse_floats(g["obs"])
    c = [c for c in g["cases"] if c["nDraws"] == 100000][0]
    m = O.adv_model(0, mean_excitation=g["mean_excitation"], n_ev_per_loop=1024, n_samples=100000)
    z = np.random.RandomState(c["seed"]).standard_normal(97 * 1024)
    got = m.lnprob(c["theta"], obs, z)
    assert rel(got, pf(c["value"])) <= 5e-6, got


def test_sweep_shape_goldens(golden, pf):
    g = golden["sweep"]
    m = O.sweep_model()
    obs = np.zeros(2048)
    obs[g["obs_nonzero_idx"]] = parse_floats(g["obs_nonzero_val"])
    z = np.random.RandomState(g["draw_seed"]).standard_normal(1024)
    xs = O.DDNXS()
    n_exact = 0
    for th, want in zip(g["thetas"], g["lnlike"]):
        got = m.lnlike(th, obs, z, xs)
        want = pf(want)
        if got == want or rel(got, want) <= 1e-12:
            n_exact += 1
        else:
            assert math.isfinite(got) == math.isfinite(want)
    assert n_exact >= len(g["thetas"]) - 1
    pdf0 = m.model_pdf(g["thetas"][0], z, xs)
    want0 = np.zeros(2048)
    want0[g["pdf0_nonzero_idx"]] = parse_floats(g["pdf0_nonzero_val"])
    np.testing.assert_allclose(pdf0, want0, rtol=1e-13, atol=0)


def test_sweep_finite_goldens(golden2, pf):
    """Benchmark shape with observables most walkers can explain (23 of 24 finite in the reference itself)."""
    g = golden2["sweep_finite"]
    m = O.sweep_model()
    obs = np.zeros(2048)
    obs[g["obs_nonzero_idx"]] = parse_floats(g["obs_nonzero_val"])
    z = np.random.RandomState(g["draw_seed"]).standard_normal(1024)
    xs = O.DDNXS()
    want = [pf(v) for v in g["lnlike"]]
    assert sum(math.isfinite(v) for v in want) >= 16
    n_ok = 0
    for th, w in zip(g["thetas"], want):
        got = m.lnlike(th, obs, z, xs)
        assert math.isfinite(got) == math.isfinite(w)
        n_ok += int(got == w or rel(got, w) <= 1e-12)
    assert n_ok >= len(want) - 1           # at most one LSODA-tolerance count flip among 24


def test_simult_goldens_small(golden, pf):
    g = golden["simult"]
    for c in g["cases"]:
        if c["n_draws"] > 10000:
            continue
        m = O.SimultModel(n_ev_per_loop=c["n_ev_per_loop"], n_samples=c["n_draws"])
        obs = [parse_floats(o) for o in c["obs"]]
        # observables themselves: regenerate through the oracle with the obs seed
        draws = O.GlobalStateDraws(np.random.RandomState(c["seed_obs"]))
        th = g["theta"]
        for r in range(5):
            ev = m.model(th[:4] + [th[4 + r]], r, draws)
            assert np.abs(np.rint(ev) - obs[r]).max() <= 1.0, r
        draws = O.GlobalStateDraws(np.random.RandomState(c["seed_eval"]))
        got = m.lnprob(th, obs, draws)
        assert rel(got, pf(c["lnprob"])) <= 1e-4, (got, c["lnprob"])


@pytest.mark.slow
def test_simult_golden_full(golden, pf):
    g = golden["simult"]
    c = [c for c in g["cases"] if c["n_draws"] == 200000][0]
    m = O.SimultModel()
    obs = [parse_floats(o) for o in c["obs"]]
    draws = O.GlobalStateDraws(np.random.RandomState(c["seed_eval"]))
    got = m.lnprob(g["theta"], obs, draws)
    assert rel(got, pf(c["lnprob"])) <= 1e-4, (got, c["lnprob"])


def test_sweep_counts_goldens(golden):
    g = golden["sweep"]
    m = O.sweep_model()
    z = np.random.RandomState(g["draw_seed"]).standard_normal(1024)
    xs = O.DDNXS()
    n_bad = 0
    for th, c in zip(g["thetas"], g["counts"]):
        want = np.zeros(2048)
        want[c["idx"]] = c["val"]
        n_bad += int(not np.array_equal(m.raw_tof(th, z, xs, density=False), want))
    assert n_bad <= 1, n_bad


def test_exact_stopping_agrees_with_rk4_and_scipy():
    """The closed-form (Ei) stopping solution against fine-step RK4 and scipy's LSODA at tight tolerance."""
    from scipy.integrate import odeint
    for exc in (19.2e-3, 19.2):
        sb = O.SimpleBethe([(1, 2, 8.565e-5, exc)])
        (A, B), = sb.reduced()
        xc = O.adv_model().x_binCenters
        E0 = np.array([250.0, 600.0, 1050.0, 1800.0, 2600.0, -5.0])
        ex = O.exact_stop(E0, xc, A, B, None)
        # the 250 keV deuteron stops to ~20 keV inside the cell, where RK4 needs many sub-steps
        rk = O.rk4_stop(E0, xc, sb.dEdx, None, 64)
        np.testing.assert_allclose(ex[:, 1:5], rk[:, 1:5], rtol=2e-13)
        np.testing.assert_allclose(ex[:, 0], rk[:, 0], rtol=1e-10)
        ls = odeint(sb.dEdx, E0[:5], xc, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(ex[:, :5], ls, rtol=2e-9)
        assert np.all(np.isnan(ex[:, 5]))
        ex0 = O.exact_stop(E0[:5], xc[:10], A, B, 0.0)
        rk0 = O.rk4_stop(E0[:5], xc[:10], sb.dEdx, 0.0, 64)
        np.testing.assert_allclose(ex0, rk0, rtol=1e-11)


def test_poisson_restatement_matches_numpy_randomstate():
    """numpy's legacy Poisson sampler restated on an explicit uniform stream == RandomState.poisson."""
    for lam in (0.0, 0.3, 5.0, 9.99, 10.0, 12.0, 57.3, 999.0):
        want = np.random.RandomState(42).poisson(lam, 300)
        st = O.UniformStream(np.random.RandomState(42).random_sample(30000))
        got = O.poisson_from_uniforms(lam, 300, st)
        assert np.array_equal(got, want), lam


def test_onebd_goldens(golden, pf):
    """tests/csi_oneBD.py: spline stopping table and seed-pinned lnprob, reproduced bit for bit."""
    g = golden["onebd"]
    tab = np.array([parse_floats(r) for r in g["stop_table"]])
    assert np.array_equal(O.OneBDModel().stop_table(), tab)
    xs = O.DDNXS()
    for c in g["cases"]:
        m = O.OneBDModel(n_ev_per_loop=c["n_ev_per_loop"], n_samples=c["n_samples"])
        obs = [parse_floats(o) for o in c["obs"]]
        rs = np.random.RandomState(c["seed_eval"])
        total = []
        for r in range(3):   # the reference consumes, per run: n_loops*n_ev normals, then poisson(bg, T)
            z_all = rs.standard_normal(m.n_loops * m.n_ev_per_loop).reshape(m.n_loops, m.n_ev_per_loop)
            p = m.run_params(c["theta"], r)
            ev, _ = m.model(p, r, z_all[-1], rs.poisson(p[4], m.tof_bins[r]), xs, tab)
            total.append(m.bin_loglike(ev, obs[r]))
        assert rel(float(np.sum(total)), pf(c["lnprob"])) <= 1e-13


def test_ppc_goldens_from_the_reference_class(golden):
    """utilities/ppcTools.py's own class (20 x 100 grid, dopri5): its TOF spectrum, the neutron spectra per x
    (eN_atEachX = rows of the integer cell counts) and the unweighted deuteron spectra of the last loop (eD_atEachX)."""
    g = golden["ppc"]
    om = O.SimultModel(x_bins=g["x_bins"], eD_bins=g["e_bins"], n_samples=g["n_samples"], n_ev_per_loop=g["n_ev_per_loop"])
    for c in g["cases"]:
        draws = lambda: O.GlobalStateDraws(np.random.RandomState(c["seed"]))      # noqa: E731
        counts, _ = om.cell_counts(c["params"], c["run"], draws())
        assert np.array_equal(counts, np.array(c["eN_atEachX"]))
        assert np.array_equal(om.deuteron_counts(c["params"], c["run"], draws()), np.array(c["eD_atEachX"]))
        np.testing.assert_allclose(om.model(c["params"], c["run"], draws(), None, True), parse_floats(c["tof"]), rtol=1e-13)
    from mcmctoffitting_b200 import ppc
    card = ppc.sdef_sia_cumulative(np.array([g["cases"][0]["eN_atEachX"]]), O.getDDneutronEnergy(om.eD_binCenters))
    assert card == g["sdef_case0"]


def test_ppc_onebd_goldens_from_the_reference_class(golden_ppc_onebd):
    """utilities/ppcTools_oneBD.py:185-268 through its own class: TOF spectrum (10 zero-degree sub-times, tau = 4
    transit taps, Gaussian timing, Poisson background), eN_atEachX (integer cell counts) and eD_atEachX (unweighted
    last-loop histogram), bit for bit; and the SDEF card of 406-431."""
    g = golden_ppc_onebd
    tab = np.array([parse_floats(r) for r in g["stop_table"]])
    assert np.array_equal(O.OneBDPPCModel().stop_table(), tab)
    assert np.array_equal(O.OneBDPPCModel.zero_deg_taps(), parse_floats(g["transit_taps"]))
    xs = O.DDNXS()
    for c in g["cases"]:
        m = O.OneBDPPCModel(n_ev_per_loop=c["n_ev_per_loop"], n_samples=c["n_samples"])
        assert (m.x_bins, m.eD_bins) == (g["x_bins"], g["e_bins"])
        rs = np.random.RandomState(c["seed"])    # the reference consumes n_loops*n_ev normals, then poisson(bg, T)
        z_all = rs.standard_normal(m.n_loops * m.n_ev_per_loop).reshape(m.n_loops, m.n_ev_per_loop)
        bg = rs.poisson(c["params"][4], m.tof_bins[c["run"]])
        tof, counts = m.model(c["params"], c["run"], z_all[-1], bg, xs, tab)
        assert np.array_equal(counts, unsparse(c["eN_atEachX"]))
        assert np.array_equal(m.deuteron_counts(c["params"], z_all[-1], tab), unsparse(c["eD_atEachX"]))
        assert np.array_equal(tof, parse_floats(c["tof"]))
    from mcmctoffitting_b200 import ppc
    m = O.OneBDPPCModel()
    card = ppc.sdef_sia_cumulative(np.array([unsparse(g["cases"][0]["eN_atEachX"])]), O.getDDneutronEnergy(m.eD_binCenters),
                                   count_format="%.3e")
    assert card == g["sdef_case0"]


def test_template_goldens_from_the_reference_script(golden):
    """tests/devShapeTemplates.py:195-268: the script's own templates (uniform initial energies per slice) are the
    adv pipeline with sigma0 = (e1 - e0)/e0 and uniform numbers in place of the normals (mcmctoffitting_b200.templates)."""
    from mcmctoffitting_b200 import templates as T
    g = golden["templates"]
    bounds = parse_floats(g["bounds"])
    thetas = T.template_thetas(bounds)
    xs = O.DDNXS()
    for c in g["cases"]:
        om = O.intermediate_model(c["run"], rho=8.565e-5, mean_excitation=19.2e-3, n_samples=g["n_samples"],
                                  n_ev_per_loop=g["n_ev_per_loop"])
        u = np.random.RandomState(c["seed"]).random_sample(g["n_samples"])          # np.random.uniform per loop, in order
        got = om.model_pdf(list(thetas[c["slice"]]), u, xs)                          # density + applySpreading
        assert np.array_equal(got, parse_floats(c["template"])), (c["run"], c["slice"])
    tpl = [parse_floats(c["template"]) for c in g["cases"][:2]]
    np.testing.assert_allclose(T.build_model_tof([2.0, 3.0, 5.0], tpl), parse_floats(g["buildModelTOF_2_3_5"]), rtol=1e-15)
