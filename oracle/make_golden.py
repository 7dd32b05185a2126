"""Generate tests/golden/reference_golden.json by running the REFERENCE's own functions.

TEST INFRASTRUCTURE.  Runs only where ``/root/reference`` exists (the build container); the
output is committed so that the GPU box (which has no reference tree) can check against it.

    python oracle/make_golden.py            # rewrites tests/golden/reference_golden.json

Recipe: ``np.random.seed(s)`` immediately before every reference call;
for the adv/intermediate scripts ``nEvPerLoop`` and ``data_x`` are patched in the exec'd
namespace so that the draw count is test-sized; the stopping model is the script's own
``simpleBethe`` (I = 19.2 keV as written) or the physical I = 19.2e-3 variant.
The oracle reproduces every value from ``RandomState(s).standard_normal`` streams.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

warnings.simplefilter("ignore")


def f(x):
    """JSON-safe float (repr round-trips float64 exactly; inf/nan as strings)."""
    x = float(x)
    if np.isnan(x):
        return "nan"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return x


def fl(a):
    return [f(v) for v in np.asarray(a, dtype=np.float64).ravel()]


def kats():
    ref = ref_loader.load_utilities()
    uu, ion = ref.utilities, ref.ionStopping
    E = np.array([250., 500., 900., 1500., 2500.])
    En = uu.getDDneutronEnergy(E)
    xs = uu.ddnXSinterpolator()
    xs_in = np.array([15, 20, 25, 95, 125, 900, 1234.5, 2950, 9999, 12000.])
    zd = uu.zeroDegreeTimingSpread()
    zt, zw = zd.getTimesAndWeights(2500.0)
    out = {
        "E": fl(E),
        "getDDneutronEnergy": fl(En),
        "getTOF_neutron_516.625": fl(uu.getTOF(939565.0, En, 516.625)),
        "getTOF_deuteron_1.43": fl(uu.getTOF(1.8756e+06, (1500 + E) / 2, 1.43)),
        "dEdx_I19.2e-3": fl(ion.ionStopping.simpleBethe([1, 2, 8.565e-5, 1, 19.2e-3]).dEdx(E)),
        "dEdx_I19.2": fl(ion.ionStopping.simpleBethe([1, 2, 8.565e-5, 1, 19.2]).dEdx(E)),
        "dEdx_oneBD": fl(ion.ionStopping.simpleBethe([1, 2, 4 * 8.565e-5, 1, 19.2e-3]).dEdx(E)),
        "xs_in": fl(xs_in),
        "xs_out": fl(xs.evaluate(xs_in.copy())),
        "xs_dense_in": fl(np.linspace(20.0, 10000.0, 2001)),
        "xs_dense_out": fl(xs.evaluate(np.linspace(20.0, 10000.0, 2001))),
        "beamTiming_taps": fl(uu._beamTiming_instance.timingDistribution),
        "gaussianTiming_2.7_4_taps": fl(uu.beamTimingShape.gaussianTiming(2.7, 4).timingDistribution),
        "zeroDeg_En2500_times": fl(zt),
        "zeroDeg_En2500_weights": fl(zw),
        "standoffs": {"mid": f(ref.constants.distances.tunlSSA_CsI.standoffMid),
                      "close": f(ref.constants.distances.tunlSSA_CsI.standoffClose),
                      "far": f(ref.constants.distances.tunlSSA_CsI.standoffFar),
                      "tunl": f(ref.constants.distances.tunlSSA_CsI.standoff_TUNLruns)},
    }
    return out


def simple():
    ns = ref_loader.load("simpleTOFmodel")
    np.random.seed(11)
    fd = ns["generateModelData"]([1100, -100, 50], 10000)
    obs = np.histogram(fd[:, 3], 25, (175, 200))[0]
    cases = []
    for seed, theta, nd in [(12, [1111, -110, 40], 1000000), (12, [1111, -110, 40], 1024),
                            (13, [1090, -95, 55], 20000), (14, [1100, -100, 50], 20000),
                            (15, [700, -100, 50], 20000), (16, [1100, -100, 99.5], 5000)]:
        np.random.seed(seed)
        if nd == 1000000:
            val = ns["lnprob"](theta, obs)
            kind = "lnprob"
        else:
            val = ns["lnlike"](theta, obs, nDraws=nd)
            kind = "lnlike"
        cases.append({"seed": seed, "theta": theta, "nDraws": nd, "kind": kind, "value": f(val)})
    return {"obs": [int(v) for v in obs], "cases": cases}


def adv(excitation, label):
    ns = ref_loader.load("advIntermediateTOFmodel")
    ref = ref_loader.load_utilities()
    n_ev = 1024
    ns["nEvPerLoop"] = n_ev
    ns["data_x"] = np.repeat(ns["x_binCenters"], n_ev)
    model = ref.ionStopping.ionStopping.simpleBethe([1, 2, 8.565e-5, 1, excitation])
    ns["stoppingModel"] = model
    standoff = ns["standoff"][0]
    np.random.seed(7)
    raw = ns["generateModelData"]([1050, .1], standoff, ns["ddnXSinstance"], model.dEdx, 1024, True)
    obs = np.rint(ns["beamTiming"].applySpreading(raw) * 5e4)
    cases = []
    for seed, theta, nd in [(7, [1050, .10], 1024), (8, [1060, .11], 1024), (9, [1040, .09], 1024),
                            (10, [1055, .12], 2048), (21, [1045.5, .105], 4096),
                            (8, [1060, .11], 100000)]:
        np.random.seed(seed)
        if nd == 100000:
            val = ns["lnprob"](theta, obs)      # default nDraws=1e5 -> 97 loops x 1024
            kind = "lnprob"
        else:
            val = ns["lnlike"](theta, obs, nDraws=nd)
            kind = "lnlike"
        cases.append({"seed": seed, "theta": theta, "nDraws": nd, "kind": kind, "value": f(val)})
    # raw spectra (generateModelData getPDF=True/False) for stage-level parity
    spectra = []
    for seed, theta in [(7, [1050, .10]), (31, [1100, .2])]:
        np.random.seed(seed)
        pdf = ns["generateModelData"](theta, standoff, ns["ddnXSinstance"], model.dEdx, 1024, True)
        np.random.seed(seed)
        cnt = ns["generateModelData"](theta, standoff, ns["ddnXSinstance"], model.dEdx, 1024, False)
        spectra.append({"seed": seed, "theta": theta, "pdf": fl(pdf), "counts": fl(cnt)})
    return {"label": label, "mean_excitation": excitation, "n_ev_per_loop": n_ev, "obs": fl(obs),
            "cases": cases, "spectra": spectra}


def intermediate(excitation, label, run=3):
    """tests/intermediateTOFmodel.py (BASELINE config 2) through its own functions: `-run 3` = far standoff, 70 TOF
    bins on (190, 260); E grid 150 bins on 200-1700 keV, rho = 8.37e-5 (intermediate:55-97); test-sized loops."""
    ns = ref_loader.load("intermediateTOFmodel", argv=["-run", str(run)])
    ref = ref_loader.load_utilities()
    n_ev = 1000
    ns["nEvPerLoop"] = n_ev
    ns["data_x"] = np.repeat(ns["x_binCenters"], n_ev)
    model = ref.ionStopping.ionStopping.simpleBethe([1, 2, 8.37e-5, 1, excitation])
    ns["stoppingModel"] = model
    standoff = ns["standoff"][run]
    np.random.seed(17)
    raw = ns["generateModelData"]([900, .12], standoff, ns["ddnXSinstance"], model.dEdx, 4000, True)
    obs = np.rint(ns["beamTiming"].applySpreading(raw) * 2e4)
    cases = []
    for seed, theta, nd in [(18, [900, .12], 4000), (19, [880, .10], 4000), (20, [950, .15], 2000), (21, [1199, .169], 3000)]:
        np.random.seed(seed)
        val = ns["lnlike"](theta, obs, nDraws=nd)
        np.random.seed(seed)
        cnt = ns["generateModelData"](theta, standoff, ns["ddnXSinstance"], model.dEdx, nd, False)
        cases.append({"seed": seed, "theta": theta, "nDraws": nd, "kind": "lnlike", "value": f(val), "counts": fl(cnt)})
    np.random.seed(22)
    outside = ns["lnprob"]([700.0, .1], obs)                       # outside the prior box (intermediate:185-189)
    return {"label": label, "run": run, "mean_excitation": excitation, "n_ev_per_loop": n_ev, "obs": fl(obs),
            "cases": cases, "lnprob_outside_prior": f(outside),
            "tof_bins": int(ns["tof_nBins"]), "tof_range": [f(v) for v in ns["tof_range"]]}


def sweep():
    """adv model at the benchmark shape: 1024 draws, 2048 TOF bins on [128,256)."""
    ns = ref_loader.load("advIntermediateTOFmodel")
    ref = ref_loader.load_utilities()
    ns["nEvPerLoop"] = 1024
    ns["data_x"] = np.repeat(ns["x_binCenters"], 1024)
    ns["tof_nBins"] = 2048
    ns["tof_range"] = (128.0, 256.0)
    model = ref.ionStopping.ionStopping.simpleBethe([1, 2, 8.565e-5, 1, 19.2e-3])
    ns["stoppingModel"] = model
    standoff = ns["standoff"][0]
    np.random.seed(7)
    raw = ns["generateModelData"]([1050, .1], standoff, ns["ddnXSinstance"], model.dEdx, 1024, True)
    obs = np.rint(1e5 * ns["beamTiming"].applySpreading(raw))
    seed = 20260101
    thetas = np.array([1050, 0.10]) + np.array([10, 1e-2]) * np.random.RandomState(1).standard_normal((24, 2))
    vals, pdfs = [], []
    for th in thetas:
        np.random.seed(seed)
        vals.append(f(ns["lnprob"](list(th), obs)) if False else f(ns["lnlike"](list(th), obs, nDraws=1024)))
    counts = []
    for th in thetas:
        np.random.seed(seed)
        c = ns["generateModelData"](list(th), standoff, ns["ddnXSinstance"], model.dEdx, 1024, False)
        nzc = np.nonzero(c)[0]
        counts.append({"idx": [int(i) for i in nzc], "val": [int(v) for v in c[nzc]]})
    np.random.seed(seed)
    pdf0 = ns["beamTiming"].applySpreading(
        ns["generateModelData"](list(thetas[0]), standoff, ns["ddnXSinstance"], model.dEdx, 1024, True))
    nz = np.nonzero(pdf0)[0]
    return {"draw_seed": seed, "obs_nonzero_idx": [int(i) for i in np.nonzero(obs)[0]],
            "obs_nonzero_val": fl(obs[np.nonzero(obs)[0]]), "thetas": [fl(t) for t in thetas],
            "lnlike": vals, "counts": counts, "pdf0_nonzero_idx": [int(i) for i in nz], "pdf0_nonzero_val": fl(pdf0[nz])}


def sweep_finite():
    """The benchmark shape again, with observables most walkers can explain: generated at sigma0 = 0.08 (narrower
    than the walkers' 0.10 +- 0.01), so that 23 of the 24 log-likelihoods are finite (the `sweep` golden above uses the
    SURVEY.md 8d observables, for which 20 of 24 are -inf).  Same thetas and draws as `sweep`: the integer TOF
    spectra are the ones stored there."""
    ns = ref_loader.load("advIntermediateTOFmodel")
    ref = ref_loader.load_utilities()
    ns["nEvPerLoop"] = 1024
    ns["data_x"] = np.repeat(ns["x_binCenters"], 1024)
    ns["tof_nBins"] = 2048
    ns["tof_range"] = (128.0, 256.0)
    model = ref.ionStopping.ionStopping.simpleBethe([1, 2, 8.565e-5, 1, 19.2e-3])
    ns["stoppingModel"] = model
    standoff = ns["standoff"][0]
    np.random.seed(7)
    raw = ns["generateModelData"]([1050, .08], standoff, ns["ddnXSinstance"], model.dEdx, 1024, True)
    obs = np.rint(1e5 * ns["beamTiming"].applySpreading(raw))
    seed = 20260101
    thetas = np.array([1050, 0.10]) + np.array([10, 1e-2]) * np.random.RandomState(1).standard_normal((24, 2))
    vals = []
    for th in thetas:
        np.random.seed(seed)
        vals.append(f(ns["lnprob"](list(th), obs)) if False else f(ns["lnlike"](list(th), obs, nDraws=1024)))
    nz = np.nonzero(obs)[0]
    return {"draw_seed": seed, "obs_theta": [1050.0, 0.08], "obs_seed": 7, "obs_nonzero_idx": [int(i) for i in nz],
            "obs_nonzero_val": fl(obs[nz]), "thetas": [fl(t) for t in thetas], "lnlike": vals,
            "n_finite": int(np.sum(np.isfinite([float(v) if not isinstance(v, str) else float(v) for v in vals])))}


def main_r2():
    """Round-2 additions, kept in their own file so that the round-1 fixture stays byte-identical."""
    gold = {
        "_about": "round-2 additions; values produced by the unmodified reference functions via oracle/ref_loader.py; "
                  "numpy %s" % np.__version__,
        "sweep_finite": sweep_finite(),
    }
    for name, fn in R2_EXTRA.items():
        gold[name] = fn()
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_r2.json")
    with open(path, "w") as fh:
        json.dump(gold, fh, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


R2_EXTRA = {}


def simult(full=True):
    ns = ref_loader.load("simultFit")
    theta = [1878.4, 850, 170, 0.5, 3e4, 2e4, 2e4, 4e4, 4e4]
    out = {"theta": theta, "cases": []}
    for n_ev, n_draws, seed_obs, seed_eval in ([(1000, 4000, 3, 4), (1000, 3500, 5, 6)] +
                                               ([(50000, 200000, 3, 4)] if full else [])):
        ns["nEvPerLoop"] = n_ev
        np.random.seed(seed_obs)
        obs = []
        for r in range(5):
            p = theta[:4] + [theta[4 + r]]
            obs.append(np.rint(ns["generateModelData"](p, ns["standoffs"][r], ns["tof_range"][r],
                                                       ns["tofRunBins"][r], ns["ddnXSinstance"],
                                                       ns["stoppingModel"].dEdx, ns["beamTiming"],
                                                       n_draws, True)))
        np.random.seed(seed_eval)
        val = ns["lnprob"](theta, [o.copy() for o in obs], ns["standoffs"], ns["tof_range"],
                           ns["tofRunBins"], n_draws)
        out["cases"].append({"n_ev_per_loop": n_ev, "n_draws": n_draws, "seed_obs": seed_obs,
                             "seed_eval": seed_eval, "obs": [fl(o) for o in obs], "lnprob": f(val)})
    return out


def main():
    gold = {
        "_about": "values produced by the unmodified reference functions via oracle/ref_loader.py; "
                  "numpy %s" % np.__version__,
        "kat": kats(),
        "simple": simple(),
        "adv_as_written": adv(19.2, "I=19.2 keV as written (adv:94)"),
        "adv_physical": adv(19.2e-3, "I=19.2e-3 keV (physical)"),
        "intermediate_as_written": intermediate(19.2, "intermediateTOFmodel.py -run 3, I=19.2 keV as written (intermediate:94)"),
        "intermediate_physical": intermediate(19.2e-3, "intermediateTOFmodel.py -run 3, I=19.2e-3 keV (physical)"),
        "sweep": sweep(),
        "simult": simult(full="--quick" not in sys.argv),
        "onebd": onebd(),
        "ppc": ppc(),
        "templates": templates(),
    }
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.json")
    with open(path, "w") as fh:
        json.dump(gold, fh, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")




def ppc():
    """utilities/ppcTools.py: the reference's own ppcTools class (its 20 x 100 grid, 1000 tracks per loop, 2 loops)
    fed with a throw-away chain file; generateModelData -> (TOF spectrum, eN_atEachX, eD_atEachX), seeded."""
    import tempfile
    from unittest import mock
    from mcmctoffitting_b200.ensemble import write_chain_step
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    stubs = {k: mock.MagicMock(name=k) for k in ("matplotlib", "matplotlib.pyplot", "corner")}
    shim = ref_loader._LinspaceShim()
    with mock.patch.dict(sys.modules, stubs):
        import utilities.ppcTools as ref_ppc
        path = os.path.join(tempfile.mkdtemp(), "chain.dat")
        rs = np.random.RandomState(0)
        for _ in range(3):
            write_chain_step(path, rs.standard_normal((18, 9)), rs.standard_normal(18))
        np.linspace = shim
        try:
            pt = ref_ppc.ppcTools(path, 1000, nBins_eD=100, nBins_x=20, nRuns=4)
        finally:
            np.linspace = shim._orig
    out = {"x_bins": 20, "e_bins": 100, "n_ev_per_loop": 1000, "n_samples": 2000, "cases": []}
    for seed, run, params in [(5, 0, [1878.4, 850, 170, 0.5, 3e4]), (6, 3, [1825.0, 1000, 300, 1.2, 4e4])]:
        np.random.seed(seed)
        tof, eN, eD = pt.generateModelData(params, pt.standoffs[run], pt.tof_range[run], pt.tofRunBins[run], pt.ddnXSinstance,
                                           pt.stoppingModel.dEdx, pt.beamTiming, 2000, True)
        out["cases"].append({"seed": seed, "run": run, "params": params, "tof": fl(tof),
                             "eN_atEachX": [[int(v) for v in r] for r in eN[1:]],      # without the leading row of zeros
                             "eD_atEachX": [[int(v) for v in r] for r in eD[1:]]})
    cells = np.array([c["eN_atEachX"] for c in out["cases"][:1]], dtype=float)
    pt.tofData, pt.neutronSpectra = [0], [[cells[0]]]
    out["sdef_case0"] = pt.makeSDEF_sia_cumulative(100)
    return out


def templates():
    """tests/devShapeTemplates.py:195-268 through its own functions: TOF templates of three of the 32 energy slices
    at two standoffs (uniform initial energies per slice; 150 x 100 grid, rho = 8.565e-5, I = 19.2e-3), test-sized
    loops, and buildModelTOF on them."""
    ns = ref_loader.load("devShapeTemplates")
    n_ev = 1000
    ns["nEvPerLoop"] = n_ev
    ns["data_x"] = np.repeat(ns["x_binCenters"], n_ev)
    bounds = ns["templateEnergyBounds"]
    out = {"bounds": fl(bounds), "n_ev_per_loop": n_ev, "n_samples": 2000, "cases": []}
    for seed, run, k in [(40, 0, 0), (41, 0, 13), (42, 3, 31), (43, 3, 20)]:
        np.random.seed(seed)
        t = ns["generateModelData"]((bounds[k], bounds[k + 1]), ns["standoffs"][run], ns["tofRunBins"][run], ns["tof_range"][run],
                                    ns["ddnXSinstance"], ns["stoppingModel"].dEdx, 2000, True)
        out["cases"].append({"seed": seed, "run": run, "slice": k, "template": fl(t)})
    tpl = [np.array(c["template"]) for c in out["cases"][:2]]
    out["buildModelTOF_2_3_5"] = fl(ns["buildModelTOF"]([2.0, 3.0, 5.0], tpl))
    return out


def onebd():
    """tests/csi_oneBD.py through the loader (prefix 700 lines, -quitEarly 0)."""
    ns = ref_loader.load("csi_oneBD", argv=["-quitEarly", "0"], prefix_lines=700)
    out = {"stop_table": [fl(r) for r in ns["stoppingApprox"].z], "cases": []}
    theta = [900.0, 170.0, 0.5, 3e4, 2e4, 4e4, 5.0, 12.0, 0.0]
    for n_ev, n_samp, seed_obs, seed_eval in [(500, 2000, 3, 4), (1000, 3000, 5, 6)]:
        ns["nEvPerLoop"] = n_ev
        np.random.seed(seed_obs)
        obs = []
        for r in range(3):
            p = [theta[0], theta[1], theta[2], theta[3 + r], theta[6 + r]]
            obs.append(np.rint(ns["generateModelData"](p, ns["standoffs"][r], ns["tof_range"][r], ns["tofRunBins"][r],
                                                       ns["ddnXSinstance"], ns["stoppingApprox"], ns["beamTiming"],
                                                       n_samp, True)))
        np.random.seed(seed_eval)
        val = ns["lnprob"](theta, [o.copy() for o in obs], ns["standoffs"], ns["tof_range"], ns["tofRunBins"], n_samp)
        out["cases"].append({"n_ev_per_loop": n_ev, "n_samples": n_samp, "seed_obs": seed_obs, "seed_eval": seed_eval,
                             "theta": theta, "obs": [fl(o) for o in obs], "lnprob": f(val)})
    return out


def ppc_onebd():
    """utilities/ppcTools_oneBD.py: the reference's own ppcTools_oneBD class (grid of initialize_oneBD: 20 x 400,
    betheApprox stopping table, tau = 4 transit taps, 10 zero-degree sub-times) fed with a throw-away chain file;
    generateModelData -> (TOF spectrum, eN_atEachX, eD_atEachX), seeded; then its SDEF writer (406-431)."""
    import tempfile
    from unittest import mock
    from mcmctoffitting_b200.ensemble import write_chain_step
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    stubs = {k: mock.MagicMock(name=k) for k in ("matplotlib", "matplotlib.pyplot", "corner")}
    shim = ref_loader._LinspaceShim()
    with mock.patch.dict(sys.modules, stubs):
        np.linspace = shim
        try:
            import utilities.ppcTools_oneBD as P
        finally:
            np.linspace = shim._orig
    path = os.path.join(tempfile.mkdtemp(), "chain.dat")
    rs = np.random.RandomState(0)
    for _ in range(3):
        write_chain_step(path, rs.standard_normal((18, 9)), rs.standard_normal(18))
    pt = P.ppcTools_oneBD(path, 1000)

    def sparse(a):
        a = np.asarray(a)
        nz = np.flatnonzero(a)
        return {"shape": list(a.shape), "idx": [int(i) for i in nz], "val": [int(v) for v in a.ravel()[nz]]}

    out = {"x_bins": int(P.x_bins), "e_bins": int(P.eD_bins), "transit_taps": fl(P.zeroDegSpread_vals),
           "stop_table": [fl(r) for r in P.stoppingApprox.z], "cases": []}
    for n_ev, n_samp, seed, run, params in [(500, 2000, 5, 0, [900.0, 170.0, 0.5, 3e4, 5.0]),
                                            (1000, 3000, 6, 2, [850.0, 300.0, 1.2, 4e4, 12.0]),
                                            (800, 1600, 7, 1, [1000.0, 120.0, 0.3, 2e4, 0.0])]:
        P.nEvPerLoop = n_ev
        np.random.seed(seed)
        tof, eN, eD = pt.generateModelData(params, P.standoffs[run], P.tof_range[run], P.tofRunBins[run], P.ddnXSinstance,
                                           P.stoppingApprox, P.beamTiming, n_samp, True)
        out["cases"].append({"n_ev_per_loop": n_ev, "n_samples": n_samp, "seed": seed, "run": run, "params": params,
                             "tof": fl(tof), "eN_atEachX": sparse(eN[1:]),     # without the leading row of zeros
                             "eD_atEachX": sparse(eD)})
        if len(out["cases"]) == 1:
            pt.tofData, pt.neutronSpectra = [0], [[eN[1:]]]
            card, _, spec = pt.makeSDEF_sia_cumulative(100)
            out["sdef_case0"] = card
    return out


def main_ppc_onebd():
    gold = {"_about": "utilities/ppcTools_oneBD.py run unmodified through its own class; numpy %s" % np.__version__,
            "ppc_onebd": ppc_onebd()}
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_ppc_onebd.json")
    with open(path, "w") as fh:
        json.dump(gold, fh, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__" and "--ppc-onebd" in sys.argv:
    main_ppc_onebd()
    sys.exit(0)

if __name__ == "__main__" and "--r2" in sys.argv:
    main_r2()
    sys.exit(0)

if __name__ == "__main__":
    main()
