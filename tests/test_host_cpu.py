"""CPU-side checks: the C ABI library loads and exports what include/tofgpu.h declares, host tables
match the oracle / reference KATs, the range tables reproduce the oracle's integer cell counts, and
the pool adapter behaves like emcee's pool seam.  No GPU compute here."""
import ctypes
import os
import warnings

import numpy as np
import pytest

import mcmctoffitting_b200 as M
from mcmctoffitting_b200 import _lib, config as C, range_tables as R
from oracle import tof_oracle as O
from conftest import parse_floats

warnings.simplefilter("ignore")


def test_library_loads_and_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert names, "no symbols parsed from include/tofgpu.h"
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == sorted(_lib.SIGNATURES)
    assert lib.tof_abi_version() == _lib.ABI_VERSION
    assert lib.tof_sizeof_config() == ctypes.sizeof(_lib.TofConfig)


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(M.TofError) as ei:
        M.TofModel(C.sweep())
    assert "no CPU fallback" in str(ei.value) or "no CUDA device" in str(ei.value)


def test_product_does_not_import_the_oracle():
    import os
    import re
    pkg = os.path.dirname(M.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_config_tables_match_oracle(golden):
    k = golden["kat"]
    for cfg, om in [(C.adv(0), O.adv_model(0)), (C.intermediate(3), O.intermediate_model(3)), (C.sweep(), O.sweep_model())]:
        assert np.array_equal(cfg.x_centers(), om.x_binCenters)
        assert np.array_equal(cfg.e_centers(), om.eD_binCenters)
        en = O.getDDneutronEnergy(om.eD_binCenters)
        assert np.array_equal(cfg.neutron_speed(), O.SPEED_OF_LIGHT * np.sqrt(2 * en / O.M_NEUTRON))
        want_dist = O.CELL_LENGTH - om.x_binCenters + O.ZERO_DEG_LENGTH / 2 + om.standoff
        assert np.array_equal(cfg.neutron_dist()[0], want_dist)
        assert cfg.n_loops == om.n_loops and cfg.prior == om.prior
    s, so = C.simult(), O.SimultModel()
    assert s.standoffs == so.standoffs and s.tof_ranges == so.tof_ranges and s.tof_bins == so.tof_bins
    assert s.prior == so.prior and s.n_loops == so.n_loops
    np.testing.assert_allclose(np.array(C.adv().taps), parse_floats(k["beamTiming_taps"]), rtol=4e-16)
    np.testing.assert_allclose(C.gaussian_timing_taps(2.7), parse_floats(k["gaussianTiming_2.7_4_taps"]), rtol=4e-16)
    zt, zw = C.zero_degree_tables(np.array([2500.0]))
    assert np.array_equal(zt[0], parse_floats(k["zeroDeg_En2500_times"]))
    assert np.array_equal(zw[0], parse_floats(k["zeroDeg_En2500_weights"]))
    assert np.array_equal(C.dd_neutron_energy(parse_floats(k["E"])), parse_floats(k["getDDneutronEnergy"]))


def test_bethe_reduced_matches_reference_dedx(golden):
    k = golden["kat"]
    E = parse_floats(k["E"])
    for key, mat in [("dEdx_I19.2e-3", (1, 2, 8.565e-5, 19.2e-3)), ("dEdx_I19.2", (1, 2, 8.565e-5, 19.2)),
                     ("dEdx_oneBD", (1, 2, 4 * 8.565e-5, 19.2e-3))]:
        A, B = C.bethe_reduced([mat])
        np.testing.assert_allclose(-(A[0] / E) * np.log(B[0] * E), parse_floats(k[key]), rtol=3e-14)


def test_cross_section_spline_matches_reference(golden):
    k = golden["kat"]
    c = C.not_a_knot_cubic(C.DDN_XS_ENERGIES, C.DDN_XS_SIGMA0)
    x = parse_floats(k["xs_dense_in"])
    i = np.clip(np.searchsorted(C.DDN_XS_ENERGIES, x, side="right") - 1, 0, 59)
    t = x - C.DDN_XS_ENERGIES[i]
    y = ((c[i, 0] * t + c[i, 1]) * t + c[i, 2]) * t + c[i, 3]
    np.testing.assert_allclose(y, parse_floats(k["xs_dense_out"]), rtol=1e-13)
    assert y.min() > 0


@pytest.mark.parametrize("excitation", [19.2e-3, 19.2])
def test_range_tables_reproduce_oracle_cell_counts(excitation):
    cfg = C.sweep(mean_excitation=excitation)
    tab = R.build_cached(cfg)
    assert tab.max_err_u <= 2e-12 and tab.max_err_omega <= 2e-13
    assert tab.sign == (-1.0 if excitation < 1 else 1.0)
    om = O.sweep_model(mean_excitation=excitation, ode_scheme="exact")
    z = np.random.RandomState(5).standard_normal(1024)
    xs = O.DDNXS()
    for th in ([1050, .1], [2000, .3], [1200, .45]):
        E0 = th[0] + (th[1] * th[0]) * z
        H = R.emulate_cell_hist(tab, cfg, E0)
        cnt = np.rint(H / np.sum(H * om.eD_binSize * om.x_binSize) * cfg.n_samples).astype(np.int64)
        assert np.array_equal(cnt, om.cell_counts(th, z, xs)), th


def test_range_tables_reject_sign_change():
    # a medium whose dE/dx changes sign inside the histogram range cannot be tabulated
    cfg = C.sweep(materials=((1, 2, 8.565e-5, 19.2e-3 * 60),))   # zero of f at ~1057 keV
    with pytest.raises(ValueError):
        R.build(cfg)


class _FakeModel:
    def __init__(self, cfg):
        self.config = cfg
        self.obs = {}
        self.calls = []

    def set_observables(self, obs, run=0):
        self.obs[run] = np.array(obs)

    def lnprob_batch(self, thetas):
        t = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        self.calls.append(t.shape[0])
        return -t.sum(axis=1)


def test_pool_adapter_batches_and_refuses_foreign_functions():
    fn = M.TofLnProb(_FakeModel(C.sweep()))
    obs = np.arange(2048.0)
    fn.bind_observables(obs)
    pool = M.BatchedPool(fn)
    pos = [np.array([1000.0 + i, 0.1]) for i in range(6)]

    class Wrapper:
        def __init__(self, f, args, kwargs):
            self.f, self.args, self.kwargs = f, args, kwargs

    out = pool.map(Wrapper(fn, [], {"observables": obs}), pos)
    assert out == [-(1000.0 + i + 0.1) for i in range(6)]
    assert fn.model.calls == [6]                  # ONE batched evaluation
    assert pool.map(fn, []) == []
    with pytest.raises(TypeError):
        pool.map(lambda p: 0.0, pos)
    assert pool.is_master() and pool.wait() is None and pool.close() is None
    # scalar call, reference signature, re-binds only when the observables change
    assert fn(pos[0], obs) == -(1000.0 + 0.1)
    n_sets = len(fn.model.obs)
    fn(pos[0], obs + 1)
    assert np.array_equal(fn.model.obs[0], obs + 1) and len(fn.model.obs) == n_sets


def test_simult_signature_geometry_is_checked():
    cfg = C.simult()
    fn = M.TofLnProb(_FakeModel(cfg))
    obs = [np.ones(n) for n in cfg.tof_bins]
    theta = np.arange(9.0)
    assert fn(theta, obs, cfg.standoffs, cfg.tof_ranges, cfg.tof_bins, cfg.n_samples) == -theta.sum()
    with pytest.raises(ValueError):
        fn(theta, obs, cfg.standoffs[::-1], cfg.tof_ranges, cfg.tof_bins)
    with pytest.raises(ValueError):
        fn(theta, obs, cfg.standoffs, cfg.tof_ranges, cfg.tof_bins, 12345)


def test_energy_distribution_shapes_match_numpy_and_the_reference():
    """a18: normal / lognormal-loss / skew-normal transforms on explicit draws."""
    import os
    from scipy.stats import lognorm, skewnorm as sp_skewnorm
    from mcmctoffitting_b200 import shapes
    n = 1000
    z = np.random.RandomState(5).standard_normal(n)
    assert np.array_equal(shapes.normal(1050.0, 0.1 * 1050.0, z), np.random.RandomState(5).normal(1050.0, 0.1 * 1050.0, n))
    want = np.repeat(1878.4, n) - lognorm.rvs(s=0.5, loc=850.0, scale=170.0, size=n, random_state=np.random.RandomState(5))
    assert np.array_equal(shapes.lognormal_loss(1878.4, 0.5, 850.0, 170.0, z), want)
    # the legacy skew-normal: restated from utilities/pdfs.py, checked against the reference itself when it is present
    rs = np.random.RandomState(9)
    z0, z1 = rs.standard_normal(n), rs.standard_normal(n)
    got = shapes.skewnorm_rvs(2.5, 900.0, 75.0, z0, z1)
    x = np.linspace(600, 1300, 50)
    np.testing.assert_allclose(shapes.skewnorm_pdf(x, 900.0, 2.5, 75.0), sp_skewnorm.pdf(x, 2.5, loc=900.0, scale=75.0), rtol=1e-12)
    assert abs(np.mean(got) - sp_skewnorm.mean(2.5, loc=900.0, scale=75.0)) < 4 * 75.0 / np.sqrt(n)
    from oracle import ref_loader
    if ref_loader.available():
        import sys
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
        import utilities.pdfs as ref_pdfs
        np.random.seed(9)
        assert np.array_equal(ref_pdfs.skewnorm().rvs(n, a=2.5, loc=900.0, scale=75.0), got)
        np.testing.assert_allclose(ref_pdfs.skewnorm().pdf(x, loc=900.0, a=2.5, scale=75.0),
                                   shapes.skewnorm_pdf(x, 900.0, 2.5, 75.0), rtol=1e-12)


def test_tof_data_file_round_trip(tmp_path):
    """utilities.readMultiStandoffTOFdata format + the window selection of adv:219-224 / simultFit.py:528-532."""
    from mcmctoffitting_b200 import dataio
    edges = np.arange(100.0, 300.0, 1.0)
    counts = np.random.RandomState(1).poisson(50, (len(edges), 5)).astype(float)
    path = str(tmp_path / "multistandoff.dat")
    dataio.write_multi_standoff_tof(path, edges, counts)
    data = dataio.read_multi_standoff_tof(path, n_runs=5)
    assert data.shape == (200, 6) and np.array_equal(data[:, 0], edges) and np.array_equal(data[:, 1:], counts)
    cfg = C.simult()
    obs = dataio.observables_for(cfg, data)
    assert [len(o) for o in obs] == list(cfg.tof_bins)            # 1-ns bins: window width == bin count
    assert np.array_equal(obs[0], counts[75:125, 0]) and np.array_equal(obs[3], counts[90:160, 3])
    ref_mod = None
    from oracle import ref_loader
    if ref_loader.available():
        ref_mod = ref_loader.load_utilities().utilities
        assert np.array_equal(ref_mod.readMultiStandoffTOFdata(path, nRuns=5), data)


def test_sdef_card_matches_the_reference_writer():
    """ppc.sdef_sia_cumulative against ppcTools.makeSDEF_sia_cumulative (utilities/ppcTools.py:397-422) itself, run as
    an unbound method on a stand-in that carries the three attributes it reads."""
    import sys
    import types
    from unittest import mock
    from oracle import ref_loader
    from mcmctoffitting_b200 import ppc
    rs = np.random.RandomState(4)
    n, X, E = 7, 10, 50
    cells = rs.poisson(30.0, size=(n, X, E)).astype(np.int64)
    e_n = np.linspace(2900.0, 4400.0, E) + rs.uniform(0, 1, E)
    got = ppc.sdef_sia_cumulative(cells, e_n, dist_number=100)
    assert got["si"].startswith("si100 a ") and got["sp"].startswith("sp100 ")
    assert len(got["si"].split()) == E + 2 and len(got["sp"].split()) == E + 1
    assert [int(v) for v in got["sp"].split()[1:]] == list(cells.sum(axis=(0, 1)))
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    stubs = {k: mock.MagicMock(name=k) for k in ("matplotlib", "matplotlib.pyplot", "corner")}
    with mock.patch.dict(sys.modules, stubs):
        import utilities.ppcTools as ref_ppc
    stand_in = types.SimpleNamespace(tofData=[0], eD_bins=E, eN_binCenters=e_n,
                                     neutronSpectra=[[cells[k].astype(float)] for k in range(n)])   # [sample][run][x, E]
    want = ref_ppc.ppcTools.makeSDEF_sia_cumulative(stand_in, 100)
    assert got == {"si": want["si"], "sp": want["sp"]}


@pytest.mark.parametrize("name", ["sweep", "adv", "intermediate", "wide_bins"])
def test_fp32_weight_records_reach_single_precision(name):
    """The FP32 mode's degree-3 weight polynomials (fitted in tof_create at the Chebyshev nodes of every T2 interval
    from the degree-7 ones) restated in numpy: <= 1e-7 on the shipped binnings, and clearly worse on a binning the
    library is expected to refuse (its own threshold is 3e-7)."""
    cfg = {"sweep": lambda: C.sweep(), "adv": lambda: C.adv(0), "intermediate": lambda: C.intermediate(0),
           "wide_bins": lambda: C.sweep(e_bins=12)}[name]()                       # 200-keV bins
    tab = R.build_cached(cfg)
    br, co = np.asarray(tab.breaks), np.asarray(tab.coefs)
    worst = 0.0
    k = np.arange(4)
    for j in range(len(br) - 1):
        w = br[j + 1] - br[j]
        nodes = 0.5 * (1.0 - np.cos((2 * k + 1) * np.pi / 8.0))                   # in s = dt / w
        p7 = np.polynomial.polynomial.polyval(nodes * w, co[j])
        b = np.linalg.solve(np.vander(nodes, 4, increasing=True), p7)
        c32 = (b / w ** k).astype(np.float32).astype(np.float64)
        d = np.linspace(0.0, w, 33)
        ref = np.polynomial.polynomial.polyval(d, co[j])
        worst = max(worst, float(np.max(np.abs(np.polynomial.polynomial.polyval(d, c32) / ref - 1.0))))
    if name == "wide_bins":
        assert worst > 3e-7, worst
    else:
        assert worst <= 1e-7, worst


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference on the host cores: one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("walker lnprob evals/sec") and line["value"] > 0 and line["dtype"] == "f64"
    assert "262144 walkers" in line["config"]["workload"] and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    # "reference": the reference's own function bodies through oracle/ref_loader.py (where /root/reference exists);
    # "port": the numpy oracle (the GPU box)
    from oracle import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "evaluations" in cb["sample"]
    # without the reference tree the same command times the port
    env = dict(os.environ, TOF_REFERENCE_ROOT="/nonexistent")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0, out.stderr[-500:]
    assert json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 do no work under torchrun
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         stdout=subprocess.PIPE, text=True, timeout=120, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_self_check_leg_counts_deviations(tmp_path):
    """bench.py --oracle-check (the leg behind `parity_sample`): oracle values pass, a perturbed one and a finite / -inf
    flip are counted."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    om, z, obs, thetas = bench.workload(O)
    xs = O.DDNXS()
    cand = np.array([om.lnprob(t, obs, z, xs) for t in thetas[:96]])     # most initial positions are -inf with these observables
    keep = np.concatenate([np.flatnonzero(np.isfinite(cand))[:3], np.flatnonzero(~np.isfinite(cand))[:2]])
    th, got = thetas[keep], cand[keep]
    assert np.isfinite(got).sum() >= 2
    fin = np.flatnonzero(np.isfinite(got))
    bad = got.copy()
    bad[fin[0]] *= 1 + 1e-6                      # outside 1e-9
    bad[fin[1]] = -np.inf                        # a flip
    out = {}
    for name, arr in (("good", got), ("bad", bad)):
        path = str(tmp_path / (name + ".npz"))
        np.savez(path, thetas=th, got=arr)
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--oracle-check", path], stdout=subprocess.PIPE,
                           text=True, timeout=300, cwd=root)
        assert r.returncode == 0
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["good"]["n_outside_1e-9"] == 0 and out["good"]["n_flips"] == 0 and out["good"]["max_rel"] <= 1e-12
    assert out["bad"]["n_outside_1e-9"] == 1 and out["bad"]["n_flips"] == 1


def test_gpu_arm_refuses_to_run_without_a_device():
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=root)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_draw_lookup_never_starts_beyond_the_answer():
    """The range kernel's draw-range search (adv_range.cuh, do_cell) starts from a per-tile lookup cell taken a hair low
    (tu_bias cells) and only walks forward.  That is sound only if the lookup never starts beyond the first draw with
    fl(u0 + delta) >= left.  The device arithmetic restated in numpy (FP64, the lookup built by the same scatter rule),
    on sorted tiles of many shapes, with thresholds placed exactly ON draw values as the adversarial case."""
    n_ulut, bias = 1024, 1.0 / 65536.0
    rs = np.random.RandomState(123)
    worst_walk, total_walk, n_cases = 0, 0, 0
    for trial in range(300):
        nt = int(rs.choice([37, 256, 1000, 1024]))
        centre = rs.uniform(0.5, 60.0)
        width = 10.0 ** rs.uniform(-2.5, 1.0)                     # tile widths from 3e-3 cm to 10 cm
        kind = trial % 3
        if kind == 0:
            u0 = centre + width * rs.standard_normal(nt)
        elif kind == 1:
            u0 = centre + width * rs.standard_t(2.5, nt)          # heavy tails: long empty stretches of lookup cells
        else:
            u0 = centre + width * np.round(rs.standard_normal(nt), 1)   # many exact duplicates
        u0 = np.sort(u0)
        tu_min, tu_max = u0[0], u0[-1]
        tu_inv = n_ulut / (tu_max - tu_min) if tu_max - tu_min > 1e-6 * n_ulut else 0.0
        cell = np.minimum(((u0 - tu_min) * tu_inv).astype(np.int64), n_ulut - 1)
        ulut = np.searchsorted(cell, np.arange(n_ulut), side="left")   # first draw whose cell is >= c (the scatter)
        for delta in rs.uniform(-3.0, 3.0, 4):
            v = u0 + delta                                         # fl(u0 + delta), as the kernel compares it
            lefts = np.concatenate([v[rs.randint(0, nt, 40)],      # thresholds ON sample values
                                    np.nextafter(v[rs.randint(0, nt, 20)], np.inf),
                                    rs.uniform(v[0] - 1.0, v[-1] + 1.0, 40)])
            for left in lefts:
                true_lb = int(np.searchsorted(v, left, side="left"))          # first draw with v >= left
                x = (left - delta) - tu_min
                c = int(np.float64(np.longdouble(x) * np.longdouble(tu_inv) - np.longdouble(bias)))   # fma, truncated
                c = min(max(c, 0), n_ulut - 1)
                start = int(ulut[c]) if c < len(ulut) else nt
                assert start <= true_lb, (trial, nt, width, delta, left, start, true_lb)
                worst_walk = max(worst_walk, true_lb - start)
                total_walk += true_lb - start
                n_cases += 1
    assert n_cases == 300 * 4 * 100
    assert total_walk / n_cases < 3.0                              # the walk stays short on average


def test_div_by_recip_restated_is_correctly_rounded():
    """tof_device.cuh div_by_recip: a*y with y = RN(1/b), then two FMA corrections, equals IEEE a/b.  Restated with
    exact rational arithmetic (Fraction -> float conversion rounds to nearest), including mantissas next to 1 and 2."""
    from fractions import Fraction as F
    import struct
    rs = np.random.RandomState(5)

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))

    def div_by_recip(a, b, y):
        q = a * y
        q = fma(fma(-b, q, a), y, q)
        return fma(fma(-b, q, a), y, q)

    cases = [(rs.uniform(0, 3), rs.uniform(0.1, 30)) for _ in range(4000)]
    cases += [(10.0 ** rs.uniform(-3, 5), 10.0 ** rs.uniform(-3, 7)) for _ in range(4000)]
    for _ in range(4000):
        a = struct.unpack("d", struct.pack("Q", (1023 << 52) | int(rs.randint(0, 256))))[0] * float(rs.choice([1, 3, 5]))
        b = struct.unpack("d", struct.pack("Q", (1023 << 52) | ((1 << 52) - 1 - int(rs.randint(0, 256)))))[0]
        cases.append((a, b))
    for a, b in cases:
        assert div_by_recip(a, b, 1.0 / b) == a / b, (a, b)


def test_shared_memory_layouts_stay_aligned(tmp_path):
    """range_layout / zrank_layout (host-computed byte offsets of the kernels' dynamic shared memory): every region keeps
    the alignment its accesses need for any cell shape -- 16 bytes where records are read with LDS.128, 8 for doubles.
    Regression: an odd x_bins used to leave the interval breaks 4-byte aligned ("misaligned address" on the GPU).
    The two functions are host code in the kernel headers: compiled into a small host program with nvcc and run here."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    csrc = os.path.join(os.path.dirname(M.__file__), "csrc")
    src = tmp_path / "layouts.cu"
    src.write_text(r'''
#include "adv_zrank.cuh"
#include <cstdio>
using namespace tof;
static int bad = 0;
static void need(unsigned off, unsigned a, const char *what, int X, int E, int T, int hcap) {
    if (off % a) { if (bad++ < 10) std::printf("%s = %u not %u-aligned (X=%d E=%d T=%d hcap=%d)\n", what, off, a, X, E, T, hcap); }
}
int main() {
    const int Es[] = {50, 120, 240, 241}, Ts[] = {50, 64, 2048, 2049, 4001};
    long n = 0;
    for (int X = 1; X <= 300; ++X) for (int E : Es) for (int T : Ts) for (int hk = 0; hk < 3; ++hk) {
        const int hcap = hk == 0 ? X * 8 : (hk == 1 ? X * E : X * 8 + 2 * hk + 1);
        for (int taps : {7, 16}) {
            const RangeLayout z = zrank_layout(X, E, T, hcap, E < 128 ? E : 128, 7, taps, E);
            need(z.pa, 16, "zrank pa", X, E, T, hcap);  need(z.rec, 16, "zrank rec", X, E, T, hcap);
            need(z.svd, 8, "zrank svd", X, E, T, hcap);  need(z.ulut, 8, "zrank ulut", X, E, T, hcap);
            need(z.staps, 8, "zrank staps", X, E, T, hcap);  need(z.scratch, 8, "zrank scratch", X, E, T, hcap);
            need(z.sdelta, 8, "zrank sdelta", X, E, T, hcap);  need(z.srow, 4, "zrank srow", X, E, T, hcap);
            need(z.hlo, 4, "zrank hlo", X, E, T, hcap);  need(z.sbrk, 8, "zrank sbrk", X, E, T, hcap);
            if (z.pa < (unsigned)hcap * 8u + ZR_MIRROR || z.total < z.sbrk + (unsigned)E * 8u) { ++bad; std::printf("zrank extents X=%d\n", X); }
            const RangeLayout r = range_layout(X, E, T, hcap, E < 128 ? E : 128, 7, taps, 1000 + X, E);
            need(r.pa, 16, "range pa", X, E, T, hcap);  need(r.rec, 16, "range rec", X, E, T, hcap);
            need(r.svd, 8, "range svd", X, E, T, hcap);  need(r.staps, 8, "range staps", X, E, T, hcap);
            need(r.scratch, 8, "range scratch", X, E, T, hcap);  need(r.sdelta, 8, "range sdelta", X, E, T, hcap);
            need(r.lut, 2, "range lut", X, E, T, hcap);  need(r.ulut, 4, "range ulut", X, E, T, hcap);
            need(r.srow, 4, "range srow", X, E, T, hcap);  need(r.hlo, 4, "range hlo", X, E, T, hcap);
            need(r.sbrk, 8, "range sbrk", X, E, T, hcap);  need(r.sbin, 2, "range sbin", X, E, T, hcap);
            if (r.total < r.sbin + (unsigned)E * 2u) { ++bad; std::printf("range extents X=%d\n", X); }
            n += 2;
        }
    }
    std::printf("%ld layouts, %d misaligned\n", n, bad);
    return bad ? 1 : 0;
}
''')
    exe = tmp_path / "layouts"
    cmd = [nvcc, "-std=c++17", "-O0", "-gencode", "arch=compute_100a,code=sm_100a", "-I", csrc, "-o", str(exe), str(src)]
    comp = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert comp.returncode == 0, comp.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout[-2000:]
    assert "0 misaligned" in run.stdout
