"""Ensemble driver on CPU: stretch-move statistics on an analytic target, emcee-shaped surface, chain
file round trip, and the N>1 sharding path over gloo (world size 2) with the numpy backend injected."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from mcmctoffitting_b200.ensemble import EnsembleSampler, write_chain_step, read_chain
from oracle.stretch_oracle import NumpyBackend, philox4x32

MU = np.array([1.0, -2.0, 0.5])
SIG = np.array([0.5, 2.0, 1.0])


def gauss_lnprob(x):
    return -0.5 * np.sum(((x - MU) / SIG) ** 2, axis=1)


def test_philox_known_answer():
    # Random123 known-answer test: philox4x32-10, counter = key = 0
    c = philox4x32(0, np.array([0], dtype=np.uint64), np.uint64(0))
    assert [int(v[0]) for v in c] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    # counter ffffffff x4, key ffffffff x2
    c = philox4x32(0xFFFFFFFFFFFFFFFF, np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64), np.uint64(0xFFFFFFFFFFFFFFFF))
    assert [int(v[0]) for v in c] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_stretch_move_samples_a_gaussian():
    k, dim = 200, 3
    rs = np.random.RandomState(0)
    p0 = MU + 0.1 * rs.standard_normal((k, dim))
    s = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=42)
    pos, lp, state = s.run_mcmc(p0, 600)
    chain = s.chain[:, 200:, :].reshape(-1, dim)
    assert s.chain.shape == (k, 600, dim) and s.lnprobability.shape == (k, 600)
    np.testing.assert_allclose(chain.mean(axis=0), MU, atol=0.08)
    np.testing.assert_allclose(chain.std(axis=0), SIG, rtol=0.08)
    af = s.acceptance_fraction
    assert 0.3 < af.mean() < 0.8
    np.testing.assert_allclose(lp, gauss_lnprob(pos))
    # resume: (pos, lnprob, rstate) handed back like the reference does between burn-in and main chain (adv:337-339)
    s.reset()
    out = s.run_mcmc(pos, 5, rstate0=state, lnprob0=lp)
    assert s.chain.shape == (k, 5, dim) and out[2] == state + 5


def test_constructor_checks_match_emcee():
    with pytest.raises(ValueError):
        EnsembleSampler(7, 2, backend=NumpyBackend(gauss_lnprob))
    with pytest.raises(ValueError):
        EnsembleSampler(4, 3, backend=NumpyBackend(gauss_lnprob))
    s = EnsembleSampler(8, 3, backend=NumpyBackend(lambda x: np.full(len(x), np.nan)))
    with pytest.raises(ValueError):
        s.run_mcmc(np.zeros((8, 3)), 1)


def test_chain_file_round_trip(tmp_path):
    k, dim, steps = 6, 9, 3
    rs = np.random.RandomState(1)
    path = str(tmp_path / "mainchain.dat")
    want_c, want_p = [], []
    for _ in range(steps):
        pos = rs.standard_normal((k, dim)) * 1e3
        lp = rs.standard_normal(k) * 1e5
        write_chain_step(path, pos, lp)             # numpy wraps 9-parameter rows over two lines
        want_c.append(pos)
        want_p.append(lp)
    chain, probs, n_params, n_walkers, n_steps = read_chain(path)
    assert (n_params, n_walkers, n_steps) == (dim, k, steps)
    np.testing.assert_allclose(chain, np.array(want_c), rtol=1e-7)   # numpy prints 8 significant digits
    np.testing.assert_allclose(probs, np.array(want_p), rtol=1e-12)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, dim, steps, p0, out_q, ckdir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=7)
        assert s.world == world and s.n_own == k // 2 // world
        pos, lp, state = s.run_mcmc(p0, steps)
        af = s.acceptance_fraction                         # collective: the owners' counters are gathered
        # sharded checkpoint: every rank calls save (counters gathered, rank 0 writes atomically, barrier), then
        # every rank resumes from the file and continues the chain
        ck = os.path.join(ckdir, "sharded_state.npz")
        s.save_checkpoint(ck, pos, lp)
        s2 = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=7)
        pos_l, lp_l, rstate = s2.load_checkpoint(ck)
        pos2, lp2, _ = s2.run_mcmc(pos_l, 5, rstate0=rstate, lnprob0=lp_l)
        out_q.put((rank, pos, lp, s.naccepted.numpy().copy(), af, pos2, lp2, s2.naccepted.numpy().copy(), s2.iterations))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_sampler_over_gloo_matches_single_process(tmp_path):
    k, dim, steps = 64, 3, 25
    p0 = MU + 0.1 * np.random.RandomState(3).standard_normal((k, dim))
    ref = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=7)
    pos1, lp1, state1 = ref.run_mcmc(p0, steps)
    af1 = ref.acceptance_fraction
    nacc1 = ref.naccepted.numpy().copy()
    pos1b, lp1b, _ = ref.run_mcmc(pos1, 5, rstate0=state1, lnprob0=lp1)      # the uninterrupted chain, 5 steps on
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, k, dim, steps, p0, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, pos2, lp2, nacc, af, pos_r, lp_r, nacc_r, iters_r in results:
        # chains do not depend on the number of ranks: counter-based RNG keyed by the global walker index
        assert np.array_equal(pos2, pos1), rank
        assert np.array_equal(lp2, lp1), rank
        # every rank reports the acceptance counters of the WHOLE ensemble (gathered from the owners), equal to
        # the single-process ones
        assert np.array_equal(nacc, nacc1), rank
        assert np.array_equal(af, af1), rank
        # resume from the sharded checkpoint: same chain and the same counters as the uninterrupted single process
        assert np.array_equal(pos_r, pos1b) and np.array_equal(lp_r, lp1b), rank
        assert np.array_equal(nacc_r, ref.naccepted.numpy()) and iters_r == steps + 5, rank


def test_chain_file_is_readable_by_the_reference_reader(tmp_path):
    """utilities.readChainFromFile (utilities.py:432-500) parses what write_chain_step writes."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    uu = ref_loader.load_utilities().utilities
    k, dim, steps = 8, 9, 4
    rs = np.random.RandomState(2)
    path = str(tmp_path / "burninchain.dat")
    for _ in range(steps):
        write_chain_step(path, rs.standard_normal((k, dim)) * 1e3, rs.standard_normal(k) * 1e4)
    chain_ref, probs_ref, n_params, n_walkers, n_steps = uu.readChainFromFile(path)
    chain, probs, p2, w2, s2 = read_chain(path)
    assert (n_params, n_walkers, n_steps) == (p2, w2, s2) == (dim, k, steps)
    assert np.array_equal(chain_ref, chain) and np.array_equal(probs_ref, probs)


def test_checkpoint_resume_continues_the_same_chain(tmp_path):
    """Binary checkpoint: stop after 7 of 15 steps, resume in a fresh sampler, same chain."""
    k, dim = 32, 3
    p0 = MU + 0.1 * np.random.RandomState(6).standard_normal((k, dim))
    full = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=11)
    pos_full, lp_full, _ = full.run_mcmc(p0, 15)
    first = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=11)
    pos7, lp7, state7 = first.run_mcmc(p0, 7)
    ck = str(tmp_path / "state.npz")
    first.save_checkpoint(ck, pos7, lp7)
    second = EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=11)
    pos, lp, rstate = second.load_checkpoint(ck)
    assert rstate == state7 == 7 and np.array_equal(pos, pos7) and np.array_equal(lp, lp7)
    pos_res, lp_res, state = second.run_mcmc(pos, 8, rstate0=rstate, lnprob0=lp)
    assert state == 15 and np.array_equal(pos_res, pos_full) and np.array_equal(lp_res, lp_full)
    assert np.array_equal(second.chain, full.chain) and np.array_equal(second.lnprobability, full.lnprobability)
    assert np.array_equal(second.naccepted.numpy(), full.naccepted.numpy()) and second.iterations == 15
    with pytest.raises(ValueError):
        EnsembleSampler(k, dim, backend=NumpyBackend(gauss_lnprob), seed=12).load_checkpoint(ck)
    with pytest.raises(ValueError):
        EnsembleSampler(k + 2, dim, backend=NumpyBackend(gauss_lnprob), seed=11).load_checkpoint(ck)


def test_chain_file_reader_agrees_with_the_reference_on_awkward_numbers(tmp_path):
    """Property check of the text chain format over shapes and magnitudes numpy prints differently: scientific
    notation, wrapped rows (ndim up to 16), negative values, -inf log-probabilities.  Where the reference tree is
    present, its own readChainFromFile (utilities.py:432-500) must parse the file identically."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from oracle import ref_loader
    uu = ref_loader.load_utilities().utilities if ref_loader.available() else None
    counter = [0]

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(k=st.integers(2, 7), dim=st.integers(1, 16), steps=st.integers(1, 4), scale=st.sampled_from([1e-6, 1.0, 1e3, 1e7]),
           seed=st.integers(0, 10 ** 6), neg_inf=st.booleans())
    def run(k, dim, steps, scale, seed, neg_inf):
        rs = np.random.RandomState(seed)
        counter[0] += 1
        path = str(tmp_path / ("chain_%d.dat" % counter[0]))
        want_c, want_p = [], []
        for _ in range(steps):
            pos = rs.standard_normal((k, dim)) * scale
            lp = rs.standard_normal(k) * 1e4
            if neg_inf:
                lp[0] = -np.inf
            write_chain_step(path, pos, lp)
            want_c.append(pos)
            want_p.append(lp)
        chain, probs, n_params, n_walkers, n_steps = read_chain(path)
        assert (n_params, n_walkers, n_steps) == (dim, k, steps)
        np.testing.assert_allclose(chain, np.array(want_c), rtol=2e-7, atol=scale * 1e-8)   # numpy prints 8 digits
        np.testing.assert_allclose(probs, np.array(want_p), rtol=1e-12)
        if uu is not None and dim > 1:        # (the reference reader infers the walker count from rows of > 1 parameter)
            c_ref, p_ref, np_ref, nw_ref, ns_ref = uu.readChainFromFile(path)
            assert (np_ref, nw_ref, ns_ref) == (dim, k, steps)
            assert np.array_equal(c_ref, chain) and np.array_equal(p_ref, probs)

    run()
