"""Ensemble driver: the emcee 2.x ``EnsembleSampler`` stretch move with walkers sharded over GPUs.

The reference builds ``emcee.EnsembleSampler(nWalkers, nDim, lnprob, kwargs=..., threads=N | pool=...)``
and iterates ``sampler.sample(p0, iterations=...)`` unpacking ``(pos, lnprob, rstate)`` each step
(adv:300-347; simultFit.py:701-786).  emcee is a third-party dependency that is not vendored in the
reference; its algorithm (Goodman & Weare 2010 affine-invariant stretch move, red/blue halves,
``a = 2``) is restated here -- sampler parity is therefore statistical, not bitwise.

B200 layout: one process per GPU.  Every rank keeps a full replica of the positions
(``k * ndim * 8`` bytes); rank ``g`` of ``G`` owns rows ``[g*h/G, (g+1)*h/G)`` of EACH half
(``h = k/2``).  Per half-step a rank proposes, evaluates ``lnprob`` and accepts for its own slice
only, then one ``all_gather`` (NCCL over NVLink) of the updated slice -- positions and
log-probabilities packed as ``[h/G, ndim+1]`` -- refreshes the replicas.  Proposal randomness is a
counter-based Philox stream keyed by ``(seed, step, half, global walker index)``, so chains do not
depend on ``G``.
"""
from __future__ import annotations

import os
from typing import Iterator, Optional, Tuple

import numpy as np
import torch

try:  # torch.distributed is optional at import time
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


class CudaBackend:
    """Stretch-move kernels + batched lnprob of one :class:`TofModel` (device pointers, current stream)."""

    def __init__(self, model):
        self.model = model
        self.device = torch.device("cuda", model.device)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def propose(self, s, walker0, comp, a, seed, step, half):
        n, ndim = s.shape
        q = torch.empty_like(s)
        log_zz = torch.empty(n, dtype=torch.float64, device=s.device)
        self.model.stretch_propose(s.data_ptr(), n, walker0, comp.data_ptr(), comp.shape[0], a, seed, step, half,
                                   q.data_ptr(), log_zz.data_ptr(), self._stream())
        return q, log_zz

    def lnprob(self, q):
        out = torch.empty(q.shape[0], dtype=torch.float64, device=q.device)
        self.model.lnprob_batch_device(q.data_ptr(), q.shape[0], out.data_ptr(), self._stream())
        return out

    def accept(self, s, lp, walker0, q, new_lp, log_zz, seed, step, half, n_accept):
        self.model.stretch_accept(s.data_ptr(), lp.data_ptr(), s.shape[0], walker0, q.data_ptr(), new_lp.data_ptr(),
                                  log_zz.data_ptr(), seed, step, half, n_accept.data_ptr(), self._stream())

    def ensemble_step(self, pos, lp, steps, a, seed, step0, n_accept):
        """Whole steps for an unsharded ensemble in one library call (no per-half-step host work)."""
        self.model.ensemble_step(pos.data_ptr(), lp.data_ptr(), pos.shape[0], steps, a, seed, step0, n_accept.data_ptr(),
                                 self._stream())

    def half_step_packed(self, state, half, own0, n_own, a, seed, step, n_accept):
        """Propose -> lnprob -> accept for the rows this rank owns, in place on the packed ``[k, dim+1]`` state
        (tof_ensemble_half_step): one library call, no temporaries on the Python side."""
        self.model.ensemble_half_step(state.data_ptr(), state.shape[0], half, own0, n_own, a, seed, step, n_accept.data_ptr(),
                                      self._stream())


class EnsembleSampler:
    """emcee-2.x-shaped sampler over a sharded walker ensemble.

    Parameters mirror ``emcee.EnsembleSampler(nwalkers, dim, lnpostfn, a=2.0)``; ``lnpostfn`` is a
    :class:`~mcmctoffitting_b200.lnprob.TofLnProb` (or any object exposing ``.model``), or a backend
    is injected directly (tests run the sharding logic on CPU/gloo with a numpy backend).

    State layout: one ``[k, dim+1]`` float64 tensor per rank -- the positions of every walker followed by its
    log-probability.  A rank updates only the rows it owns; after each half-step ONE in-place all-gather of the
    half's ``[h, dim+1]`` slab (each rank contributes its ``[h/G, dim+1]`` rows) refreshes positions and
    log-probabilities on every replica.
    """

    def __init__(self, nwalkers: int, dim: int, lnpostfn=None, a: float = 2.0, seed: int = 0, backend=None,
                 group=None, store_chain: bool = True):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")            # emcee's own checks
        if nwalkers < 2 * dim:
            raise ValueError("The number of walkers needs to be more than twice the dimension of your parameter space.")
        self.k, self.dim, self.a, self.seed = int(nwalkers), int(dim), float(a), int(seed)
        self.backend = backend if backend is not None else CudaBackend(lnpostfn.model)
        self.device = self.backend.device
        self.group = group
        self.distributed = dist is not None and dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.h = self.k // 2
        if self.h % self.world != 0:
            raise ValueError("half-ensemble size %d must divide evenly over %d ranks" % (self.h, self.world))
        self.n_own = self.h // self.world
        # NCCL gathers in place (sendbuff == recvbuff + rank*count); other backends get a private copy of the input
        self._inplace_gather = self.distributed and dist.get_backend(group) == "nccl"
        self.store_chain = store_chain
        self.reset()
        self._step = 0

    # -- emcee surface ---------------------------------------------------------------------------------
    def reset(self) -> None:
        self._chain = []
        self._lnprob = []
        # every rank counts the acceptances of the rows it owns; the other rows are filled in by _gather_naccepted
        self._nacc = torch.zeros(self.k, dtype=torch.int64, device=self.device)
        self.iterations = 0

    @property
    def chain(self) -> np.ndarray:
        """[nwalkers, steps, dim] like emcee."""
        if not self._chain:
            return np.empty((self.k, 0, self.dim))
        return np.stack(self._chain, axis=1)

    @property
    def lnprobability(self) -> np.ndarray:
        if not self._lnprob:
            return np.empty((self.k, 0))
        return np.stack(self._lnprob, axis=1)

    @property
    def flatchain(self) -> np.ndarray:
        c = self.chain
        return c.reshape(-1, self.dim)

    def _own(self, half: int) -> slice:
        lo = half * self.h + self.rank * self.n_own
        return slice(lo, lo + self.n_own)

    def _gather_rows(self, full: torch.Tensor, half: int) -> None:
        """All-gather the rows of one half: every rank contributes the rows it owns."""
        lo = half * self.h
        slab, mine = full[lo:lo + self.h], full[self._own(half)]
        dist.all_gather_into_tensor(slab, mine if self._inplace_gather else mine.clone(), group=self.group)

    @property
    def naccepted(self) -> torch.Tensor:
        """Per-walker acceptance counts of the WHOLE ensemble.  With several ranks this is a collective (every rank
        must read it): each rank only counts for the walkers it owns."""
        if self.world > 1:
            for half in (0, 1):
                self._gather_rows(self._nacc, half)
        return self._nacc

    @naccepted.setter
    def naccepted(self, value: torch.Tensor) -> None:
        self._nacc = value

    @property
    def acceptance_fraction(self) -> np.ndarray:
        """emcee's ``acceptance_fraction`` (collective when sharded, see :attr:`naccepted`)."""
        return self.naccepted.cpu().numpy() / max(self.iterations, 1)

    # -- packed state ----------------------------------------------------------------------------------
    def pack_state(self, pos, lnprob) -> torch.Tensor:
        """``[k, dim+1]`` device tensor: positions, then the log-probability of each walker."""
        st = torch.empty((self.k, self.dim + 1), dtype=torch.float64, device=self.device)
        st[:, :self.dim] = self._as_device(pos, (self.k, self.dim))
        st[:, self.dim] = self._as_device(lnprob, (self.k,))
        return st

    # -- one red/blue half-step on the packed state --------------------------------------------------------
    def _half_step(self, state: torch.Tensor, half: int) -> None:
        h, n, g, d = self.h, self.n_own, self.rank, self.dim
        if hasattr(self.backend, "half_step_packed"):
            self.backend.half_step_packed(state, half, g * n, n, self.a, self.seed, self._step, self._nacc)
        else:
            own = self._own(half)
            comp = state[(1 - half) * h:(2 - half) * h, :d]
            s, lps = state[own, :d], state[own, d]
            q, log_zz = self.backend.propose(s, own.start, comp, self.a, self.seed, self._step, half)
            new_lp = self.backend.lnprob(q)
            self.backend.accept(s, lps, own.start, q, new_lp, log_zz, self.seed, self._step, half, self._nacc[own])
        if self.world > 1:
            self._gather_rows(state, half)             # positions AND log-probabilities, one collective, in place

    def _as_device(self, a, shape) -> torch.Tensor:
        t = torch.as_tensor(np.asarray(a, dtype=np.float64) if not torch.is_tensor(a) else a, dtype=torch.float64)
        return t.reshape(shape).to(self.device).contiguous().clone()

    def initial_lnprob(self, pos: torch.Tensor) -> torch.Tensor:
        """lnprob of every walker, each rank evaluating its share, gathered."""
        per = self.k // self.world
        mine = pos[self.rank * per:(self.rank + 1) * per].contiguous()
        lp_mine = self.backend.lnprob(mine)
        if self.world == 1:
            return lp_mine
        out = torch.empty(self.k, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(out, lp_mine, group=self.group)
        return out

    def sample(self, p0, lnprob0=None, rstate0=None, iterations: int = 1, storechain: Optional[bool] = None
               ) -> Iterator[Tuple[np.ndarray, np.ndarray, int]]:
        """Generator yielding ``(pos, lnprob, rstate)`` per step, like emcee 2.x ``sample``.  ``rstate``
        is the step counter of the counter-based generator (pass it back as ``rstate0`` to resume)."""
        store = self.store_chain if storechain is None else storechain
        pos = self._as_device(p0, (self.k, self.dim))
        if rstate0 is not None:
            self._step = int(rstate0)
        lp = self._as_device(lnprob0, (self.k,)) if lnprob0 is not None else self.initial_lnprob(pos)
        if bool(torch.isnan(lp).any()):
            raise ValueError("The initial lnprob was NaN.")                         # emcee raises the same
        state = self.pack_state(pos, lp)
        for _ in range(int(iterations)):
            self._half_step(state, 0)
            self._half_step(state, 1)
            self._step += 1
            self.iterations += 1
            host = state.cpu().numpy()
            p_host, lp_host = np.ascontiguousarray(host[:, :self.dim]), np.ascontiguousarray(host[:, self.dim])
            if store:
                self._chain.append(p_host.copy())
                self._lnprob.append(lp_host.copy())
            yield p_host, lp_host, self._step

    def run_mcmc(self, p0, N: int, rstate0=None, lnprob0=None):
        out = None
        for out in self.sample(p0, lnprob0, rstate0, iterations=N):
            pass
        return out

    # -- checkpoint / resume -------------------------------------------------------------------------------
    def save_checkpoint(self, path: str, pos, lnprob) -> None:
        """Binary state for an exact resume (the reference only has its append-only text chain, adv:314-317, and
        hands burn-in over to the main chain in memory, adv:337-339): positions and log-probabilities at full
        precision, the counter of the proposal generator, the seed, acceptance counters and the stored chain.
        Collective when sharded: every rank calls it (the acceptance counters are gathered), rank 0 writes to a
        temporary file and renames it over ``path`` (a crash mid-write leaves the previous checkpoint intact), and
        all ranks leave together so that a following :meth:`load_checkpoint` sees the finished file."""
        nacc = self.naccepted.cpu().numpy()                       # collective gather when world > 1
        if self.rank == 0:
            final = path if str(path).endswith(".npz") else str(path) + ".npz"
            tmp = final + ".tmp.npz"
            np.savez(tmp, pos=np.asarray(pos, dtype=np.float64), lnprob=np.asarray(lnprob, dtype=np.float64),
                     rstate=np.int64(self._step), seed=np.int64(self.seed), a=np.float64(self.a),
                     naccepted=nacc, iterations=np.int64(self.iterations),
                     chain=self.chain, lnprobability=self.lnprobability)
            os.replace(tmp, final)
        if self.world > 1:
            dist.barrier(group=self.group)

    def load_checkpoint(self, path: str):
        """Restore what :meth:`save_checkpoint` wrote; returns ``(pos, lnprob, rstate)`` ready for
        ``sample(pos, lnprob0=lnprob, rstate0=rstate)``.  The continued chain equals the uninterrupted one.  Every
        rank loads the full acceptance counters and keeps counting for the rows it owns."""
        with np.load(path if str(path).endswith(".npz") else str(path) + ".npz") as f:
            pos, lnprob = f["pos"], f["lnprob"]
            if pos.shape != (self.k, self.dim):
                raise ValueError("checkpoint holds %s positions, sampler expects %s" % (pos.shape, (self.k, self.dim)))
            if int(f["seed"]) != self.seed or float(f["a"]) != self.a:
                raise ValueError("checkpoint was written with seed=%d a=%g" % (int(f["seed"]), float(f["a"])))
            self._step = int(f["rstate"])
            self.iterations = int(f["iterations"])
            self._nacc = torch.from_numpy(f["naccepted"].copy()).to(self.device)
            self._chain = [c.copy() for c in np.moveaxis(f["chain"], 1, 0)]
            self._lnprob = [c.copy() for c in np.moveaxis(f["lnprobability"], 1, 0)]
            return pos.copy(), lnprob.copy(), self._step

    # -- device-resident stepping for throughput runs (no per-step host copies) ------------------------------
    def run_state(self, state: torch.Tensor, steps: int) -> None:
        """``steps`` whole steps on a packed state (see :meth:`pack_state`), in place, nothing copied to the host:
        per half-step one library call (propose -> lnprob -> accept for the own rows) and, when sharded, one
        in-place all-gather."""
        for _ in range(int(steps)):
            self._half_step(state, 0)
            self._half_step(state, 1)
            self._step += 1
            self.iterations += 1

    def run_device(self, pos: torch.Tensor, lp: torch.Tensor, steps: int) -> None:
        if self.world == 1 and hasattr(self.backend, "ensemble_step") and pos.is_contiguous() and lp.is_contiguous():
            # the whole loop inside the library: same kernels, same counters, same chain
            self.backend.ensemble_step(pos, lp, int(steps), self.a, self.seed, self._step, self._nacc)
            self._step += int(steps)
            self.iterations += int(steps)
            return
        state = self.pack_state(pos, lp)
        self.run_state(state, steps)
        pos.copy_(state[:, :self.dim])
        lp.copy_(state[:, self.dim])


# ---- chain files in the reference's text format (adv:314-317; simultFit.py:737-740) -----------------------
def write_chain_step(path: str, pos: np.ndarray, lnprob: Optional[np.ndarray] = None) -> None:
    """Append one step as ``"{k} {pos[k]} {lnprob[k]}\\n"`` per walker -- numpy's own array ``str`` with
    its line wrapping, which ``utilities.readChainFromFile`` (utilities.py:432-500) parses."""
    with open(path, "a") as fout:
        for k in range(pos.shape[0]):
            if lnprob is None:
                fout.write("{} {}\n".format(k, pos[k]))                              # adv:345-346
            else:
                fout.write("{} {} {}\n".format(k, pos[k], lnprob[k]))                # adv:315-316


def read_chain(path: str):
    """Reader for the format above: ``(chain[step, walker, param], probs[step, walker], nParams,
    nWalkers, nSteps)`` -- the return contract of utilities.readChainFromFile (utilities.py:432-500)."""
    idx, vals, probs = [], [], []
    with open(path, "r") as f:
        text = f.read()
    pos = 0
    n = len(text)
    while pos < n:
        lb = text.find("[", pos)
        if lb < 0:
            break
        rb = text.find("]", lb)
        idx.append(int(float(text[pos:lb])))
        vals.append([float(v) for v in text[lb + 1:rb].split()])
        nl = text.find("\n", rb)
        nl = n if nl < 0 else nl
        tail = text[rb + 1:nl].strip()
        probs.append(float(tail) if tail else float("nan"))
        pos = nl + 1
    n_walkers = max(idx) + 1
    n_steps = len(idx) // n_walkers
    chain = np.array(vals[:n_steps * n_walkers]).reshape(n_steps, n_walkers, -1)
    pr = np.array(probs[:n_steps * n_walkers]).reshape(n_steps, n_walkers)
    return chain, pr, chain.shape[2], n_walkers, n_steps
