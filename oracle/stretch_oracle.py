"""numpy restatement of the stretch-move kernels (TEST INFRASTRUCTURE; never imported by the product).

emcee 2.x ``EnsembleSampler._propose_stretch`` (third-party, un-vendored; Goodman & Weare 2010):
``zz = ((a-1) U + 1)^2 / a``, partner ``j = randint(n_comp)``, ``q = c_j - zz (c_j - s)``,
accept when ``(ndim-1) ln zz + lnprob(q) - lnprob(s) > ln U``.  Randomness: Philox4x32-10 with
key = seed and counter = (global walker index, step*4 + half*2 + kind), exactly as
``mcmctoffitting_b200/csrc/tof_device.cuh`` does, so the CUDA kernels can be checked bit for bit and the
sharding logic of ``ensemble.EnsembleSampler`` can run on CPU (gloo) with this backend injected.
"""
from __future__ import annotations

import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(seed: int, ctr_lo: np.ndarray, ctr_hi: np.ndarray):
    """Vectorised Philox4x32-10.  Returns the four 32-bit outputs as uint64 arrays."""
    ctr_lo = np.asarray(ctr_lo, dtype=np.uint64)
    ctr_hi = np.broadcast_to(np.asarray(ctr_hi, dtype=np.uint64), ctr_lo.shape)
    c0, c1 = ctr_lo & MASK, ctr_lo >> np.uint64(32)
    c2, c3 = ctr_hi & MASK, ctr_hi >> np.uint64(32)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def uniforms(seed, walker_idx, step, half, kind):
    c0, c1, c2, c3 = philox4x32(seed, walker_idx, np.uint64(step * 4 + half * 2 + kind))
    scale = 1.0 / 9007199254740992.0
    u0 = (((c1 << np.uint64(32)) | c0) >> np.uint64(11)).astype(np.float64) * scale
    u1 = (((c3 << np.uint64(32)) | c2) >> np.uint64(11)).astype(np.float64) * scale
    return u0, u1


def propose(s, walker0, comp, a, seed, step, half):
    n, ndim = s.shape
    idx = np.arange(walker0, walker0 + n, dtype=np.uint64)
    u0, u1 = uniforms(seed, idx, step, half, 0)
    r = (a - 1.0) * u0 + 1.0
    zz = r * r / a
    j = np.minimum((u1 * comp.shape[0]).astype(np.int64), comp.shape[0] - 1)
    c = comp[j]
    q = c - zz[:, None] * (c - s)
    return q, (ndim - 1) * np.log(zz)


def accept(s, lp, walker0, q, new_lp, log_zz, seed, step, half, n_accept):
    n = s.shape[0]
    idx = np.arange(walker0, walker0 + n, dtype=np.uint64)
    u0, _ = uniforms(seed, idx, step, half, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ok = (log_zz + new_lp - lp) > np.log(u0)
    s[ok] = q[ok]
    lp[ok] = new_lp[ok]
    if n_accept is not None:
        n_accept[ok] += 1
    return ok


class NumpyBackend:
    """CPU stand-in for ``ensemble.CudaBackend`` (tests only): torch CPU tensors in, numpy inside."""

    def __init__(self, lnprob_fn):
        self.fn = lnprob_fn
        self.device = torch.device("cpu")

    def propose(self, s, walker0, comp, a, seed, step, half):
        q, lz = propose(s.numpy(), walker0, comp.numpy(), a, seed, step, half)
        return torch.from_numpy(q), torch.from_numpy(lz)

    def lnprob(self, q):
        return torch.from_numpy(np.asarray(self.fn(q.numpy()), dtype=np.float64))

    def accept(self, s, lp, walker0, q, new_lp, log_zz, seed, step, half, n_accept):
        accept(s.numpy(), lp.numpy(), walker0, q.numpy(), new_lp.numpy(), log_zz.numpy(), seed, step, half,
               n_accept.numpy())
