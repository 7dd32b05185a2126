// Ensemble stretch-move kernels and the FP64 peak microbenchmark.
#pragma once
#include "tof_common.cuh"

namespace tof {

// ================================================================================================
// ensemble stretch move (emcee 2.x EnsembleSampler._propose_stretch, restated from Goodman & Weare)
// ================================================================================================
// counter layout: ctr_lo = global walker index, ctr_hi = step*4 + half*2 + kind (kind 0 propose, 1 accept)
// `ld_s` / `ld_c`: doubles between consecutive rows of s / comp (ndim for plain position arrays, ndim + 1 for the
// packed [positions, lnprob] state the sharded driver keeps so that ONE all-gather refreshes both)
__global__ void stretch_propose_kernel(const double *__restrict__ s, long long n, long long walker0,
                                       const double *__restrict__ comp, long long n_comp, int ndim, int ld_s, int ld_c, double a,
                                       unsigned long long seed, long long step, int half, double *__restrict__ q,
                                       double *__restrict__ log_zz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox rng(seed, (uint64_t)(walker0 + i), (uint64_t)step * 4ull + (uint64_t)half * 2ull);
    const double r = (a - 1.0) * rng.u0() + 1.0;
    const double zz = r * r / a;
    long long j = (long long)(rng.u1() * (double)n_comp);
    if (j >= n_comp) j = n_comp - 1;
    for (int p = 0; p < ndim; ++p) {
        const double c = comp[j * ld_c + p];
        q[i * ndim + p] = c - zz * (c - s[i * ld_s + p]);
    }
    log_zz[i] = (double)(ndim - 1) * log(zz);
}

__global__ void stretch_accept_kernel(double *__restrict__ s, double *__restrict__ lnprob, long long n, long long walker0,
                                      const double *__restrict__ q, const double *__restrict__ new_lnprob,
                                      const double *__restrict__ log_zz, int ndim, int ld_s, int ld_lp, unsigned long long seed,
                                      long long step, int half, long long *__restrict__ n_accept) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox rng(seed, (uint64_t)(walker0 + i), (uint64_t)step * 4ull + (uint64_t)half * 2ull + 1ull);
    const double lnpdiff = log_zz[i] + new_lnprob[i] - lnprob[i * ld_lp];
    if (lnpdiff > log(rng.u0())) {  // NaN and -inf proposals compare false: rejected
        for (int p = 0; p < ndim; ++p) s[i * ld_s + p] = q[i * ndim + p];
        lnprob[i * ld_lp] = new_lnprob[i];
        if (n_accept) n_accept[i] += 1;
    }
}

// ================================================================================================
// FP64 FMA peak microbenchmark: the roofline denominator for this path
// ================================================================================================
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}

}  // namespace tof
