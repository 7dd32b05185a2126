// Model kernels (sm_100a).  One CTA evaluates one walker (one (walker, run) pair for the
// simultaneous fit): Monte-Carlo draws are spread over the threads, every histogram lives in
// shared memory, and only theta (in) and lnprob (out) touch HBM.
#pragma once
#include "tof_device.cuh"

namespace tof {

// Outputs requested from a model kernel.  Production: only `lnprob`.
struct ModelOut {
    double *lnprob;       // [n] (adv/simple) or [n][n_runs] partials (simult)
    double *spectra;      // optional [n][T] at `stage`
    long long *cells;     // optional [n][X][E] integer cell counts
    int stage;
    unsigned long long *work;  // optional global work counter (persistent CTAs take walkers dynamically)
    // range kernel, banded launch: capacity of the cell histogram (cells) and of the staged T2 records
    int hcap, rcap;
    int *queue_out;                  // walkers that do not fit the banded layout ...
    unsigned long long *queue_count; // ... and how many
    const int *queue_in;             // full-size launch: process queue_in[0 .. *queue_count)
};

// ================================================================================================
// adv / intermediate model:  tests/advIntermediateTOFmodel.py:115-199
// ================================================================================================
//
// shared memory layout (doubles unless noted):
//   H[X*E]  weighted (x,E) histogram | tofc[T] (u64) | sx[X] | sdist[X] | svd[E] | svn[E]
//   xs_bp[n_xs] | xs_cf[(n_xs-1)*4] | staps[n_taps] | scratch[40] | xs_lut bytes
struct AdvSmem {
    double *H;
    unsigned long long *tofc;
    double *sx, *sdist, *svd, *svn, *xs_bp, *xs_cf, *staps, *scratch;
    unsigned char *xs_lut;
};

__host__ __device__ inline size_t adv_smem_bytes(int X, int E, int T, int n_xs, int n_taps, int lut_n) {
    size_t d = (size_t)X * E + T + 2 * (size_t)X + 2 * (size_t)E + n_xs + (size_t)(n_xs - 1) * 4 + n_taps + 40;
    return d * 8 + (((size_t)lut_n + 15) / 16) * 16;
}

__device__ __forceinline__ AdvSmem adv_carve(unsigned char *base, const DevModel &m, int T) {
    AdvSmem s;
    double *p = reinterpret_cast<double *>(base);
    s.H = p;            p += (size_t)m.x_bins * m.e_bins;
    s.tofc = reinterpret_cast<unsigned long long *>(p); p += T;
    s.sx = p;           p += m.x_bins;
    s.sdist = p;        p += m.x_bins;
    s.svd = p;          p += m.e_bins;
    s.svn = p;          p += m.e_bins;
    s.xs_bp = p;        p += m.n_xs;
    s.xs_cf = p;        p += (size_t)(m.n_xs - 1) * 4;
    s.staps = p;        p += m.n_taps;
    s.scratch = p;      p += 40;
    s.xs_lut = reinterpret_cast<unsigned char *>(p);
    return s;
}

// One RK4 step of size h for DPT independent energies (interleaved for ILP).
template <int DPT, int NMAT>
__device__ __forceinline__ void rk4_step(double (&E)[DPT], double h, const double *A, const double *B, int nmat) {
    double k1[DPT], k2[DPT], k3[DPT], k4[DPT];
    const double hh = 0.5 * h, h6 = h / 6.0;
#pragma unroll
    for (int k = 0; k < DPT; ++k) k1[k] = bethe<NMAT>(E[k], A, B, nmat);
#pragma unroll
    for (int k = 0; k < DPT; ++k) k2[k] = bethe<NMAT>(E[k] + hh * k1[k], A, B, nmat);
#pragma unroll
    for (int k = 0; k < DPT; ++k) k3[k] = bethe<NMAT>(E[k] + hh * k2[k], A, B, nmat);
#pragma unroll
    for (int k = 0; k < DPT; ++k) k4[k] = bethe<NMAT>(E[k] + h * k3[k], A, B, nmat);
#pragma unroll
    for (int k = 0; k < DPT; ++k) E[k] = E[k] + h6 * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
}

// Add the DPT samples of one thread at cell row `Hrow`; equal consecutive bins are merged in
// registers first (with sorted draws neighbouring samples share a bin), so that fewer shared
// memory atomics are issued.
template <int DPT>
__device__ __forceinline__ void hist_row(double *Hrow, const double (&E)[DPT], const DevModel &m, double e_step,
                                         double e_scale, const XsTab &xs) {
    int cur = -1;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int b = np_bin(E[k], m.e_bins, m.e_min, m.e_max, e_step, e_scale);
        if (b >= 0) {
            const double w = xs_eval(E[k], xs);  // adv:131 (only in-range samples are ever binned)
            if (b == cur) {
                acc += w;
            } else {
                if (cur >= 0) atomicAdd(Hrow + cur, acc);
                cur = b;
                acc = w;
            }
        }
    }
    if (cur >= 0) atomicAdd(Hrow + cur, acc);
}

// Stage common to adv and simult: timing-response convolution evaluated at bin n,
//   np.convolve(pdf, taps, 'same')[n] = sum_k taps[k] * pdf[n + shift - k],  shift = (n_taps-1)/2
// with pdf[t] = counts[t] / db[t] / total (np.histogram density=True, _histograms_impl.py).
template <typename CountT>
__device__ __forceinline__ double spread_at(int n, const CountT *cnt, double total, int T, double lo, double hi,
                                            double step, const double *taps, int n_taps, int shift) {
    double acc = 0.0;
    for (int k = 0; k < n_taps; ++k) {
        const int t = n + shift - k;
        if (t >= 0 && t < T) {
            const double db = __dsub_rn(np_edge(t + 1, T, lo, hi, step), np_edge(t, T, lo, hi, step));
            const double pdf = __ddiv_rn(__ddiv_rn((double)cnt[t], db), total);
            acc += taps[k] * pdf;
        }
    }
    return acc;
}

template <int NT, int DPT, int NMAT>
__global__ void __launch_bounds__(NT) adv_lnprob_kernel(const DevModel m, const DevRun run, const double *__restrict__ theta,
                                                        long long n_walkers, ModelOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = run.tof_bins;
    const AdvSmem s = adv_carve(smem_raw, m, T);
    const int tid = threadIdx.x;
    const int X = m.x_bins, EB = m.e_bins;
    const long long w = blockIdx.x;
    if (w >= n_walkers) return;

    const double e0 = theta[w * m.ndim + 0];
    const double sigma0 = theta[w * m.ndim + 1];

    // ---- lnprior (adv:185-189): outside the box -> -inf, no model evaluation (adv:196-198) ----
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = theta[w * m.ndim + p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside && out.spectra == nullptr && out.cells == nullptr) {
        if (tid == 0) out.lnprob[w] = -CUDART_INF;
        return;
    }

    // ---- stage tables, zero histograms -------------------------------------------------------------
    for (int i = tid; i < X * EB; i += NT) s.H[i] = 0.0;
    for (int i = tid; i < T; i += NT) s.tofc[i] = 0ull;
    for (int i = tid; i < X; i += NT) {
        s.sx[i] = m.x_centers[i];
        s.sdist[i] = run.neutron_dist[i];
    }
    for (int j = tid; j < EB; j += NT) {
        // adv:151-152: velocity of the deuteron at the mean of e0 and the bin centre
        const double eff = __ddiv_rn(__dadd_rn(e0, m.e_centers[j]), 2.0);
        s.svd[j] = speed_of(m.c, eff, m.m_d);
        s.svn[j] = m.neutron_speed[j];
    }
    for (int i = tid; i < m.n_xs; i += NT) s.xs_bp[i] = m.xs_breaks[i];
    for (int i = tid; i < (m.n_xs - 1) * 4; i += NT) s.xs_cf[i] = m.xs_coefs[i];
    for (int i = tid; i < m.n_taps; i += NT) s.staps[i] = m.taps[i];
    for (int i = tid; i < m.xs_lut_n; i += NT) s.xs_lut[i] = m.xs_lut[i];
    __syncthreads();

    XsTab xs;
    xs.bp = s.xs_bp; xs.cf = s.xs_cf; xs.lut = s.xs_lut; xs.n = m.n_xs; xs.lut_n = m.xs_lut_n;
    xs.lut_lo = m.xs_lut_lo; xs.lut_inv = m.xs_lut_inv;

    const double e_step = (m.e_max - m.e_min) / (double)EB;   // np.linspace step
    const double e_scale = (double)EB / (m.e_max - m.e_min);
    const double spread = __dmul_rn(sigma0, e0);                // np.random.normal(e0, sigma0*e0), adv:128

    // ---- phase 1: energy loss through the cell + cross-section weighted (x,E) histogram ------------
    for (long long base = (long long)tid * DPT; base < m.n_draws; base += (long long)NT * DPT) {
        double E[DPT];
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
            const long long d = base + k;
            E[k] = (d < m.n_draws) ? __dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + d))) : CUDART_NAN;
        }
        double x_prev = m.ode_from_zero ? 0.0 : s.sx[0];
        for (int i = 0; i < X; ++i) {
            if (i > 0 || m.ode_from_zero) {
                const double h = (s.sx[i] - x_prev) / (double)m.ode_substeps;
                for (int ss = 0; ss < m.ode_substeps; ++ss) rk4_step<DPT, NMAT>(E, h, m.bethe_A, m.bethe_B, m.n_materials);
                x_prev = s.sx[i];
            }
            hist_row<DPT>(s.H + (size_t)i * EB, E, m, e_step, e_scale, xs);
        }
    }
    __syncthreads();

    // ---- phase 2: normalise (adv:143) and quantise (adv:146) ---------------------------------------
    const double de = (m.e_max - m.e_min) / (double)EB;        // eD_binSize, adv:60
    const double dx = (m.x_max - m.x_min) / (double)X;         // x_binSize,  adv:70
    double part = 0.0;
    for (int i = tid; i < X * EB; i += NT) part += __dmul_rn(__dmul_rn(s.H[i], de), dx);
    const double S = block_sum<double>(part, s.scratch);

    // ---- phase 3: every non-empty cell becomes `count` events at one flight time (adv:149-158) ------
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    const double nsamp = (double)m.n_samples;
    for (int idx = tid; idx < X * EB; idx += NT) {
        const double cnt = rint(__dmul_rn(__ddiv_rn(s.H[idx], S), nsamp));  // NaN when S == 0, like numpy
        if (out.cells) out.cells[(size_t)w * X * EB + idx] = (cnt == cnt) ? (long long)cnt : LLONG_MIN;
        if (cnt != 0.0 && cnt == cnt) {
            const int i = idx / EB, j = idx - i * EB;
            const double tof_d = __ddiv_rn(s.sx[i], s.svd[j]);
            const double tof_n = __ddiv_rn(s.sdist[i], s.svn[j]);
            const int b = np_bin(__dadd_rn(tof_d, tof_n), T, run.tof_min, run.tof_max, t_step, t_scale);
            if (b >= 0) atomicAdd(s.tofc + b, (unsigned long long)(long long)cnt);
        }
    }
    __syncthreads();

    // ---- phase 4: density normalisation constant n.sum() (np.histogram density=True) -----------------
    long long cpart = 0;
    for (int t = tid; t < T; t += NT) cpart += (long long)s.tofc[t];
    const long long total_i = block_sum<long long>(cpart, reinterpret_cast<long long *>(s.scratch));
    // S == 0 or NaN (no in-range sample): numpy divides by zero -> every bin NaN -> lnlike NaN
    const bool degenerate = !(S > 0.0) || total_i == 0;
    const double total = (double)total_i;
    const long long *cnts = reinterpret_cast<const long long *>(s.tofc);

    if (out.spectra) {
        double *sp = out.spectra + (size_t)w * T;
        for (int t = tid; t < T; t += NT) {
            double v;
            if (out.stage == TOF_STAGE_COUNTS) {
                v = (double)cnts[t];
            } else if (degenerate) {
                v = CUDART_NAN;
            } else if (out.stage == TOF_STAGE_PDF) {
                const double db = __dsub_rn(np_edge(t + 1, T, run.tof_min, run.tof_max, t_step),
                                            np_edge(t, T, run.tof_min, run.tof_max, t_step));
                v = __ddiv_rn(__ddiv_rn((double)cnts[t], db), total);
            } else {
                v = spread_at(t, cnts, total, T, run.tof_min, run.tof_max, t_step, s.staps, m.n_taps, m.conv_shift);
            }
            sp[t] = v;
        }
    }

    // ---- phase 5: ln L = sum_{obs>0} obs * ln(model)  (adv:173-181) ----------------------------------
    double lp = 0.0;
    if (!degenerate) {
        for (int q = tid; q < run.n_obs_nz; q += NT) {
            const int t = run.obs_nz_idx[q];
            const double ev = spread_at(t, cnts, total, T, run.tof_min, run.tof_max, t_step, s.staps, m.n_taps, m.conv_shift);
            lp += run.obs_nz_val[q] * log(ev);  // ev == 0 -> -inf, as np.log does
        }
    }
    lp = block_sum<double>(lp, s.scratch);
    if (tid == 0 && out.lnprob) {
        double r = degenerate ? CUDART_NAN : lp;
        if (!inside) r = -CUDART_INF;
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
}

// ================================================================================================
// simple model: tests/simpleTOFmodel.py:57-120  (every sample is histogrammed directly)
// ================================================================================================
// grid = (chunks, walkers).  counts[n][T] (u64, zeroed by the caller) accumulate across chunks.
template <int NT>
__global__ void __launch_bounds__(NT) simple_hist_kernel(const DevModel m, const DevRun run, const double *__restrict__ theta,
                                                         long long n_walkers, unsigned long long *__restrict__ counts,
                                                         int ignore_prior) {
    __shared__ unsigned int sh[1024];
    const int T = run.tof_bins;
    const long long w = blockIdx.y;
    const int tid = threadIdx.x;
    const double e0 = theta[w * 3 + 0], e1 = theta[w * 3 + 1], sigma = theta[w * 3 + 2];
    bool inside = true;
    for (int p = 0; p < 3; ++p) {
        const double v = theta[w * 3 + p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside && !ignore_prior) return;  // lnprob never evaluates the model outside the prior (simple:117-119)
    for (int t = tid; t < T; t += NT) sh[t] = 0u;
    __syncthreads();

    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    // getDDneutronEnergy constants in the reference's order (simple:37-43)
    const double k_mm = __dmul_rn(m.m_d, m.m_n);
    const double k_den = __dadd_rn(m.m_n, m.m_he3);
    const double k_dm = __dsub_rn(m.m_he3, m.m_d);
    const double k_q = __dmul_rn(m.q_ddn, m.m_he3);

    const long long per = (m.n_draws + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per;
    const long long hi = (lo + per < m.n_draws) ? lo + per : m.n_draws;
    for (long long d = lo + tid; d < hi; d += NT) {
        const double x = __dmul_rn(m.cell_length, __ldg(run.z1 + d));                       // simple:62
        const double ed = __dadd_rn(__dadd_rn(e0, __dmul_rn(e1, x)), __dmul_rn(sigma, __ldg(run.z + d)));  // simple:64
        const double rv = __ddiv_rn(__dsqrt_rn(__dmul_rn(k_mm, ed)), k_den);               // rVal (cos 0 = 1)
        const double sv = __ddiv_rn(__dadd_rn(__dmul_rn(ed, k_dm), k_q), k_den);           // sVal
        const double sq = __dadd_rn(rv, __dsqrt_rn(__dadd_rn(__dmul_rn(rv, rv), sv)));
        const double en = __dmul_rn(sq, sq);
        const double dist = __dadd_rn(m.simple_neutron_base, __dsub_rn(m.cell_length, x));  // simple:66
        const double tof_n = __ddiv_rn(dist, speed_of(m.c, en, m.m_n));
        const double eff = __ddiv_rn(__dadd_rn(e0, ed), 2.0);
        const double tof_d = __ddiv_rn(x, speed_of(m.c, eff, m.m_d));
        const int b = np_bin(__dadd_rn(tof_n, tof_d), T, run.tof_min, run.tof_max, t_step, t_scale);
        if (b >= 0) atomicAdd(&sh[b], 1u);
    }
    __syncthreads();
    for (int t = tid; t < T; t += NT)
        if (sh[t]) atomicAdd(counts + (size_t)w * T + t, (unsigned long long)sh[t]);
}

// One CTA per walker: density, log, dot with the observations (simple:78-102).
template <int NT>
__global__ void __launch_bounds__(NT) simple_finish_kernel(const DevModel m, const DevRun run, const double *__restrict__ theta,
                                                           long long n_walkers, const unsigned long long *__restrict__ counts,
                                                           ModelOut out) {
    __shared__ double scratch[40];
    const int T = run.tof_bins;
    const long long w = blockIdx.x;
    const int tid = threadIdx.x;
    bool inside = true;
    for (int p = 0; p < 3; ++p) {
        const double v = theta[w * 3 + p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    const unsigned long long *cw = counts + (size_t)w * T;
    long long cpart = 0;
    for (int t = tid; t < T; t += NT) cpart += (long long)cw[t];
    const long long total_i = block_sum<long long>(cpart, reinterpret_cast<long long *>(scratch));
    const double total = (double)total_i;
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    if (out.spectra) {
        for (int t = tid; t < T; t += NT) {
            const double db = __dsub_rn(np_edge(t + 1, T, run.tof_min, run.tof_max, t_step),
                                        np_edge(t, T, run.tof_min, run.tof_max, t_step));
            out.spectra[(size_t)w * T + t] = (out.stage == TOF_STAGE_COUNTS) ? (double)cw[t]
                                                                               : __ddiv_rn(__ddiv_rn((double)cw[t], db), total);
        }
    }
    double lp = 0.0;
    for (int q = tid; q < run.n_obs_nz; q += NT) {
        const int t = run.obs_nz_idx[q];
        const double db = __dsub_rn(np_edge(t + 1, T, run.tof_min, run.tof_max, t_step),
                                    np_edge(t, T, run.tof_min, run.tof_max, t_step));
        const double pdf = __ddiv_rn(__ddiv_rn((double)cw[t], db), total);
        lp += run.obs_nz_val[q] * log(pdf);
    }
    lp = block_sum<double>(lp, scratch);
    if (tid == 0 && out.lnprob) {
        double r = (total_i == 0) ? CUDART_NAN : lp;
        if (!inside) r = -CUDART_INF;
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
}

// ================================================================================================
// ensemble stretch move (emcee 2.x EnsembleSampler._propose_stretch, restated from Goodman & Weare)
// ================================================================================================
// counter layout: ctr_lo = global walker index, ctr_hi = step*4 + half*2 + kind (kind 0 propose, 1 accept)
__global__ void stretch_propose_kernel(const double *__restrict__ s, long long n, long long walker0,
                                       const double *__restrict__ comp, long long n_comp, int ndim, double a,
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
        const double c = comp[j * ndim + p];
        q[i * ndim + p] = c - zz * (c - s[i * ndim + p]);
    }
    log_zz[i] = (double)(ndim - 1) * log(zz);
}

__global__ void stretch_accept_kernel(double *__restrict__ s, double *__restrict__ lnprob, long long n, long long walker0,
                                      const double *__restrict__ q, const double *__restrict__ new_lnprob,
                                      const double *__restrict__ log_zz, int ndim, unsigned long long seed, long long step,
                                      int half, long long *__restrict__ n_accept) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox rng(seed, (uint64_t)(walker0 + i), (uint64_t)step * 4ull + (uint64_t)half * 2ull + 1ull);
    const double lnpdiff = log_zz[i] + new_lnprob[i] - lnprob[i];
    if (lnpdiff > log(rng.u0())) {  // NaN and -inf proposals compare false: rejected
        for (int p = 0; p < ndim; ++p) s[i * ndim + p] = q[i * ndim + p];
        lnprob[i] = new_lnprob[i];
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

namespace tof {

// ================================================================================================
// adv / intermediate model, range-table formulation (TOF_ODE_RANGE)
// ================================================================================================
// The stopping ODE is autonomous, so u(E) = int dE/|f| turns "integrate every draw through every x"
// into v = u0_d + sgn*(x_i - x_start).  With the draws sorted, v is monotone along d for a fixed row,
// so one thread walks a run of consecutive draws with a pointer into the T2 table (bin + polynomial
// of the cross-section weight), sums whole runs in a register and touches the (x,E) histogram once
// per run instead of once per sample.
constexpr int RANGE_TILE = 1024;   // draws staged in shared memory at a time
constexpr int RANGE_ULUT = 1024;   // cells of the per-tile draw-index lookup table
constexpr int RANGE_STREAM_MIN = 8192;  // draws per walker from which the warp-private streaming walk is used
constexpr int RANGE_SPLIT = 6;      // long runs: pieces per warp when a tile has few (row, interval) tasks
constexpr int SIMULT_ULUT = 256;   // same for the 10-row simultaneous fit (fewer lookups per tile)

// hcap: cells of the (possibly banded) histogram; rcap: staged T2 records
__host__ __device__ inline size_t range_smem_bytes(int X, int E, int T, int hcap, int rcap, int P, int n_taps, int lut_n,
                                                   int rng_n) {
    size_t region_a = (size_t)T * 4 > (size_t)RANGE_TILE * 8 ? (size_t)T * 4 : (size_t)RANGE_TILE * 8;
    region_a = (region_a + 15) / 16 * 16;
    size_t d = (size_t)hcap + (size_t)rcap * (P + 3) + E + n_taps + 40 + X /* per-row offsets */;
    return d * 8 + region_a + (((size_t)lut_n * 2 + 15) / 16) * 16 + RANGE_ULUT * 2 + (((size_t)X * 8 + 15) / 16) * 16 +
           (size_t)rng_n * 8 + (((size_t)rng_n * 2 + 15) / 16) * 16 + 16;
}

constexpr int RANGE_CH = 40;        // runs longer than this are summed by the whole warp

// Interval of the T2 table that holds v (0 <= v <= u_max); uniform lookup cell, then edge compares.
// brk[j] = break that ends interval j (brk[M-1] = +inf).
__device__ __forceinline__ int range_interval(double v, const double *brk, const unsigned short *lut, double lut_inv, int lut_n,
                                              int M) {
    int c = (int)(v * lut_inv);
    c = c < 0 ? 0 : (c > lut_n - 1 ? lut_n - 1 : c);
    int j = lut[c];
    while (j + 1 < M && v >= brk[j]) ++j;
    while (j > 0 && v < brk[j - 1]) --j;
    return j;
}

// u0 = u(E0): T1 cell from the exponent/mantissa bits, degree-7 Horner in t in [-1, 1].
__device__ __forceinline__ double t1_eval(double E0, const DevModel &m) {
    double t;
    int idx;
    if (!(E0 >= m.e_tab_lo)) {                 // below the table, non-positive or NaN
        if (m.rng_sign > 0.0 && E0 > 0.0) {    // rising energies: clamp tiny E0 to the table start
            t = -1.0;
            idx = 0;
        } else {
            return -CUDART_INF;
        }
    } else if (E0 >= m.e_tab_hi) {
        return CUDART_INF;
    } else {
        const int hi = __double2hiint(E0), lo = __double2loint(E0);
        const int key = hi >> (20 - m.t1_q);
        idx = key - m.t1_key_lo;
        const double mant = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);   // [1, 2)
        const double c = (double)(key & ((1 << m.t1_q) - 1));
        t = (mant - 1.0) * (double)(1 << (m.t1_q + 1)) - (2.0 * c + 1.0);             // exact
    }
    const double *k = m.t1_coefs + 8 * idx;
    double acc = __ldg(k + 7);
#pragma unroll
    for (int q = 6; q >= 0; --q) acc = fma(acc, t, __ldg(k + q));
    return acc;
}

// Phase 1 for one tile of sorted u0 values (shared memory): add the cross-section weights of every (draw, row)
// sample to the (x,E) histogram H.  Called by all threads of the CTA (contains barriers).
template <int NT, int P>
// `brk`: the ends of all T2 intervals (shared memory) for interval searches; `rec`: the shared-memory copy of records
// jbase.. used by the tasks; H has `hstride` bins per row; row i starts at E-bin hlo[i] (banded layout; hlo == nullptr:
// every row starts at bin 0).
__device__ __forceinline__ void range_accumulate_tile(const double *u0, int nt, const double *brk, const double *rec, int jbase,
                                                      const unsigned short *lut, unsigned short *ulut, int n_ulut,
                                                      const double *sdelta, int *srow, double *H, int hstride, const int *hlo,
                                                      int X, int M, double umax, double lut_inv, int lut_n, int &bin_lo_all,
                                                      int &bin_hi_all) {
    constexpr int RW = P + 3;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsteps = 32 - __clz(nt);                 // binary-search iterations for [0, nt]
    // valid (finite) part of the sorted tile: -inf (never in range) first, +inf last
    int v_lo = 0, v_hi = nt;
    {
        int lo = 0, hi = nt, lo2 = 0, hi2 = nt;
        for (int it = 0; it < nsteps; ++it) {
            const int mid = (lo + hi) >> 1, mid2 = (lo2 + hi2) >> 1;
            const bool ge = u0[mid < nt ? mid : nt - 1] > -CUDART_INF;
            const bool gt = u0[mid2 < nt ? mid2 : nt - 1] >= CUDART_INF;
            const bool go = lo < hi, go2 = lo2 < hi2;
            hi = (go && ge) ? mid : hi;
            lo = (go && !ge) ? mid + 1 : lo;
            hi2 = (go2 && gt) ? mid2 : hi2;
            lo2 = (go2 && !gt) ? mid2 + 1 : lo2;
        }
        v_lo = lo;
        v_hi = lo2;
    }
    if (v_hi <= v_lo) return;                        // uniform: no usable draw in this tile
    const double tu_min = u0[v_lo], tu_max = u0[v_hi - 1];
    const double tu_inv = (tu_max > tu_min) ? (double)n_ulut / (tu_max - tu_min) : 0.0;
    // per-tile lookup: ulut[c] = first draw with u0 >= tu_min + c*cell
    for (int c = tid; c < n_ulut; c += NT) {
        const double x = tu_min + (double)c * ((tu_max - tu_min) / (double)n_ulut);
        int lo = v_lo, hi = v_hi;
        for (int it = 0; it < nsteps; ++it) {
            const int mid = (lo + hi) >> 1;
            const bool ge = u0[mid < nt ? mid : nt - 1] >= x;
            const bool go = lo < hi;
            hi = (go && ge) ? mid : hi;
            lo = (go && !ge) ? mid + 1 : lo;
        }
        ulut[c] = (unsigned short)lo;
    }
    // per-row interval of the tile's median draw: rows are processed along the trajectory (interval j = k + shift(row)),
    // so that the 32 lanes of a task look at the same slice of the draw distribution and have runs of similar length
    {
        const double u_med = u0[(v_lo + v_hi) >> 1];
        for (int i = tid; i < X; i += NT) {
            double vm = __dadd_rn(u_med, sdelta[i]);
            vm = vm < 0.0 ? 0.0 : (vm > umax ? umax : vm);
            srow[i] = range_interval(vm, brk, lut, lut_inv, lut_n, M);
        }
    }
    // band of T2 intervals any row of this tile can touch
    double dmin = sdelta[0], dmax = sdelta[0];
    {
        const double dl = sdelta[X - 1];
        dmin = dl < dmin ? dl : dmin;
        dmax = dl > dmax ? dl : dmax;                  // delta is monotone in the row index
    }
    const double vmin = __dadd_rn(tu_min, dmin), vmax = __dadd_rn(tu_max, dmax);
    __syncthreads();
    if (!(vmax >= 0.0) || vmin > umax) return;         // uniform
    const int band_lo = range_interval(vmin > 0.0 ? vmin : 0.0, brk, lut, lut_inv, lut_n, M);
    const int band_hi = range_interval(vmax < umax ? vmax : umax, brk, lut, lut_inv, lut_n, M);
    bin_lo_all = min(bin_lo_all, __double2loint(rec[(band_lo - jbase) * RW + 1]));
    bin_hi_all = max(bin_hi_all, __double2loint(rec[(band_hi - jbase) * RW + 1]));
    // One task = 32 (row, interval) cells.  Type A: one T2 interval x 32 consecutive rows (lane = row; all
    // lanes use the same polynomial).  Type B, for the X % 32 leftover rows: R rows x (32/R) consecutive
    // intervals.  A lane's draws are the contiguous range [lb, ub) found through the per-tile lookup.
    // Cell (row, bin) is produced by exactly one lane: plain read-modify-write, fixed summation order.
    const int Gf = X >> 5, R = X & 31;
    const int s_ref = srow[0];
    const int s_a = srow[0] - s_ref, s_b = srow[X - 1] - s_ref;       // shift is monotone in the row index
    const int s_min = s_a < s_b ? s_a : s_b, s_max = s_a < s_b ? s_b : s_a;
    const int k_lo = band_lo - s_max;
    const int n_iv = (band_hi - s_min) - k_lo + 1;
    const int per_b = R ? 32 / R : 1;
    const int nA = n_iv * Gf, nB = R ? (n_iv + per_b - 1) / per_b : 0;
    // One (row, interval) cell for this lane; with nch > 1 the lane takes piece `piece` of the run and the pieces
    // are combined with atomics (long runs: tiles of a big draw set cover few intervals).
    auto do_cell = [&](int row, int j, bool active, int piece, int nch) {
        active = active && j >= band_lo && j <= band_hi;
        j = j < band_lo ? band_lo : (j > band_hi ? band_hi : j);
        const double2 *rj = reinterpret_cast<const double2 *>(rec + (j - jbase) * RW);
        const double2 hd = rj[0];
        const double left = j ? rec[(j - 1 - jbase) * RW] : 0.0;
        const bool last = (j == M - 1);
        const double right = last ? umax : hd.x;
        const int bin = __double2loint(hd.y);
        double a[P + 1];
#pragma unroll
        for (int k = 0; k <= P; k += 2) {
            const double2 c2 = rj[1 + (k >> 1)];
            a[k] = c2.x;
            a[k + 1] = c2.y;
        }
        if (!active) return;
        const double delta = sdelta[row];
        // first draw with v >= left
        int c = (int)((left - delta - tu_min) * tu_inv);
        c = c < 0 ? 0 : (c > n_ulut - 1 ? n_ulut - 1 : c);
        int lb = ulut[c];
        while (lb > v_lo && __dadd_rn(u0[lb - 1], delta) >= left) --lb;
        while (lb < v_hi && !(__dadd_rn(u0[lb], delta) >= left)) ++lb;
        // first draw beyond the interval: v >= right (v > u_max for the last interval, which is closed)
        c = (int)((right - delta - tu_min) * tu_inv);
        c = c < 0 ? 0 : (c > n_ulut - 1 ? n_ulut - 1 : c);
        int ub = ulut[c];
        if (last) {
            while (ub > v_lo && __dadd_rn(u0[ub - 1], delta) > right) --ub;
            while (ub < v_hi && !(__dadd_rn(u0[ub], delta) > right)) ++ub;
        } else {
            while (ub > v_lo && __dadd_rn(u0[ub - 1], delta) >= right) --ub;
            while (ub < v_hi && !(__dadd_rn(u0[ub], delta) >= right)) ++ub;
        }
        if (nch > 1) {
            const int len = (ub - lb + nch - 1) / nch;
            lb += piece * len;
            ub = (lb + len < ub) ? lb + len : ub;
        }
        if (ub <= lb) return;
        double acc = 0.0;
        for (int d = lb; d < ub; ++d) {
            const double dt = __dadd_rn(u0[d], delta) - left;
            double wgt = a[P];
#pragma unroll
            for (int k = P - 1; k >= 0; --k) wgt = fma(wgt, dt, a[k]);
            acc += wgt;
        }
        const int col = bin - (hlo ? hlo[row] : 0);
        if ((unsigned)col >= (unsigned)hstride) return;       // cannot happen: the band has an interval of slack
        double *cell = H + (size_t)row * hstride + col;
        if (nch > 1 || __double2hiint(hd.y) < 0) atomicAdd(cell, acc);   // shared cell: pieces / bin split over intervals
        else *cell += acc;
    };
    const int GfD = Gf > 0 ? Gf : 1;
    const int n_tasks = nA + nB;
    // few tasks (a tile of a big draw set spans few intervals): split every run so that all warps have work
    int nch = 1;
    if (n_tasks < 2 * NW) {
        nch = (RANGE_SPLIT * NW + n_tasks - 1) / (n_tasks > 0 ? n_tasks : 1);   // ~RANGE_SPLIT pieces per warp
        nch = nch > 64 ? 64 : nch;
    }
    if (nch == 1) {
        // (jj, g) of type-A task `task` without a division in the loop
        int a_jj = warp / GfD, a_g = warp - a_jj * GfD;
        const int step_j = NW / GfD, step_g = NW - step_j * GfD;
        for (int task = warp; task < n_tasks; task += NW) {
            if (task < nA) {
                const int row = (a_g << 5) + lane;
                do_cell(row, k_lo + a_jj + (srow[row] - s_ref), true, 0, 1);
                a_jj += step_j;
                a_g += step_g;
                if (a_g >= GfD) {
                    a_g -= GfD;
                    ++a_jj;
                }
            } else {
                const int isub = lane / R;
                const int row = (Gf << 5) + (lane - isub * R);
                do_cell(row, k_lo + (task - nA) * per_b + isub + (srow[row] - s_ref), isub < per_b, 0, 1);
            }
        }
    } else {
        for (int t2 = warp; t2 < n_tasks * nch; t2 += NW) {
            const int task = t2 / nch, piece = t2 - task * nch;
            if (task < nA) {
                const int jj = task / GfD;
                const int row = ((task - jj * GfD) << 5) + lane;
                do_cell(row, k_lo + jj + (srow[row] - s_ref), true, piece, nch);
            } else {
                const int isub = lane / R;
                const int row = (Gf << 5) + (lane - isub * R);
                do_cell(row, k_lo + (task - nA) * per_b + isub + (srow[row] - s_ref), isub < per_b, piece, nch);
            }
        }
    }
}

template <int NT, int P>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) adv_range_kernel(const DevModel m, const DevRun run, const double *__restrict__ theta,
                                                       long long n_walkers, ModelOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RW = P + 3;
    const int T = run.tof_bins, X = m.x_bins, EB = m.e_bins, M = m.rng_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    // ---- carve --------------------------------------------------------------------------------------
    // banded launch (out.hcap < X*EB): the cell histogram holds only the E-bins this walker can touch and only the
    // matching T2 records are staged, so that two 512-thread CTAs fit one SM; walkers that do not fit are queued
    // for the full-size launch
    const bool banded = out.hcap < X * EB;
    double *H = reinterpret_cast<double *>(smem_raw);
    size_t region_a = (size_t)T * 4 > (size_t)RANGE_TILE * 8 ? (size_t)T * 4 : (size_t)RANGE_TILE * 8;
    region_a = (region_a + 15) / 16 * 16;
    unsigned char *pa = reinterpret_cast<unsigned char *>(H + (size_t)out.hcap);
    unsigned int *tofc = reinterpret_cast<unsigned int *>(pa);
    double *u0 = reinterpret_cast<double *>(pa);                       // aliases tofc (phase 1 only)
    double *rec = reinterpret_cast<double *>(pa + region_a);
    double *svd = rec + (size_t)out.rcap * RW;
    double *staps = svd + EB;
    double *scratch = staps + m.n_taps;
    double *sdelta = scratch + 40;                                      // [X] sgn*(x_i - x_start)
    unsigned short *lut = reinterpret_cast<unsigned short *>(sdelta + X);
    unsigned short *ulut = lut + ((m.rng_lut_n + 7) / 8) * 8;           // [RANGE_ULUT]
    int *srow = reinterpret_cast<int *>(ulut + RANGE_ULUT);             // [X]
    int *hlo_s = srow + X;                                              // [X] first E-bin of each row (banded launch)
    double *sbrk = reinterpret_cast<double *>(hlo_s + X + (X & 1));     // [M] interval ends
    unsigned short *sbin = reinterpret_cast<unsigned short *>(sbrk + M);  // [M] E-bin of each interval
    __shared__ int s_band[3];                                           // widest row, first / last interval of the walker

    // ---- walker-independent tables: staged once per CTA (persistent CTAs loop over walkers) ------------------
    if (!banded)
        for (int i = tid; i < M * RW; i += NT) rec[i] = m.rng_rec[i];
    const double *recf = m.rng_rec;                        // full table in global memory
    for (int j = tid; j < M; j += NT) {
        sbrk[j] = recf[(size_t)j * RW];
        sbin[j] = (unsigned short)__double2loint(recf[(size_t)j * RW + 1]);
    }
    for (int i = tid; i < m.rng_lut_n; i += NT) lut[i] = m.rng_lut[i];
    for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
    const double sgn = m.rng_sign, umax = m.rng_u_max;
    const double x_start = m.ode_from_zero ? 0.0 : m.x_centers[0];
    for (int i = tid; i < X; i += NT) sdelta[i] = sgn * (m.x_centers[i] - x_start);
    __shared__ long long s_next;
    for (long long iter = 0;; ++iter) {
    __syncthreads();                                       // the previous walker is done with shared memory
    if (tid == 0)
        s_next = out.work ? (long long)atomicAdd(out.work, 1ull) : (long long)blockIdx.x + iter * (long long)gridDim.x;
    __syncthreads();
    const long long item = s_next;
    // the full-size launch of a banded call works through the queue the banded launch filled
    const long long n_items = out.queue_in ? (long long)*out.queue_count : n_walkers;
    if (item >= n_items) break;
    const long long w = out.queue_in ? (long long)out.queue_in[item] : item;
    const double e0 = theta[w * m.ndim + 0];
    const double sigma0 = theta[w * m.ndim + 1];
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = theta[w * m.ndim + p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside && out.spectra == nullptr && out.cells == nullptr) {
        if (tid == 0) out.lnprob[w] = -CUDART_INF;
        continue;
    }

    const double spread = __dmul_rn(sigma0, e0);          // adv:128
    const bool rev = spread < 0.0;                         // draws are sorted ascending: E0 ascends unless the spread is negative
    // ---- per walker: E-bins it can touch (the draws are sorted: first and last give the extremes) ------------
    int hstride = EB, jbase = 0;
    const int *hlo = nullptr;
    if (banded) {
        const double u_lo = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 : 0)))), m);
        const double u_hi = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? 0 : m.n_draws - 1)))), m);
        if (tid == 0) {
            s_band[0] = 0;
            s_band[1] = M;
            s_band[2] = -1;
        }
        __syncthreads();
        // every row has its own window of E-bins: [u_lo + delta_i, u_hi + delta_i], one interval of slack on both
        // sides (T1 is only monotone up to its 2e-13 cm fit error)
        for (int i = tid; i < X; i += NT) {
            double vmin = (u_lo > -CUDART_INF ? u_lo : 0.0) + sdelta[i];   // -inf draws: the lowest in-range v is 0
            double vmax = u_hi + sdelta[i];
            vmin = vmin > 0.0 ? vmin : 0.0;
            vmax = vmax < umax ? vmax : umax;
            int j_lo = 0, j_hi = 0;
            if (vmax >= vmin) {                               // otherwise this row gets nothing: any window will do
                j_lo = range_interval(vmin, sbrk, lut, m.rng_lut_inv, m.rng_lut_n, M);
                j_hi = range_interval(vmax, sbrk, lut, m.rng_lut_inv, m.rng_lut_n, M);
                j_lo = j_lo > 0 ? j_lo - 1 : 0;
                j_hi = j_hi < M - 1 ? j_hi + 1 : M - 1;
                atomicMin(&s_band[1], j_lo);
                atomicMax(&s_band[2], j_hi);
            }
            const int b_lo = sbin[j_lo];
            hlo_s[i] = b_lo;
            atomicMax(&s_band[0], (int)sbin[j_hi] - b_lo + 1);
        }
        __syncthreads();
        hstride = s_band[0];
        const int j_lo_all = s_band[2] >= 0 ? s_band[1] : 0, j_hi_all = s_band[2] >= 0 ? s_band[2] : 0;
        jbase = j_lo_all > 0 ? j_lo_all - 1 : 0;
        hlo = hlo_s;
        const bool fits = (long long)X * hstride <= out.hcap && (j_hi_all - jbase + 1) <= out.rcap && T <= out.hcap;
        if (!fits) {                                          // queue for the full-size launch
            if (tid == 0) out.queue_out[atomicAdd(out.queue_count, 1ull)] = (int)w;
            continue;
        }
        for (int i = tid; i < (j_hi_all - jbase + 1) * RW; i += NT) rec[i] = recf[(size_t)jbase * RW + i];
    }
    // ---- per walker: zero the cell histogram, deuteron speeds ------------------------------------------------
    for (int i = tid; i < X * hstride; i += NT) H[i] = 0.0;
    for (int j = tid; j < EB; j += NT) {
        const double eff = __ddiv_rn(__dadd_rn(e0, m.e_centers[j]), 2.0);   // adv:151
        svd[j] = speed_of(m.c, eff, m.m_d);
    }

    // ---- phase 1: (x,E) histogram of cross-section weights through the range tables ---------------------
    int bin_lo_all = EB, bin_hi_all = -1;                  // E-bins any draw of any tile can have touched (uniform)
    if (m.n_draws >= RANGE_STREAM_MIN) {
        // Big draw sets: the sorted draws of one interval are hundreds of consecutive values, so a lane that walks
        // draws in order changes interval rarely.  Warp-private streaming, no barriers: a warp takes 128 consecutive
        // draws (4 per lane, T1 evaluated once, kept in registers and broadcast by shuffle) and, for every group of
        // 32 rows, lane = row walks the 128 samples with an interval pointer; runs go to H with atomics.
        __syncthreads();                                   // staging done
        bin_lo_all = 0;
        bin_hi_all = EB - 1;
        const int n_groups = (X + 31) >> 5;
        for (long long base = (long long)warp * 128; base < m.n_draws; base += (long long)NW * 128) {
            double ur[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long d = base + q * 32 + lane;
                ur[q] = (d < m.n_draws) ? t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 - d : d)))), m)
                                        : CUDART_INF;
            }
            for (int g = 0; g < n_groups; ++g) {
                const int row = (g << 5) + lane;
                const bool rowok = row < X;
                const double delta = rowok ? sdelta[row] : 0.0;
                const int row_lo = (rowok && hlo) ? hlo[row] : 0;
                double *Hrow = H + (size_t)(rowok ? row : 0) * hstride;
                int bin = -1;
                double next = -CUDART_INF, brk = CUDART_INF, acc = 0.0;   // forces a lookup at the first in-range sample
                double a[P + 1];
#pragma unroll
                for (int k = 0; k <= P; ++k) a[k] = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    for (int k = 0; k < 32; ++k) {
                        const double v = __dadd_rn(__shfl_sync(FULL, ur[q], k), delta);
                        if (rowok && v >= 0.0 && v <= umax) {
                            if (v >= next || v < brk) {          // another interval (rare: runs are long)
                                const int j = range_interval(v, sbrk, lut, m.rng_lut_inv, m.rng_lut_n, M);
                                const double2 *rj = reinterpret_cast<const double2 *>(rec + (j - jbase) * RW);
                                const double2 hd = rj[0];
                                next = hd.x;
                                brk = j ? sbrk[j - 1] : 0.0;
                                const int nb = __double2loint(hd.y);
                                if (nb != bin) {
                                    const int col = bin - row_lo;
                                    if (bin >= 0 && (unsigned)col < (unsigned)hstride) atomicAdd(Hrow + col, acc);
                                    acc = 0.0;
                                    bin = nb;
                                }
#pragma unroll
                                for (int c = 0; c <= P; c += 2) {
                                    const double2 c2 = rj[1 + (c >> 1)];
                                    a[c] = c2.x;
                                    a[c + 1] = c2.y;
                                }
                            }
                            const double dt = v - brk;
                            double wgt = a[P];
#pragma unroll
                            for (int c = P - 1; c >= 0; --c) wgt = fma(wgt, dt, a[c]);
                            acc += wgt;
                        }
                    }
                }
                const int col = bin - row_lo;
                if (bin >= 0 && (unsigned)col < (unsigned)hstride) atomicAdd(Hrow + col, acc);
            }
        }
    } else {
        for (long long tile = 0; tile < m.n_draws; tile += RANGE_TILE) {
            const int nt = (int)((m.n_draws - tile < RANGE_TILE) ? (m.n_draws - tile) : RANGE_TILE);
            __syncthreads();                               // previous tile fully consumed / staging done
            for (int d = tid; d < nt; d += NT)
                u0[d] = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 - (tile + d) : tile + d)))), m);
            __syncthreads();
            range_accumulate_tile<NT, P>(u0, nt, sbrk, rec, jbase, lut, ulut, RANGE_ULUT, sdelta, srow, H, hstride, hlo, X, M,
                                         umax, m.rng_lut_inv, m.rng_lut_n, bin_lo_all, bin_hi_all);
        }
    }
    __syncthreads();

    // ---- phase 2: normalise (adv:143) ---------------------------------------------------------------------
    for (int i = tid; i < T; i += NT) tofc[i] = 0u;        // u0 is dead now
    const double de = (m.e_max - m.e_min) / (double)EB;
    const double dx = (m.x_max - m.x_min) / (double)X;
    // banded launch: every row holds `hstride` bins from hlo[row]; full-size launch: only bins bin_lo_all..bin_hi_all
    // can be non-zero
    const int nbw = hlo ? hstride : bin_hi_all - bin_lo_all + 1;
    double part = 0.0;
    for (int row = warp; row < X; row += NW) {
        const double *Hr = H + (size_t)row * hstride + (hlo ? 0 : bin_lo_all);
        for (int jb = lane; jb < nbw; jb += 32) part += __dmul_rn(__dmul_rn(Hr[jb], de), dx);
    }
    const double S = block_sum<double>(part, scratch);     // includes the barrier that publishes tofc = 0

    // ---- phase 3: quantise (adv:146) and scatter every non-empty cell to its flight time (adv:149-158) ----
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    const double nsamp = (double)m.n_samples;
    if (out.cells) {                                        // debug output: every cell, zeros included
        for (int idx = tid; idx < X * EB; idx += NT) {
            const double cnt = rint(__dmul_rn(__ddiv_rn(H[idx], S), nsamp));
            out.cells[(size_t)w * X * EB + idx] = (cnt == cnt) ? (long long)cnt : LLONG_MIN;
        }
    }
    for (int row = warp; row < X; row += NW) {
        const double xi = __ldg(m.x_centers + row), di = __ldg(run.neutron_dist + row);
        const int row_lo = hlo ? hlo[row] : bin_lo_all;
        const double *Hr = H + (size_t)row * hstride + (hlo ? 0 : bin_lo_all);
        for (int jb = lane; jb < nbw; jb += 32) {
            const int j = row_lo + jb;
            if (j >= EB) break;
            const double h = Hr[jb];
            if (h != 0.0 && S > 0.0) {
                const double cnt = rint(__dmul_rn(__ddiv_rn(h, S), nsamp));
                if (cnt > 0.0) {
                    const double tof_d = __ddiv_rn(xi, svd[j]);
                    const double tof_n = __ddiv_rn(di, __ldg(m.neutron_speed + j));
                    const int b = np_bin(__dadd_rn(tof_d, tof_n), T, run.tof_min, run.tof_max, t_step, t_scale);
                    if (b >= 0) atomicAdd(tofc + b, (unsigned int)cnt);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 4: density (np.histogram density=True) into the (now free) H region -------------------------
    long long cpart = 0;
    for (int t = tid; t < T; t += NT) cpart += (long long)tofc[t];
    const long long total_i = block_sum<long long>(cpart, reinterpret_cast<long long *>(scratch));
    const bool degenerate = !(S > 0.0) || total_i == 0;
    const double total = (double)total_i;
    double *pdf = H;
    for (int t = tid; t < T; t += NT) {
        const unsigned int cn = tofc[t];
        double v = 0.0;
        if (cn) {
            const double db = __dsub_rn(np_edge(t + 1, T, run.tof_min, run.tof_max, t_step),
                                        np_edge(t, T, run.tof_min, run.tof_max, t_step));
            v = __ddiv_rn(__ddiv_rn((double)cn, db), total);
        }
        pdf[t] = v;
    }
    __syncthreads();

    if (out.spectra) {
        double *sp = out.spectra + (size_t)w * T;
        for (int t = tid; t < T; t += NT) {
            double v;
            if (out.stage == TOF_STAGE_COUNTS) {
                v = (double)tofc[t];
            } else if (degenerate) {
                v = CUDART_NAN;
            } else if (out.stage == TOF_STAGE_PDF) {
                v = pdf[t];
            } else {
                v = 0.0;
                for (int k = 0; k < m.n_taps; ++k) {
                    const int tt = t + m.conv_shift - k;
                    if (tt >= 0 && tt < T) v += staps[k] * pdf[tt];
                }
            }
            sp[t] = v;
        }
    }

    // ---- phase 5: timing response at the observed bins + log-likelihood (adv:173-181) ------------------------
    double lp = 0.0;
    if (!degenerate) {
        for (int q = tid; q < run.n_obs_nz; q += NT) {
            const int t = run.obs_nz_idx[q];
            double ev = 0.0;
            for (int k = 0; k < m.n_taps; ++k) {
                const int tt = t + m.conv_shift - k;
                if (tt >= 0 && tt < T) ev += staps[k] * pdf[tt];
            }
            lp += run.obs_nz_val[q] * log(ev);
        }
    }
    lp = block_sum<double>(lp, scratch);
    if (tid == 0 && out.lnprob) {
        double r = degenerate ? CUDART_NAN : lp;
        if (!inside) r = -CUDART_INF;
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
    }   // persistent walker loop
}

}  // namespace tof

namespace tof {

// ================================================================================================
// simultaneous multi-standoff fit: tests/simultFit.py:223-300 (model), 380-469 (likelihood)
// ================================================================================================
// One CTA per (walker, run).  Shared memory: one (x,E) histogram copy per warp (contention), the run's
// TOF histogram, cross-section table.
struct DevRunSet {
    DevRun r[TOF_MAX_RUNS];
};

__host__ __device__ inline size_t simult_smem_bytes(int NT, int X, int E, int T, int n_xs, int n_taps, int lut_n) {
    size_t d = (size_t)(NT / 32) * X * E + 2 * (size_t)T + X + E + n_xs + (size_t)(n_xs - 1) * 4 + n_taps + 48;
    return d * 8 + (((size_t)lut_n + 15) / 16) * 16;
}

// Everything after the (x,E) histogram of one (walker, run): normalise, quantise, flight times with the
// zero-degree sub-times, density, timing response, per-bin likelihood (simultFit.py:279-300, 389-409).
template <int NT>
__device__ __forceinline__ void simult_tail(const DevModel &m, const DevRun &run, int r, long long w, const ModelOut &out,
                                            double *H, double *tofh, double *pdf, const double *sx, double *svd,
                                            const double *staps, double *scratch, double sum_e0_last, double sf,
                                            bool exhausted) {
    const int T = run.tof_bins, X = m.x_bins, EB = m.e_bins, CELLS = X * EB;
    const int tid = threadIdx.x;
    // ---- normalise, quantise (simultFit.py:279-283) --------------------------------------------------------
    const double de = (m.e_max - m.e_min) / (double)EB, dx = (m.x_max - m.x_min) / (double)X;
    double part = 0.0;
    for (int c = tid; c < CELLS; c += NT) part += __dmul_rn(__dmul_rn(H[c], de), dx);
    const double S = block_sum<double>(part, scratch);
    const double e0mean = __ddiv_rn(sum_e0_last, (double)m.n_ev_per_loop);
    for (int j = tid; j < EB; j += NT) {
        const double eff = __ddiv_rn(__dadd_rn(e0mean, m.e_centers[j]), 2.0);             // simultFit.py:288
        svd[j] = speed_of(m.c, eff, m.m_d);
    }
    __syncthreads();

    // ---- cells -> flight times, 10 zero-degree sub-times each (simultFit.py:286-299) -----------------------
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    const double nsamp = (double)m.n_samples;
    const int NZ = m.n_zero_deg;
    for (int idx = tid; idx < CELLS; idx += NT) {
        const double cnt = rint(__dmul_rn(__ddiv_rn(H[idx], S), nsamp));
        if (out.cells) out.cells[(size_t)w * CELLS + idx] = (cnt == cnt) ? (long long)cnt : LLONG_MIN;
        if (cnt != 0.0 && cnt == cnt) {
            const int i = idx / EB, j = idx - i * EB;
            const double tof_d = __ddiv_rn(sx[i], svd[j]);
            const double tof_n = __ddiv_rn(__ldg(run.neutron_dist + i), __ldg(m.neutron_speed + j));
            const double base = __dadd_rn(tof_d, tof_n);
            if (NZ == 0) {
                const int b = np_bin(base, T, run.tof_min, run.tof_max, t_step, t_scale);
                if (b >= 0) atomicAdd(tofh + b, cnt);
            } else {
                for (int k = 0; k < NZ; ++k) {
                    const double tof = __dadd_rn(base, __ldg(m.zd_times + j * NZ + k));
                    const int b = np_bin(tof, T, run.tof_min, run.tof_max, t_step, t_scale);
                    if (b >= 0) atomicAdd(tofh + b, __dmul_rn(cnt, __ldg(m.zd_weights + j * NZ + k)));
                }
            }
        }
    }
    __syncthreads();

    // ---- density, timing response, per-bin likelihood (simultFit.py:298-300, 389-409) ------------------------
    double tpart = 0.0;
    for (int t = tid; t < T; t += NT) tpart += tofh[t];
    const double total = block_sum<double>(tpart, scratch);
    const bool degenerate = exhausted || !(S > 0.0) || !(total != 0.0);
    for (int t = tid; t < T; t += NT) {
        const double db = __dsub_rn(np_edge(t + 1, T, run.tof_min, run.tof_max, t_step),
                                    np_edge(t, T, run.tof_min, run.tof_max, t_step));
        pdf[t] = __ddiv_rn(__ddiv_rn(tofh[t], db), total);
    }
    __syncthreads();
    double lp = 0.0;
    for (int t = tid; t < T; t += NT) {
        double acc = 0.0;
        for (int k = 0; k < m.n_taps; ++k) {
            const int tt = t + m.conv_shift - k;
            if (tt >= 0 && tt < T) acc += staps[k] * pdf[tt];
        }
        double ev = __dmul_rn(sf, acc);                                                    // simultFit.py:300
        if (out.spectra) {
            const double v = out.stage == TOF_STAGE_COUNTS ? tofh[t] : (out.stage == TOF_STAGE_PDF ? pdf[t] : ev);
            out.spectra[(size_t)w * T + t] = (degenerate && out.stage != TOF_STAGE_COUNTS) ? CUDART_NAN : v;
        }
        const double o = run.obs ? run.obs[t] : 1.0;                                       // 0 -> 1 done at upload
        if (ev == 0.0) ev = 1.0;                                                           // simultFit.py:393-394
        double poi = -o - lgamma(trunc(ev) + 1.0);                                         // simultFit.py:397
        if (ev > 0.0) poi += ev * log(o);                                                  // simultFit.py:398-399
        lp += o * poi;                                                                     // simultFit.py:400
    }
    lp = block_sum<double>(lp, scratch);
    if (tid == 0 && out.lnprob) out.lnprob[w * m.n_runs + r] = degenerate ? CUDART_NAN : lp;
}

template <int NT>
__global__ void __launch_bounds__(NT) simult_run_kernel(const DevModel m, const DevRunSet runs, const double *__restrict__ theta,
                                                        long long n_walkers, ModelOut out, int only_run) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const int n_launch_runs = (only_run >= 0) ? 1 : m.n_runs;
    const long long w = blockIdx.x / n_launch_runs;
    const int r = (only_run >= 0) ? only_run : (int)(blockIdx.x % n_launch_runs);
    if (w >= n_walkers) return;
    const DevRun &run = runs.r[r];
    const int T = run.tof_bins, X = m.x_bins, EB = m.e_bins, CELLS = X * EB;
    const int tid = threadIdx.x, warp = tid >> 5;

    double *Hw = reinterpret_cast<double *>(smem_raw);          // [NW][CELLS]
    double *tofh = Hw + (size_t)NW * CELLS;                      // [T]
    double *pdf = tofh + T;                                      // [T]
    double *sx = pdf + T;                                        // [X]
    double *svd = sx + X;                                        // [E]
    double *xs_bp = svd + EB;
    double *xs_cf = xs_bp + m.n_xs;
    double *staps = xs_cf + (size_t)(m.n_xs - 1) * 4;
    double *scratch = staps + m.n_taps;                          // [48]
    unsigned char *xs_lut = reinterpret_cast<unsigned char *>(scratch + 48);

    const double *th = theta + w * m.ndim;
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = th[p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside && out.spectra == nullptr && out.cells == nullptr) return;   // the finish kernel writes -inf

    const double beamE = th[0], eLoss = th[1], scale = th[2], sshape = th[3], sf = th[4 + r];

    for (int i = tid; i < NW * CELLS; i += NT) Hw[i] = 0.0;
    for (int i = tid; i < T; i += NT) tofh[i] = 0.0;
    for (int i = tid; i < X; i += NT) sx[i] = m.x_centers[i];
    for (int i = tid; i < m.n_xs; i += NT) xs_bp[i] = m.xs_breaks[i];
    for (int i = tid; i < (m.n_xs - 1) * 4; i += NT) xs_cf[i] = m.xs_coefs[i];
    for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
    for (int i = tid; i < m.xs_lut_n; i += NT) xs_lut[i] = m.xs_lut[i];
    __syncthreads();
    XsTab xs;
    xs.bp = xs_bp; xs.cf = xs_cf; xs.lut = xs_lut; xs.n = m.n_xs; xs.lut_n = m.xs_lut_n;
    xs.lut_lo = m.xs_lut_lo; xs.lut_inv = m.xs_lut_inv;
    const double e_step = (m.e_max - m.e_min) / (double)EB;
    const double e_scale = (double)EB / (m.e_max - m.e_min);
    double *Hmine = Hw + (size_t)warp * CELLS;

    // ---- draws -> initial energies -> stopping -> weighted (x,E) histogram -------------------------------
    // simultFit.py:243-252: E0 = beamE - lognorm.rvs(s, loc=eLoss, scale); entries <= 0 are redrawn, the whole
    // bad list at once, until none is left.  Whether a replacement is bad depends only on its own value, so the
    // final multiset of energies is: the good main draws, the good ones among the next N0 replacement draws
    // (N0 = bad main draws), the good ones among the next N1 (N1 = bad ones among those), ...  No ranks needed.
    long long extra_pos = 0;
    double sum_e0_last = 0.0;
    bool exhausted = false;
    for (long long loop = 0; loop < m.n_loops; ++loop) {
        const double *src = run.z + loop * m.n_ev_per_loop;
        long long count = m.n_ev_per_loop;
        double loop_sum = 0.0;
        while (count > 0) {
            long long nbad = 0;
            double part = 0.0;
            for (long long d = tid; d < count; d += NT) {
                const double z = __ldg(src + d);
                double E = __dsub_rn(beamE, __dadd_rn(__dmul_rn(exp(__dmul_rn(sshape, z)), scale), eLoss));
                if (E <= 0.0) {
                    ++nbad;
                } else if (E == E) {
                    part += E;
                    double x_prev = m.ode_from_zero ? 0.0 : sx[0];
                    for (int i = 0; i < X; ++i) {
                        if (i > 0 || m.ode_from_zero) {
                            const double h = (sx[i] - x_prev) / (double)m.ode_substeps;
                            double Ev[1] = {E};
                            for (int ss = 0; ss < m.ode_substeps; ++ss) rk4_step<1, 0>(Ev, h, m.bethe_A, m.bethe_B, m.n_materials);
                            E = Ev[0];
                            x_prev = sx[i];
                        }
                        const int b = np_bin(E, EB, m.e_min, m.e_max, e_step, e_scale);       // simultFit.py:264
                        if (b >= 0) atomicAdd(Hmine + i * EB + b, xs_eval(E, xs));           // simultFit.py:263
                    }
                }
            }
            const long long nbad_tot = block_sum<long long>(nbad, reinterpret_cast<long long *>(scratch));
            loop_sum += block_sum<double>(part, scratch);
            if (nbad_tot == 0) break;
            if (extra_pos + nbad_tot > run.n_z1) {          // replacement stream exhausted
                exhausted = true;
                break;
            }
            src = run.z1 + extra_pos;
            extra_pos += nbad_tot;
            count = nbad_tot;
        }
        if (exhausted) break;
        if (loop == m.n_loops - 1) sum_e0_last = loop_sum;   // e0mean uses the LAST loop only (simultFit.py:282)
    }
    __syncthreads();
    double *H = Hw;                                           // fold the per-warp copies into copy 0
    for (int c = tid; c < CELLS; c += NT) {
        double v = Hw[c];
        for (int k = 1; k < NW; ++k) v += Hw[(size_t)k * CELLS + c];
        H[c] = v;
    }
    __syncthreads();

    simult_tail<NT>(m, run, r, w, out, H, tofh, pdf, sx, svd, staps, scratch, sum_e0_last, sf, exhausted);
}

// Ascending bitonic sort of n <= cap doubles in shared memory (cap a power of two, tail padded with +inf).
template <int NT>
__device__ __forceinline__ void smem_sort(double *a, int n, int cap) {
    for (int i = n + threadIdx.x; i < cap; i += NT) a[i] = CUDART_INF;
    __syncthreads();
    for (int k = 2; k <= cap; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap; i += NT) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__host__ __device__ inline size_t simult_range_smem_bytes(int X, int E, int T, int rng_n, int P, int n_taps, int lut_n) {
    size_t d = (size_t)X * E + 2 * (size_t)T + RANGE_TILE + (size_t)rng_n * (P + 3) + X + E + n_taps + 48 + X;
    return d * 8 + (((size_t)lut_n * 2 + 15) / 16) * 16 + SIMULT_ULUT * 2 + (((size_t)X * 4 + 15) / 16) * 16 + (size_t)rng_n * 8 + 32;
}

// Range-table formulation of the simultaneous fit: same model as simult_run_kernel, stopping through T1/T2.
template <int NT, int P>
__global__ void __launch_bounds__(NT) simult_range_kernel(const DevModel m, const DevRunSet runs, const double *__restrict__ theta,
                                                          long long n_walkers, ModelOut out, int only_run) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RW = P + 3;
    const int n_launch_runs = (only_run >= 0) ? 1 : m.n_runs;
    const long long w = blockIdx.x / n_launch_runs;
    const int r = (only_run >= 0) ? only_run : (int)(blockIdx.x % n_launch_runs);
    if (w >= n_walkers) return;
    const DevRun &run = runs.r[r];
    const int T = run.tof_bins, X = m.x_bins, EB = m.e_bins, CELLS = X * EB, M = m.rng_n;
    const int tid = threadIdx.x;

    double *H = reinterpret_cast<double *>(smem_raw);            // [CELLS]
    double *tofh = H + CELLS;                                    // [T]
    double *pdf = tofh + T;                                      // [T]
    double *u0 = pdf + T;                                        // [RANGE_TILE]
    double *rec = u0 + RANGE_TILE;                               // [M][RW]
    double *sx = rec + (size_t)M * RW;                           // [X]
    double *svd = sx + X;                                        // [E]
    double *staps = svd + EB;
    double *scratch = staps + m.n_taps;                          // [48]
    double *sdelta = scratch + 48;                               // [X]
    unsigned short *lut = reinterpret_cast<unsigned short *>(sdelta + X);
    unsigned short *ulut = lut + ((m.rng_lut_n + 7) / 8) * 8;    // [SIMULT_ULUT]
    int *srow = reinterpret_cast<int *>(ulut + SIMULT_ULUT);     // [X]
    double *sbrk = reinterpret_cast<double *>(srow + X + (X & 1) + 2);   // [M] interval ends

    const double *th = theta + w * m.ndim;
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = th[p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside && out.spectra == nullptr && out.cells == nullptr) return;
    const double beamE = th[0], eLoss = th[1], scale = th[2], sshape = th[3], sf = th[4 + r];

    const double x_start = m.ode_from_zero ? 0.0 : m.x_centers[0];
    for (int i = tid; i < CELLS; i += NT) H[i] = 0.0;
    for (int i = tid; i < T; i += NT) tofh[i] = 0.0;
    for (int i = tid; i < X; i += NT) {
        sx[i] = m.x_centers[i];
        sdelta[i] = m.rng_sign * (m.x_centers[i] - x_start);
    }
    for (int i = tid; i < M * RW; i += NT) rec[i] = m.rng_rec[i];
    for (int j = tid; j < M; j += NT) sbrk[j] = m.rng_rec[(size_t)j * RW];
    for (int i = tid; i < m.rng_lut_n; i += NT) lut[i] = m.rng_lut[i];
    for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
    __syncthreads();

    long long extra_pos = 0;
    double sum_e0_last = 0.0;
    bool exhausted = false;
    int bin_lo_all = EB, bin_hi_all = -1;
    for (long long loop = 0; loop < m.n_loops; ++loop) {
        const double *src = run.z + loop * m.n_ev_per_loop;     // sorted by the library: E0 ascending
        long long count = m.n_ev_per_loop;
        bool sorted = true;
        double loop_sum = 0.0;
        while (count > 0) {
            long long nbad = 0;
            double part = 0.0;
            for (long long tile = 0; tile < count; tile += RANGE_TILE) {
                const int nt = (int)((count - tile < RANGE_TILE) ? (count - tile) : RANGE_TILE);
                __syncthreads();
                for (int d = tid; d < nt; d += NT) {
                    const double z = __ldg(src + tile + d);
                    const double E = __dsub_rn(beamE, __dadd_rn(__dmul_rn(exp(__dmul_rn(sshape, z)), scale), eLoss));
                    double u = -CUDART_INF;                      // redrawn (E <= 0) or NaN: contributes nothing
                    if (E <= 0.0) {
                        ++nbad;
                    } else if (E == E) {
                        part += E;
                        u = t1_eval(E, m);
                    }
                    u0[d] = u;
                }
                __syncthreads();
                if (!sorted) {
                    int cap = 1;
                    while (cap < nt) cap <<= 1;
                    smem_sort<NT>(u0, nt, cap);
                }
                range_accumulate_tile<NT, P>(u0, nt, sbrk, rec, 0, lut, ulut, SIMULT_ULUT, sdelta, srow, H, EB, nullptr, X, M, m.rng_u_max,
                                             m.rng_lut_inv, m.rng_lut_n, bin_lo_all, bin_hi_all);
            }
            const long long nbad_tot = block_sum<long long>(nbad, reinterpret_cast<long long *>(scratch));
            loop_sum += block_sum<double>(part, scratch);
            if (nbad_tot == 0) break;
            if (extra_pos + nbad_tot > run.n_z1) {
                exhausted = true;
                break;
            }
            src = run.z1 + extra_pos;                            // replacement draws: arbitrary order
            extra_pos += nbad_tot;
            count = nbad_tot;
            sorted = false;
        }
        if (exhausted) break;
        if (loop == m.n_loops - 1) sum_e0_last = loop_sum;
    }
    __syncthreads();
    simult_tail<NT>(m, run, r, w, out, H, tofh, pdf, sx, svd, staps, scratch, sum_e0_last, sf, exhausted);
}

// lnprob = lnprior + sum of the per-run log-likelihoods in run order (simultFit.py:412-420, 444-469).
__global__ void simult_finish_kernel(const DevModel m, const double *__restrict__ theta, long long n_walkers,
                                     const double *__restrict__ partial, double *__restrict__ lnprob) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = theta[w * m.ndim + p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    double r = -CUDART_INF;
    if (inside) {
        r = 0.0;
        for (int k = 0; k < m.n_runs; ++k) r += partial[w * m.n_runs + k];
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;                                    // simultFit.py:463-468
    }
    lnprob[w] = r;
}

}  // namespace tof

namespace tof {

// ================================================================================================
// oneBD production model: tests/csi_oneBD.py:415-521 (model), 543-649 (likelihood)
// ================================================================================================
// numpy's legacy Poisson sampler (numpy/random/src/distributions/distributions.c: random_poisson_mult /
// random_poisson_ptrs / random_loggam) on an explicit uniform stream -- np.random.poisson(bgLevel, T) at
// csi_oneBD.py:521.  Sequential by construction: one thread draws the T values of a run.
__device__ inline double np_loggam(double x) {
    const double a[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04, -5.952380952380952e-04,
                          8.417508417508418e-04, -1.917526917526918e-03, 6.410256410256410e-03, -2.955065359477124e-02,
                          1.796443723688307e-01, -1.39243221690590e+00};
    if (x == 1.0 || x == 2.0) return 0.0;
    const long long n = (x < 7.0) ? (long long)(7 - x) : 0;
    double x0 = x + (double)n;
    const double x2 = __dmul_rn(1.0 / x0, 1.0 / x0);
    double gl0 = a[9];
    for (int k = 8; k >= 0; --k) gl0 = __dadd_rn(__dmul_rn(gl0, x2), a[k]);
    double gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * log(x0) - x0;
    if (x < 7.0)
        for (long long k = 1; k <= n; ++k) {
            gl -= log(x0 - 1.0);
            x0 -= 1.0;
        }
    return gl;
}

// Returns false when the uniform stream runs out.
__device__ inline bool np_poisson(double lam, const double *u, long long n_u, long long &pos, double &out) {
    if (lam == 0.0) {
        out = 0.0;
        return true;
    }
    if (lam >= 10.0) {
        const double slam = sqrt(lam), loglam = log(lam);
        const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
        const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
        while (true) {
            if (pos + 2 > n_u) return false;
            const double U = u[pos] - 0.5, V = u[pos + 1];
            pos += 2;
            const double us = 0.5 - fabs(U);
            const double k = floor((2 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) { out = k; return true; }
            if (k < 0 || (us < 0.013 && V > us)) continue;
            if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - np_loggam(k + 1))) {
                out = k;
                return true;
            }
        }
    }
    const double enlam = exp(-lam);
    double X = 0.0, prod = 1.0;
    while (true) {
        if (pos + 1 > n_u) return false;
        prod *= u[pos++];
        if (prod > enlam) X += 1.0; else break;
    }
    out = X;
    return true;
}

__host__ __device__ inline size_t onebd_smem_bytes(int NT, int X, int E, int T, int n_xs, int n_taps, int n_taps2, int stop_n,
                                                   int lut_n) {
    size_t d = (size_t)(NT / 32) * X * E + 4 * (size_t)T + X + E + n_xs + (size_t)(n_xs - 1) * 4 + n_taps + n_taps2 +
               (size_t)X * (stop_n - 1) * 4 + X + 48;
    return d * 8 + (((size_t)lut_n + 15) / 16) * 16;
}

template <int NT>
__global__ void __launch_bounds__(NT) onebd_run_kernel(const DevModel m, const DevRunSet runs, const double *__restrict__ theta,
                                                       long long n_walkers, ModelOut out, int only_run) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const int n_launch_runs = (only_run >= 0) ? 1 : m.n_runs;
    const long long w = blockIdx.x / n_launch_runs;
    const int r = (only_run >= 0) ? only_run : (int)(blockIdx.x % n_launch_runs);
    if (w >= n_walkers) return;
    const DevRun &run = runs.r[r];
    const int T = run.tof_bins, X = m.x_bins, EB = m.e_bins, CELLS = X * EB, NS = m.stop_n - 1;
    const int tid = threadIdx.x, warp = tid >> 5;

    double *Hw = reinterpret_cast<double *>(smem_raw);          // [NW][CELLS]
    double *tofh = Hw + (size_t)NW * CELLS;                      // [T]
    double *pdf = tofh + T;                                      // [T]
    double *c1 = pdf + T;                                        // [T] after the causal transit convolution
    double *bg = c1 + T;                                         // [T] Poisson background realisation
    double *sx = bg + T;                                         // [X]
    double *svd = sx + X;                                        // [E]
    double *xs_bp = svd + EB;
    double *xs_cf = xs_bp + m.n_xs;
    double *staps = xs_cf + (size_t)(m.n_xs - 1) * 4;
    double *staps2 = staps + m.n_taps;
    double *sstop = staps2 + m.n_taps2;                          // [X][NS][4]
    double *satt = sstop + (size_t)X * NS * 4;                   // [X]
    double *scratch = satt + X;                                  // [48]
    unsigned char *xs_lut = reinterpret_cast<unsigned char *>(scratch + 48);

    const double *th = theta + w * m.ndim;
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = th[p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside && out.spectra == nullptr && out.cells == nullptr) return;
    // csi_oneBD.py:581: [eLoss, scale, s, scaleFactor_r, bgLevel_r]
    const double eLoss = th[0], scale = th[1], sshape = th[2], sf = th[3 + r], bg_level = th[m.ndim - m.n_runs + r];

    for (int i = tid; i < NW * CELLS; i += NT) Hw[i] = 0.0;
    for (int i = tid; i < T; i += NT) tofh[i] = 0.0;
    for (int i = tid; i < X; i += NT) {
        sx[i] = m.x_centers[i];
        satt[i] = m.attenuation[i];
    }
    for (int i = tid; i < X * NS * 4; i += NT) sstop[i] = m.stop_coefs[i];
    for (int i = tid; i < m.n_xs; i += NT) xs_bp[i] = m.xs_breaks[i];
    for (int i = tid; i < (m.n_xs - 1) * 4; i += NT) xs_cf[i] = m.xs_coefs[i];
    for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
    for (int i = tid; i < m.n_taps2; i += NT) staps2[i] = m.taps2[i];
    for (int i = tid; i < m.xs_lut_n; i += NT) xs_lut[i] = m.xs_lut[i];
    __syncthreads();
    XsTab xs;
    xs.bp = xs_bp; xs.cf = xs_cf; xs.lut = xs_lut; xs.n = m.n_xs; xs.lut_n = m.xs_lut_n;
    xs.lut_lo = m.xs_lut_lo; xs.lut_inv = m.xs_lut_inv;
    const double e_step = (m.e_max - m.e_min) / (double)EB;
    const double e_scale = (double)EB / (m.e_max - m.e_min);
    double *Hmine = Hw + (size_t)warp * CELLS;

    // ---- last loop only: the script ASSIGNS dataHist[idx,:] = hist (csi_oneBD.py:465), so earlier loops are
    //      overwritten, and e0mean is the mean of the last eZeros (489).  No redraw of E0 <= 0 here (440-447). ----
    const double *z = run.z + (m.n_loops - 1) * m.n_ev_per_loop;
    const double stop_hi = m.stop_lo + m.stop_step * (double)(m.stop_n - 1);
    const double inv_step = 1.0 / m.stop_step;
    double part = 0.0;
    for (long long d = tid; d < m.n_ev_per_loop; d += NT) {
        const double E0 = __dsub_rn(m.beam_energy, __dadd_rn(__dmul_rn(exp(__dmul_rn(sshape, __ldg(z + d))), scale), eLoss));
        part += E0;
        // betheApprox.evalStopped (ionStopping.py:132-136): FITPACK clamps the argument to the grid
        double a = E0 < m.stop_lo ? m.stop_lo : (E0 > stop_hi ? stop_hi : E0);
        if (!(a == a)) continue;
        int k = (int)((a - m.stop_lo) * inv_step);
        k = k < 0 ? 0 : (k > NS - 1 ? NS - 1 : k);
        const double dx0 = a - (m.stop_lo + m.stop_step * (double)k);
        for (int i = 0; i < X; ++i) {
            const double *c = sstop + ((size_t)i * NS + k) * 4;
            const double E = ((c[0] * dx0 + c[1]) * dx0 + c[2]) * dx0 + c[3];
            const int b = np_bin(E, EB, m.e_min, m.e_max, e_step, e_scale);                  // csi_oneBD.py:463
            if (b >= 0) atomicAdd(Hmine + i * EB + b, __dmul_rn(xs_eval(E, xs), satt[i]));  // csi_oneBD.py:462
        }
    }
    const double sum_e0 = block_sum<double>(part, scratch);
    const double e0mean = __ddiv_rn(sum_e0, (double)m.n_ev_per_loop);
    double *H = Hw;
    for (int c = tid; c < CELLS; c += NT) {
        double v = Hw[c];
        for (int k = 1; k < NW; ++k) v += Hw[(size_t)k * CELLS + c];
        H[c] = v;
    }
    for (int j = tid; j < EB; j += NT) {
        const double eff = __ddiv_rn(__dadd_rn(e0mean, m.e_centers[j]), 2.0);               // csi_oneBD.py:499
        svd[j] = speed_of(m.c, eff, m.m_d);
    }
    // background realisation: np.random.poisson(bgLevel, T) (csi_oneBD.py:521), one thread, in bin order
    bool bg_ok = true;
    if (tid == 0) {
        long long pos = 0;
        for (int t = 0; t < T; ++t) {
            double k = 0.0;
            if (!np_poisson(bg_level, run.z1, run.n_z1, pos, k)) {
                bg_ok = false;
                k = CUDART_NAN;
            }
            bg[t] = k;
        }
    }
    __syncthreads();

    // ---- cells (no normalisation: drawHist2d = rint(dataHist * nSamples), csi_oneBD.py:490) -> flight times ----
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    const double nsamp = (double)m.n_samples;
    for (int idx = tid; idx < CELLS; idx += NT) {
        const double cnt = rint(__dmul_rn(H[idx], nsamp));
        if (out.cells) out.cells[(size_t)w * CELLS + idx] = (cnt == cnt) ? (long long)cnt : LLONG_MIN;
        if (cnt != 0.0 && cnt == cnt) {
            const int i = idx / EB, j = idx - i * EB;
            const double tof_d = __ddiv_rn(sx[i], svd[j]);
            const double tof_n = __ddiv_rn(__ldg(run.neutron_dist + i), __ldg(m.neutron_speed + j));
            const int b = np_bin(__dadd_rn(tof_d, tof_n), T, run.tof_min, run.tof_max, t_step, t_scale);
            if (b >= 0) atomicAdd(tofh + b, cnt);            // integer-valued doubles: exact in any order
        }
    }
    __syncthreads();
    double tpart = 0.0;
    for (int t = tid; t < T; t += NT) tpart += tofh[t];
    const double total = block_sum<double>(tpart, scratch);
    for (int t = tid; t < T; t += NT) {
        const double db = __dsub_rn(np_edge(t + 1, T, run.tof_min, run.tof_max, t_step),
                                    np_edge(t, T, run.tof_min, run.tof_max, t_step));
        pdf[t] = __ddiv_rn(__ddiv_rn(tofh[t], db), total);   // NaN when nothing landed in the window, like numpy
    }
    __syncthreads();
    // causal transit-time smearing: np.convolve(pdf, taps2, 'full')[:T] (csi_oneBD.py:519)
    for (int t = tid; t < T; t += NT) {
        double acc = 0.0;
        for (int k = 0; k < m.n_taps2; ++k)
            if (t - k >= 0) acc += staps2[k] * pdf[t - k];
        c1[t] = acc;
    }
    __syncthreads();
    double lp = 0.0;
    for (int t = tid; t < T; t += NT) {
        double acc = 0.0;
        for (int k = 0; k < m.n_taps; ++k) {
            const int tt = t + m.conv_shift - k;
            if (tt >= 0 && tt < T) acc += staps[k] * c1[tt];
        }
        double ev = __dadd_rn(__dmul_rn(sf, acc), bg[t]);                                   // csi_oneBD.py:521
        if (out.spectra) {
            const double v = out.stage == TOF_STAGE_COUNTS ? tofh[t] : (out.stage == TOF_STAGE_PDF ? pdf[t] : ev);
            out.spectra[(size_t)w * T + t] = v;
        }
        if (ev != ev) {
            lp += -CUDART_INF;                                                              // csi_oneBD.py:554-555
        } else {
            const double o = run.obs ? run.obs[t] : 1.0;
            if (ev == 0.0) ev = 1.0;
            double poi = -o - lgamma(trunc(ev) + 1.0);
            if (ev > 0.0) poi += ev * log(o);
            lp += o * poi;
        }
    }
    lp = block_sum<double>(lp, scratch);
    if (tid == 0 && out.lnprob) out.lnprob[w * m.n_runs + r] = lp;
    (void)bg_ok;
}

}  // namespace tof
