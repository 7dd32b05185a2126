// simple model (config 1): tests/simpleTOFmodel.py:57-120.
#pragma once
#include "tof_common.cuh"

namespace tof {

// ================================================================================================
// simple model: tests/simpleTOFmodel.py:57-120  (every sample is histogrammed directly)
// ================================================================================================
// grid = (walkers, chunks).  counts[n][T] (u64, zeroed by the caller) accumulate across chunks.
template <int NT>
__global__ void __launch_bounds__(NT) simple_hist_kernel(const DevModel m, const DevRun run, const double *__restrict__ theta,
                                                         long long n_walkers, unsigned long long *__restrict__ counts,
                                                         int ignore_prior) {
    __shared__ unsigned int sh[1024];
    const int T = run.tof_bins;
    const long long w = blockIdx.x;
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

    const long long per = (m.n_draws + gridDim.y - 1) / gridDim.y;
    const long long lo = (long long)blockIdx.y * per;
    const long long hi = (lo + per < m.n_draws) ? lo + per : m.n_draws;
    for (long long d = lo + tid; d < hi; d += NT) {
        // draws: the bound streams, or this (call, walker)'s own (tof_set_draw_mode; the reference draws per call)
        const double ud = run.fresh ? fresh_uniform(run, w, 0, d) : __ldg(run.z1 + d);
        const double zd = run.fresh ? fresh_normal(run, w, 0, d) : __ldg(run.z + d);
        const double x = __dmul_rn(m.cell_length, ud);                                      // simple:62
        const double ed = __dadd_rn(__dadd_rn(e0, __dmul_rn(e1, x)), __dmul_rn(sigma, zd));  // simple:64
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
        if (r != r && out.nan_count) atomicAdd(out.nan_count, 1ull);
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
}

}  // namespace tof
