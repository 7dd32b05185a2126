// adv / intermediate model, RK4 formulation (TOF_ODE_RK4): tests/advIntermediateTOFmodel.py:115-199.
#pragma once
#include "tof_common.cuh"

namespace tof {

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
            const double zd = (d < m.n_draws) ? (run.fresh ? fresh_normal(run, w, 0, d) : __ldg(run.z + d)) : 0.0;
            E[k] = (d < m.n_draws) ? __dadd_rn(e0, __dmul_rn(spread, zd)) : CUDART_NAN;
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
        if (r != r && out.nan_count) atomicAdd(out.nan_count, 1ull);
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
}

}  // namespace tof
