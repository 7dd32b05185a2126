// oneBD production model: tests/csi_oneBD.py, and its posterior-predictive twin utilities/ppcTools_oneBD.py.
#pragma once
#include "simult_model.cuh"

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

// Returns false when the uniform stream runs out.  With per-evaluation draws (run.fresh) the uniforms are this (call,
// walker, run)'s own Philox stream and never run out.
__device__ inline bool np_poisson(double lam, const double *u_bound, long long n_bound, long long &pos, double &out, const DevRun &run,
                                  long long walker, int run_idx) {
    const long long n_u = run.fresh ? (1ll << 62) : n_bound;
    auto u = [&](long long p) { return run.fresh ? fresh_uniform(run, walker, run_idx, p) : u_bound[p]; };
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
            const double U = u(pos) - 0.5, V = u(pos + 1);
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
        prod *= u(pos++);
        if (prod > enlam) X += 1.0; else break;
    }
    out = X;
    return true;
}

// `copies`: private (x,E) histogram copies (one per group of warps; NT/32 when shared memory allows, fewer for the big
// grid of the posterior-predictive variant: 20 x 400 cells)
__host__ __device__ inline size_t onebd_smem_bytes(int copies, int X, int E, int T, int n_xs, int n_taps, int n_taps2, int stop_n,
                                                   int lut_n) {
    size_t d = (size_t)copies * X * E + 4 * (size_t)T + X + E + n_xs + (size_t)(n_xs - 1) * 4 + n_taps + n_taps2 +
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

    const int NC = m.onebd_copies;                               // histogram copies, 1..NW
    double *Hw = reinterpret_cast<double *>(smem_raw);          // [NC][CELLS]
    double *tofh = Hw + (size_t)NC * CELLS;                      // [T]
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

    for (int i = tid; i < NC * CELLS; i += NT) Hw[i] = 0.0;
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
    double *Hmine = Hw + (size_t)(warp % NC) * CELLS;

    // ---- last loop only: the script ASSIGNS dataHist[idx,:] = hist (csi_oneBD.py:465), so earlier loops are
    //      overwritten, and e0mean is the mean of the last eZeros (489).  No redraw of E0 <= 0 here (440-447). ----
    const double *z = run.z + (m.n_loops - 1) * m.n_ev_per_loop;
    const double stop_hi = m.stop_lo + m.stop_step * (double)(m.stop_n - 1);
    const double inv_step = 1.0 / m.stop_step;
    double part = 0.0;
    for (long long d = tid; d < m.n_ev_per_loop; d += NT) {
        const double zd = run.fresh ? fresh_normal(run, w, r, (m.n_loops - 1) * m.n_ev_per_loop + d) : __ldg(z + d);
        const double E0 = __dsub_rn(m.beam_energy, __dadd_rn(__dmul_rn(exp(__dmul_rn(sshape, zd)), scale), eLoss));
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
            if (b < 0) continue;
            if (out.unweighted) atomicAdd(Hmine + i * EB + b, 1.0);                          // ppcTools_oneBD.py:223-224 eD_atEachX
            else atomicAdd(Hmine + i * EB + b, __dmul_rn(xs_eval(E, xs), satt[i]));          // csi_oneBD.py:462
        }
    }
    const double sum_e0 = block_sum<double>(part, scratch);
    const double e0mean = __ddiv_rn(sum_e0, (double)m.n_ev_per_loop);
    double *H = Hw;
    for (int c = tid; c < CELLS; c += NT) {
        double v = Hw[c];
        for (int k = 1; k < NC; ++k) v += Hw[(size_t)k * CELLS + c];
        H[c] = v;
    }
    if (out.unweighted) {                                    // raw per-cell counts of the last loop, nothing else
        __syncthreads();
        if (out.cells)
            for (int c = tid; c < CELLS; c += NT) out.cells[(size_t)w * CELLS + c] = (long long)H[c];
        return;
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
            if (!np_poisson(bg_level, run.z1, run.n_z1, pos, k, run, w, r)) {
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
    const int NZ = m.n_zero_deg;
    for (int idx = tid; idx < CELLS; idx += NT) {
        const double cnt = rint(__dmul_rn(H[idx], nsamp));
        if (out.cells) out.cells[(size_t)w * CELLS + idx] = (cnt == cnt) ? (long long)cnt : LLONG_MIN;
        if (cnt != 0.0 && cnt == cnt) {
            const int i = idx / EB, j = idx - i * EB;
            const double tof_d = __ddiv_rn(sx[i], svd[j]);
            const double tof_n = __ddiv_rn(__ldg(run.neutron_dist + i), __ldg(m.neutron_speed + j));
            const double base = __dadd_rn(tof_d, tof_n);
            if (NZ == 0) {
                const int b = np_bin(base, T, run.tof_min, run.tof_max, t_step, t_scale);
                if (b >= 0) atomicAdd(tofh + b, cnt);        // integer-valued doubles: exact in any order
            } else {                                         // ppcTools_oneBD.py:246-248: 10 zero-degree sub-times per cell
                for (int k = 0; k < NZ; ++k) {
                    const double tof = __dadd_rn(base, __ldg(m.zd_times + j * NZ + k));
                    const int b = np_bin(tof, T, run.tof_min, run.tof_max, t_step, t_scale);
                    if (b >= 0) atomicAdd(tofh + b, __dmul_rn(cnt, __ldg(m.zd_weights + j * NZ + k)));
                }
            }
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
