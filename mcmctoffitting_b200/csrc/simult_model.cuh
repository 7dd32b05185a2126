// simultaneous multi-standoff fit (config 4): tests/simultFit.py.
#pragma once
#include "adv_range.cuh"

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
    if (out.unweighted) {                                   // raw per-cell counts of the last loop, nothing else
        if (out.cells)
            for (int c = tid; c < CELLS; c += NT) out.cells[(size_t)w * CELLS + c] = (long long)H[c];
        return;
    }
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
        long long fresh_base = loop * m.n_ev_per_loop;       // per-evaluation draws (tof_set_draw_mode): index into the stream
        int fresh_stream = 0;
        long long count = m.n_ev_per_loop;
        double loop_sum = 0.0;
        while (count > 0) {
            long long nbad = 0;
            double part = 0.0;
            for (long long d = tid; d < count; d += NT) {
                const double z = run.fresh ? fresh_normal(run, w, r, fresh_base + d, fresh_stream) : __ldg(src + d);
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
                        if (out.unweighted) {                                                 // ppcTools.py:151-154 eD_atEachX
                            if (b >= 0 && loop == m.n_loops - 1) atomicAdd(Hmine + i * EB + b, 1.0);
                        } else if (b >= 0) {
                            atomicAdd(Hmine + i * EB + b, xs_eval(E, xs));                    // simultFit.py:263
                        }
                    }
                }
            }
            const long long nbad_tot = block_sum<long long>(nbad, reinterpret_cast<long long *>(scratch));
            loop_sum += block_sum<double>(part, scratch);
            if (nbad_tot == 0) break;
            if (!run.fresh && extra_pos + nbad_tot > run.n_z1) {   // replacement stream exhausted
                exhausted = true;
                break;
            }
            src = run.z1 + extra_pos;
            fresh_base = extra_pos;
            fresh_stream = 3;
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

// draws staged at a time by simult_range_kernel: with ten rows a warp packs 10 rows x 3 sub-slices, 24 sub-slices per row
// over the CTA's eight warps -- 2048 draws give every lane a run of ~85 samples (1024: the per-tile searches cost as much
// as the polynomial work)
constexpr int SIMULT_TILE = 2048;

// Phase 1 for one tile of a BIG sorted draw set (the simultaneous fit draws 50 000 per loop): the tile is a narrow slice
// of the energy distribution, a row's slice of it lies in one E-bin unless a bin edge falls inside, so runs are long and
// the polynomial loop is nearly all there is.  Same scheme as adv_zrank_multi_kernel (adv_zrank.cuh): lane = row (with
// fewer than 32 rows: R rows x (32/R) sub-slices per warp), every lane of a warp sums the first segment of its slice in
// lockstep, leftover segments are summed one after the other by all 32 lanes; partial sums of a cell meet with atomics.
// Membership rule RN(u0[d] + delta) >= edge as everywhere.  Full-width histogram H[X][EB], all T2 records staged.
// Called by all threads of the CTA; no barrier inside.
template <int NT, int P>
__device__ __forceinline__ void simult_tile_long(const double *u0, int nt, const double *brk, const double *rec, const unsigned short *lut,
                                                 const double *sdelta, double *H, int EB, int X, int M, double umax, double lut_inv,
                                                 int lut_n) {
    constexpr int RW = P + 3;
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // finite part of the sorted tile: -inf (redrawn / NaN) first, +inf last
    int v_lo = 0, v_hi = nt;
    if (!(u0[0] > -CUDART_INF) || u0[nt - 1] >= CUDART_INF) {
        int lo = 0, hi = nt;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (u0[mid] > -CUDART_INF) hi = mid;
            else lo = mid + 1;
        }
        v_lo = lo;
        hi = nt;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (u0[mid] >= CUDART_INF) hi = mid;
            else lo = mid + 1;
        }
        v_hi = lo;
    }
    if (v_hi <= v_lo) return;                              // uniform
    const int Gf = X >> 5, R = X & 31;
    const int wB = R ? (Gf ? 1 : NW) : 0, wA = NW - wB;
    int row, parts, part;
    bool lane_ok = true;
    if (warp < wA) {
        const int g = warp % Gf, q = warp / Gf;
        row = (g << 5) + lane;
        parts = (wA - g + Gf - 1) / Gf;
        part = q;
        if (g >= wA) return;                               // (more groups of rows than warps: not a simultFit shape)
    } else {
        const int per_b = 32 / R, sub = lane / R;
        row = (Gf << 5) + (lane - sub * R);
        parts = wB * per_b;
        part = (warp - wA) * per_b + sub;
        lane_ok = sub < per_b;
    }
    const double delta = sdelta[row];
    const int hbase = row * EB;
    const double umax_next = __longlong_as_double(__double_as_longlong(umax) + 1);
    const unsigned u0_s32 = (unsigned)__cvta_generic_to_shared(u0);
    auto edge_of = [&](int j) -> double { return j == 0 ? 0.0 : (j >= M ? umax_next : brk[j - 1]); };
    auto first_ge = [&](int lo, int hi, double edge, double dl) -> int {
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__dadd_rn(u0[mid], dl) >= edge) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    auto load_record = [&](int j, double (&a)[P + 1], double &a0, int &bin) {
        const double2 *r2 = reinterpret_cast<const double2 *>(rec + j * RW);
        bin = __double2loint(r2[0].y);
        const double2 c01 = r2[1];
        a0 = c01.x;
        a[0] = 0.0;
        a[1] = c01.y;
#pragma unroll
        for (int k = 2; k <= P; k += 2) {
            const double2 c2 = r2[1 + (k >> 1)];
            a[k] = c2.x;
            a[k + 1] = c2.y;
        }
    };
    const int nf = v_hi - v_lo;
    int s = v_lo + (part * nf) / parts, sB = v_lo + ((part + 1) * nf) / parts;
    if (!lane_ok) sB = s;
    if (s < sB && !(__dadd_rn(u0[s], delta) >= 0.0)) s = first_ge(s, sB, 0.0, delta);   // below the histogram range
    int j = 0;
    if (s < sB) {
        const double v = __dadd_rn(u0[s], delta);
        if (v > umax) s = sB;
        else j = range_interval(v, brk, lut, lut_inv, lut_n, M);
    }
    {   // main pass: the first segment of every lane's slice, in lockstep
        const bool act = s < sB;
        int s1 = sB;
        if (act && __dadd_rn(u0[sB - 1], delta) >= edge_of(j + 1)) s1 = first_ge(s, sB, edge_of(j + 1), delta);
        const int n = act ? s1 - s : 0;
        const int nmax = __reduce_max_sync(FULL, n);
        if (nmax > 0) {
            const int nmin = __reduce_min_sync(FULL, n);
            double a[P + 1], a0;
            int bin;
            load_record(act ? j : 0, a, a0, bin);
            const double off = delta - (act ? edge_of(j) : 0.0);
            double acc = 0.0;
            unsigned addr = u0_s32 + (unsigned)s * 8u;
            const int tfull = nmin >> 2;
#pragma unroll 1
            for (int t = tfull; t > 0; --t) {
                poly_full4<P>(acc, addr, off, a);
                addr += 32u;
            }
            int rem = n - (tfull << 2);
#pragma unroll 1
            for (int t = ((nmax + 3) >> 2) - tfull; t > 0; --t) {
                poly_run4<P>(acc, addr, rem, off, a);
                addr += 32u;
                rem -= 4;
            }
            TOF_CHECK(n == 0 || (bin >= 0 && bin < EB && s >= v_lo && s1 <= v_hi));
            if (n > 0) atomicAdd(H + hbase + bin, fma((double)n, a0, acc));
        }
        s = s1;
        ++j;
    }
    for (;;) {   // leftover segments (an interval edge inside the slice), one at a time, all 32 lanes on each
        const bool more = s < sB && j < M;
        const unsigned pending = __ballot_sync(FULL, more);
        if (pending == 0u) break;
        const int src = __ffs(pending) - 1;
        int s1 = 0;
        if (lane == src) s1 = first_ge(s, sB, edge_of(j + 1), delta);
        const int seg_s = __shfl_sync(FULL, s, src), seg_e = __shfl_sync(FULL, s1, src), seg_j = __shfl_sync(FULL, j, src);
        const double seg_delta = __shfl_sync(FULL, delta, src);
        const int seg_h = __shfl_sync(FULL, hbase, src);
        const int seg_n = seg_e - seg_s;
        if (seg_n > 0) {
            double a[P + 1], a0;
            int bin;
            load_record(seg_j, a, a0, bin);
            const double off = seg_delta - edge_of(seg_j);
            const int ch = (seg_n + 31) >> 5;
            const int la = seg_s + lane * ch;
            int n = seg_e - la;
            n = n < 0 ? 0 : (n > ch ? ch : n);
            double acc = 0.0;
            unsigned addr = u0_s32 + (unsigned)(la < seg_e ? la : seg_s) * 8u;
            int rem = n;
#pragma unroll 1
            for (int t = (ch + 3) >> 2; t > 0; --t) {
                poly_run4<P>(acc, addr, rem, off, a);
                addr += 32u;
                rem -= 4;
            }
            acc = fma((double)n, a0, acc);
            acc = warp_sum(acc);
            TOF_CHECK(bin >= 0 && bin < EB && seg_s >= v_lo && seg_e <= v_hi);
            if (lane == 0) atomicAdd(H + seg_h + bin, acc);
        }
        if (lane == src) {
            s = s1;
            ++j;
        }
    }
}

__host__ __device__ inline size_t simult_range_smem_bytes(int X, int E, int T, int rng_n, int P, int n_taps, int lut_n) {
    size_t d = (size_t)X * E + 2 * (size_t)T + SIMULT_TILE + (size_t)rng_n * (P + 3) + X + E + n_taps + 48 + X;
    return d * 8 + (((size_t)lut_n * 2 + 15) / 16) * 16 + SIMULT_ULUT * 2 + (((size_t)X * 4 + 15) / 16) * 16 + (size_t)rng_n * 8 + 32;
}

// Range-table formulation of the simultaneous fit: same model as simult_run_kernel, stopping through T1/T2.
template <int NT, int P>
__global__ void __launch_bounds__(NT, 4) simult_range_kernel(const DevModel m, const DevRunSet runs, const double *__restrict__ theta,
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
    double *u0 = pdf + T;                                        // [SIMULT_TILE]
    double *rec = u0 + SIMULT_TILE;                              // [M][RW]
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
            for (long long tile = 0; tile < count; tile += SIMULT_TILE) {
                const int nt = (int)((count - tile < SIMULT_TILE) ? (count - tile) : SIMULT_TILE);
                __syncthreads();
                // the tile's draws are fetched first (independent loads in flight), then transformed
                constexpr int ZPT = SIMULT_TILE / NT;
                double zz[ZPT];
#pragma unroll
                for (int q = 0; q < ZPT; ++q) zz[q] = (tid + q * NT < nt) ? __ldg(src + tile + tid + q * NT) : 0.0;
#pragma unroll
                for (int q = 0; q < ZPT; ++q) {
                    const int d = tid + q * NT;
                    if (d >= nt) break;
                    const double z = zz[q];
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
                // long runs (a tile of a 50 000-draw loop is a narrow energy slice): the lean routine; short tiles of
                // replacement draws and split E-bins go through the general one
                if (nt >= 256 && X <= 32 * (NT / 32))
                    simult_tile_long<NT, P>(u0, nt, sbrk, rec, lut, sdelta, H, EB, X, M, m.rng_u_max, m.rng_lut_inv, m.rng_lut_n);
                else
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
                                     const double *__restrict__ partial, double *__restrict__ lnprob,
                                     unsigned long long *nan_count) {
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
        if (r != r && nan_count) atomicAdd(nan_count, 1ull);                               // the dump the reference prints, as a counter
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;                                    // simultFit.py:463-468
    }
    lnprob[w] = r;
}

}  // namespace tof
