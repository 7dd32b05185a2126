// adv / intermediate model, range-table formulation: the kernel for "many walkers, one tile of draws each".
//
// Same model, tables, shared-memory layout and arithmetic as adv_range_kernel (adv_range.cuh); what differs is how the
// code is cut.  adv_range_kernel is one function: every phase of a walker is inlined into the persistent loop, the
// compiler keeps the union of their live values in 64 registers, spills, re-derives pointers inside hot loops and
// serialises the polynomial's Horner chains.  Here the hot loop -- range_exec_cells -- is a function of its own
// (__noinline__ on purpose: a call is a register-allocation firewall) and the few per-walker scalars cross phases through
// a small frame in shared memory, so that it is allocated almost alone and ptxas keeps four Horner chains in flight per
// lane (dependent DFMA latency on B200: 8.8 cycles, issue interval 2.2: tools/dfma_latency.cu).  The set-up and finish
// phases are inlined into the kernel: out of line they read the kernel parameters through generic pointers (LD.E
// instead of constant-bank operands), which cost adv_zrank_kernel 8 % (measured there).
//
// Used for: FP64, n_draws <= RANGE_TILE, one T2 interval per E-bin (rng_identity), production output (lnprob only).
// Everything else (debug spectra / cell counts, big or multi-tile draw sets, FP32 mode, draw splits, split E-bins) and
// the walkers whose E-band does not fit the banded layout run through adv_range_kernel as before.
#pragma once
#include "adv_range.cuh"

namespace tof {

// per-walker scalars handed from phase to phase (static shared memory)
struct PlannedFrame {
    long long next;          // work item fetched by thread 0
    long long w;             // walker index
    double e0;
    int hstride, jbase;
    int band[3];             // widest row, first / last interval of the walker
    long long t_mark;        // stage timing (PROF)
};

enum { PLANNED_DONE = 0, PLANNED_SKIP = 1, PLANNED_RUN = 2 };

template <bool PROF>
__device__ __forceinline__ void planned_stage_done(PlannedFrame *f, unsigned long long *stage_cycles, int k) {
    if constexpr (PROF) {
        if (threadIdx.x == 0) {
            const long long t = clock64();
            atomicAdd(stage_cycles + k, (unsigned long long)(t - f->t_mark));
            f->t_mark = t;
        }
    }
}

// Work fetch, prior, per-row E-band, record staging, histogram reset, energy-loss lookup of the draws (adv:128-129).
// Returns PLANNED_DONE when the work counter is exhausted, PLANNED_SKIP when the walker needs nothing more (outside the
// prior: -inf written; band too wide: queued for the full-size launch), PLANNED_RUN with the frame filled and the tile
// of u0 values staged otherwise.  Uniform over the CTA; ends with a barrier.
template <int NT, int P, bool PROF>
__device__ __forceinline__ int planned_setup(const DevModel *mp, const DevRun *rp, const double *__restrict__ theta, long long n_walkers,
                                          const ModelOut *op, unsigned char *smem_raw, PlannedFrame *f) {
    __builtin_assume(__isShared(smem_raw));
    __builtin_assume(__isShared(f));
    const DevModel &m = *mp;
    const DevRun &run = *rp;
    const ModelOut &out = *op;
    constexpr int RW = P + 3;
    const int tid = threadIdx.x;
    const int X = m.x_bins, M = m.rng_n, T = run.tof_bins;
    double *H = reinterpret_cast<double *>(smem_raw);
    double *u0 = reinterpret_cast<double *>(smem_raw + out.lay.pa);
    double *rec = reinterpret_cast<double *>(smem_raw + out.lay.rec);
    const double *sdelta = reinterpret_cast<const double *>(smem_raw + out.lay.sdelta);
    const unsigned short *lut = reinterpret_cast<const unsigned short *>(smem_raw + out.lay.lut);
    int *hlo_s = reinterpret_cast<int *>(smem_raw + out.lay.hlo);
    const double *sbrk = reinterpret_cast<const double *>(smem_raw + out.lay.sbrk);
    __syncthreads();                                       // the previous walker is done with shared memory
    if (tid == 0) {
        f->next = (long long)atomicAdd(out.work, 1ull);
        f->band[0] = 0;
        f->band[1] = M;
        f->band[2] = -1;
    }
    __syncthreads();
    const long long w = f->next;
    if (w >= n_walkers) return PLANNED_DONE;
    const double e0 = theta[w * m.ndim + 0];
    const double sigma0 = theta[w * m.ndim + 1];
    bool inside = true;
    for (int p = 0; p < m.ndim; ++p) {
        const double v = theta[w * m.ndim + p];
        inside = inside && (m.prior_strict ? (m.prior_lo[p] < v && v < m.prior_hi[p])
                                           : !(v < m.prior_lo[p] || v > m.prior_hi[p]));
    }
    if (!inside) {                                         // adv:191-195: the model is never evaluated outside the prior
        if (tid == 0) out.lnprob[w] = -CUDART_INF;
        return PLANNED_SKIP;
    }
    const double spread = __dmul_rn(sigma0, e0);          // adv:128
    const bool rev = spread < 0.0;                         // draws are sorted ascending: E0 ascends unless the spread is negative
    const double umax = m.rng_u_max;
    // E-bins the walker can touch: the draws are sorted, first and last give the extremes
    double u_lo, u_hi;
    const int nt = (int)m.n_draws;                         // one tile
    if (run.fresh) {
        // per-evaluation draws (tof_set_draw_mode): this walker's own sorted normals, generated in place of the tile;
        // the sign of the spread does not matter (the normal law is symmetric): E0 ascends with |spread|
        double *scr = reinterpret_cast<double *>(smem_raw + out.lay.scratch);
        fresh_sorted_normals<NT>(u0, nt, run, w, 0, scr);
        const double sp = fabs(spread);
        double za = 0.0, zb = 0.0;
        if (tid < nt) za = u0[tid];
        if (tid + NT < nt) zb = u0[tid + NT];
        __syncthreads();
        if (tid < nt) u0[tid] = t1_eval(__dadd_rn(e0, __dmul_rn(sp, za)), m);
        if (tid + NT < nt) u0[tid + NT] = t1_eval(__dadd_rn(e0, __dmul_rn(sp, zb)), m);
        __syncthreads();
        u_lo = u0[0];
        u_hi = u0[nt - 1];
    } else {
        u_lo = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 : 0)))), m);
        u_hi = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? 0 : m.n_draws - 1)))), m);
    }
    // every row has its own window of E-bins: [u_lo + delta_i, u_hi + delta_i], one interval of slack on both sides
    // (T1 is only monotone up to its 2e-13 cm fit error); interval j == E-bin j on this path
    for (int i = tid; i < X; i += NT) {
        double vmin = u_lo > -CUDART_INF ? u_lo + sdelta[i] : 0.0;     // -inf draws: the lowest in-range v is 0 (whatever the sign of delta)
        double vmax = u_hi + sdelta[i];
        vmin = vmin > 0.0 ? vmin : 0.0;
        vmax = vmax < umax ? vmax : umax;
        int j_lo = 0, j_hi = 0;
        if (vmax >= vmin) {                               // otherwise this row gets nothing: any window will do
            j_lo = range_interval(vmin, sbrk, lut, m.rng_lut_inv, m.rng_lut_n, M);
            j_hi = range_interval(vmax, sbrk, lut, m.rng_lut_inv, m.rng_lut_n, M);
            j_lo = j_lo > 0 ? j_lo - 1 : 0;
            j_hi = j_hi < M - 1 ? j_hi + 1 : M - 1;
            atomicMin(&f->band[1], j_lo);
            atomicMax(&f->band[2], j_hi);
        }
        hlo_s[i] = j_lo;
        atomicMax(&f->band[0], j_hi - j_lo + 1);
    }
    __syncthreads();
    const int hstride = f->band[0];
    const int j_lo_all = f->band[2] >= 0 ? f->band[1] : 0, j_hi_all = f->band[2] >= 0 ? f->band[2] : 0;
    const int jbase = j_lo_all > 0 ? j_lo_all - 1 : 0;
    const bool fits = (long long)X * hstride <= out.hcap && (j_hi_all - jbase + 1) <= out.rcap && T <= out.hcap;
    if (!fits) {                                           // queue for the full-size launch
        if (tid == 0) out.queue_out[atomicAdd(out.queue_count, 1ull)] = (int)w;
        return PLANNED_SKIP;
    }
    TOF_CHECK(j_hi_all - jbase + 1 <= out.rcap && j_hi_all < M && jbase >= 0 && X * hstride <= out.hcap);
    const double *recg = m.rng_rec;
    for (int i = tid; i < (j_hi_all - jbase + 1) * RW; i += NT) rec[i] = recg[(size_t)jbase * RW + i];
    for (int i = tid; i < X * hstride; i += NT) H[i] = 0.0;
    if (!run.fresh)
        for (int d = tid; d < nt; d += NT)
            u0[d] = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? nt - 1 - d : d)))), m);
    if (tid == 0) {
        f->w = w;
        f->e0 = e0;
        f->hstride = hstride;
        f->jbase = jbase;
    }
    __syncthreads();
    planned_stage_done<PROF>(f, out.stage_cycles, 0);
    return PLANNED_RUN;
}

// Phases 2-5 of a walker: normalise (adv:143), np.rint + flight-time scatter (adv:146-159), density, timing response at
// the observed bins and log-likelihood (adv:160-181).  Same arithmetic, in the same order, as adv_range_kernel.
template <int NT, int P, bool PROF>
__device__ __forceinline__ void planned_finish(const DevModel *mp, const DevRun *rp, const ModelOut *op, unsigned char *smem_raw,
                                            PlannedFrame *f) {
    __builtin_assume(__isShared(smem_raw));
    __builtin_assume(__isShared(f));
    const DevModel &m = *mp;
    const DevRun &run = *rp;
    const ModelOut &out = *op;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X = m.x_bins, EB = m.e_bins, T = run.tof_bins;
    double *H = reinterpret_cast<double *>(smem_raw);
    unsigned int *tofc = reinterpret_cast<unsigned int *>(smem_raw + out.lay.pa);
    double *svd = reinterpret_cast<double *>(smem_raw + out.lay.svd);
    const double *staps = reinterpret_cast<const double *>(smem_raw + out.lay.staps);
    double *scratch = reinterpret_cast<double *>(smem_raw + out.lay.scratch);
    double *rvd = reinterpret_cast<double *>(smem_raw + out.lay.ulut);   // [EB] 1/svd: the draw lookup is dead now
    const int *hlo = reinterpret_cast<const int *>(smem_raw + out.lay.hlo);
    const long long w = f->w;
    const double e0 = f->e0;
    const int hstride = f->hstride;
    planned_stage_done<PROF>(f, out.stage_cycles, 1);

    // ---- phase 2: normalise (adv:143) ---------------------------------------------------------------------
    for (int i = tid; i < T; i += NT) tofc[i] = 0u;        // the draw tile is dead now
    for (int j = tid; j < EB; j += NT) {                   // deuteron speeds and their reciprocals
        const double eff = __ddiv_rn(__dadd_rn(e0, m.e_centers[j]), 2.0);   // adv:151
        const double v = speed_of(m.c, eff, m.m_d);
        svd[j] = v;
        rvd[j] = __ddiv_rn(1.0, v);
    }
    const double de = (m.e_max - m.e_min) / (double)EB;
    const double dx = (m.x_max - m.x_min) / (double)X;
    double part = 0.0;
    for (int row = warp; row < X; row += NW) {
        const double *Hr = H + (size_t)row * hstride;
        for (int jb = lane; jb < hstride; jb += 32) part += __dmul_rn(__dmul_rn(Hr[jb], de), dx);
    }
    const double S = block_sum<double>(part, scratch);     // includes the barrier that publishes tofc = 0
    planned_stage_done<PROF>(f, out.stage_cycles, 2);

    // ---- phase 3: quantise (adv:146) and scatter every non-empty cell to its flight time (adv:149-158) ----
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    const double nsamp = (double)m.n_samples;
    const double rS = __ddiv_rn(1.0, S);                    // IEEE quotients below come from this reciprocal (div_by_recip)
    if (S > 0.0) {
        for (int row = warp; row < X; row += NW) {
            const double xi = __ldg(m.x_centers + row), di = __ldg(run.neutron_dist + row);
            const int row_lo = hlo[row];
            const double *Hr = H + (size_t)row * hstride;
            for (int jb = lane; jb < hstride; jb += 32) {
                const int j = row_lo + jb;
                if (j >= EB) break;
                const double h = Hr[jb];
                if (h != 0.0) {
                    const double cnt = rint(__dmul_rn(div_by_recip(h, S, rS), nsamp));
                    if (cnt > 0.0) {
                        const double tof_d = div_by_recip(xi, svd[j], rvd[j]);
                        const double tof_n = div_by_recip(di, __ldg(m.neutron_speed + j), __ldg(m.neutron_rspeed + j));
                        const int b = np_bin(__dadd_rn(tof_d, tof_n), T, run.tof_min, run.tof_max, t_step, t_scale);
                        TOF_CHECK(b < T && j < EB);
                        if (b >= 0) atomicAdd(tofc + b, (unsigned int)cnt);
                    }
                }
            }
        }
    }
    __syncthreads();
    planned_stage_done<PROF>(f, out.stage_cycles, 3);

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
    if (tid == 0) {
        double r = degenerate ? CUDART_NAN : lp;
        if (r != r && out.nan_count) atomicAdd(out.nan_count, 1ull);
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
    planned_stage_done<PROF>(f, out.stage_cycles, 4);
    if (PROF && tid == 0) atomicAdd(out.stage_cycles + TOF_N_STAGES, 1ull);
}

template <int NT, int P, bool PROF = false>
__global__ void __launch_bounds__(NT, 2) adv_planned_kernel(const __grid_constant__ DevModel m, const __grid_constant__ DevRun run,
                                                            const double *__restrict__ theta, long long n_walkers,
                                                            const __grid_constant__ ModelOut out) {
    extern __shared__ __align__(16) unsigned char smem_sym[];
    unsigned char *smem_raw = smem_sym;
    asm volatile("" : "+l"(smem_raw));                     // opaque base, still known to be shared (see adv_range_kernel)
    __builtin_assume(__isShared(smem_raw));
    __shared__ PlannedFrame frame;
    constexpr int RW = P + 3;
    const int tid = threadIdx.x;
    const int X = m.x_bins, M = m.rng_n;
    // ---- walker-independent tables: staged once per CTA (persistent CTAs loop over walkers) ------------------
    {
        double *staps = reinterpret_cast<double *>(smem_raw + out.lay.staps);
        double *sdelta = reinterpret_cast<double *>(smem_raw + out.lay.sdelta);
        unsigned short *lut = reinterpret_cast<unsigned short *>(smem_raw + out.lay.lut);
        double *sbrk = reinterpret_cast<double *>(smem_raw + out.lay.sbrk);
        const double *recg = m.rng_rec;
        for (int j = tid; j < M; j += NT) sbrk[j] = recg[(size_t)j * RW];
        for (int i = tid; i < m.rng_lut_n; i += NT) lut[i] = m.rng_lut[i];
        for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
        const double x_start = m.ode_from_zero ? 0.0 : m.x_centers[0];
        for (int i = tid; i < X; i += NT) sdelta[i] = m.rng_sign * (m.x_centers[i] - x_start);
        if (PROF && tid == 0) frame.t_mark = clock64();
    }
    for (;;) {
        const int st = planned_setup<NT, P, PROF>(&m, &run, theta, n_walkers, &out, smem_raw, &frame);
        if (st == PLANNED_DONE) break;
        if (st == PLANNED_SKIP) continue;
        // plan: per cell (first draw, count) into H; with one tile the band it returns is the walker's
        int band_lo = M, band_hi = -1;
        const bool any = range_tile_planned<NT, P, false>(
            reinterpret_cast<const double *>(smem_raw + out.lay.pa), (int)m.n_draws,
            reinterpret_cast<const double *>(smem_raw + out.lay.sbrk), reinterpret_cast<const double *>(smem_raw + out.lay.rec),
            frame.jbase, reinterpret_cast<const unsigned short *>(smem_raw + out.lay.lut),
            reinterpret_cast<unsigned short *>(smem_raw + out.lay.ulut), RANGE_ULUT,
            reinterpret_cast<const double *>(smem_raw + out.lay.sdelta), reinterpret_cast<int *>(smem_raw + out.lay.srow),
            reinterpret_cast<double *>(smem_raw), frame.hstride, reinterpret_cast<const int *>(smem_raw + out.lay.hlo), X, M,
            m.rng_u_max, m.rng_lut_inv, m.rng_lut_n, band_lo, band_hi);
        // execute: called from here, the leanest frame there is, so that the polynomial loop gets the registers
        if (any)
            range_exec_cells<NT, P>(reinterpret_cast<const double *>(smem_raw + out.lay.pa),
                                    reinterpret_cast<const double *>(smem_raw + out.lay.sbrk),
                                    reinterpret_cast<const double *>(smem_raw + out.lay.rec), frame.jbase,
                                    reinterpret_cast<const double *>(smem_raw + out.lay.sdelta),
                                    reinterpret_cast<const int *>(smem_raw + out.lay.srow), reinterpret_cast<double *>(smem_raw),
                                    frame.hstride, reinterpret_cast<const int *>(smem_raw + out.lay.hlo), X, band_lo, band_hi);
        __syncthreads();
        planned_finish<NT, P, PROF>(&m, &run, &out, smem_raw, &frame);
    }
}

}  // namespace tof
