// adv / intermediate model, range-table formulation: the shipped kernel for "many walkers, one tile of draws each".
//
// Same model, tables and arithmetic as adv_range_kernel / adv_planned_kernel (adv_range.cuh, adv_planned.cuh): a
// (draw, row) sample is v = u0_d + delta_i, the cell (row, E-bin j) sums the degree-7 weight polynomial of interval j
// over the contiguous run of sorted draws whose v lies in [brk[j-1], brk[j]), membership decided by the same compare
// RN(u0[d] + delta) >= edge.  What is new is how a run is FOUND and what happens after it is summed:
//
//  * rank hints instead of a per-walker plan.  "First draw with u(E0_d) + delta_i >= U_j" is "first draw with
//    E0_d >= Theta[i][j]", Theta[i][j] = u^-1(U_j - delta_i) -- the initial energy that reaches row i with the lower
//    edge energy of E-bin j.  Theta does not depend on the walker; and because E0_d = e0 + spread * z_d with the SAME
//    sorted normals z_d for every walker, the rank of a threshold is a lookup in a walker-independent table over z
//    (built once per draw set: zlut[c] = first draw with z >= z_lo + c/z_inv).  One float FMA turns Theta into the
//    lookup cell (per-walker constants a = z_inv/spread, b), taken a quarter cell low so that the exact answer is
//    reached by a short forward walk with the usual compare.  The hint only has to be low and close; the membership
//    rule is untouched, so the cells are the ones adv_range_kernel / adv_planned_kernel produce, bit for bit.
//    This removes the per-walker lookup build, the plan pass (18 % of the instructions of adv_planned_kernel), its
//    barrier and the slot hand-over through the histogram.  In memory Theta is interval-major (a warp's lanes = 32
//    rows at nearby intervals share cache lines; row-major was measured 17 % slower), with pad rows around it and, for
//    cells of up to ZR_TPITCH rows, a compile-time pitch: the three thresholds of a visit are one pointer + immediates.
//  * the normalisation sum S = sum(H * dE * dx) (adv:143) is accumulated by the lane that produces a cell (no second
//    pass over the histogram), and every cell of a row's window is written, zero or not (no histogram reset).
//  * scatter (adv:146-159): rint(H/S * N) and the TOF bin of a cell are first formed from reciprocals (two products);
//    unless the value sits within 1e-6 of a rounding / bin boundary that IS the correctly rounded reference result,
//    otherwise the cell takes the exact path (div_by_recip quotients + numpy's edge rule) -- same integers always.
//  * the density (np.histogram density=True) is only evaluated at the bins the timing response of a non-zero observed
//    bin reads.
//  * walkers whose E-band does not fit the banded histogram (sigma0 >~ 0.2) no longer wait for a second, full-size
//    launch at 1 CTA/SM: the same CTA keeps their cell sums in an L2-resident scratch histogram (WIDE instantiation
//    of the same phase functions) -- one launch per call, no serial tail.
//  * the phases are __forceinline__ functions of one kernel body: out of line they would receive the kernel parameters
//    by pointer and read every field with a generic load (measured: 8 % slower, with a stack frame for the calls).
//    What keeps the polynomial loop's registers free instead: per-walker scalars are re-read from shared memory where
//    they are used -- from a small mirror at fixed offsets below the draw tile (ZR_MIRROR), whose address is live anyway.
//
// Used for: FP64, n_draws <= RANGE_TILE, one T2 interval per E-bin (rng_identity), production output (lnprob only),
// bound draws.  The second kernel of this file, adv_zrank_multi_kernel, serves bigger draw sets (the reference's own
// 1e5 - 1e6 draws per evaluation) with the same phase functions after the cell sums.  Debug outputs, the FP32 mode,
// split E-bins and per-evaluation draws stay with adv_range_kernel / adv_planned_kernel.
#pragma once
#include "adv_planned.cuh"
#ifdef TOF_ZR_DEBUG
#include <cstdio>
#endif

namespace tof {

constexpr int ZR_LUT = 8192;             // cells of the draw-rank lookup over z (entries: ZR_LUT + 1); 4096: -0.5 %, 16384: same
constexpr float ZR_BIAS = 0.25f;         // cells the hint is lowered by (float rounding of the hint << 0.25 cell)
// keV; below this (and for reversed / degenerate spreads) hints are off.  Error budget of a hint in lookup cells, worst case
// at spread = 8 keV with thresholds near 2600 keV: float rounding of Theta (1.6e-4 keV = 0.025 cells), of b (|b| ~ 4e5:
// 0.016 cells) and of the FMA (0.016) -- together < 0.06 cells against the 0.25-cell bias.
constexpr double ZR_MIN_SPREAD = 8.0;
// Scalars the cell sums re-read at every use (they do not fit in registers next to four Horner chains) sit right below
// the draw tile, whose shared-memory address is live in a register anyway: one LDS with an immediate offset each.
constexpr int ZR_TPITCH = 128;           // floats per row of the threshold table when the cell has at most that many rows
constexpr int ZR_MIRROR = 32;            // bytes: de f64 @-32 | dx f64 @-24 | jbase s32 @-16 | hint_a, hint_b f32 @-8

// Byte offsets of the regions of adv_zrank_kernel's dynamic shared memory (host-computed).  Order:
//   H [hcap] f64 (later the density [T]) | ZR_MIRROR bytes | draw tile u0 [RANGE_TILE + 1] f64, later the TOF counters [T] u32 |
//   staged T2 records [rcap][P+3] f64, later deuteron speeds, their reciprocals and 1/neutron speed [3][E] | taps |
//   40 doubles of scratch |
//   delta [X] | srow [X] int | hlo [X] int | 0.0 | interval ends [M] f64 (the last one replaced by nextafter(u_max):
//   range_interval never reads it, the edge lookup of the cell sums does)
inline RangeLayout zrank_layout(int X, int E, int T, int hcap, int rcap, int P, int n_taps, int rng_n) {
    RangeLayout L{};
    size_t region_a = (size_t)T * 4 > (size_t)(RANGE_TILE + 1) * 8 ? (size_t)T * 4 : (size_t)(RANGE_TILE + 1) * 8;   // + the +inf sentinel behind the draws
    region_a = (region_a + 15) / 16 * 16;
    size_t rec_b = (size_t)rcap * (P + 3) * 8;
    rec_b = rec_b > (size_t)3 * E * 8 ? rec_b : (size_t)3 * E * 8;
    size_t o = ((size_t)hcap * 8 + 15) / 16 * 16;         // what follows is read with 16-byte loads
    o += ZR_MIRROR;                                       // per-walker scalars of the cell sums, at fixed offsets below u0
    L.pa = (unsigned)o;        o += region_a;
    L.rec = (unsigned)o;
    L.svd = (unsigned)o;                          // aliases the records (dead after the cell sums)
    L.ulut = (unsigned)(o + (size_t)E * 8);       // 1/speed [E], followed by 1/neutron speed [E]
    o += (rec_b + 15) / 16 * 16;
    L.staps = (unsigned)o;     o += (size_t)n_taps * 8;
    L.scratch = (unsigned)o;   o += 40 * 8;
    L.sdelta = (unsigned)o;    o += (size_t)X * 8;
    L.srow = (unsigned)o;      o += (size_t)(X + (X & 1)) * 4;   // (odd X: keeps what follows 8-byte aligned)
    L.hlo = (unsigned)o;       o += (size_t)(X + (X & 1)) * 4;
    o += 8;                                       // sbrk[-1] = 0: sbrk[j - 1] is the lower edge of interval j for every j in [0, M]
    L.sbrk = (unsigned)o;      o += (size_t)rng_n * 8;
    L.lut = L.sbin = 0;
    L.total = (unsigned)(o + 16);
    return L;
}

// per-walker scalars handed from phase to phase (static shared memory)
struct ZrFrame {
    long long idx[2];        // work items: idx[it & 1] is this iteration's walker, the other slot receives the next one
    long long w;             // walker index
    double e0;
    int hstride, jbase;
    int band[3];             // widest row window, first / last interval of the walker
    int wide;                // 1: cell sums live in the CTA's global scratch histogram, records are read from global
    float hint_a, hint_b;    // lookup cell of a threshold energy: Theta * a + b
    int s_ref, k_lo, n_iv, nB, wB;   // trajectory-aligned visit grid of the cell sums (see zr_exec)
    double spread;           // sigma0 * e0 (multi-tile kernel: the draws are staged tile by tile)
    int split;               // multi-tile kernel with few walkers: which share of the tiles this CTA sums
    int last;                // ... whether this CTA arrived last and finishes the walker
    double de, dx;           // bin widths of the (x,E) histogram (set once per CTA)
    const unsigned short *zlut;   // draw-rank lookup of this run (set once per CTA)
    long long t_mark;        // stage timing (PROF)
};

// Sum over the CTA with ONE barrier: warp partials to `slot`, then every warp adds them with the same xor butterfly
// (fixed order: the result is bit-identical in all threads and from run to run).  `slot` (>= NT/32 elements) must not be
// written again before another barrier: callers alternate between two slots.
template <typename T>
__device__ __forceinline__ T block_sum1(T v, T *slot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) slot[warp] = v;
    __syncthreads();
    T t = (lane < nw) ? slot[lane] : T(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    return t;
}

__device__ __forceinline__ float ldg_stream_f32(const float *p) {   // read-only, do not allocate in L1
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Work fetch, prior, per-row E-window, record staging, energy-loss lookup of the draws (adv:128-129), hint constants.
// Returns PLANNED_DONE / PLANNED_SKIP (outside the prior: -inf written) / PLANNED_RUN.  Uniform; ends with a barrier.
template <int NT, int P, bool PROF>
__device__ __forceinline__ int zr_setup(const DevModel *mp, const DevRun *rp, const double *__restrict__ theta, long long n_walkers,
                                     const ModelOut *op, unsigned char *smem_raw, ZrFrame *f, int it) {
    __builtin_assume(__isShared(smem_raw));
    __builtin_assume(__isShared(f));
    const DevModel &m = *mp;
    const DevRun &run = *rp;
    const ModelOut &out = *op;
    constexpr int RW = P + 3;
    const int tid = threadIdx.x;
    const int X = m.x_bins, M = m.rng_n, T = run.tof_bins;
    double *u0 = reinterpret_cast<double *>(smem_raw + out.lay.pa);
    double *rec = reinterpret_cast<double *>(smem_raw + out.lay.rec);
    const double *sdelta = reinterpret_cast<const double *>(smem_raw + out.lay.sdelta);
    int *srow = reinterpret_cast<int *>(smem_raw + out.lay.srow);
    int *hlo_s = reinterpret_cast<int *>(smem_raw + out.lay.hlo);
    const double *sbrk = reinterpret_cast<const double *>(smem_raw + out.lay.sbrk);
    if (tid == 0) {                                        // (every thread read the previous walker's band at the top of zr_finish)
        f->band[0] = 0;
        f->band[1] = M;
        f->band[2] = -1;
    }
    __syncthreads();                                       // the previous walker is done with shared memory
    const long long w = f->idx[it & 1];
    if (w >= n_walkers) return PLANNED_DONE;
    // the next work item is fetched now (a global atomic: ~1 us) and read after the next iteration's first barrier
    // (measured: dealing the first 85 % of the walkers out round-robin, without the atomic and with the next parameter
    // vector copied in by cp.async, is 1.7 % slower -- the CTAs drift into lockstep phases)
    if (tid == 0) f->idx[(it + 1) & 1] = (long long)atomicAdd(out.work, 1ull);
    // the draws do not depend on the walker: fetched before the parameter vector is looked at, so that both loads fly together
    const int nt = (int)m.n_draws;                         // one tile
    constexpr int DPT = (1024 + NT - 1) / NT;              // draws per thread (zrank_layout caps the tile at 1024 draws)
    double zf[DPT], zb[DPT];                               // z[d] and z[nt - 1 - d]: the order depends on the sign of the spread
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
        const int d = tid + q * NT;
        zf[q] = d < nt ? __ldg(run.z + d) : 0.0;
        zb[q] = d < nt ? __ldg(run.z + (nt - 1 - d)) : 0.0;
    }
    const double z_first = __ldg(run.z), z_last = __ldg(run.z + (nt - 1)), z_mid_f = __ldg(run.z + (nt >> 1)),
                 z_mid_b = __ldg(run.z + (nt - 1 - (nt >> 1)));
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
    // energy-loss lookup of every draw (adv:128-129): u0[d] = u(e0 + spread * z_d), ascending
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
        const int d = tid + q * NT;
        if (d < nt) u0[d] = t1_eval(__dadd_rn(e0, __dmul_rn(spread, rev ? zb[q] : zf[q])), m);
    }
    if (tid == 0) u0[nt] = CUDART_INF;                     // sentinel: the forward walks of the cell sums stop here
    // E-bins the walker can touch: the draws are sorted, first and last give the extremes.  The row threads evaluate
    // these three draws themselves (the same expression as the staged copy: identical bits) instead of waiting for
    // the staged array behind a barrier.
    // every row has its own window of E-bins: [u_lo + delta_i, u_hi + delta_i], one interval of slack on both sides
    // (T1 is only monotone up to its 2e-13 cm fit error); interval j == E-bin j on this path.  srow: interval of the
    // median draw -- rows are processed along the trajectory so that the lanes of a warp have runs of similar length.
    double u_lo = 0.0, u_hi = 0.0, u_med = 0.0;
    if (tid < X) {
        u_lo = t1_eval(__dadd_rn(e0, __dmul_rn(spread, rev ? z_last : z_first)), m);
        u_hi = t1_eval(__dadd_rn(e0, __dmul_rn(spread, rev ? z_first : z_last)), m);
        u_med = t1_eval(__dadd_rn(e0, __dmul_rn(spread, rev ? z_mid_b : z_mid_f)), m);
    }
    for (int i = tid; i < X; i += NT) {
        const double dl = sdelta[i];
        double vmin = u_lo > -CUDART_INF ? u_lo + dl : 0.0;     // -inf draws: the lowest in-range v is 0 (whatever the sign of delta)
        double vmax = u_hi + dl;
        vmin = vmin > 0.0 ? vmin : 0.0;
        vmax = vmax < umax ? vmax : umax;
        double vm = __dadd_rn(u_med, dl);
        vm = vm < 0.0 ? 0.0 : (vm > umax ? umax : vm);    // NaN (median draw outside the table) -> any interval
        // the three table lookups of the row fly together (range_interval = lookup + refinement against the breaks)
        const int g_lo = range_lut_guess(vmin, m.rng_lut, m.rng_lut_inv, m.rng_lut_n);
        const int g_hi = range_lut_guess(vmax, m.rng_lut, m.rng_lut_inv, m.rng_lut_n);
        const int g_md = range_lut_guess(vm, m.rng_lut, m.rng_lut_inv, m.rng_lut_n);
        int j_lo = 0, j_hi = 0;
        if (vmax >= vmin) {                               // otherwise this row gets nothing: any window will do
            j_lo = range_refine(vmin, g_lo, sbrk, M);
            j_hi = range_refine(vmax, g_hi, sbrk, M);
            j_lo = j_lo > 0 ? j_lo - 1 : 0;
            j_hi = j_hi < M - 1 ? j_hi + 1 : M - 1;
            atomicMin(&f->band[1], j_lo);
            atomicMax(&f->band[2], j_hi);
        }
        hlo_s[i] = j_lo;
        atomicMax(&f->band[0], j_hi - j_lo + 1);
        srow[i] = (vm == vm) ? range_refine(vm, g_md, sbrk, M) : 0;
    }
    if (tid == NT - 1) {                                   // (a thread of the last warp: the row loop keeps the first ones busy)
        // rank hint of a threshold energy Th: cell = ((Th - e0)/spread - z_lo) * z_inv - bias
        float ha = 0.0f, hb = 0.0f;
        if (spread >= ZR_MIN_SPREAD && spread < 1e30 && run.zlut != nullptr) {
            const double inv = 1.0 / spread;
            ha = (float)(run.zlut_inv * inv);
            hb = (float)((-e0 * inv - run.zlut_lo) * run.zlut_inv - (double)ZR_BIAS);
            if (!(ha < 1e6f) || !(fabsf(hb) < 1e9f)) ha = hb = 0.0f;   // a degenerate draw set: no hints (walk from draw 0)
        }
        f->hint_a = ha;
        f->hint_b = hb;
        reinterpret_cast<float *>(smem_raw + out.lay.pa - 8)[0] = ha;   // ZR_MIRROR
        reinterpret_cast<float *>(smem_raw + out.lay.pa - 8)[1] = hb;
    }
    __syncthreads();
    const int hstride = f->band[0];
    const int j_lo_all = f->band[2] >= 0 ? f->band[1] : 0, j_hi_all = f->band[2] >= 0 ? f->band[2] : 0;
    const int jbase = j_lo_all > 0 ? j_lo_all - 1 : 0;
    const bool fits = (long long)X * hstride <= out.hcap && (j_hi_all - jbase + 1) <= out.rcap;
    TOF_CHECK(T <= out.hcap && hstride <= M);
    if (fits) {
        const double *recg = m.rng_rec;
        for (int i = tid; i < (j_hi_all - jbase + 1) * RW; i += NT) rec[i] = recg[(size_t)jbase * RW + i];
    }
    if (tid == 0) {
        f->w = w;
        f->e0 = e0;
        f->hstride = hstride;
        f->jbase = fits ? jbase : 0;
        f->wide = fits ? 0 : 1;
        if (f->band[2] < 0) {                             // no row can be reached: nothing to visit
            f->band[1] = 0;
            f->band[2] = -1;
        }
        *reinterpret_cast<int *>(smem_raw + out.lay.pa - 16) = fits ? jbase : 0;   // ZR_MIRROR
        // visit grid: rows are walked along the trajectory, interval j = k + (srow[row] - srow[0]); the shift is
        // monotone in the row index.  Leftover rows (X % 32): R rows x (32/R) offsets per visit, on wB warps of their own
        // (in proportion to their share of the visits, at least one when there are any).
        {
            constexpr int NW = NT / 32;
            const int Gf = X >> 5, R = X & 31;
            const int s_ref = srow[0], s_b = srow[X - 1] - s_ref;
            const int s_min = s_b < 0 ? s_b : 0, s_max = s_b < 0 ? 0 : s_b;
            const int jl = f->band[1], jh = f->band[2];
            const int k_lo = jl - s_max;
            const int n_iv = jh >= jl ? (jh - s_min) - k_lo + 1 : 0;
            const int per_b = R ? 32 / R : 1;
            const int nB = R ? (n_iv + per_b - 1) / per_b : 0;
            int wB = 0;
            if (R) {
                const int den = n_iv * Gf + nB;
                wB = (Gf && den > 0) ? (NW * nB + den / 2) / den : NW;
                wB = wB < 1 ? 1 : (wB > NW - 1 && Gf ? NW - 1 : wB);
            }
            f->s_ref = s_ref;
            f->k_lo = k_lo;
            f->n_iv = n_iv;
            f->nB = nB;
            f->wB = wB;
        }
        if (!fits && out.queue_count) atomicAdd(out.queue_count, 1ull);   // statistics: walkers kept in the scratch histogram
    }
    __syncthreads();
    return PLANNED_RUN;
}

// shared-memory accessors on 32-bit shared addresses (the phase functions receive generic pointers; spelling the
// address space out keeps record loads and histogram stores LDS / STS without 64-bit address arithmetic)
__device__ __forceinline__ double2 zr_lds_v2(unsigned a) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double zr_lds(unsigned a) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void zr_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// The (x,E) histogram of cross-section weights (adv:128-138) for one walker, lane = row, one (row, interval) cell per
// lane and visit.  Returns this thread's share of sum(H * dE * dx) (adv:143).  WIDE: H and rec are global pointers.
// Work split as in range_exec_cells: a warp keeps ONE group of 32 rows and walks the trajectory-aligned interval
// offsets of that group with a stride; the X % 32 leftover rows get warps of their own that pack R rows x (32/R)
// offsets per visit.  The rank hints of the next visit are fetched (L2) while the current one is summed.
template <int NT, int P, bool WIDE, int TP>
__device__ __forceinline__ double zr_exec(const DevModel *mp, const DevRun *rp, const ModelOut *op, unsigned char *smem_raw,
                                       const ZrFrame *f, double *Hglobal) {
    __builtin_assume(__isShared(smem_raw));
    __builtin_assume(__isShared(f));
    const DevModel &m = *mp;
    constexpr int RW = P + 3;
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int X = m.x_bins, M = m.rng_n;
    const double *u0 = reinterpret_cast<const double *>(smem_raw + op->lay.pa);
    const double *brk = reinterpret_cast<const double *>(smem_raw + op->lay.sbrk);
    const double *sdelta = reinterpret_cast<const double *>(smem_raw + op->lay.sdelta);
    const int *srow = reinterpret_cast<const int *>(smem_raw + op->lay.srow);
    const int *hlo = reinterpret_cast<const int *>(smem_raw + op->lay.hlo);
    const double *rec_g = m.rng_rec;                       // WIDE: records straight from global memory
    const int hstride = f->hstride;
    const int j_lo_all = f->band[1], j_hi_all = f->band[2];
    // Per-walker constants are re-read from the frame where they are used (one LDS each) instead of living in
    // registers across the polynomial loop: at 64 registers per thread that is what lets ptxas keep the four Horner
    // chains of a trip interleaved.
    const volatile ZrFrame *fv = f;
    const int nt = (int)m.n_draws;
    const float *theta_t = m.rank_theta;
    const int tstride = TP > 0 ? TP : m.rank_stride;      // TP: the pitch as a compile-time constant (immediate load offsets)
    double part = 0.0;
    if (j_hi_all < j_lo_all) return part;                  // uniform: no row can be reached
    const int Gf = X >> 5, R = X & 31;
    const int s_ref = f->s_ref, k_lo = f->k_lo, n_iv = f->n_iv, nB = f->nB, wB = f->wB;
    const int per_b = R ? 32 / R : 1;
    const int wA = NW - wB;
    const unsigned smem_s32 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned u0_s32 = smem_s32 + op->lay.pa;
    const unsigned rec_s32 = smem_s32 + op->lay.rec;
    // ---- finding a cell's run of draws: hint -> lookup -> short forward walk -------------------------------------
    // lower edge of interval j in u (j == M: the upper end of the last, closed, interval: v > u_max <=> v >= next(u_max))
    auto edge_of = [&](int j) -> double { return brk[j - 1]; };   // 0 <= j <= M (see the staging of sbrk)
    const unsigned short *zlut = rp->zlut;
    auto hint = [&](float th) -> int {
        float ha, hb;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+-8];" : "=f"(ha), "=f"(hb) : "r"(u0_s32));
        int c = __float2int_rz(fmaf(th, ha, hb));
        c = c < 0 ? 0 : (c > ZR_LUT ? ZR_LUT : c);
        return (int)__ldg(zlut + c);
    };
    // the hints are low by construction: forward walks with the membership compare of the other range kernels
    // (measured: loading three candidates at once instead, branch-free, is 1.5 % slower -- more instructions)
    // (u0[nt] = +inf ends every walk: no bound check)
    auto walk = [&](int d, double edge, double delta) -> int {
        while (!(__dadd_rn(u0[d], delta) >= edge)) ++d;
        return d;
    };
    // ---- summing a cell: the weight polynomial over draws [d0, d0 + n), the cell store, the normalisation sum -------
    auto sum_cell = [&](int j, bool active, int d0, int n, double left, double delta, unsigned hrow_s32, double *hrow_g) {
        const int nmax = __reduce_max_sync(FULL, n);
        if (nmax == 0) {                                   // uniform
            if (active) {
                if constexpr (WIDE) hrow_g[j] = 0.0;
                else zr_sts(hrow_s32 + (unsigned)j * 8u, 0.0);
            }
            return;
        }
        const int nmin = __reduce_min_sync(FULL, n);
        // idle lanes read the first record (valid memory, finite numbers) and multiply it by t = 0
        int jb;
        asm volatile("ld.shared.s32 %0, [%1+-16];" : "=r"(jb) : "r"(u0_s32));
        const int ridx = active ? (j - jb) * RW + 2 : 2;
        double a[P + 1];
        double a0;
        if constexpr (WIDE) {
            const double2 *r2 = reinterpret_cast<const double2 *>(rec_g + ridx);
            const double2 c01 = r2[0];
            a0 = c01.x;
            a[1] = c01.y;
#pragma unroll
            for (int k = 2; k <= P; k += 2) {
                const double2 c2 = r2[k >> 1];
                a[k] = c2.x;
                a[k + 1] = c2.y;
            }
        } else {
            const unsigned ra = rec_s32 + (unsigned)ridx * 8u;
            const double2 c01 = zr_lds_v2(ra);
            a0 = c01.x;
            a[1] = c01.y;
#pragma unroll
            for (int k = 2; k <= P; k += 2) {
                const double2 c2 = zr_lds_v2(ra + (unsigned)k * 8u);
                a[k] = c2.x;
                a[k + 1] = c2.y;
            }
        }
        a[0] = 0.0;                                        // the constant term is added once per run, after the loop
        const double off = delta - left;
        double acc = 0.0;
        unsigned addr = u0_s32 + (unsigned)d0 * 8u;
        const int tfull = nmin >> 2;                       // trips in which every lane still has four samples
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
        if (active) {
            const double val = fma((double)n, a0, acc);
            if constexpr (WIDE) hrow_g[j] = val;
            else zr_sts(hrow_s32 + (unsigned)j * 8u, val);
            double de, dx;
            asm volatile("ld.shared.f64 %0, [%1+-32];" : "=d"(de) : "r"(u0_s32));
            asm volatile("ld.shared.f64 %0, [%1+-24];" : "=d"(dx) : "r"(u0_s32));
            part += __dmul_rn(__dmul_rn(val, de), dx);             // adv:143
        }
    };
    // one (row, interval) cell per lane and visit
    auto cell = [&](int j, bool active, float th0, float th1, double delta, unsigned hrow_s32, double *hrow_g) {
        int d0 = 0, n = 0;
        double left = 0.0;
        if (active) {
            left = edge_of(j);
            d0 = walk(hint(th0), left, delta);
            const int h1 = hint(th1);
            const int d1 = walk(h1 < d0 ? d0 : h1, edge_of(j + 1), delta);
            n = d1 - d0;
            TOF_CHECK(d0 >= 0 && d1 <= nt && n >= 0 && (WIDE || j - fv->jbase >= 0));
        }
        sum_cell(j, active, d0, n, left, delta, hrow_s32, hrow_g);
    };
    // a lane bound to one row: the row's window of cells [jw_lo, jw_lo + jw_n) (its E-window, the walker's interval band,
    // and `ok`), its addresses
    struct RowState {
        double delta;
        int jw_lo;
        unsigned jw_n, hrow_s32;
        double *hrow_g;
        const float *th_row;
    };
    auto row_state = [&](int row, bool ok) {
        RowState r;
        const int row_lo = hlo[row];
        r.delta = sdelta[row];
        r.jw_lo = row_lo > j_lo_all ? row_lo : j_lo_all;
        int jw_hi = row_lo + hstride - 1;
        jw_hi = jw_hi < j_hi_all ? jw_hi : j_hi_all;
        if (!ok) jw_hi = r.jw_lo - 1;
        r.jw_n = (unsigned)(jw_hi - r.jw_lo + 1);          // 0 when the window is empty
        r.hrow_s32 = smem_s32 + (unsigned)(row * hstride - row_lo) * 8u;
        r.hrow_g = WIDE ? Hglobal + (size_t)row * hstride - row_lo : nullptr;
        r.th_row = theta_t + row;
        return r;
    };
    // visits j = j_first, j_first + j_step, ... (n_vis of them), one cell each
    auto run_row = [&](int row, bool ok, int j_first, int j_step, int n_vis) {
        const RowState r = row_state(row, ok);
        int j = j_first;
        bool act = (unsigned)(j - r.jw_lo) < r.jw_n && n_vis > 0;
        float th0 = 0.0f, th1 = 0.0f;
        const float *pt = r.th_row + (ptrdiff_t)j * tstride;
        const int pstep = j_step * tstride;
        if (act) {
            th0 = ldg_stream_f32(pt);
            th1 = ldg_stream_f32(pt + tstride);
        }
        for (int v = 0; v < n_vis; ++v) {
            const int jn = j + j_step;
            pt += pstep;
            const bool actn = (unsigned)(jn - r.jw_lo) < r.jw_n && v + 1 < n_vis;
            float th0n = 0.0f, th1n = 0.0f;
            if (actn) {                                    // next visit's hints: in flight while this cell is summed
                th0n = ldg_stream_f32(pt);
                th1n = ldg_stream_f32(pt + tstride);
            }
            cell(j, act, th0, th1, r.delta, r.hrow_s32, r.hrow_g);
            j = jn;
            act = actn;
            th0 = th0n;
            th1 = th1n;
        }
    };
    // visits of PAIRS of adjacent cells (j, j + 1), j = j_first, j_first + j_step, ...: the end of the first run is the
    // start of the second, so a pair costs three edge searches instead of four, and they are independent of each other
    // (+1.5 %; a generic K-cells-per-visit version with K = 2, 3, 4 was measured 2.5 - 3 % slower than this one)
    auto run_row_pairs = [&](int row, int j_first, int j_step, int n_vis) {
        const RowState r = row_state(row, true);
        int j = j_first;
        bool actA = (unsigned)(j - r.jw_lo) < r.jw_n && n_vis > 0, actB = (unsigned)(j + 1 - r.jw_lo) < r.jw_n && n_vis > 0;
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
        const float *pt = r.th_row + (ptrdiff_t)j * tstride;
        const int pstep = j_step * tstride;
        if (actA || actB) {
            t0 = ldg_stream_f32(pt);
            t1 = ldg_stream_f32(pt + tstride);
            t2 = ldg_stream_f32(pt + 2 * tstride);
        }
        for (int v = 0; v < n_vis; ++v) {
            const int jn = j + j_step;
            pt += pstep;
            const bool more = v + 1 < n_vis;
            const bool actAn = (unsigned)(jn - r.jw_lo) < r.jw_n && more, actBn = (unsigned)(jn + 1 - r.jw_lo) < r.jw_n && more;
            float t0n = 0.0f, t1n = 0.0f, t2n = 0.0f;
            if (actAn || actBn) {                          // next visit's hints: in flight while these cells are summed
                t0n = ldg_stream_f32(pt);
                t1n = ldg_stream_f32(pt + tstride);
                t2n = ldg_stream_f32(pt + 2 * tstride);
            }
            int e0 = 0, e1 = 0, e2 = 0;
            double edge0 = 0.0, edge1 = 0.0;
            if (actA || actB) {
                edge1 = edge_of(j + 1);
                int h1 = hint(t1);
                if (actA) {
                    edge0 = edge_of(j);
                    e0 = walk(hint(t0), edge0, r.delta);
                    h1 = h1 < e0 ? e0 : h1;
                }
                e1 = walk(h1, edge1, r.delta);
                if (actB) {
                    const int h2 = hint(t2);
                    e2 = walk(h2 < e1 ? e1 : h2, edge_of(j + 2), r.delta);
                }
            }
            sum_cell(j, actA, e0, actA ? e1 - e0 : 0, edge0, r.delta, r.hrow_s32, r.hrow_g);
            sum_cell(j + 1, actB, e1, actB ? e2 - e1 : 0, edge1, r.delta, r.hrow_s32, r.hrow_g);
            j = jn;
            actA = actAn;
            actB = actBn;
            t0 = t0n;
            t1 = t1n;
            t2 = t2n;
        }
    };
    if (warp < wA) {
        if (Gf <= wA) {                                    // the usual case: this warp keeps one group of rows
            const int g = warp % Gf, idx = warp / Gf;
            const int cnt = (wA - g + Gf - 1) / Gf;        // warps sharing group g
            const int row = (g << 5) + lane;
            const int n_pr = (n_iv + 1) >> 1;              // pairs of adjacent trajectory offsets
            run_row_pairs(row, k_lo + (srow[row] - s_ref) + 2 * idx, 2 * cnt, idx < n_pr ? (n_pr - idx + cnt - 1) / cnt : 0);
        } else {                                           // more groups than warps: stride over (offset, group) pairs
            for (int task = warp; task < n_iv * Gf; task += wA) {
                const int jj = task / Gf;
                const int row = ((task - jj * Gf) << 5) + lane;
                run_row(row, true, k_lo + jj + (srow[row] - s_ref), 1, 1);
            }
        }
    } else {
        const int isub = lane / R;
        const int row = (Gf << 5) + (lane - isub * R);
        const int tb0 = warp - wA;
        run_row(row, isub < per_b, k_lo + isub + (srow[row] - s_ref) + tb0 * per_b, wB * per_b,
                tb0 < nB ? (nB - tb0 + wB - 1) / wB : 0);
    }
    return part;
}

// Exact paths of the scatter, taken by the few cells that sit within 1e-6 of a rounding / bin boundary: functions of
// their own so that the common path stays short and branch-free.
__device__ __noinline__ double zr_exact_count(double h, double S, double rS, double nsamp) {
    return rint(__dmul_rn(div_by_recip(h, S, rS), nsamp));                                   // adv:146
}
__device__ __noinline__ int zr_exact_bin(double xi, double di, double vd, double rvd, double vn, double rvn, int T, double t_min,
                                         double t_max, double t_step, double t_scale) {
    const double tof_d = div_by_recip(xi, vd, rvd);                                          // adv:151-152
    const double tof_n = div_by_recip(di, vn, rvn);                                          // adv:153-156
    return np_bin(__dadd_rn(tof_d, tof_n), T, t_min, t_max, t_step, t_scale);               // adv:159
}

// Normalise (adv:143), np.rint + flight-time scatter (adv:146-159), density, timing response at the observed bins and
// log-likelihood (adv:160-181).  Same integers and the same floating-point results as adv_range_kernel's phases 2-5
// given the same cell sums; `part` is this thread's share of the normalisation sum.
template <int NT, int P, bool PROF, bool WIDE>
__device__ __forceinline__ void zr_finish(const DevModel *mp, const DevRun *rp, const ModelOut *op, unsigned char *smem_raw,
                                       ZrFrame *f, double *Hglobal, double part) {
    __builtin_assume(__isShared(smem_raw));
    __builtin_assume(__isShared(f));
    const DevModel &m = *mp;
    const DevRun &run = *rp;
    const ModelOut &out = *op;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X = m.x_bins, T = run.tof_bins;
    const double *H = WIDE ? Hglobal : reinterpret_cast<const double *>(smem_raw);
    unsigned int *tofc = reinterpret_cast<unsigned int *>(smem_raw + out.lay.pa);
    double *svd = reinterpret_cast<double *>(smem_raw + out.lay.svd);     // [E] deuteron speeds (records are dead now)
    double *rvd = reinterpret_cast<double *>(smem_raw + out.lay.ulut);    // [E] their reciprocals
    const double *staps = reinterpret_cast<const double *>(smem_raw + out.lay.staps);
    double *scratch = reinterpret_cast<double *>(smem_raw + out.lay.scratch);
    const int *hlo = reinterpret_cast<const int *>(smem_raw + out.lay.hlo);
    const long long w = f->w;
    const double e0 = f->e0;
    const int hstride = f->hstride;
    const int j_lo_all = f->band[1], j_hi_all = f->band[2];

    // ---- phase 2: normalise (adv:143); the products were formed where the cells were summed -----------------
    // (the caller's barrier after the cell sums: every warp is done with the draw tile and the records)
    for (int i = tid; i < T; i += NT) tofc[i] = 0u;
    double *rvn_s = rvd + m.e_bins;                        // [E] 1/neutron speed of the E-bins in reach
    {
        const double *__restrict__ ec = m.e_centers;
        const double *__restrict__ rvn_g = m.neutron_rspeed;
        for (int j = j_lo_all + tid; j <= j_hi_all; j += NT) { // deuteron speeds of the E-bins in reach, and reciprocals
            const double eff = __ddiv_rn(__dadd_rn(e0, ec[j]), 2.0);   // adv:151
            const double v = speed_of(m.c, eff, m.m_d);
            svd[j] = v;
            rvd[j] = __ddiv_rn(1.0, v);
            rvn_s[j] = rvn_g[j];
        }
    }
    static_assert(NT <= 640, "the two block_sum1 slots share 40 doubles of scratch");
    const double S = block_sum1<double>(part, scratch);    // includes the barrier that publishes tofc = 0 and the speeds
    if (PROF && tid == 0) {
        const long long t = clock64();
        atomicAdd(out.stage_cycles + 2, (unsigned long long)(t - f->t_mark));
        f->t_mark = t;
    }

    // ---- phase 3: quantise (adv:146) and scatter every non-empty cell to its flight time (adv:149-158) ----
    const double t_min = run.tof_min, t_max = run.tof_max;
    const double t_step = (t_max - t_min) / (double)T;
    const double t_scale = (double)T / (t_max - t_min);
    const double nsamp = (double)m.n_samples;
    const double rS = __ddiv_rn(1.0, S);                    // IEEE quotients below come from this reciprocal (div_by_recip)
    long long cpart = 0;                                    // counts this thread put into the TOF window (np.histogram's n.sum())
    if (S > 0.0 && S < CUDART_INF) {
        constexpr double MAGIC = 6755399441055744.0;        // 2^52 + 2^51: x + MAGIC - MAGIC = rint(x), low word = (int)rint(x)
        constexpr double SURE = 0.499999;                   // farther than 1e-6 from a rounding boundary
        const double k1 = __dmul_rn(rS, nsamp);
        const double q_off = __dmul_rn(-t_min, t_scale) - 0.5;
        const double *__restrict__ xc = m.x_centers;
        const double *__restrict__ nd = run.neutron_dist;
        const double *__restrict__ vn_g = m.neutron_speed;
        for (int row = warp; row < X; row += NW) {
            const double xi = xc[row], di = nd[row];
            const int row_lo = hlo[row];
            const double *Hr = H + (size_t)row * hstride;
            int jb_hi = j_hi_all - row_lo + 1;              // cells beyond the walker's last interval were never written
            jb_hi = jb_hi < hstride ? jb_hi : hstride;
            // (two cells per lane and trip, unrolled, measured 1 % slower: more registers, same latency chain per cell)
            for (int jb = lane + (row_lo < j_lo_all ? j_lo_all - row_lo : 0); jb < jb_hi; jb += 32) {
                const int j = row_lo + jb;
                const double h = Hr[jb];
                if (h != 0.0) {
                    // Two independent chains, both formed before any branch so that they overlap:
                    // (1) cnt = rint(RN(RN(h/S) * N)) (adv:146).  Fast: c = h * RN(rS*N) is within 2 ulp of the product the
                    //     reference rounds; if it is not within 1e-6 of a half-integer both round to the same integer.
                    const double c = __dmul_rn(h, k1);
                    const double cm = __dadd_rn(c, MAGIC);
                    double cnt = __dsub_rn(cm, MAGIC);
                    unsigned int ci = (unsigned int)__double2loint(cm);
                    const bool sure_c = fabs(__dsub_rn(c, cnt)) < SURE && c < 1e9;
                    // (2) TOF bin (adv:149-159).  Fast: t = (tof - tof_min) * T/(max - min) - 1/2 from reciprocals; if t is
                    //     not within 1e-6 of a half-integer, rint(t) is numpy's bin (its edges are within 1e-12 bins of the
                    //     uniform grid); otherwise the exact quotients and numpy's edge rule decide.
                    const double rvn = rvn_s[j];
                    const double tof = fma(xi, rvd[j], __dmul_rn(di, rvn));
                    const double t = fma(tof, t_scale, q_off);
                    const double tm = __dadd_rn(t, MAGIC);
                    int b = __double2loint(tm);
                    const bool sure_b = fabs(__dsub_rn(t, __dsub_rn(tm, MAGIC))) < SURE && fabs(t) < 1e9;
                    if (__builtin_expect(!sure_c, 0)) {
                        cnt = zr_exact_count(h, S, rS, nsamp);
                        ci = (unsigned int)cnt;
                    }
                    if (cnt > 0.0) {
                        if (__builtin_expect(!sure_b, 0))
                            b = zr_exact_bin(xi, di, svd[j], rvd[j], vn_g[j], rvn, T, t_min, t_max, t_step, t_scale);
                        TOF_CHECK(j < m.e_bins);
                        if ((unsigned)b < (unsigned)T) {
                            atomicAdd(tofc + b, ci);
                            cpart += (long long)ci;
                        }
                    }
                }
            }
        }
    }
    // (the barrier inside the next reduction is also the one that completes the scatter)
    const long long total_i = block_sum1<long long>(cpart, reinterpret_cast<long long *>(scratch) + NT / 32);
    if (PROF && tid == 0) {
        const long long t = clock64();
        atomicAdd(out.stage_cycles + 3, (unsigned long long)(t - f->t_mark));
        f->t_mark = t;
    }

    // ---- phase 4: density (np.histogram density=True), only where the timing response of an observed bin reads it ----
    const bool degenerate = !(S > 0.0) || total_i == 0;
    const double total = (double)total_i;
#ifdef TOF_ZR_DEBUG
    if (tid == 0 && w < 4) printf("zr w=%lld wide=%d S=%g total=%lld hstride=%d jlo=%d jhi=%d jbase=%d ha=%g hb=%g\n", w, f->wide, S, total_i, hstride, j_lo_all, j_hi_all, f->jbase, (double)f->hint_a, (double)f->hint_b);
#endif
    double *pdf = reinterpret_cast<double *>(smem_raw);    // the banded histogram region (a wide walker's cells are in global)
    int need_lo = 0, need_hi = -1;
    if (run.n_obs_nz > 0) {                                // obs_nz_idx ascends
        need_lo = run.obs_nz_idx[0] + m.conv_shift - (m.n_taps - 1);
        need_hi = run.obs_nz_idx[run.n_obs_nz - 1] + m.conv_shift;
        need_lo = need_lo < 0 ? 0 : need_lo;
        need_hi = need_hi > T - 1 ? T - 1 : need_hi;
    }
    for (int t = need_lo + tid; t <= need_hi; t += NT) {    // (the scatter has finished reading H: barrier in the reduction above)
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
    lp = block_sum1<double>(lp, scratch);
    if (tid == 0) {
        double r = degenerate ? CUDART_NAN : lp;
        if (r != r && out.nan_count) atomicAdd(out.nan_count, 1ull);
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
    if (PROF && tid == 0) {
        const long long t = clock64();
        atomicAdd(out.stage_cycles + 4, (unsigned long long)(t - f->t_mark));
        atomicAdd(out.stage_cycles + TOF_N_STAGES, 1ull);
        f->t_mark = t;
    }
}

template <int NT, int P, bool PROF = false, int TP = 0>
__global__ void __launch_bounds__(NT, 2) adv_zrank_kernel(const __grid_constant__ DevModel m, const __grid_constant__ DevRun run,
                                                          const double *__restrict__ theta, long long n_walkers,
                                                          const __grid_constant__ ModelOut out) {
    extern __shared__ __align__(16) unsigned char smem_sym[];
    unsigned char *smem_raw = smem_sym;
    asm volatile("" : "+l"(smem_raw));                     // opaque base, still known to be shared (see adv_range_kernel)
    __builtin_assume(__isShared(smem_raw));
    __shared__ ZrFrame frame;
    constexpr int RW = P + 3;
    const int tid = threadIdx.x;
    const int X = m.x_bins, M = m.rng_n;
    // ---- walker-independent tables: staged once per CTA (persistent CTAs loop over walkers) ------------------
    {
        double *staps = reinterpret_cast<double *>(smem_raw + out.lay.staps);
        double *sdelta = reinterpret_cast<double *>(smem_raw + out.lay.sdelta);
        double *sbrk = reinterpret_cast<double *>(smem_raw + out.lay.sbrk);
        const double *recg = m.rng_rec;
        // lower edge of interval j = sbrk[j - 1], j in [0, M]: 0 in front, and the upper end of the last (closed) interval
        // as the first double above it (v > u_max <=> v >= next(u_max))
        for (int j = tid; j < M; j += NT)
            sbrk[j] = j == M - 1 ? __longlong_as_double(__double_as_longlong(m.rng_u_max) + 1) : recg[(size_t)j * RW];
        if (tid == 0) sbrk[-1] = 0.0;
        for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
        const double x_start = m.ode_from_zero ? 0.0 : m.x_centers[0];
        for (int i = tid; i < X; i += NT) sdelta[i] = m.rng_sign * (m.x_centers[i] - x_start);
        if (PROF && tid == 0) frame.t_mark = clock64();
        if (tid == 0) {
            frame.de = (m.e_max - m.e_min) / (double)m.e_bins;
            frame.dx = (m.x_max - m.x_min) / (double)X;
            frame.zlut = run.zlut;
            reinterpret_cast<double *>(smem_raw + out.lay.pa - 32)[0] = frame.de;     // ZR_MIRROR
            reinterpret_cast<double *>(smem_raw + out.lay.pa - 32)[1] = frame.dx;
            frame.idx[0] = (long long)atomicAdd(out.work, 1ull);
        }
    }
    double *Hglobal = out.wide_scratch ? out.wide_scratch + (size_t)blockIdx.x * (size_t)out.split_stride : nullptr;
    auto stage = [&](int k) {
        if constexpr (PROF) {
            if (tid == 0) {
                const long long t = clock64();
                atomicAdd(out.stage_cycles + k, (unsigned long long)(t - frame.t_mark));
                frame.t_mark = t;
            }
        }
    };
    for (int it = 0;; ++it) {
        const int st = zr_setup<NT, P, PROF>(&m, &run, theta, n_walkers, &out, smem_raw, &frame, it);
        if (st == PLANNED_DONE) break;
        stage(0);
        if (st == PLANNED_SKIP) continue;
        if (frame.wide) {
            const double part = zr_exec<NT, P, true, TP>(&m, &run, &out, smem_raw, &frame, Hglobal);
            __threadfence_block();
            __syncthreads();
            stage(1);
            zr_finish<NT, P, PROF, true>(&m, &run, &out, smem_raw, &frame, Hglobal, part);
        } else {
            const double part = zr_exec<NT, P, false, TP>(&m, &run, &out, smem_raw, &frame, nullptr);
            __syncthreads();
            stage(1);
            zr_finish<NT, P, PROF, false>(&m, &run, &out, smem_raw, &frame, nullptr, part);
        }
    }
}

// ================================================================================================================
// Many draws per walker (n_draws > RANGE_TILE; the reference's own default is 1e5 - 1e6): adv_zrank_multi_kernel
// ================================================================================================================
// The sorted draws are taken a tile of RANGE_TILE at a time.  A tile of a big draw set is a narrow slice of the energy
// distribution: for a given row it touches one to three E-bins, so a (row, E-bin) run is hundreds of consecutive draws
// and the polynomial loop is all there is -- 11 instructions per 32 samples against 34 for the warp-private streaming
// walk of adv_range_kernel (which re-derives the interval of every sample; ncu: 11.0 M warp-instructions per walker at
// 1e5 draws).  Lane = row as in zr_exec; the runs of a tile are found by binary search in the tile (no hints needed:
// the search is ~1 % of a run's polynomial work); every run is split over the warps that share the group of rows and
// the partial sums meet in the cell with atomics (as in the streaming walk, the summation order of those few partials
// is not fixed).  Normalisation, scatter, density and likelihood are zr_finish.
template <int NT, int P>
__device__ __forceinline__ int zrm_setup(const DevModel *mp, const DevRun *rp, const double *__restrict__ theta, long long n_walkers,
                                      const ModelOut *op, unsigned char *smem_raw, ZrFrame *f, int it, double *Hglobal) {
    __builtin_assume(__isShared(smem_raw));
    __builtin_assume(__isShared(f));
    const DevModel &m = *mp;
    const DevRun &run = *rp;
    const ModelOut &out = *op;
    constexpr int RW = P + 3;
    const int tid = threadIdx.x;
    const int X = m.x_bins, M = m.rng_n;
    double *H = reinterpret_cast<double *>(smem_raw);
    double *rec = reinterpret_cast<double *>(smem_raw + out.lay.rec);
    const double *sdelta = reinterpret_cast<const double *>(smem_raw + out.lay.sdelta);
    int *hlo_s = reinterpret_cast<int *>(smem_raw + out.lay.hlo);
    const double *sbrk = reinterpret_cast<const double *>(smem_raw + out.lay.sbrk);
    if (tid == 0) {
        f->band[0] = 0;
        f->band[1] = M;
        f->band[2] = -1;
    }
    __syncthreads();                                       // the previous walker is done with shared memory
    // a work item is (walker, split): with few walkers and a big draw set, n_split CTAs share one walker's tiles
    const int NS = out.n_split > 1 ? out.n_split : 1;
    const long long item = f->idx[it & 1];
    const long long w = item / NS;
    const int split = (int)(item - w * NS);
    if (w >= n_walkers) return PLANNED_DONE;
    if (tid == 0) f->idx[(it + 1) & 1] = (long long)atomicAdd(out.work, 1ull);
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
    const bool rev = spread < 0.0;
    const double umax = m.rng_u_max;
    // E-bins the walker can touch: the draws are sorted, first and last give the extremes
    const double u_lo = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 : 0)))), m);
    const double u_hi = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? 0 : m.n_draws - 1)))), m);
    for (int i = tid; i < X; i += NT) {
        const double dl = sdelta[i];
        double vmin = u_lo > -CUDART_INF ? u_lo + dl : 0.0;     // -inf draws: the lowest in-range v is 0
        double vmax = u_hi + dl;
        vmin = vmin > 0.0 ? vmin : 0.0;
        vmax = vmax < umax ? vmax : umax;
        int j_lo = 0, j_hi = 0;
        if (vmax >= vmin) {
            j_lo = range_interval(vmin, sbrk, m.rng_lut, m.rng_lut_inv, m.rng_lut_n, M);
            j_hi = range_interval(vmax, sbrk, m.rng_lut, m.rng_lut_inv, m.rng_lut_n, M);
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
    const bool fits = (long long)X * hstride <= out.hcap && (j_hi_all - jbase + 1) <= out.rcap;
    if (fits) {
        const double *recg = m.rng_rec;
        for (int i = tid; i < (j_hi_all - jbase + 1) * RW; i += NT) rec[i] = recg[(size_t)jbase * RW + i];
        for (int i = tid; i < X * hstride; i += NT) H[i] = 0.0;
    } else {
        for (int i = tid; i < X * hstride; i += NT) Hglobal[i] = 0.0;
    }
    if (tid == 0) {
        f->w = w;
        f->e0 = e0;
        f->spread = spread;
        f->split = split;
        f->hstride = hstride;
        f->jbase = fits ? jbase : 0;
        f->wide = fits ? 0 : 1;
        if (f->band[2] < 0) {
            f->band[1] = 0;
            f->band[2] = -1;
        }
        if (!fits && split == 0 && out.queue_count) atomicAdd(out.queue_count, 1ull);
    }
    __syncthreads();
    return PLANNED_RUN;
}

// All tiles of one walker: (x,E) histogram of cross-section weights (adv:128-138), accumulated into H.
//
// Work split inside a tile: the warps that share a group of 32 rows each take a contiguous slice [A, B) of the tile's
// draws -- the SAME slice for every lane (= row) of the warp, so all lanes have the same number of samples.  A row's
// slice lies in one E-bin unless one of the bin's edges falls inside it (a quarter of the rows per tile, one slice in
// five): every lane sums the first segment of its slice in lockstep (the main pass: full four-sample trips only), and
// the few leftover segments are then summed one after the other by all 32 lanes together (strided over the segment's
// draws, coefficients broadcast by shuffle, warp sum).  Letting those rows run a second lockstep pass instead doubled
// the polynomial work (measured: 10.0 M warp-instructions per walker at 1e5 draws, against 11.0 M for the streaming walk).
// Measured alternatives, both slower than this (81.6 K evals/s at 4096 walkers x 1e5 draws): double-buffered tiles with
// one barrier per tile and the next tile staged ahead (77.4 K); warp-private chunks of 128 draws with no CTA barrier at
// all in the draw loop (76.6 K: the energy-loss lookups are repeated by every group of rows, per-chunk overheads); ten
// slices per group handed out through a shared counter instead of one static slice per warp (80.5 K).
template <int NT, int P, bool WIDE>
__device__ __forceinline__ void zrm_exec(const DevModel &m, const DevRun &run, const ModelOut &out, unsigned char *smem_raw,
                                         const ZrFrame *f, double *Hglobal) {
    constexpr int RW = P + 3;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X = m.x_bins, M = m.rng_n;
    double *u0 = reinterpret_cast<double *>(smem_raw + out.lay.pa);
    double *Hs = reinterpret_cast<double *>(smem_raw);
    const double *brk = reinterpret_cast<const double *>(smem_raw + out.lay.sbrk);
    const double *sdelta = reinterpret_cast<const double *>(smem_raw + out.lay.sdelta);
    const int *hlo = reinterpret_cast<const int *>(smem_raw + out.lay.hlo);
    const double *rec_s = reinterpret_cast<const double *>(smem_raw + out.lay.rec);
    const double *rec_g = m.rng_rec;
    const volatile ZrFrame *fv = f;
    const int hstride = f->hstride;
    const int j_lo_all = f->band[1], j_hi_all = f->band[2];
    if (j_hi_all < j_lo_all) return;                       // uniform: no row can be reached
    const double e0 = f->e0, spread = f->spread, umax = m.rng_u_max;
    const bool rev = spread < 0.0;
    const unsigned smem_s32 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned u0_s32 = smem_s32 + out.lay.pa;
    // warps -> groups of 32 rows; the X % 32 leftover rows get one warp that packs R rows x (32/R) sub-slices
    const int Gf = X >> 5, R = X & 31;
    const int wB = R ? (Gf ? 1 : NW) : 0, wA = NW - wB;
    int row, parts, part;
    bool lane_ok = true;
    if (warp < wA) {
        const int g = warp % Gf, q = warp / Gf;
        row = (g << 5) + lane;
        parts = (wA - g + Gf - 1) / Gf;                    // warps sharing group g
        part = q;
    } else {
        const int per_b = 32 / R, sub = lane / R;
        row = (Gf << 5) + (lane - sub * R);
        parts = wB * per_b;
        part = (warp - wA) * per_b + sub;
        lane_ok = sub < per_b;
    }
    const double delta = sdelta[row];
    const int hbase = row * hstride - hlo[row];            // H index of (row, E-bin j) is hbase + j
    const int row_lo = hlo[row];
    auto edge_of = [&](int j) -> double { return brk[j - 1]; };   // 0 <= j <= M (see the staging of sbrk)
    // first draw d in [lo, hi) with RN(u0[d] + dl) >= edge (hi when there is none): the membership compare of every range
    // kernel, by bisection
    auto first_ge = [&](int lo, int hi, double edge, double dl) -> int {
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__dadd_rn(u0[mid], dl) >= edge) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    auto load_record = [&](int j, double (&a)[P + 1], double &a0) {
        const int ridx = (j - fv->jbase) * RW + 2;
        const double2 *r2 = reinterpret_cast<const double2 *>((WIDE ? rec_g : rec_s) + ridx);
        const double2 c01 = r2[0];
        a0 = c01.x;
        a[0] = 0.0;                                        // the constant term is added once per run
        a[1] = c01.y;
#pragma unroll
        for (int k = 2; k <= P; k += 2) {
            const double2 c2 = r2[k >> 1];
            a[k] = c2.x;
            a[k + 1] = c2.y;
        }
    };
    auto add_cell = [&](int hidx, double val) {
        if constexpr (WIDE) atomicAdd(Hglobal + hidx, val);
        else atomicAdd(Hs + hidx, val);
    };
    const int NS = out.n_split > 1 ? out.n_split : 1;
    for (long long tile = (long long)f->split * RANGE_TILE; tile < m.n_draws; tile += (long long)NS * RANGE_TILE) {
        const int nt = (int)((m.n_draws - tile < RANGE_TILE) ? (m.n_draws - tile) : RANGE_TILE);
        __syncthreads();                                   // the previous tile is consumed
        int n_neg = 0, n_pos = 0;                          // draws outside the energy table: -inf first, +inf last
        for (int d = tid; d < RANGE_TILE; d += NT) {
            double u = CUDART_INF;
            if (d < nt) {
                u = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 - (tile + d) : tile + d)))), m);
                u0[d] = u;
            }
            n_neg += __syncthreads_count(d < nt && !(u > -CUDART_INF));
            n_pos += __syncthreads_count(d < nt && u >= CUDART_INF);
        }
        const int v_lo = n_neg, v_hi = nt - n_pos;         // finite part of the tile (NaN counts as -inf)
        if (v_hi <= v_lo) continue;                        // uniform
        // this lane's slice of the tile's draws, and its first in-range draw
        const int nf = v_hi - v_lo;
        int sB = v_lo + ((part + 1) * nf) / parts;
        int s = v_lo + (part * nf) / parts;
        if (!lane_ok) sB = s;
        s = first_ge(s, sB, 0.0, delta);                   // samples below the histogram range (v < 0) are skipped
        // E-bin of the first sample of the slice (edges: the compare rule, so range_interval's answer is the cell)
        int j = 0;
        if (s < sB) {
            const double v = __dadd_rn(u0[s], delta);
            if (v > umax) s = sB;                          // the whole slice is beyond the range (sorted)
            else j = range_interval(v, brk, m.rng_lut, m.rng_lut_inv, m.rng_lut_n, M);
        }
        // ---- main pass: every lane sums the first segment of its slice (up to the next edge) in lockstep --------
        {
            const bool act = s < sB && (unsigned)(j - row_lo) < (unsigned)hstride;
            const int s1 = (s < sB) ? first_ge(s, sB, edge_of(j + 1), delta) : sB;
            const int n = act ? s1 - s : 0;
            const int nmax = __reduce_max_sync(FULL, n);
            if (nmax > 0) {
                const int nmin = __reduce_min_sync(FULL, n);
                double a[P + 1], a0;
                load_record(act ? j : fv->jbase, a, a0);
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
                TOF_CHECK(n == 0 || (s >= v_lo && s1 <= v_hi && hbase + j >= 0 && hbase + j < X * hstride && (WIDE || j - fv->jbase >= 0)));
                if (n > 0) add_cell(hbase + j, fma((double)n, a0, acc));
            }
            s = s1;
            ++j;
        }
        // ---- leftover segments (a bin edge inside the slice), one at a time, all 32 lanes on each ------------------
        for (;;) {
            const bool more = s < sB && j < M;
            const unsigned pending = __ballot_sync(FULL, more);
            if (pending == 0u) break;                      // uniform
            const int src = __ffs(pending) - 1;
            // the owner finds the end of its segment, then hands the segment to the warp
            int s1 = 0;
            if (lane == src) s1 = first_ge(s, sB, edge_of(j + 1), delta);
            const int seg_s = __shfl_sync(FULL, s, src), seg_e = __shfl_sync(FULL, s1, src), seg_j = __shfl_sync(FULL, j, src);
            const double seg_delta = __shfl_sync(FULL, delta, src);
            const int seg_h = __shfl_sync(FULL, hbase, src), seg_row_lo = __shfl_sync(FULL, row_lo, src);
            const int seg_n = seg_e - seg_s;
            if (seg_n > 0 && (unsigned)(seg_j - seg_row_lo) < (unsigned)hstride) {
                double a[P + 1], a0;
                load_record(seg_j, a, a0);
                const double off = seg_delta - edge_of(seg_j);
                const int ch = (seg_n + 31) >> 5;          // consecutive samples per lane
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
                TOF_CHECK(seg_s >= v_lo && seg_e <= v_hi && seg_h + seg_j >= 0 && seg_h + seg_j < X * hstride);
                if (lane == 0) add_cell(seg_h + seg_j, acc);
            }
            if (lane == src) {
                s = s1;
                ++j;
            }
        }
    }
}

template <int NT, int P>
__global__ void __launch_bounds__(NT, 2) adv_zrank_multi_kernel(const __grid_constant__ DevModel m, const __grid_constant__ DevRun run,
                                                                const double *__restrict__ theta, long long n_walkers,
                                                                const __grid_constant__ ModelOut out) {
    extern __shared__ __align__(16) unsigned char smem_sym[];
    unsigned char *smem_raw = smem_sym;
    asm volatile("" : "+l"(smem_raw));
    __builtin_assume(__isShared(smem_raw));
    __shared__ ZrFrame frame;
    constexpr int RW = P + 3;
    const int tid = threadIdx.x;
    const int X = m.x_bins, M = m.rng_n;
    {
        double *staps = reinterpret_cast<double *>(smem_raw + out.lay.staps);
        double *sdelta = reinterpret_cast<double *>(smem_raw + out.lay.sdelta);
        double *sbrk = reinterpret_cast<double *>(smem_raw + out.lay.sbrk);
        const double *recg = m.rng_rec;
        // lower edge of interval j = sbrk[j - 1], j in [0, M]: 0 in front, and the upper end of the last (closed) interval
        // as the first double above it (v > u_max <=> v >= next(u_max))
        for (int j = tid; j < M; j += NT)
            sbrk[j] = j == M - 1 ? __longlong_as_double(__double_as_longlong(m.rng_u_max) + 1) : recg[(size_t)j * RW];
        if (tid == 0) sbrk[-1] = 0.0;
        for (int i = tid; i < m.n_taps; i += NT) staps[i] = m.taps[i];
        const double x_start = m.ode_from_zero ? 0.0 : m.x_centers[0];
        for (int i = tid; i < X; i += NT) sdelta[i] = m.rng_sign * (m.x_centers[i] - x_start);
        if (tid == 0) {
            frame.de = (m.e_max - m.e_min) / (double)m.e_bins;
            frame.dx = (m.x_max - m.x_min) / (double)X;
            frame.zlut = nullptr;
            frame.idx[0] = (long long)atomicAdd(out.work, 1ull);
        }
    }
    double *Hglobal = out.wide_scratch ? out.wide_scratch + (size_t)blockIdx.x * (size_t)out.split_stride : nullptr;
    for (int it = 0;; ++it) {
        const int st = zrm_setup<NT, P>(&m, &run, theta, n_walkers, &out, smem_raw, &frame, it, Hglobal);
        if (st == PLANNED_DONE) break;
        if (st == PLANNED_SKIP) continue;
        const bool wide = frame.wide != 0;
        if (wide) zrm_exec<NT, P, true>(m, run, out, smem_raw, &frame, Hglobal);
        else zrm_exec<NT, P, false>(m, run, out, smem_raw, &frame, nullptr);
        __threadfence_block();
        __syncthreads();
        const int ncell = X * frame.hstride;
        if (out.n_split > 1) {
            // partial histogram -> L2-resident scratch; the CTA that arrives last adds the partials in split order (a fixed
            // order: the sum does not depend on which CTA finishes last) and carries on alone
            const int NS = out.n_split;
            const long long w = frame.w;
            double *Hmine = wide ? Hglobal : reinterpret_cast<double *>(smem_raw);
            double *part_out = out.split_scratch + ((size_t)w * NS + frame.split) * (size_t)out.split_stride;
            for (int i = tid; i < ncell; i += NT) part_out[i] = Hmine[i];
            __threadfence();
            __syncthreads();
            if (tid == 0) frame.last = (atomicAdd(out.split_tickets + w, 1u) == (unsigned)(NS - 1)) ? 1 : 0;
            __syncthreads();
            if (!frame.last) continue;
            __threadfence();
            const double *base_in = out.split_scratch + (size_t)w * NS * (size_t)out.split_stride;
            for (int i = tid; i < ncell; i += NT) {
                double v = 0.0;
                for (int k = 0; k < NS; ++k) v += __ldcg(base_in + (size_t)k * out.split_stride + i);
                Hmine[i] = v;
            }
            if (tid == 0) out.split_tickets[w] = 0u;       // ready for the next call
            __syncthreads();
        }
        // normalisation sum (adv:143) over the walker's cells, then phases 3-5
        const double *H = wide ? Hglobal : reinterpret_cast<const double *>(smem_raw);
        double part = 0.0;
        for (int i = tid; i < ncell; i += NT) part += __dmul_rn(__dmul_rn(H[i], frame.de), frame.dx);
        if (wide) zr_finish<NT, P, false, true>(&m, &run, &out, smem_raw, &frame, Hglobal, part);
        else zr_finish<NT, P, false, false>(&m, &run, &out, smem_raw, &frame, nullptr, part);
    }
}

}  // namespace tof
