// adv / intermediate model, range-table formulation (TOF_ODE_RANGE).
#pragma once
#include <type_traits>

#include "tof_common.cuh"
#include "adv_rk4.cuh"

namespace tof {

// ================================================================================================
// adv / intermediate model, range-table formulation (TOF_ODE_RANGE)
// ================================================================================================
// The stopping ODE (ionStopping.py:78-97) is autonomous, so u(E) = int dE/|f| turns "integrate every draw through
// every x" (adv:129) into v = u0_d + sgn*(x_i - x_start): one T1 lookup per draw, and for every (draw, x) sample an
// interval of the T2 table that gives both its E-bin and a polynomial for its cross-section weight
// (range_tables.py).  The library keeps the draws sorted, so for a fixed row the samples of one interval are a
// contiguous range of draws.  Three phase-1 strategies share the tables:
//   * tiles (n_draws < RANGE_STREAM_MIN): 1024 draws staged in shared memory, a per-tile lookup from u to draw
//     index, tasks of 32 lanes x up to RANGE_PAIR consecutive (row, interval) cells with lane = row -- every cell is
//     summed by exactly one lane, no atomics, fixed order; the polynomial loop is warp-uniform and predicated
//     (range_accumulate_tile, poly_trip4).  Optional FP32 weights: see "FP32 mode" below;
//   * pieces: when a tile spans few intervals each run is split over several lanes (atomics on the cell);
//   * streaming (big draw sets): warp-private walk over 128 draws at a time, values broadcast by shuffle.
// The cell histogram is either full (X*E doubles, 1 CTA/SM) or banded per row (2 CTAs/SM), see adv_range_kernel.
constexpr int RANGE_TILE = 1024;   // draws staged in shared memory at a time
constexpr int RANGE_ULUT = 1024;   // cells of the per-tile draw-index lookup table
constexpr int RANGE_STREAM_MIN = 8192;  // draws per walker from which the warp-private streaming walk is used
constexpr int RANGE_PAIR = 4;       // consecutive intervals per lane in a type-A task (fewer when tasks are scarce)
constexpr int RANGE_SPLIT = 6;      // long runs: pieces per warp when a tile has few (row, interval) tasks
constexpr int RANGE_PF = 3;         // FP32 mode: degree of the per-interval weight polynomial (float records)
constexpr int RANGE_RWF = 8;        // FP32 mode: floats per record = c0..c3, E-bin, shared-cell flag, c0 as a double
constexpr int SIMULT_ULUT = 256;   // same for the 10-row simultaneous fit (fewer lookups per tile)

// bytes of the per-tile lookup region; after phase 1 it holds the reciprocal deuteron speeds ([E] doubles)
__host__ __device__ inline size_t range_ulut_bytes(int E) {
    const size_t a = (size_t)RANGE_ULUT * 2, b = (size_t)E * 8;
    return ((a > b ? a : b) + 15) / 16 * 16;
}

// doubles reserved for the staged T2 records: FP64 records, or (FP32 mode) float records followed by the float tile
__host__ __device__ inline size_t range_rec_doubles(int rcap, int P) {
    const size_t a = (size_t)rcap * (P + 3), b = (size_t)rcap * (RANGE_RWF / 2) + RANGE_TILE / 2;
    return a > b ? a : b;
}

// hcap: cells of the (possibly banded) histogram; rcap: staged T2 records.  Regions, in order:
//   H [hcap] f64 | region A: draw tile u0 [RANGE_TILE] f64, later the TOF counters [T] u32 | staged T2 records |
//   deuteron speeds [E] | taps | 40 doubles of reduction scratch | delta [X] | interval lookup (u16) | per-tile draw
//   lookup (u16; later 1/speed [E]) | srow [X] int | hlo [X] int | interval ends [M] f64 | E-bin of each interval [M] u16
inline RangeLayout range_layout(int X, int E, int T, int hcap, int rcap, int P, int n_taps, int lut_n, int rng_n) {
    RangeLayout L;
    size_t region_a = (size_t)T * 4 > (size_t)RANGE_TILE * 8 ? (size_t)T * 4 : (size_t)RANGE_TILE * 8;
    region_a = (region_a + 15) / 16 * 16;
    size_t o = ((size_t)hcap * 8 + 15) / 16 * 16;         // what follows is read with 16-byte loads
    L.pa = (unsigned)o;        o += region_a;
    L.rec = (unsigned)o;       o += range_rec_doubles(rcap, P) * 8;
    L.svd = (unsigned)o;       o += (size_t)E * 8;
    L.staps = (unsigned)o;     o += (size_t)n_taps * 8;
    L.scratch = (unsigned)o;   o += 40 * 8;
    L.sdelta = (unsigned)o;    o += (size_t)X * 8;
    L.lut = (unsigned)o;       o += (size_t)((lut_n + 7) / 8) * 8 * 2;
    L.ulut = (unsigned)o;      o += range_ulut_bytes(E);
    L.srow = (unsigned)o;      o += (size_t)(X + (X & 1)) * 4;   // (odd X: keeps what follows 8-byte aligned)
    L.hlo = (unsigned)o;       o += (size_t)(X + (X & 1)) * 4;
    L.sbrk = (unsigned)o;      o += (size_t)rng_n * 8;
    L.sbin = (unsigned)o;      o += (((size_t)rng_n * 2 + 15) / 16) * 16;
    L.total = (unsigned)(o + 16);
    return L;
}

inline size_t range_smem_bytes(int X, int E, int T, int hcap, int rcap, int P, int n_taps, int lut_n, int rng_n) {
    return range_layout(X, E, T, hcap, rcap, P, n_taps, lut_n, rng_n).total;
}


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

// range_interval in two steps, so that a caller can put the table loads of several lookups in flight together
__device__ __forceinline__ int range_lut_guess(double v, const unsigned short *lut, double lut_inv, int lut_n) {
    int c = (int)(v * lut_inv);
    c = c < 0 ? 0 : (c > lut_n - 1 ? lut_n - 1 : c);
    return lut[c];
}
__device__ __forceinline__ int range_refine(double v, int j, const double *brk, int M) {
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

// Four consecutive samples of one (row, interval) run: acc += dt*q(dt) with dt = u0[d] + off and
// q = a1 + a2 dt + ... + aP dt^(P-1) (the a0 term is added once per run as n*a0).  Sample k counts only if k < r.
// Written in PTX so that the tile pointer stays a 32-bit shared address in a register and the accumulate is a
// predicated DFMA: 1 LDS + 1 DADD + P DFMA per sample.
template <int P>
__device__ __forceinline__ void poly_trip4(double &acc, unsigned addr, int r, double off, const double (&a)[P + 1]) {
    static_assert(P == 7, "poly_trip4 is written for degree-7 records");
    asm("{\n\t"
        ".reg .pred p0, p1, p2, p3;\n\t"
        ".reg .f64 t0, t1, t2, t3, q0, q1, q2, q3;\n\t"
        "ld.shared.f64 t0, [%1];\n\t"
        "ld.shared.f64 t1, [%1+8];\n\t"
        "ld.shared.f64 t2, [%1+16];\n\t"
        "ld.shared.f64 t3, [%1+24];\n\t"
        "setp.gt.s32 p0, %2, 0;\n\t"
        "setp.gt.s32 p1, %2, 1;\n\t"
        "setp.gt.s32 p2, %2, 2;\n\t"
        "setp.gt.s32 p3, %2, 3;\n\t"
        "add.rn.f64 t0, t0, %3;\n\t"
        "add.rn.f64 t1, t1, %3;\n\t"
        "add.rn.f64 t2, t2, %3;\n\t"
        "add.rn.f64 t3, t3, %3;\n\t"
        "fma.rn.f64 q0, %10, t0, %9;\n\t"
        "fma.rn.f64 q1, %10, t1, %9;\n\t"
        "fma.rn.f64 q2, %10, t2, %9;\n\t"
        "fma.rn.f64 q3, %10, t3, %9;\n\t"
        "fma.rn.f64 q0, q0, t0, %8;\n\t"
        "fma.rn.f64 q1, q1, t1, %8;\n\t"
        "fma.rn.f64 q2, q2, t2, %8;\n\t"
        "fma.rn.f64 q3, q3, t3, %8;\n\t"
        "fma.rn.f64 q0, q0, t0, %7;\n\t"
        "fma.rn.f64 q1, q1, t1, %7;\n\t"
        "fma.rn.f64 q2, q2, t2, %7;\n\t"
        "fma.rn.f64 q3, q3, t3, %7;\n\t"
        "fma.rn.f64 q0, q0, t0, %6;\n\t"
        "fma.rn.f64 q1, q1, t1, %6;\n\t"
        "fma.rn.f64 q2, q2, t2, %6;\n\t"
        "fma.rn.f64 q3, q3, t3, %6;\n\t"
        "fma.rn.f64 q0, q0, t0, %5;\n\t"
        "fma.rn.f64 q1, q1, t1, %5;\n\t"
        "fma.rn.f64 q2, q2, t2, %5;\n\t"
        "fma.rn.f64 q3, q3, t3, %5;\n\t"
        "fma.rn.f64 q0, q0, t0, %4;\n\t"
        "fma.rn.f64 q1, q1, t1, %4;\n\t"
        "fma.rn.f64 q2, q2, t2, %4;\n\t"
        "fma.rn.f64 q3, q3, t3, %4;\n\t"
        "@p0 fma.rn.f64 %0, q0, t0, %0;\n\t"
        "@p1 fma.rn.f64 %0, q1, t1, %0;\n\t"
        "@p2 fma.rn.f64 %0, q2, t2, %0;\n\t"
        "@p3 fma.rn.f64 %0, q3, t3, %0;\n\t"
        "}"
        : "+d"(acc)
        : "r"(addr), "r"(r), "d"(off), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]));
}

// ================================================================================================
// FP32 mode (tof_config.precision = TOF_PRECISION_FP32): the per-sample arithmetic of phase 1 in single precision
// ================================================================================================
// Which draws fall into which (row, interval) cell is decided exactly as in FP64 mode (same searches on the FP64
// tile), so the two modes bin every sample identically.  The cross-section weight of a sample is evaluated in single
// precision: a second copy of the tile holds u0 - center as floats (center = u of the walker's mean energy: a few cm
// at most), dt = u0f - (float)(left - delta - center), and a degree-3 polynomial from float records fitted to the
// degree-7 ones at tof_create (fit + rounding <= 1e-7 relative).  A cell's sum is n*c0 in FP64 plus the float sum of
// the dt-dependent part; everything after phase 1 (normalisation, np.rint, flight times, density, timing response,
// likelihood) is the FP64 code.  Contract: 1e-4 relative on lnprob (BASELINE.json north_star).
__device__ __forceinline__ void poly_trip4_f32(float &acc, unsigned addr, int r, float thr, float c1, float c2, float c3) {
    asm("{\n\t"
        ".reg .pred p0, p1, p2, p3;\n\t"
        ".reg .f32 t0, t1, t2, t3, q0, q1, q2, q3;\n\t"
        "ld.shared.f32 t0, [%1];\n\t"
        "ld.shared.f32 t1, [%1+4];\n\t"
        "ld.shared.f32 t2, [%1+8];\n\t"
        "ld.shared.f32 t3, [%1+12];\n\t"
        "setp.gt.s32 p0, %2, 0;\n\t"
        "setp.gt.s32 p1, %2, 1;\n\t"
        "setp.gt.s32 p2, %2, 2;\n\t"
        "setp.gt.s32 p3, %2, 3;\n\t"
        "sub.rn.f32 t0, t0, %3;\n\t"
        "sub.rn.f32 t1, t1, %3;\n\t"
        "sub.rn.f32 t2, t2, %3;\n\t"
        "sub.rn.f32 t3, t3, %3;\n\t"
        "fma.rn.f32 q0, %6, t0, %5;\n\t"
        "fma.rn.f32 q1, %6, t1, %5;\n\t"
        "fma.rn.f32 q2, %6, t2, %5;\n\t"
        "fma.rn.f32 q3, %6, t3, %5;\n\t"
        "fma.rn.f32 q0, q0, t0, %4;\n\t"
        "fma.rn.f32 q1, q1, t1, %4;\n\t"
        "fma.rn.f32 q2, q2, t2, %4;\n\t"
        "fma.rn.f32 q3, q3, t3, %4;\n\t"
        "@p0 fma.rn.f32 %0, q0, t0, %0;\n\t"
        "@p1 fma.rn.f32 %0, q1, t1, %0;\n\t"
        "@p2 fma.rn.f32 %0, q2, t2, %0;\n\t"
        "@p3 fma.rn.f32 %0, q3, t3, %0;\n\t"
        "}"
        : "+f"(acc)
        : "r"(addr), "r"(r), "f"(thr), "f"(c1), "f"(c2), "f"(c3));
}

// Phase 1 for one tile of sorted u0 values (shared memory): add the cross-section weights of every (draw, row)
// sample to the (x,E) histogram H.  Called by all threads of the CTA (contains barriers).
template <int NT, int P, bool F32 = false>
// `brk`: the ends of all T2 intervals (shared memory) for interval searches; `rec`: the shared-memory copy of records
// jbase.. used by the tasks; H has `hstride` bins per row; row i starts at E-bin hlo[i] (banded layout; hlo == nullptr:
// every row starts at bin 0).
__device__ __forceinline__ void range_accumulate_tile(const double *u0, int nt, const double *brk, const double *rec, int jbase,
                                                      const unsigned short *lut, unsigned short *ulut, int n_ulut,
                                                      const double *sdelta, int *srow, double *H, int hstride, const int *hlo,
                                                      int X, int M, double umax, double lut_inv, int lut_n, int &bin_lo_all,
                                                      int &bin_hi_all, const float *u0f = nullptr, double center = 0.0) {
    // F32: `rec` points at float records (RANGE_RWF floats each), `u0f` is the float tile (u0 - center)
    constexpr int RW = P + 3;
    const float *recf = reinterpret_cast<const float *>(rec);
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
    // The draw-range searches start from a lookup cell taken a hair low (tu_bias cells) so that only a forward walk is
    // needed; that is sound while tu_bias * cell width (>= 1.5e-11 cm) dwarfs the rounding of v (< 1e-12 cm for
    // u < 1000 cm).  Narrower tiles (degenerate spread) are walked from their first draw.
    constexpr double tu_bias = 1.0 / 65536.0;
    const double tu_inv = (tu_max - tu_min > 1e-6 * (double)n_ulut) ? (double)n_ulut / (tu_max - tu_min) : 0.0;
    // per-tile lookup: ulut[c] = first draw whose cell (uniform in u between the tile's extremes) is >= c.  The draws
    // are sorted, so draw d owns the cells after its predecessor's up to its own: a scatter, no searches.  (A narrow
    // tile, tu_inv == 0, only ever looks at cell 0.)
    for (int d = v_lo + tid; d < v_hi; d += NT) {
        int c1 = (int)((u0[d] - tu_min) * tu_inv);
        c1 = c1 > n_ulut - 1 ? n_ulut - 1 : c1;
        int c0 = -1;
        if (d > v_lo) {
            c0 = (int)((u0[d - 1] - tu_min) * tu_inv);
            c0 = c0 > n_ulut - 1 ? n_ulut - 1 : c0;
        }
        TOF_CHECK(c1 >= 0 && c1 < n_ulut && c0 >= -1 && c0 <= c1);
        for (int c = c0 + 1; c <= c1; ++c) ulut[c] = (unsigned short)d;
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
    TOF_CHECK(band_lo >= jbase && band_lo <= band_hi && band_hi < M);
    if (F32) {
        bin_lo_all = min(bin_lo_all, __float_as_int(recf[(band_lo - jbase) * RANGE_RWF + 4]));
        bin_hi_all = max(bin_hi_all, __float_as_int(recf[(band_hi - jbase) * RANGE_RWF + 4]));
    } else {
        bin_lo_all = min(bin_lo_all, __double2loint(rec[(band_lo - jbase) * RW + 1]));
        bin_hi_all = max(bin_hi_all, __double2loint(rec[(band_hi - jbase) * RW + 1]));
    }
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
    // `right` of the last (closed) interval: v > u_max  <=>  v >= nextafter(u_max)
    const double umax_next = __longlong_as_double(__double_as_longlong(umax) + 1);
    const unsigned u0_s32 = (unsigned)__cvta_generic_to_shared(u0);
    const unsigned u0f_s32 = F32 ? (unsigned)__cvta_generic_to_shared(u0f) : 0u;
    // `ncell` consecutive intervals j0, j0+1, .. of one row: the end of one run is the start of the next, so every
    // further cell costs one search instead of two.
    auto do_cell = [&](int row, int j0, bool row_ok, int piece, int nch, auto ncell_c) {
        constexpr int ncell = decltype(ncell_c)::value;
        const double delta = sdelta[row];
        int carry = -1;                                  // first draw beyond the previous interval of this lane
        for (int q = 0; q < ncell; ++q) {
            int j = j0 + q;
            const bool active = row_ok && j >= band_lo && j <= band_hi;
            j = j < band_lo ? band_lo : (j > band_hi ? band_hi : j);
            TOF_CHECK(j - jbase >= 0 && (j == 0 || j - 1 - jbase >= 0 || F32));   // staged records cover the band
            double left, right;
            int bin;
            bool shared_cell;
            double a[F32 ? 1 : P + 1];                     // FP64: a0..aP; FP32: a0 only (FP64 copy in the float record)
            float cf1 = 0.0f, cf2 = 0.0f, cf3 = 0.0f;
            if (F32) {
                const float *rj = recf + (j - jbase) * RANGE_RWF;
                const float4 c4 = *reinterpret_cast<const float4 *>(rj);
                const int2 hb = *reinterpret_cast<const int2 *>(rj + 4);
                left = j ? brk[j - 1] : 0.0;              // brk[j] = break that ends interval j (+inf for the last)
                right = (j == M - 1) ? umax_next : brk[j];
                a[0] = *reinterpret_cast<const double *>(rj + 6);
                cf1 = c4.y;
                cf2 = c4.z;
                cf3 = c4.w;
                bin = hb.x;
                shared_cell = hb.y != 0;
            } else {
                const double2 *rj = reinterpret_cast<const double2 *>(rec + (j - jbase) * RW);
                const double2 hd = rj[0];
                left = j ? rec[(j - 1 - jbase) * RW] : 0.0;
                right = (j == M - 1) ? umax_next : hd.x;
                bin = __double2loint(hd.y);
                shared_cell = __double2hiint(hd.y) < 0;
#pragma unroll
                for (int k = 0; k <= P; k += 2) {
                    const double2 c2 = rj[1 + (k >> 1)];
                    a[k] = c2.x;
                    a[F32 ? 0 : k + 1] = c2.y;
                }
            }
            int lb = 0, n = 0;
            double off = 0.0;
            if (active) {
                off = delta - left;
                // first draw with v >= left.  The lookup cell is taken a hair low (tu_bias cells), so the cell's
                // first draw can only be at or before the answer: one forward walk, no backward fix-up.
                if (carry >= 0) {
                    lb = carry;
                } else {
                    int c = (int)fma(left - delta - tu_min, tu_inv, -tu_bias);
                    c = c < 0 ? 0 : (c > n_ulut - 1 ? n_ulut - 1 : c);
                    lb = ulut[c];
                    while (lb < v_hi && !(__dadd_rn(u0[lb], delta) >= left)) ++lb;
                }
                // first draw beyond the interval: v >= right
                int c = (int)fma(right - delta - tu_min, tu_inv, -tu_bias);
                c = c < 0 ? 0 : (c > n_ulut - 1 ? n_ulut - 1 : c);
                int ub = ulut[c];
                ub = ub < lb ? lb : ub;
                while (ub < v_hi && !(__dadd_rn(u0[ub], delta) >= right)) ++ub;
                carry = ub;
                TOF_CHECK(lb >= v_lo && lb <= ub && ub <= v_hi && v_hi <= nt);
                if (nch > 1) {
                    const int len = (ub - lb + nch - 1) / nch;
                    lb += piece * len;
                    ub = (lb + len < ub) ? lb + len : ub;
                }
                n = ub > lb ? ub - lb : 0;
            } else {
                carry = -1;
            }
            // sum_d poly(dt_d) = n*a0 + sum_d dt_d*q(dt_d): all lanes run to the longest run of the warp, four
            // samples per trip, lanes past their own run predicated off (no divergent loop, no remainder loops).
            // A finished lane reads u0[0..3] instead (always inside the tile buffer) and discards the result.
            const int nmax = __reduce_max_sync(FULL, n);
            double acc = 0.0;
            if constexpr (F32) {
                // dt in single precision from the float tile; the n*a0 term stays in FP64
                const float thr = (float)(left - delta - center);
                const unsigned p32 = u0f_s32 + (unsigned)lb * 4u;
                float accf = 0.0f;
                for (int i = 0; i < nmax; i += 4) {
                    const int r = n - i;
                    poly_trip4_f32(accf, r > 0 ? p32 + (unsigned)i * 4u : u0f_s32, r, thr, cf1, cf2, cf3);
                }
                acc = (double)accf;
            } else {
                const unsigned p32 = u0_s32 + (unsigned)lb * 8u;
                for (int i = 0; i < nmax; i += 4) {
                    const int r = n - i;
                    if constexpr (!F32) poly_trip4<P>(acc, r > 0 ? p32 + (unsigned)i * 8u : u0_s32, r, off, a);
                }
            }
            if (n == 0) continue;
            acc = fma((double)n, a[0], acc);
            const int col = bin - (hlo ? hlo[row] : 0);
            TOF_CHECK((unsigned)col < (unsigned)hstride && row >= 0 && row < X);
            if ((unsigned)col >= (unsigned)hstride) continue;     // cannot happen: the band has an interval of slack
            double *cell = H + (size_t)row * hstride + col;
            if (nch > 1 || shared_cell) atomicAdd(cell, acc);    // shared cell: pieces / bin split over intervals
            else *cell += acc;
        }
    };
    const int GfD = Gf > 0 ? Gf : 1;
    const int n_tasks = nA + nB;
    constexpr std::integral_constant<int, 1> one_c{};
    // few tasks (a tile of a big draw set spans few intervals): split every run so that all warps have work
    int nch = 1;
    if (n_tasks < 2 * NW) {
        nch = (RANGE_SPLIT * NW + n_tasks - 1) / (n_tasks > 0 ? n_tasks : 1);   // ~RANGE_SPLIT pieces per warp
        nch = nch > 64 ? 64 : nch;
    }
    if (nch == 1) {
        // type-A tasks take up to RANGE_PAIR consecutive intervals per lane (measured: 1 -> 2 +3 %, 2 -> 4 +1 %).  Static striding over the tasks (a shared work
        // counter with heaviest-first order was measured slower); (jj, g) of a type-A task without a division in the
        // loop
        auto run_tasks = [&](auto pair_c) {
            constexpr int pair = decltype(pair_c)::value;
            const int n_pr = (n_iv + pair - 1) / pair;
            const int nA2 = n_pr * Gf;
            int a_jj = warp / GfD, a_g = warp - a_jj * GfD;
            const int step_j = NW / GfD, step_g = NW - step_j * GfD;
            for (int task = warp; task < nA2 + nB; task += NW) {
                if (task < nA2) {
                    const int row = (a_g << 5) + lane;
                    do_cell(row, k_lo + a_jj * pair + (srow[row] - s_ref), true, 0, 1, pair_c);
                    a_jj += step_j;
                    a_g += step_g;
                    if (a_g >= GfD) {
                        a_g -= GfD;
                        ++a_jj;
                    }
                } else {
                    const int isub = lane / R;
                    const int row = (Gf << 5) + (lane - isub * R);
                    do_cell(row, k_lo + (task - nA2) * per_b + isub + (srow[row] - s_ref), isub < per_b, 0, 1, one_c);
                }
            }
        };
        if (nA >= 6 * NW) run_tasks(std::integral_constant<int, RANGE_PAIR>{});   // enough tasks left for every warp
        else run_tasks(one_c);
    } else {
        for (int t2 = warp; t2 < n_tasks * nch; t2 += NW) {
            const int task = t2 / nch, piece = t2 - task * nch;
            if (task < nA) {
                const int jj = task / GfD;
                const int row = ((task - jj * GfD) << 5) + lane;
                do_cell(row, k_lo + jj + (srow[row] - s_ref), true, piece, nch, one_c);
            } else {
                const int isub = lane / R;
                const int row = (Gf << 5) + (lane - isub * R);
                do_cell(row, k_lo + (task - nA) * per_b + isub + (srow[row] - s_ref), isub < per_b, piece, nch, one_c);
            }
        }
    }
}

// ================================================================================================
// Phase 1, planned form (FP64, a single tile of sorted draws, one T2 interval per E-bin)
// ================================================================================================
// The general routine above finds a cell's draws and sums its polynomial in one loop; with the search state, the
// record and the polynomial all live at once the compiler runs out of registers at 64 per thread (spills, the four
// Horner chains of a trip serialised).  When every E-bin is exactly one T2 interval (interval j == bin j: no
// cross-section knot falls inside a bin) and the walker's draws are one tile, the two jobs separate cleanly and the
// cell histogram itself carries the hand-over:
//   plan     lane = interval, consecutive lanes = consecutive intervals of one row.  Lane finds the first draw at or
//            beyond the left edge of its interval (per-tile lookup + short forward walk); the neighbour's answer is
//            the end of its run (shuffle).  (first draw, count) is written INTO the cell's 8-byte slot of H.
//   execute  lane = row, 32 rows of one trajectory-aligned interval offset (runs of similar length).  Lane reads its
//            slot, loads the polynomial, sums its run -- four samples per trip, warp-uniform trip count, the four
//            Horner chains interleaved -- and overwrites the slot with the sum.
// Every cell is planned by one lane and summed by one lane: no atomics, fixed summation order, same membership rule
// (RN(u0 + delta) >= edge) as the general routine, so both produce the same cells from the same draws.
template <int P>
__device__ __forceinline__ void poly_run4(double &acc, unsigned addr, int r, double off, const double (&a)[P + 1]) {
    static_assert(P == 7, "poly_run4 is written for degree-7 records");
    // Samples past the lane's run (k >= r) get t = (-off) + off = 0 and add q(0)*0 = 0: the loads are predicated (a lane
    // past its run keeps -off), the arithmetic is not -- no select before or after the Horner chains.
    asm("{\n\t"
        ".reg .pred p0, p1, p2, p3;\n\t"
        ".reg .f64 t0, t1, t2, t3, q0, q1, q2, q3;\n\t"
        "setp.gt.s32 p0, %2, 0;\n\t"
        "setp.gt.s32 p1, %2, 1;\n\t"
        "setp.gt.s32 p2, %2, 2;\n\t"
        "setp.gt.s32 p3, %2, 3;\n\t"
        "mov.f64 t0, %11;\n\t"
        "mov.f64 t1, %11;\n\t"
        "mov.f64 t2, %11;\n\t"
        "mov.f64 t3, %11;\n\t"
        "@p0 ld.shared.f64 t0, [%1];\n\t"
        "@p1 ld.shared.f64 t1, [%1+8];\n\t"
        "@p2 ld.shared.f64 t2, [%1+16];\n\t"
        "@p3 ld.shared.f64 t3, [%1+24];\n\t"
        "add.rn.f64 t0, t0, %3;\n\t"
        "add.rn.f64 t1, t1, %3;\n\t"
        "add.rn.f64 t2, t2, %3;\n\t"
        "add.rn.f64 t3, t3, %3;\n\t"
        "fma.rn.f64 q0, %10, t0, %9;\n\t"
        "fma.rn.f64 q1, %10, t1, %9;\n\t"
        "fma.rn.f64 q2, %10, t2, %9;\n\t"
        "fma.rn.f64 q3, %10, t3, %9;\n\t"
        "fma.rn.f64 q0, q0, t0, %8;\n\t"
        "fma.rn.f64 q1, q1, t1, %8;\n\t"
        "fma.rn.f64 q2, q2, t2, %8;\n\t"
        "fma.rn.f64 q3, q3, t3, %8;\n\t"
        "fma.rn.f64 q0, q0, t0, %7;\n\t"
        "fma.rn.f64 q1, q1, t1, %7;\n\t"
        "fma.rn.f64 q2, q2, t2, %7;\n\t"
        "fma.rn.f64 q3, q3, t3, %7;\n\t"
        "fma.rn.f64 q0, q0, t0, %6;\n\t"
        "fma.rn.f64 q1, q1, t1, %6;\n\t"
        "fma.rn.f64 q2, q2, t2, %6;\n\t"
        "fma.rn.f64 q3, q3, t3, %6;\n\t"
        "fma.rn.f64 q0, q0, t0, %5;\n\t"
        "fma.rn.f64 q1, q1, t1, %5;\n\t"
        "fma.rn.f64 q2, q2, t2, %5;\n\t"
        "fma.rn.f64 q3, q3, t3, %5;\n\t"
        "fma.rn.f64 q0, q0, t0, %4;\n\t"
        "fma.rn.f64 q1, q1, t1, %4;\n\t"
        "fma.rn.f64 q2, q2, t2, %4;\n\t"
        "fma.rn.f64 q3, q3, t3, %4;\n\t"
        "fma.rn.f64 %0, q0, t0, %0;\n\t"
        "fma.rn.f64 %0, q1, t1, %0;\n\t"
        "fma.rn.f64 %0, q2, t2, %0;\n\t"
        "fma.rn.f64 %0, q3, t3, %0;\n\t"
        "}"
        : "+d"(acc)
        : "r"(addr), "r"(r), "d"(off), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(-off));
}

// Four samples, all of them inside the lane's run: no predicates at all.
template <int P>
__device__ __forceinline__ void poly_full4(double &acc, unsigned addr, double off, const double (&a)[P + 1]) {
    static_assert(P == 7, "poly_full4 is written for degree-7 records");
    asm("{\n\t"
        ".reg .f64 t0, t1, t2, t3, q0, q1, q2, q3;\n\t"
        "ld.shared.f64 t0, [%1];\n\t"
        "ld.shared.f64 t1, [%1+8];\n\t"
        "ld.shared.f64 t2, [%1+16];\n\t"
        "ld.shared.f64 t3, [%1+24];\n\t"
        "add.rn.f64 t0, t0, %2;\n\t"
        "add.rn.f64 t1, t1, %2;\n\t"
        "add.rn.f64 t2, t2, %2;\n\t"
        "add.rn.f64 t3, t3, %2;\n\t"
        "fma.rn.f64 q0, %9, t0, %8;\n\t"
        "fma.rn.f64 q1, %9, t1, %8;\n\t"
        "fma.rn.f64 q2, %9, t2, %8;\n\t"
        "fma.rn.f64 q3, %9, t3, %8;\n\t"
        "fma.rn.f64 q0, q0, t0, %7;\n\t"
        "fma.rn.f64 q1, q1, t1, %7;\n\t"
        "fma.rn.f64 q2, q2, t2, %7;\n\t"
        "fma.rn.f64 q3, q3, t3, %7;\n\t"
        "fma.rn.f64 q0, q0, t0, %6;\n\t"
        "fma.rn.f64 q1, q1, t1, %6;\n\t"
        "fma.rn.f64 q2, q2, t2, %6;\n\t"
        "fma.rn.f64 q3, q3, t3, %6;\n\t"
        "fma.rn.f64 q0, q0, t0, %5;\n\t"
        "fma.rn.f64 q1, q1, t1, %5;\n\t"
        "fma.rn.f64 q2, q2, t2, %5;\n\t"
        "fma.rn.f64 q3, q3, t3, %5;\n\t"
        "fma.rn.f64 q0, q0, t0, %4;\n\t"
        "fma.rn.f64 q1, q1, t1, %4;\n\t"
        "fma.rn.f64 q2, q2, t2, %4;\n\t"
        "fma.rn.f64 q3, q3, t3, %4;\n\t"
        "fma.rn.f64 q0, q0, t0, %3;\n\t"
        "fma.rn.f64 q1, q1, t1, %3;\n\t"
        "fma.rn.f64 q2, q2, t2, %3;\n\t"
        "fma.rn.f64 q3, q3, t3, %3;\n\t"
        "fma.rn.f64 %0, q0, t0, %0;\n\t"
        "fma.rn.f64 %0, q1, t1, %0;\n\t"
        "fma.rn.f64 %0, q2, t2, %0;\n\t"
        "fma.rn.f64 %0, q3, t3, %0;\n\t"
        "}"
        : "+d"(acc)
        : "r"(addr), "d"(off), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]));
}

// Execute half of the planned form (see range_tile_planned below): a function of its own so that the polynomial
// loop is register-allocated without the planner's state.  `hlo` is never null here (all zeros for the full-size
// launch).  Work split: a warp keeps ONE group of 32 rows for the whole tile (row-dependent values are loaded once)
// and walks the interval offsets of that group with a stride; the X % 32 leftover rows get a warp of their own that
// packs R rows x (32/R) interval offsets per visit.
template <int NT, int P>
__device__ __noinline__ void range_exec_cells(const double *u0, const double *brk, const double *rec, int jbase,
                                              const double *sdelta, const int *srow, double *H, int hstride, const int *hlo,
                                              int X, int band_lo, int band_hi) {
    __builtin_assume(__isShared(u0));
    __builtin_assume(__isShared(brk));
    __builtin_assume(__isShared(rec));
    __builtin_assume(__isShared(sdelta));
    __builtin_assume(__isShared(srow));
    __builtin_assume(__isShared(H));
    __builtin_assume(__isShared(hlo));
    constexpr int RW = P + 3;
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Gf = X >> 5, R = X & 31;
    const int s_ref = srow[0];
    const int s_b = srow[X - 1] - s_ref;               // the shift is monotone in the row index
    const int s_min = s_b < 0 ? s_b : 0, s_max = s_b < 0 ? 0 : s_b;
    const int k_lo = band_lo - s_max;
    const int n_iv = (band_hi - s_min) - k_lo + 1;
    const int per_b = R ? 32 / R : 1;
    const int nB = R ? (n_iv + per_b - 1) / per_b : 0;
    const unsigned u0_s32 = (unsigned)__cvta_generic_to_shared(u0);
    // warps for the leftover rows: in proportion to their share of the visits, at least one when there are any
    int wB = 0;
    if (R) {
        wB = Gf ? (NW * nB + (n_iv * Gf + nB) / 2) / (n_iv * Gf + nB) : NW;
        wB = wB < 1 ? 1 : (wB > NW - 1 && Gf ? NW - 1 : wB);
    }
    const int wA = NW - wB;                            // warps that own a full group of 32 rows
    // one (row, interval) cell of this lane: sum its run and overwrite the slot
    auto cell = [&](int j, bool ok, double delta, double *Hrow /* row base minus its first E-bin */, int row_lo) {
        const int col = j - row_lo;
        const bool active = ok && j >= band_lo && j <= band_hi && (unsigned)col < (unsigned)hstride;
        int lb = 0, n = 0;
        if (active) {
            const long long dn = *reinterpret_cast<const long long *>(Hrow + j);   // 0.0 (untouched) reads as n == 0
            lb = (int)(unsigned)dn;
            n = (int)(dn >> 32);
        }
        const int nmax = __reduce_max_sync(FULL, n);
        if (nmax == 0) return;                         // uniform
        const int nmin = __reduce_min_sync(FULL, n);
        double a[P + 1];
        double off = 0.0;
        const double *rj = rec + (n > 0 ? (j - jbase) * RW + 2 : 2);
        if (n > 0) {
            TOF_CHECK(j - jbase >= 0 && lb >= 0);
            const double2 *r2 = reinterpret_cast<const double2 *>(rj);
            a[1] = r2[0].y;
#pragma unroll
            for (int k = 2; k <= P; k += 2) {
                const double2 c2 = r2[k >> 1];
                a[k] = c2.x;
                a[k + 1] = c2.y;
            }
            off = delta - (j ? brk[j - 1] : 0.0);
        } else {
#pragma unroll
            for (int k = 1; k <= P; ++k) a[k] = 0.0;
        }
        a[0] = 0.0;                                    // the constant term is added once per run, after the loop
        double acc = 0.0;
        unsigned addr = u0_s32 + (unsigned)lb * 8u;
        const int tfull = nmin >> 2;                   // trips in which every lane still has four samples
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
        if (n > 0) Hrow[j] = fma((double)n, rj[0], acc);
    };
    if (warp < wA) {
        if (Gf <= wA) {                                // the usual case: this warp keeps one group of rows
            const int g = warp % Gf, idx = warp / Gf;
            const int cnt = (wA - g + Gf - 1) / Gf;    // warps sharing group g
            const int row = (g << 5) + lane;
            const int row_lo = hlo[row];
            const double delta = sdelta[row];
            double *Hrow = H + (size_t)row * hstride - row_lo;
            const int jrow = k_lo + (srow[row] - s_ref);
            for (int jj = idx; jj < n_iv; jj += cnt) cell(jrow + jj, true, delta, Hrow, row_lo);
        } else {                                       // more groups than warps: stride over (offset, group) pairs
            for (int task = warp; task < n_iv * Gf; task += wA) {
                const int jj = task / Gf;
                const int row = ((task - jj * Gf) << 5) + lane;
                const int row_lo = hlo[row];
                cell(k_lo + jj + (srow[row] - s_ref), true, sdelta[row], H + (size_t)row * hstride - row_lo, row_lo);
            }
        }
    } else {
        const int isub = lane / R;
        const int row = (Gf << 5) + (lane - isub * R);
        const int row_lo = hlo[row];
        const double delta = sdelta[row];
        double *Hrow = H + (size_t)row * hstride - row_lo;
        const int jrow = k_lo + isub + (srow[row] - s_ref);
        for (int tb = warp - wA; tb < nB; tb += wB) cell(jrow + tb * per_b, isub < per_b, delta, Hrow, row_lo);
    }
}

// Called by all threads of the CTA (contains barriers).  H must be zero where no draw can land (the kernel clears it
// per walker); `rec` holds the staged records jbase.. (coefficients at +2 doubles), `brk` the ends of all intervals.
// __noinline__ on purpose: the call is a register-allocation firewall.  Inlined into the kernel, the routine competes
// with the kernel's own live state for 64 registers and ptxas serialises the Horner chains; as a function it is
// allocated on its own (the caller's live registers are saved once per tile).
// EXEC = false: plan only; the caller runs range_exec_cells itself with the band returned in bin_lo_all / bin_hi_all
// (returns false when the tile has nothing to execute).  A callee is register-allocated in what its callers leave
// free, so the hot loop wants to be called from the leanest frame available (adv_planned_kernel calls it directly).
template <int NT, int P, bool EXEC = true>
__device__ __noinline__ bool range_tile_planned(const double *u0, int nt, const double *brk, const double *rec, int jbase,
                                                   const unsigned short *lut, unsigned short *ulut, int n_ulut,
                                                   const double *sdelta, int *srow, double *H, int hstride, const int *hlo,
                                                   int X, int M, double umax, double lut_inv, int lut_n, int &bin_lo_all,
                                                   int &bin_hi_all) {
    // every pointer is a shared-memory address (the call boundary hides that from the compiler: without the hints it
    // emits generic LD.E / ST.E instead of LDS / STS)
    __builtin_assume(__isShared(u0));
    __builtin_assume(__isShared(brk));
    __builtin_assume(__isShared(rec));
    __builtin_assume(__isShared(lut));
    __builtin_assume(__isShared(ulut));
    __builtin_assume(__isShared(sdelta));
    __builtin_assume(__isShared(srow));
    __builtin_assume(__isShared(H));
    __builtin_assume(__isShared(hlo));
    constexpr int RW = P + 3;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsteps = 32 - __clz(nt);
    int v_lo = 0, v_hi = nt;                          // finite part of the sorted tile: -inf first, +inf last
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
    if (v_hi <= v_lo) return false;                   // uniform: no usable draw
    const double tu_min = u0[v_lo], tu_max = u0[v_hi - 1];
    constexpr double tu_bias = 1.0 / 65536.0;         // lookup cell taken a hair low: forward walks only (see above)
    const double tu_inv = (tu_max - tu_min > 1e-6 * (double)n_ulut) ? (double)n_ulut / (tu_max - tu_min) : 0.0;
    for (int d = v_lo + tid; d < v_hi; d += NT) {     // ulut[c] = first draw whose lookup cell is >= c (scatter)
        int c1 = (int)((u0[d] - tu_min) * tu_inv);
        c1 = c1 > n_ulut - 1 ? n_ulut - 1 : c1;
        int c0 = -1;
        if (d > v_lo) {
            c0 = (int)((u0[d - 1] - tu_min) * tu_inv);
            c0 = c0 > n_ulut - 1 ? n_ulut - 1 : c0;
        }
        TOF_CHECK(c1 >= 0 && c1 < n_ulut && c0 >= -1 && c0 <= c1);
        for (int c = c0 + 1; c <= c1; ++c) ulut[c] = (unsigned short)d;
    }
    {                                                 // trajectory alignment: interval of the median draw per row
        const double u_med = u0[(v_lo + v_hi) >> 1];
        for (int i = tid; i < X; i += NT) {
            double vm = __dadd_rn(u_med, sdelta[i]);
            vm = vm < 0.0 ? 0.0 : (vm > umax ? umax : vm);
            srow[i] = range_interval(vm, brk, lut, lut_inv, lut_n, M);
        }
    }
    double dmin = sdelta[0], dmax = sdelta[0];
    {
        const double dl = sdelta[X - 1];
        dmin = dl < dmin ? dl : dmin;
        dmax = dl > dmax ? dl : dmax;
    }
    const double vmin = __dadd_rn(tu_min, dmin), vmax = __dadd_rn(tu_max, dmax);
    __syncthreads();
    if (!(vmax >= 0.0) || vmin > umax) return false;  // uniform
    const int band_lo = range_interval(vmin > 0.0 ? vmin : 0.0, brk, lut, lut_inv, lut_n, M);
    const int band_hi = range_interval(vmax < umax ? vmax : umax, brk, lut, lut_inv, lut_n, M);
    TOF_CHECK(band_lo >= jbase && band_lo <= band_hi && band_hi < M);
    bin_lo_all = min(bin_lo_all, band_lo);            // interval j == E-bin j on this path
    bin_hi_all = max(bin_hi_all, band_hi);
    const double umax_next = __longlong_as_double(__double_as_longlong(umax) + 1);

    // ---- plan: (first draw, count) of every cell into its slot of H ------------------------------------------
    for (int row = warp; row < X; row += NW) {
        const double delta = sdelta[row];
        const int row_lo = hlo[row];
        int ja = row_lo > band_lo ? row_lo : band_lo;
        int jb = row_lo + hstride - 1;
        jb = jb < band_hi ? jb : band_hi;
        long long *Hrow = reinterpret_cast<long long *>(H + (size_t)row * hstride) - row_lo;
        for (int j0 = ja; j0 <= jb; j0 += 31) {       // lanes 0..30 own a cell each, lane 31 supplies the last edge
            int j = j0 + lane;
            j = j <= jb + 1 ? j : jb + 1;
            // left edge of interval j (the right edge of the last, closed, interval is nextafter(u_max))
            const double edge = (j == 0) ? 0.0 : (j >= M ? umax_next : brk[j - 1]);
            int c = (int)fma(edge - delta - tu_min, tu_inv, -tu_bias);
            c = c < 0 ? 0 : (c > n_ulut - 1 ? n_ulut - 1 : c);
            int d = ulut[c];
            while (d < v_hi && !(__dadd_rn(u0[d], delta) >= edge)) ++d;
            const int d_next = __shfl_down_sync(FULL, d, 1);
            const int n = d_next - d;
            TOF_CHECK(d >= v_lo && d <= v_hi && (lane == 31 || n >= 0));
            if (lane < 31 && j0 + lane <= jb && n > 0) {
                TOF_CHECK(j - row_lo >= 0 && j - row_lo < hstride && j - jbase >= 0);
                Hrow[j] = (long long)(unsigned)d | ((long long)n << 32);
            }
        }
    }
    __syncthreads();
    if constexpr (EXEC) range_exec_cells<NT, P>(u0, brk, rec, jbase, sdelta, srow, H, hstride, hlo, X, band_lo, band_hi);
    return true;
}

template <int NT, int P, bool F32 = false, bool PROF = false>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) adv_range_kernel(const DevModel m, const DevRun run, const double *__restrict__ theta,
                                                       long long n_walkers, ModelOut out) {
    extern __shared__ __align__(16) unsigned char smem_sym[];
    // The base of dynamic shared memory is made opaque (but still known to be shared): the compiler otherwise
    // re-derives it from the symbol (S2UR CgaCtaId + 4 uniform ops) at every use inside the hot loops instead of
    // keeping it in a register.
    unsigned char *smem_raw = smem_sym;
    asm volatile("" : "+l"(smem_raw));
    __builtin_assume(__isShared(smem_raw));
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
    unsigned char *pa = smem_raw + out.lay.pa;
    unsigned int *tofc = reinterpret_cast<unsigned int *>(pa);
    double *u0 = reinterpret_cast<double *>(pa);                       // aliases tofc (phase 1 only)
    double *rec = reinterpret_cast<double *>(smem_raw + out.lay.rec);
    double *svd = reinterpret_cast<double *>(smem_raw + out.lay.svd);
    double *staps = reinterpret_cast<double *>(smem_raw + out.lay.staps);
    double *scratch = reinterpret_cast<double *>(smem_raw + out.lay.scratch);
    double *sdelta = reinterpret_cast<double *>(smem_raw + out.lay.sdelta);   // [X] sgn*(x_i - x_start)
    unsigned short *lut = reinterpret_cast<unsigned short *>(smem_raw + out.lay.lut);
    unsigned short *ulut = reinterpret_cast<unsigned short *>(smem_raw + out.lay.ulut);   // [RANGE_ULUT]
    int *srow = reinterpret_cast<int *>(smem_raw + out.lay.srow);       // [X]
    double *rvd = reinterpret_cast<double *>(ulut);                      // [EB] 1/svd, aliases ulut (phases 2-3 only)
    int *hlo_s = reinterpret_cast<int *>(smem_raw + out.lay.hlo);       // [X] first E-bin of each row (banded launch)
    double *sbrk = reinterpret_cast<double *>(smem_raw + out.lay.sbrk); // [M] interval ends
    unsigned short *sbin = reinterpret_cast<unsigned short *>(smem_raw + out.lay.sbin);  // [M] E-bin of each interval
    __shared__ int s_band[3];                                           // widest row, first / last interval of the walker

    // ---- walker-independent tables: staged once per CTA (persistent CTAs loop over walkers) ------------------
    float *rec_f = reinterpret_cast<float *>(rec);        // FP32 mode: float records live in the same region
    if (!banded) {
        if (F32)
            for (int i = tid; i < M * RANGE_RWF; i += NT) rec_f[i] = m.rng_rec_f32[i];
        else
            for (int i = tid; i < M * RW; i += NT) rec[i] = m.rng_rec[i];
    }
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
    if (!banded)
        for (int i = tid; i < X; i += NT) hlo_s[i] = 0;      // full-size launch: every row starts at E-bin 0
    __shared__ long long s_next;
    // stage timing (tof_set_stage_timing; PROF instantiations only, the shipped kernels carry none of this): thread 0
    // charges the SM clock between stage boundaries to the stage that ends there -- it leaves a barrier when the CTA
    // does, so this is CTA time per stage, summed over CTAs
    long long t_mark = PROF ? clock64() : 0;
    auto stage_done = [&](int k) {
        if constexpr (PROF) {
            if (tid == 0) {
                const long long t = clock64();
                atomicAdd(out.stage_cycles + k, (unsigned long long)(t - t_mark));
                t_mark = t;
            }
        }
    };
    for (long long iter = 0;; ++iter) {
    __syncthreads();                                       // the previous walker is done with shared memory
    if (tid == 0)
        s_next = out.work ? (long long)atomicAdd(out.work, 1ull) : (long long)blockIdx.x + iter * (long long)gridDim.x;
    __syncthreads();
    // a work item is (walker, split): with few walkers and a big draw set, `n_split` CTAs share one walker's draws
    const int NS = out.n_split > 1 ? out.n_split : 1;
    const long long item = s_next / NS;
    const int split = (int)(s_next - item * NS);
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
            if (tid == 0 && split == 0) out.queue_out[atomicAdd(out.queue_count, 1ull)] = (int)w;
            continue;
        }
        TOF_CHECK(j_hi_all - jbase + 1 <= out.rcap && j_hi_all < M && jbase >= 0 && X * hstride <= out.hcap);
        if (F32) {
            for (int i = tid; i < (j_hi_all - jbase + 1) * RANGE_RWF; i += NT) rec_f[i] = m.rng_rec_f32[(size_t)jbase * RANGE_RWF + i];
        } else {
            for (int i = tid; i < (j_hi_all - jbase + 1) * RW; i += NT) rec[i] = recf[(size_t)jbase * RW + i];
        }
    }
    // ---- per walker: zero the cell histogram ------------------------------------------------------------------
    for (int i = tid; i < X * hstride; i += NT) H[i] = 0.0;

    stage_done(0);                                         // work fetch, prior, band, record staging, histogram reset
    // ---- phase 1: (x,E) histogram of cross-section weights through the range tables ---------------------
    int bin_lo_all = EB, bin_hi_all = -1;                  // E-bins any draw of any tile can have touched (uniform)
    if (!F32 && m.n_draws >= RANGE_STREAM_MIN) {           // (FP32 contexts with big draw sets are run by the FP64 kernels)
        // Big draw sets: the sorted draws of one interval are hundreds of consecutive values, so a lane that walks
        // draws in order changes interval rarely.  Warp-private streaming, no barriers: a warp takes 128 consecutive
        // draws (4 per lane, T1 evaluated once, kept in registers and broadcast by shuffle) and, for every group of
        // 32 rows, lane = row walks the 128 samples with an interval pointer; runs go to H with atomics.
        __syncthreads();                                   // staging done
        bin_lo_all = 0;
        bin_hi_all = EB - 1;
        const int n_groups = (X + 31) >> 5;
        for (long long base = ((long long)split * NW + warp) * 128; base < m.n_draws; base += (long long)NS * NW * 128) {
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
            if (F32) {
                // FP64 tile for the searches + float copy relative to u(mean energy) for the weights
                float *u0f = rec_f + (size_t)out.rcap * RANGE_RWF;
                double center = t1_eval(e0, m);
                center = (center > -CUDART_INF && center < CUDART_INF) ? center : 0.0;
                for (int d = tid; d < nt; d += NT) {
                    const double u = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 - (tile + d) : tile + d)))), m);
                    u0[d] = u;
                    u0f[d] = (float)(u - center);
                }
                __syncthreads();
                range_accumulate_tile<NT, P, true>(u0, nt, sbrk, rec, jbase, lut, ulut, RANGE_ULUT, sdelta, srow, H, hstride, hlo,
                                                   X, M, umax, m.rng_lut_inv, m.rng_lut_n, bin_lo_all, bin_hi_all, u0f, center);
                continue;
            }
            if (run.fresh) {
                // per-evaluation draws (tof_set_draw_mode; one tile only, checked by the host): this walker's own sorted
                // normals, generated in place; E0 ascends with |spread| (the normal law is symmetric)
                fresh_sorted_normals<NT>(u0, nt, run, w, 0, scratch);
                const double sp = fabs(spread);
                double zv[(RANGE_TILE + NT - 1) / NT];
#pragma unroll
                for (int q = 0; q < (RANGE_TILE + NT - 1) / NT; ++q) zv[q] = (tid + q * NT < nt) ? u0[tid + q * NT] : 0.0;
                __syncthreads();
#pragma unroll
                for (int q = 0; q < (RANGE_TILE + NT - 1) / NT; ++q)
                    if (tid + q * NT < nt) u0[tid + q * NT] = t1_eval(__dadd_rn(e0, __dmul_rn(sp, zv[q])), m);
            } else {
                for (int d = tid; d < nt; d += NT)
                    u0[d] = t1_eval(__dadd_rn(e0, __dmul_rn(spread, __ldg(run.z + (rev ? m.n_draws - 1 - (tile + d) : tile + d)))), m);
            }
            __syncthreads();
            if (m.rng_identity && m.n_draws <= RANGE_TILE)      // one tile, one interval per E-bin: plan / execute
                range_tile_planned<NT, P>(u0, nt, sbrk, rec, jbase, lut, ulut, RANGE_ULUT, sdelta, srow, H, hstride, hlo_s, X, M,
                                          umax, m.rng_lut_inv, m.rng_lut_n, bin_lo_all, bin_hi_all);
            else
                range_accumulate_tile<NT, P>(u0, nt, sbrk, rec, jbase, lut, ulut, RANGE_ULUT, sdelta, srow, H, hstride, hlo, X, M,
                                             umax, m.rng_lut_inv, m.rng_lut_n, bin_lo_all, bin_hi_all);
        }
    }
    __syncthreads();
    stage_done(1);                                         // phase 1: T1, lookup, (row, interval) tasks

    if (NS > 1) {
        // partial histogram -> L2-resident scratch; the CTA that arrives last adds the NS partials in split order
        // (fixed order: the sum does not depend on which CTA finishes last) and carries on alone
        const int ncell = X * hstride;
        double *part_out = out.split_scratch + ((size_t)w * NS + split) * (size_t)out.split_stride;
        for (int i = tid; i < ncell; i += NT) part_out[i] = H[i];
        __threadfence();
        __syncthreads();
        __shared__ int s_last;
        if (tid == 0) s_last = (atomicAdd(out.split_tickets + w, 1u) == (unsigned)(NS - 1));
        __syncthreads();
        if (!s_last) continue;
        __threadfence();
        const double *base_in = out.split_scratch + (size_t)w * NS * (size_t)out.split_stride;
        for (int i = tid; i < ncell; i += NT) {
            double v = 0.0;
            for (int k = 0; k < NS; ++k) v += __ldcg(base_in + (size_t)k * out.split_stride + i);
            H[i] = v;
        }
        if (tid == 0) out.split_tickets[w] = 0u;           // ready for the next call
        __syncthreads();
    }

    // ---- phase 2: normalise (adv:143) ---------------------------------------------------------------------
    for (int i = tid; i < T; i += NT) tofc[i] = 0u;        // u0 is dead now
    for (int j = tid; j < EB; j += NT) {                   // deuteron speeds and their reciprocals (the lookup is dead too)
        const double eff = __ddiv_rn(__dadd_rn(e0, m.e_centers[j]), 2.0);   // adv:151
        const double v = speed_of(m.c, eff, m.m_d);
        svd[j] = v;
        rvd[j] = __ddiv_rn(1.0, v);
    }
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
    stage_done(2);                                         // normalisation sum (+ split merge when several CTAs share a walker)

    // ---- phase 3: quantise (adv:146) and scatter every non-empty cell to its flight time (adv:149-158) ----
    const double t_step = (run.tof_max - run.tof_min) / (double)T;
    const double t_scale = (double)T / (run.tof_max - run.tof_min);
    const double nsamp = (double)m.n_samples;
    const double rS = __ddiv_rn(1.0, S);                    // IEEE quotients below come from this reciprocal (div_by_recip)
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
    __syncthreads();
    stage_done(3);                                         // quantise + scatter to flight times

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
        if (r != r && out.nan_count) atomicAdd(out.nan_count, 1ull);
        if (m.nan_to_neginf && r != r) r = -CUDART_INF;
        out.lnprob[w] = r;
    }
    stage_done(4);                                         // density, timing response, likelihood
    if (PROF && tid == 0) atomicAdd(out.stage_cycles + TOF_N_STAGES, 1ull);
    }   // persistent walker loop
}

}  // namespace tof
