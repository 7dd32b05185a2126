// Device-side building blocks shared by the model kernels (sm_100a).
//
// Everything that feeds an integer decision (a histogram bin, np.rint) is written with explicit
// round-to-nearest intrinsics in the reference's operation order, so that those decisions are
// bit-identical to numpy's; everything else may contract to FMA.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/tofgpu.h"

// Checked build (python -m mcmctoffitting_b200.build --checked -> libtofgpu_checked.so): every shared-memory index
// the range kernels compute is asserted; a violation prints its site and traps (the call then fails with a CUDA
// error).  compute-sanitizer is not available on the GPU pool, so this is how out-of-bounds indexing is hunted:
// the parity suite is run once against the checked library (profiles/).  Compiled out of the shipped library.
#ifdef TOF_CHECKED
#include <cstdio>
#define TOF_CHECK(cond)                                                                                     \
    do {                                                                                                    \
        if (!(cond)) {                                                                                      \
            printf("TOF_CHECK failed: %s at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                       \
            __trap();                                                                                       \
        }                                                                                                   \
    } while (0)
#else
#define TOF_CHECK(cond) ((void)0)
#endif

namespace tof {

constexpr unsigned FULL = 0xffffffffu;

// ---- tables in device memory ------------------------------------------------------------------
struct DevModel {
    int model, ode_mode, ode_substeps, ode_from_zero, prior_strict, nan_to_neginf;
    int ndim, n_runs, x_bins, e_bins, n_taps, conv_shift, n_zero_deg, n_materials, n_xs;
    long long n_samples, n_ev_per_loop, n_loops, n_draws;
    double x_min, x_max, e_min, e_max;
    double c, m_d, m_n, m_he3, q_ddn, cell_length, simple_neutron_base;
    double bethe_A[TOF_MAX_MATERIALS], bethe_B[TOF_MAX_MATERIALS];
    double prior_lo[TOF_MAX_DIM], prior_hi[TOF_MAX_DIM];
    const double *x_centers, *e_centers, *neutron_speed, *neutron_rspeed /* 1/neutron_speed */, *xs_breaks, *xs_coefs, *taps, *zd_times, *zd_weights;
    const unsigned char *xs_lut;
    int xs_lut_n;
    double xs_lut_lo, xs_lut_inv;
    // range-energy tables (TOF_ODE_RANGE), see range_tables.py
    const double *t1_coefs;        // [t1_n][8]
    const double *rng_rec;         // [rng_n][P+3]: next break, bin, a0..aP
    const float *rng_rec_f32;      // FP32 mode: [rng_n][8] floats: c0..c3, bin, shared-cell flag, c0 as a double (adv_range.cuh)
    const unsigned short *rng_lut; // [rng_lut_n]
    int t1_q, t1_key_lo, t1_n, rng_degree, rng_n, rng_lut_n;
    int rng_identity;              // 1: T2 interval j is exactly E-bin j (no cross-section knot inside a bin)
    double rng_sign, rng_u_max, rng_lut_inv, e_tab_lo, e_tab_hi;
    // rank hints of adv_zrank_kernel: rank_theta[j][i] = initial energy (keV) that reaches row i at the lower edge of
    // T2 interval j (j = 0..rng_n), row stride rank_stride floats; walker-independent (adv_zrank.cuh)
    const float *rank_theta;
    int rank_stride;
    // oneBD: spline stopping table, attenuation, causal transit taps
    const double *stop_coefs, *attenuation, *taps2;
    int stop_n, n_taps2;
    int onebd_copies;              // private (x,E) histogram copies per CTA of the oneBD kernel
    double stop_lo, stop_step, beam_energy;
};

struct DevRun {
    int tof_bins;
    double tof_min, tof_max;
    const double *neutron_dist;  // [x_bins]
    const double *z;             // stream 0
    long long n_z;
    const double *z1;            // stream 1
    long long n_z1;
    // per-evaluation draws generated on the device (tof_set_draw_mode): every (call, walker) has its own stream
    int fresh;
    unsigned long long fresh_seed, fresh_epoch;   // epoch: one per model call (or 2*step + half inside the ensemble entry points)
    long long fresh_walker0;                       // global index of the call's first walker
    // adv_zrank_kernel: zlut[c] = first (sorted) draw with z >= zlut_lo + c / zlut_inv, c = 0..ZR_LUT (zlut[ZR_LUT] = n_z)
    const unsigned short *zlut;
    double zlut_lo, zlut_inv;
    const double *obs;           // [tof_bins] (private copy, simult: 0 -> 1 applied)
    const int *obs_nz_idx;       // compacted bins with obs != 0
    const double *obs_nz_val;
    int n_obs_nz;
};

// ---- reductions --------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    return v;
}

// Sum over the CTA, result broadcast to every thread.  scratch: >= 33 elements of T in smem.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T t = (lane < nw) ? scratch[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---- numpy-compatible uniform binning ------------------------------------------------------------
// np.linspace(lo, hi, n+1)[k]: arange(k)*step + start with separately rounded mul and add, last
// element forced to `hi` (numpy/_core/function_base.py).
__device__ __forceinline__ double np_edge(int k, int n, double lo, double hi, double step) {
    return (k == n) ? hi : __dadd_rn(__dmul_rn((double)k, step), lo);
}

// Bin of v in n uniform bins on [lo, hi] with numpy's semantics (np.histogram fast path and
// np.histogramdd's searchsorted give the same answer): edges[k] <= v < edges[k+1], last bin closed,
// NaN / out-of-range -> -1.  `scale` = n/(hi-lo) is only an estimate; the edge compares decide.
__device__ __forceinline__ int np_bin(double v, int n, double lo, double hi, double step, double scale) {
    if (!(v >= lo && v <= hi)) return -1;
    int b = (int)((v - lo) * scale);
    b = b < 0 ? 0 : (b > n - 1 ? n - 1 : b);
    if (v < np_edge(b, n, lo, hi, step)) {
        --b;
    } else if (b != n - 1 && v >= np_edge(b + 1, n, lo, hi, step)) {
        ++b;
    }
    return b;
}

// ---- D(d,n) cross section: piecewise cubic in shared memory (utilities.py:412-429) ---------------
struct XsTab {
    const double *bp;          // [n]
    const double *cf;          // [n-1][4]
    const unsigned char *lut;  // [lut_n]
    int n, lut_n;
    double lut_lo, lut_inv;
};

__device__ __forceinline__ double xs_eval(double E, const XsTab &t) {
    const double lo = t.bp[0], hi = t.bp[t.n - 1];
    if (E <= lo) E = lo;  // utilities.py:425-428 clamps in place
    if (E >= hi) E = hi;
    int c = (int)((E - t.lut_lo) * t.lut_inv);
    c = c < 0 ? 0 : (c > t.lut_n - 1 ? t.lut_n - 1 : c);
    int i = t.lut[c];
    while (i + 2 < t.n && E >= t.bp[i + 1]) ++i;
    while (i > 0 && E < t.bp[i]) --i;
    const double x = E - t.bp[i];
    const double *k = t.cf + 4 * i;
    return ((k[0] * x + k[1]) * x + k[2]) * x + k[3];
}

// ---- Bethe stopping power, reduced form of ionStopping.py:78-97 ------------------------------------
// dE/dx = -(1/E) * sum_k A_k ln(B_k E).  E <= 0 gives NaN, like the reference's sqrt of a negative.
template <int NMAT>
__device__ __forceinline__ double bethe(double E, const double *A, const double *B, int nmat) {
    double s;
    if (NMAT == 1) {
        s = A[0] * log(B[0] * E);
    } else {
        s = 0.0;
        for (int k = 0; k < nmat; ++k) s += A[k] * log(B[k] * E);
    }
    return -s / E;
}

// ---- flight time: utilities.py:64-73 in the reference's operation order ----------------------------
__device__ __forceinline__ double speed_of(double c, double energy, double mass) {
    return __dmul_rn(c, __dsqrt_rn(__ddiv_rn(__dmul_rn(2.0, energy), mass)));
}

// a / b correctly rounded (== __ddiv_rn(a, b)) given y = RN(1/b): product, then two Markstein corrections
// (residual by FMA is exact once q is within an ulp; Markstein 1990).  Five FP64 instructions, no branch -- used where
// many quotients share a divisor (cell / S, distance / speed_j).  b = 0 or non-finite gives NaN instead of inf; the
// callers drop both (np_bin).
__device__ __forceinline__ double div_by_recip(double a, double b, double y) {
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
}

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011) ------------------------------------
struct Philox {
    uint32_t c[4];
    __device__ __forceinline__ Philox(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        c[0] = (uint32_t)ctr_lo;
        c[1] = (uint32_t)(ctr_lo >> 32);
        c[2] = (uint32_t)ctr_hi;
        c[3] = (uint32_t)(ctr_hi >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
            const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
            c[0] = n0;
            c[1] = lo1;
            c[2] = n2;
            c[3] = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
    // two uniforms in [0,1) with 53 random bits each
    __device__ __forceinline__ double u0() const {
        return (double)((((uint64_t)c[1] << 32) | c[0]) >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double u1() const {
        return (double)((((uint64_t)c[3] << 32) | c[2]) >> 11) * (1.0 / 9007199254740992.0);
    }
};

}  // namespace tof

namespace tof {

// ---- per-evaluation Monte-Carlo draws generated on the device (tof_set_draw_mode) ---------------------------------
// The reference draws inside every lnlike call (adv:128, simple:62-64): each walker at each step sees its own noise.
// Philox4x32-10, key = seed, counter = (pair index, epoch << 32 | global walker << 6 | run << 3 | stream): the draws of a
// (call, walker) do not depend on how walkers are batched or sharded.
__device__ __forceinline__ Philox fresh_philox(const DevRun &run, long long walker, int run_idx, int stream, unsigned long long pair) {
    const unsigned long long gw = (unsigned long long)(run.fresh_walker0 + walker);
    return Philox(run.fresh_seed, pair, (run.fresh_epoch << 32) | (gw << 6) | ((unsigned long long)run_idx << 3) | (unsigned long long)stream);
}
// uniform on the open interval (0, 1): (k + 1/2) * 2^-52, k < 2^52 (exact in binary64; never 0 or 1)
__device__ __forceinline__ double philox_open01(uint32_t lo, uint32_t hi) {
    return ((double)((((uint64_t)hi << 32) | lo) >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}
// iid standard normal number d of this (call, walker, run): inverse-CDF of one open uniform.  stream 0: the model's
// draws (adv:128, simultFit.py:244, csi_oneBD.py:438); stream 3: the simultaneous fit's replacement draws (245-252)
__device__ __forceinline__ double fresh_normal(const DevRun &run, long long walker, int run_idx, long long d, int stream = 0) {
    const Philox ph = fresh_philox(run, walker, run_idx, stream, (unsigned long long)d >> 1);
    return normcdfinv((d & 1) ? philox_open01(ph.c[2], ph.c[3]) : philox_open01(ph.c[0], ph.c[1]));
}
// iid uniform [0, 1) number d (stream 1 of the simple model, simple:62)
__device__ __forceinline__ double fresh_uniform(const DevRun &run, long long walker, int run_idx, long long d) {
    const Philox ph = fresh_philox(run, walker, run_idx, 1, (unsigned long long)d >> 1);
    return (d & 1) ? ph.u1() : ph.u0();
}

// nt standard normals in ASCENDING order for this (call, walker), into zs (shared memory), by all NT threads of the CTA:
// the order statistics of nt iid uniforms are S_k / S_{nt+1} with S the running sums of nt + 1 iid exponentials
// (Renyi), so a prefix sum replaces the sort; z_k = Phi^-1(U_(k)).  Exactly the joint law of nt sorted iid normals.
// Needs nt <= 2 * NT; scratch: >= NT/32 + 2 doubles of shared memory.  Ends with a barrier.  The result does not depend
// on NT (two spacings per thread, the same summation order), so the banded and the full-size launch of a call, and
// tof_generate_draws, see the same draws.
template <int NT>
__device__ __forceinline__ void fresh_sorted_normals(double *zs, int nt, const DevRun &run, long long walker, int run_idx,
                                                     double *scratch) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Philox ph = fresh_philox(run, walker, run_idx, 2, (unsigned long long)tid);
    // the (nt+1)-th spacing: a fixed counter, so that every thread -- and every kernel, whatever its NT -- gets the same
    const Philox pl = fresh_philox(run, walker, run_idx, 2, 0xFFFFFFFFull);
    const int i0 = 2 * tid, i1 = 2 * tid + 1;              // this thread's two exponentials (indices 0..nt-1)
    const double ea = (i0 < nt) ? -log(philox_open01(ph.c[0], ph.c[1])) : 0.0;
    const double eb = (i1 < nt) ? -log(philox_open01(ph.c[2], ph.c[3])) : 0.0;
    const double e_last = -log(philox_open01(pl.c[0], pl.c[1]));
    double incl = ea + eb;                                  // inclusive scan of the per-thread sums, fixed order
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += up;
    }
    __syncthreads();                                        // scratch may still be in use by the caller
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    double base = 0.0;
    for (int k = 0; k < warp; ++k) base += scratch[k];
    const double s_a = base + (incl - (ea + eb)) + ea;      // S at index i0
    const double s_b = s_a + eb;                            // S at index i1
    if (i0 == nt - 1) scratch[NW] = s_a;                    // the sum of the first nt spacings
    if (i1 == nt - 1) scratch[NW] = s_b;
    __syncthreads();
    const double total = scratch[NW] + e_last;
    constexpr double TOP = 1.0 - 1.0 / 9007199254740992.0;
    if (i0 < nt) {
        const double u = s_a / total;
        zs[i0] = normcdfinv(u < TOP ? u : TOP);
    }
    if (i1 < nt) {
        const double u = s_b / total;
        zs[i1] = normcdfinv(u < TOP ? u : TOP);
    }
    __syncthreads();
}

}  // namespace tof
