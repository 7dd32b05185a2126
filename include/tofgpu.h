/*
 * tofgpu.h -- C ABI of the B200-native lnprob path for neutron time-of-flight MCMC fitting.
 *
 * Drop-in boundary for ONE path of gcrich/mcmcTOFfitting: the per-walker forward model +
 * log-likelihood (`lnprob`) that emcee evaluates on every step.  The reference is pure Python,
 * so there is no existing FFI; every entry point cites the reference function(s) whose work it
 * replaces (paths relative to the reference tree).  The reference-side binding is the ctypes
 * stub shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * tof_status; the message is available from tof_last_error().  All floating-point data is
 * IEEE-754 binary64.  The library never writes caller memory except the documented outputs.
 * There is no CPU fallback: tof_create() fails when no sm_100 device is present.
 */
#ifndef TOFGPU_H
#define TOFGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOF_ABI_VERSION 3

#define TOF_MAX_DIM 16       /* parameters per walker (reference max: 9, simultFit.py:444-448) */
#define TOF_MAX_RUNS 8       /* simultaneous standoff runs (reference max: 5, simultFit.py:127-131) */
#define TOF_MAX_MATERIALS 8  /* simpleBethe material list (ionStopping.py:71-76) */

typedef enum tof_status {
    TOF_OK = 0,
    TOF_ERR_INVALID = -1,    /* bad argument / inconsistent config */
    TOF_ERR_NO_DEVICE = -2,  /* no CUDA device of compute capability 10.x */
    TOF_ERR_CUDA = -3,       /* CUDA runtime error, see tof_last_error */
    TOF_ERR_STATE = -4,      /* observables / draws not set */
    TOF_ERR_CAPACITY = -5    /* configuration does not fit the device (shared memory) */
} tof_status;

typedef enum tof_model_kind {
    TOF_MODEL_SIMPLE = 1, /* tests/simpleTOFmodel.py:57-120 (and mpiTOFmodel.py:40-128)        */
    TOF_MODEL_ADV = 2,    /* tests/advIntermediateTOFmodel.py:115-199 = intermediateTOFmodel.py */
    TOF_MODEL_SIMULT = 3, /* tests/simultFit.py:223-300, 380-469                                */
    TOF_MODEL_ONEBD = 4   /* tests/csi_oneBD.py:415-521, 543-649 (production "one-BD" model); with n_zero_deg = 10
                           * and tau = 4 transit taps: its posterior-predictive twin, utilities/ppcTools_oneBD.py:185-268 */
} tof_model_kind;

typedef enum tof_ode_mode {
    TOF_ODE_RK4 = 0,   /* classical RK4, ode_substeps per x-interval (the oracle's scheme)      */
    TOF_ODE_RANGE = 1  /* range-energy table of the autonomous Bethe ODE staged in shared memory */
} tof_ode_mode;

/* arithmetic of the Monte-Carlo sample stage (BASELINE.json north_star: 1e-9 in FP64, 1e-4 in an optional FP32 mode) */
typedef enum tof_precision {
    TOF_PRECISION_FP64 = 0, /* everything in FP64: integer stages bit-exact, lnprob <= 1e-9 relative               */
    TOF_PRECISION_FP32 = 1  /* adv/intermediate model with TOF_ODE_RANGE: per-sample arithmetic (energy-loss lookup
                             * offsets, interval search, cross-section weight) in FP32, every reduction and every
                             * stage after the (x,E) histogram in FP64; lnprob <= 1e-4 relative.  Draw sets of >= 8192
                             * per walker are run by the FP64 kernels. */
} tof_precision;

/* spectrum stages returned by tof_model_batch */
typedef enum tof_stage {
    TOF_STAGE_COUNTS = 0, /* np.histogram(tofs, weights=cell counts)          adv:159 density=False */
    TOF_STAGE_PDF = 1,    /* ... density=True                                 adv:159-160          */
    TOF_STAGE_SPREAD = 2  /* beamTiming.applySpreading(pdf) [x scaleFactor]   adv:173, simultFit:300 */
} tof_stage;

/*
 * Everything the reference scripts hold as module-level literals / objects, flattened.
 * Tables are HOST pointers, copied by tof_create (the caller may free them afterwards).
 */
typedef struct tof_config {
    int32_t abi_version;   /* = TOF_ABI_VERSION */
    int32_t model;         /* tof_model_kind */
    int32_t device;        /* CUDA device ordinal */
    int32_t ode_mode;      /* tof_ode_mode */
    int32_t ode_substeps;  /* RK4 steps per x-interval (>=1) */
    int32_t ode_from_zero; /* 0: E(x_0)=E0 like odeint (adv:129); 1: E(0)=E0 like dopri5 (simultFit.py:256) */
    int32_t prior_strict;  /* 1: lo < v < hi (adv:187); 0: reject v<lo or v>hi (simultFit.py:440) */
    int32_t nan_to_neginf; /* 1: NaN lnprob -> -inf (simultFit.py:463-468); 0: NaN is returned */
    int32_t ndim;          /* parameters per walker */
    int32_t n_runs;        /* 1 for simple/adv; 5 in simultFit */
    int32_t x_bins;        /* adv:66 */
    int32_t e_bins;        /* adv:56 */
    int32_t n_taps;        /* timing-response taps, 16 for beamTimingShape (utilities.py:247-262) */
    int32_t n_zero_deg;    /* 0-degree detector sub-times per cell; 0 = off, 10 in simultFit (utilities.py:161) */
    int32_t n_materials;   /* Bethe materials */
    int32_t n_xs;          /* cross-section data points = spline breakpoints (61, utilities.py:338-409) */
    int64_t n_samples;     /* the multiplier in np.rint(dataHist * nSamples) (adv:146) */
    int64_t n_ev_per_loop; /* adv:76 / simultFit.py:178; draws per run = n_loops * n_ev_per_loop */
    int64_t n_loops;       /* adv:126 / simultFit.py:239 */
    double x_min, x_max;   /* adv:67-68 */
    double e_min, e_max;   /* adv:57-58 */
    double speed_of_light; /* constants.py:13 */
    double mass_deuteron;  /* constants.py:22 */
    double mass_neutron;   /* constants.py:23 */
    double mass_he3;       /* constants.py:25 */
    double q_ddn;          /* constants.py:93 */
    double cell_length;    /* constants.py:44 */
    double simple_neutron_base; /* simple model: cellToZero (constants.py:43, simple:66) */
    /* simpleBethe reduced to dE/dx = -(1/E) * sum_k A_k * ln(B_k * E)  (ionStopping.py:78-97) */
    double bethe_A[TOF_MAX_MATERIALS];
    double bethe_B[TOF_MAX_MATERIALS];
    double prior_lo[TOF_MAX_DIM];
    double prior_hi[TOF_MAX_DIM];
    int32_t tof_bins[TOF_MAX_RUNS]; /* constants.py:105 */
    double tof_min[TOF_MAX_RUNS];   /* constants.py:107 */
    double tof_max[TOF_MAX_RUNS];   /* constants.py:106 */
    const double *x_centers;      /* [x_bins]            adv:71-73 */
    const double *e_centers;      /* [e_bins]            adv:61-63 */
    const double *neutron_speed;  /* [e_bins]  c*sqrt(2*En_j/m_n), En_j = getDDneutronEnergy(e_centers) adv:100,110 */
    const double *neutron_dist;   /* [n_runs][x_bins]    adv:153-155 / simultFit.py:290-291 */
    const double *xs_breaks;      /* [n_xs]              utilities.py:338-346 */
    const double *xs_coefs;       /* [n_xs-1][4] power basis, highest order first, about xs_breaks[i] */
    const double *taps;           /* [n_taps]            utilities.py:262 */
    const double *zero_deg_times;   /* [e_bins][n_zero_deg] utilities.py:185 */
    const double *zero_deg_weights; /* [e_bins][n_zero_deg] utilities.py:188-190 */
    /* ---- TOF_ODE_RANGE only: range-energy tables of the autonomous stopping ODE ------------------
     * u(E) = int_{e_min}^{E} dE'/|dE/dx|; along a track u(E(x)) = u(E0) + rng_sign*(x - x_start).
     * T1 gives u(E0) per draw; T2 gives, for v = u of a (draw, x) sample, the E-bin and the
     * cross-section weight XS(E(v)) (ionStopping.py:78-97 + utilities.py:412-429 composed). */
    int32_t t1_q;        /* T1 cells per octave of E = 2^t1_q, indexed by exponent/mantissa bits */
    int32_t t1_key_lo;   /* (high word of E) >> (20 - t1_q) of the first T1 cell */
    int32_t t1_n;        /* T1 cells */
    int32_t rng_degree;  /* T2 polynomial degree (7) */
    int32_t rng_n;       /* T2 intervals */
    int32_t rng_lut_n;   /* uniform lookup cells over [0, rng_u_max] */
    double rng_sign;     /* sign of dE/dx on the table domain */
    double rng_u_max;    /* u(e_max); u(e_min) = 0 */
    double e_tab_lo, e_tab_hi;  /* T1 domain [lo, hi) */
    const double *t1_coefs;     /* [t1_n][8] monomials in t in [-1,1] over the cell, lowest order first */
    const double *rng_breaks;   /* [rng_n+1] T2 breakpoints in u (every E-bin edge and XS knot is one) */
    const int32_t *rng_bins;    /* [rng_n] E-bin of each interval */
    const double *rng_coefs;    /* [rng_n][rng_degree+1] monomials in (v - break), lowest order first */
    const uint16_t *rng_lut;    /* [rng_lut_n] interval holding the left edge of each lookup cell */
    /* ---- TOF_MODEL_ONEBD only ----------------------------------------------------------------------
     * betheApprox (ionStopping.py:102-136): E(E0, x_i) from a spline table instead of an ODE.  Evaluated at
     * the table's own x nodes the bicubic spline reduces to one not-a-knot cubic in E0 per x column. */
    int32_t stop_n;             /* E0 grid points (23: arange(100, 2400, 100), csi_oneBD.py:293) */
    int32_t n_taps2;            /* causal 0-degree transit taps (7, csi_oneBD.py:407-408) */
    double stop_lo, stop_step;  /* E0 grid origin and spacing; arguments are clamped to the grid (FITPACK) */
    double beam_energy;         /* experimentConsts.csi_oneBD.beamReferenceEnergy (constants.py:128) */
    const double *stop_coefs;   /* [x_bins][stop_n-1][4] power basis, highest order first, about the left node */
    const double *attenuation;  /* [x_bins] exp(-x/20 cm) (initialization.py:35-40) */
    const double *taps2;        /* [n_taps2] np.convolve(pdf, taps2, 'full')[:T] (csi_oneBD.py:519) */
    int32_t precision;          /* tof_precision; 0 = FP64 (the reference computes in FP64 throughout) */
} tof_config;

typedef struct tof_ctx tof_ctx;

/* Build a context on cfg->device: copies tables, sizes kernels.  Replaces the module-level
 * setup of the model scripts (adv:44-100; simultFit.py:121-205). */
int tof_create(const tof_config *cfg, tof_ctx **out);
void tof_destroy(tof_ctx *ctx);

/* Last error text for ctx (or for the failed tof_create when ctx == NULL). */
const char *tof_last_error(const tof_ctx *ctx);

/* Observed TOF histogram of one run: the `observables` kwarg of lnprob (adv:300-302;
 * simultFit.py:713-718).  Copied; the 0 -> 1 substitution of simultFit.py:391-392 is applied to
 * the private copy only. */
int tof_set_observables(tof_ctx *ctx, int run, const double *counts, int nbins);

/* Explicit Monte-Carlo draws replacing the reference's global np.random stream.
 *   stream 0: standard normals z[n] (adv:128; simple:64; simultFit.py:244), n = n_loops*n_ev_per_loop
 *   stream 1: simple model: uniforms u[n] in [0,1) (simple:62);
 *             simult model: replacement normals for the E0<=0 rejection loop (simultFit.py:245-252);
 *             oneBD model: the uniforms numpy's legacy Poisson sampler consumes for
 *             np.random.poisson(bgLevel, T) (csi_oneBD.py:521), taken sequentially
 * Host pointers; copied to the device. */
int tof_set_draws(tof_ctx *ctx, int run, int stream, const double *values, int64_t n);

/* Where the Monte-Carlo draws of an evaluation come from.
 *   TOF_DRAWS_BOUND (default): the draw set bound with tof_set_draws, shared by all walkers and all calls (common
 *     random numbers): bit-reproducible, the mode every parity test runs in.
 *   TOF_DRAWS_PER_EVALUATION: what the reference does -- every lnlike call draws its own numbers from the global stream
 *     (adv:128 np.random.normal inside generateModelData; simple:62-64), so each walker at each step sees its own noise.
 *     The draws are generated on the device: Philox4x32-10, key = seed, counter = (index, epoch, global walker index);
 *     `epoch` advances by one per model call made through the batch entry points (first call: epoch0, < 2^31) and is
 *     2^31 + 2*step + half inside tof_ensemble_step / tof_ensemble_half_step, so chains do not depend on the sharding.  Normals
 *     are inverse-CDF transforms of open uniforms; the range kernels get each walker's normals already sorted (order
 *     statistics from exponential spacings: no sort).  tof_set_draws is not needed in this mode.
 * Built for every model; adv/intermediate with TOF_ODE_RANGE needs FP64 and n_draws <= 1024 (one sorted tile per
 * walker), the simultaneous fit needs TOF_ODE_RK4. */
typedef enum tof_draw_mode { TOF_DRAWS_BOUND = 0, TOF_DRAWS_PER_EVALUATION = 1 } tof_draw_mode;
int tof_set_draw_mode(tof_ctx *ctx, int mode, uint64_t seed, uint64_t epoch0);

/* The draws walker `walker` of model call `epoch` uses for run `run` in TOF_DRAWS_PER_EVALUATION mode (HOST buffer
 * out[n]): stream 0 the model's normals (sorted != 0: ascending, as the range kernels consume them; n <= 1024), stream 1
 * uniforms (the simple model's x positions, simple:62; the oneBD model's Poisson background, csi_oneBD.py:521), stream 3
 * the simultaneous fit's replacement normals (simultFit.py:245-252).  For parity checks: feed them to a CPU evaluation. */
int tof_generate_draws(tof_ctx *ctx, uint64_t epoch, int64_t walker, int run, int stream, int sorted, double *out, int64_t n);

/* lnprob for n walkers: theta[n][ndim] row-major -> out[n].  HOST buffers; the call copies in,
 * launches, copies out and synchronises.  Replaces n calls of lnprob (adv:191-199, simple:112-120,
 * simultFit.py:444-469), i.e. one emcee `_get_lnprob` map over a half-ensemble. */
int tof_lnprob_batch(tof_ctx *ctx, const double *theta, int64_t n, double *out);

/* Same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t, NULL = default stream). */
int tof_lnprob_batch_device(tof_ctx *ctx, const double *d_theta, int64_t n, double *d_out, void *stream);

/* Model spectra for n parameter vectors (HOST buffers): spectra[n][tof_bins[run]] at `stage`.
 * Replaces generateModelData (adv:115-161; simultFit.py:223-300) for parity checks and
 * posterior-predictive generation. */
int tof_model_batch(tof_ctx *ctx, const double *theta, int64_t n, int run, int stage, double *spectra);

/* Integer (x, E) cell counts drawHist2d (adv:146; simultFit.py:283): counts[n][x_bins][e_bins]. */
int tof_cell_counts_batch(tof_ctx *ctx, const double *theta, int64_t n, int run, int64_t *counts);

/* Unweighted (x, E) histogram of the stopped deuteron energies of the LAST loop, [n][x_bins][e_bins]: the
 * `eD_atEachX` rows that utilities/ppcTools.py:140-157 collects for posterior-predictive checks (its leading row
 * of zeros omitted).  Built for TOF_MODEL_SIMULT (the model ppcTools re-runs; RK4 energies in either ode_mode) and for
 * TOF_MODEL_ONEBD (utilities/ppcTools_oneBD.py:214, 223-224). */
int tof_deuteron_counts_batch(tof_ctx *ctx, const double *theta, int64_t n, int run, int64_t *counts);

/* ---- ensemble driver: emcee 2.x EnsembleSampler stretch move (a = 2), red/blue halves -------- */

/* Propose for a half-ensemble (DEVICE buffers, async on stream):
 *   q[i] = c[j] - zz*(c[j] - s[i]),  zz = ((a-1)*U + 1)^2 / a,  j = randint(n_comp)
 * Counter-based Philox keyed by (seed, step, half, global walker index) so that the chain does not
 * depend on how walkers are sharded over GPUs.  s: [n][ndim] current positions of the slice whose
 * first walker has global index walker0; comp: [n_comp][ndim] the complementary half.
 * Outputs q[n][ndim], log_zz[n] = (ndim-1)*ln zz. */
int tof_stretch_propose(tof_ctx *ctx, const double *d_s, int64_t n, int64_t walker0, const double *d_comp,
                        int64_t n_comp, double a, uint64_t seed, int64_t step, int half, double *d_q,
                        double *d_log_zz, void *stream);

/* Accept/reject in place: accept when log_zz + new_lnprob - lnprob > ln U.  Updates s, lnprob and
 * the per-walker acceptance counters n_accept[n] (may be NULL). */
int tof_stretch_accept(tof_ctx *ctx, double *d_s, double *d_lnprob, int64_t n, int64_t walker0,
                       const double *d_q, const double *d_new_lnprob, const double *d_log_zz,
                       uint64_t seed, int64_t step, int half, int64_t *d_n_accept, void *stream);

/* n_steps whole ensemble steps (both red/blue half-steps each: propose -> lnprob -> accept) for an ensemble that lives
 * on this GPU, enqueued on `stream` without returning to the host in between -- what emcee's
 * EnsembleSampler.sample(p0, iterations=n_steps) loop does (adv:311-317; simultFit.py:733-740) for the kwargs bound
 * with tof_set_observables / tof_set_draws.  d_pos [n_walkers][ndim] and d_lnprob [n_walkers] are DEVICE buffers
 * updated in place (d_lnprob must hold the log-probabilities of d_pos on entry, e.g. from tof_lnprob_batch_device);
 * first half = walkers [0, n/2), like emcee.  Randomness as in tof_stretch_propose with step = step0 + s, so the
 * chain equals the one produced half-step by half-step (and by any sharding).  d_n_accept [n_walkers] may be NULL.
 * emcee's own constructor checks apply: n_walkers even and >= 2*ndim. */
int tof_ensemble_step(tof_ctx *ctx, double *d_pos, double *d_lnprob, int64_t n_walkers, int64_t n_steps, double a,
                      uint64_t seed, int64_t step0, int64_t *d_n_accept, void *stream);

/* One red/blue half-step for the slice of the ensemble THIS GPU owns, on the packed state the sharded driver keeps:
 * d_state [n_walkers][ndim + 1] = positions followed by the log-probability of each walker (so that one all-gather of
 * the rows a rank owns refreshes both on every replica).  Rows [half*h + own0, half*h + own0 + n_own)
 * (h = n_walkers/2) are proposed against the whole complementary half, evaluated and accepted in place; every other
 * row is only read.  Randomness as in tof_stretch_propose (keyed by the global walker index), so the chain is the one
 * tof_ensemble_step produces on a single GPU.  The caller then all-gathers rows [half*h, (half+1)*h) over its ranks
 * (emcee's EnsembleSampler.sample loop, adv:311-317, with walkers sharded instead of pool.map'ed). */
int tof_ensemble_half_step(tof_ctx *ctx, double *d_state, int64_t n_walkers, int half, int64_t own0, int64_t n_own, double a,
                           uint64_t seed, int64_t step, int64_t *d_n_accept, void *stream);

/* ---- diagnostics ----------------------------------------------------------------------------- */
typedef struct tof_stats {
    int64_t kernel_launches; /* kernels launched by this context since creation */
    int64_t evaluations;     /* walker lnprob evaluations issued */
    int64_t nan_results;     /* walkers inside the prior whose log-probability came out NaN since creation (returned as
                              * NaN, or as -inf when nan_to_neginf is set: the case simultFit.py:463-468 prints a dump for) */
    int32_t sm_count;
    int32_t smem_bytes;      /* dynamic shared memory of the main model kernel */
    int32_t threads;         /* threads per CTA of the main model kernel */
    int32_t ctas_per_sm;     /* resident CTAs per SM of the main model kernel */
    int32_t band_ctas_per_sm; /* range kernel, banded launch (512 threads): resident CTAs per SM; 0 = disabled */
    int32_t band_cells;       /* ... cell-histogram capacity of the banded launch */
    int64_t band_queued_last; /* ... walkers of the most recent call that were queued for the second, full-size launch
                               * (always 0 with the single-launch kernel, see model_launches_per_call) */
    int32_t fp32_active;      /* 1 when the FP32 sample stage is what the model kernel runs (see tof_precision) */
    int32_t model_launches_per_call; /* model kernels one tof_lnprob_batch call of the adv/intermediate range path launches:
                               * 1 with the single-launch kernel (adv_zrank_kernel: every walker, wide E-band or not, is
                               * handled by the CTA that fetched it), 2 with the banded + overflow pair */
    int64_t wide_last;        /* single-launch kernel: walkers of the most recent call whose (x,E) histogram did not fit
                               * shared memory and was kept in the CTA's L2-resident scratch slice instead */
} tof_stats;
int tof_get_stats(const tof_ctx *ctx, tof_stats *out);

/* Device time (ms, CUDA events on the launching stream) of the model kernel(s) of the most recent
 * tof_lnprob_batch / tof_lnprob_batch_device call made with timing enabled. */
int tof_set_timing(tof_ctx *ctx, int enabled);
int tof_last_kernel_ms(tof_ctx *ctx, float *ms);

/* Per-stage timing of the shipped adv/intermediate kernel (TOF_ODE_RANGE): while enabled, every CTA charges its SM
 * clock cycles to the stage that just ended (thread 0, between barriers).  tof_get_stage_cycles returns the sums since
 * the last call and resets them: cycles[k] for the TOF_N_STAGES stages below, then the number of walkers processed.
 *   0 setup      work fetch, prior, per-row E-band, record staging, histogram reset
 *   1 histogram  energy-loss lookup per draw, draw-index lookup, (row, interval) cell sums      adv:128-138
 *   2 normalise  sum(H * dE * dx)                                                                adv:143
 *   3 scatter    np.rint counts, flight time of every non-empty cell, TOF histogram             adv:146-159
 *   4 likelihood density, timing response at the observed bins, sum(obs * log)                  adv:160-181
 * Off by default (one predicated branch per stage); this is the profiling hook the reference lacks
 * (testStoppingApproximation.py:5-6 "really we should actually profile the sampling"). */
#define TOF_N_STAGES 5
int tof_set_stage_timing(tof_ctx *ctx, int enabled);
int tof_get_stage_cycles(tof_ctx *ctx, uint64_t cycles[TOF_N_STAGES + 1]);

/* Peak FP64 FMA rate of the device measured with a register-resident DFMA loop (TFLOP/s):
 * the roofline denominator for this FP64-pipe-bound path. */
int tof_measure_fp64_peak(tof_ctx *ctx, double *tflops);

int tof_abi_version(void);
/* sizeof(tof_config) as compiled, so that FFI struct mirrors can be checked at load time. */
int tof_sizeof_config(void);

#ifdef __cplusplus
}
#endif
#endif /* TOFGPU_H */
