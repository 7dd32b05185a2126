// C ABI of libtofgpu.so (see include/tofgpu.h).  Host-side context management + kernel launches.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tof_kernels.cuh"

using namespace tof;

namespace {

thread_local std::string g_create_error;

struct DeviceBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct tof_ctx {
    tof_config cfg{};
    DevModel dm{};
    DevRun runs[TOF_MAX_RUNS]{};
    std::vector<void *> owned;  // device allocations freed in tof_destroy
    DeviceBuf d_theta, d_out, d_spectra, d_cells, d_counts, d_partial, d_work, d_queue, d_split, d_tickets, d_ens, d_nan, d_stage;
    DeviceBuf d_wide, d_zlut[TOF_MAX_RUNS];   // adv_zrank_kernel: scratch histograms of wide walkers; draw-rank lookup
    // rebindable inputs: reused across tof_set_draws / tof_set_observables calls (no growth when draws are refreshed)
    DeviceBuf d_z[TOF_MAX_RUNS][2], d_obs[TOF_MAX_RUNS], d_obs_idx[TOF_MAX_RUNS], d_obs_val[TOF_MAX_RUNS];
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // recorded after every model launch on the launching stream: rebinding inputs and reading counters wait for
    // THIS event only (no device-wide synchronisation: other contexts / ranks of the process keep running)
    cudaEvent_t ev_busy = nullptr;
    bool busy = false;
    unsigned long long *h_counters = nullptr;   // pinned staging for tof_get_stats: {nan results, queued walkers}
    bool timing = false, timed = false, stage_timing = false;
    std::string err;
    tof_stats stats{};
    int adv_nt = 1024, adv_dpt = 1;
    size_t adv_smem = 0;
    int rng_nt = 1024;
    bool f32 = false;        // FP32 sample stage active (tof_config.precision, tiled draw sets only)
    // banded launch of the range kernel: 512 threads, 2 CTAs/SM
    int band_hcap = 0, band_rcap = 0, band_ctas = 0, band_nt = 512;
    size_t band_smem = 0;
    bool band_enabled = false;
    bool planned = false;    // banded launch runs adv_planned_kernel (FP64, <= one tile of draws, interval == E-bin)
    // ... or adv_zrank_kernel (same conditions; shipped): one launch per call, wide walkers in a global scratch histogram
    bool zrank = false;
    bool zrank_multi = false;   // ... adv_zrank_multi_kernel for draw sets of more than one tile (no draw split)
    int last_model_launches = 0;   // model kernels the most recent adv/intermediate range call launched
    int zr_hcap = 0, zr_rcap = 0, zr_nt = 512;
    size_t zr_smem = 0;
    RangeLayout lay_zr{};
    RangeLayout lay_full{}, lay_band{};   // shared-memory layouts of the two launches (host-computed offsets)
    size_t simult_smem = 0, onebd_smem = 0;
    size_t simult_rk4_smem = 0;   // range-mode contexts: the RK4 kernel still serves the unweighted deuteron histograms
    int max_smem_optin = 0;
    // per-evaluation draws (tof_set_draw_mode): epoch of the next model call, key of the call being launched
    bool fresh = false;
    uint64_t fresh_seed = 0, fresh_epoch = 0;
    uint64_t cur_epoch = 0;
    int64_t cur_walker0 = 0;
    bool have_obs[TOF_MAX_RUNS]{};
    bool have_z[TOF_MAX_RUNS][2]{};
};

namespace {

int fail(tof_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, TOF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));      \
    } while (0)

template <typename T>
int upload(tof_ctx *ctx, const T *host, size_t count, const T **dev) {
    void *p = nullptr;
    CU(ctx, cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    ctx->owned.push_back(p);
    if (count) CU(ctx, cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T *>(p);
    return TOF_OK;
}

int ensure(tof_ctx *ctx, DeviceBuf &b, size_t bytes);

// Copy host data into a reusable device buffer (grown when needed); *dev receives the device pointer.
template <typename T>
int upload_into(tof_ctx *ctx, DeviceBuf &b, const T *host, size_t count, const T **dev) {
    if (int rc = ensure(ctx, b, std::max<size_t>(count, 1) * sizeof(T))) return rc;
    // make sure no kernel of this context still reads the previous contents: wait for the context's last model
    // launch (whatever stream it went to), not for the device
    if (ctx->busy) CU(ctx, cudaEventSynchronize(ctx->ev_busy));
    if (count) {
        CU(ctx, cudaMemcpyAsync(b.p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));       // `host` may be a temporary of the caller
    }
    *dev = static_cast<const T *>(b.p);
    return TOF_OK;
}

int ensure(tof_ctx *ctx, DeviceBuf &b, size_t bytes) {
    if (b.bytes >= bytes) return TOF_OK;
    if (b.p) CU(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
    CU(ctx, cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return TOF_OK;
}

// ---- adv kernel variants ---------------------------------------------------------------------
using AdvKernel = void (*)(const DevModel, const DevRun, const double *, long long, ModelOut);

template <int NT, int DPT>
AdvKernel adv_pick(int nmat) {
    return nmat == 1 ? adv_lnprob_kernel<NT, DPT, 1> : adv_lnprob_kernel<NT, DPT, 0>;
}

AdvKernel adv_variant(int nt, int dpt, int nmat) {
    if (nt == 128 && dpt == 8) return adv_pick<128, 8>(nmat);
    if (nt == 256 && dpt == 4) return adv_pick<256, 4>(nmat);
    if (nt == 256 && dpt == 2) return adv_pick<256, 2>(nmat);
    if (nt == 512 && dpt == 2) return adv_pick<512, 2>(nmat);
    if (nt == 512 && dpt == 1) return adv_pick<512, 1>(nmat);
    if (nt == 1024 && dpt == 1) return adv_pick<1024, 1>(nmat);
    return nullptr;
}

AdvKernel range_variant_f32(int nt, int degree) {
    if (nt == 1024 && degree == 7) return adv_range_kernel<1024, 7, true>;
    if (nt == 512 && degree == 7) return adv_range_kernel<512, 7, true>;
    return nullptr;
}

AdvKernel range_variant_prof(int nt, int degree) {          // stage-timing instantiations (tof_set_stage_timing)
    if (nt == 1024 && degree == 7) return adv_range_kernel<1024, 7, false, true>;
    if (nt == 512 && degree == 7) return adv_range_kernel<512, 7, false, true>;
    return nullptr;
}

AdvKernel range_variant(int nt, int degree, bool f32 = false, bool prof = false) {
    if (prof) return range_variant_prof(nt, degree);
    if (f32) return range_variant_f32(nt, degree);
    if (nt == 1024 && degree == 7) return adv_range_kernel<1024, 7>;
    if (nt == 800 && degree == 7) return adv_range_kernel<800, 7>;
    if (nt == 640 && degree == 7) return adv_range_kernel<640, 7>;
    if (nt == 512 && degree == 7) return adv_range_kernel<512, 7>;
    if (nt == 448 && degree == 7) return adv_range_kernel<448, 7>;
    if (nt == 384 && degree == 7) return adv_range_kernel<384, 7>;
    return nullptr;
}

int check_run(tof_ctx *ctx, int run) {
    if (run < 0 || run >= ctx->cfg.n_runs) return fail(ctx, TOF_ERR_INVALID, "run index out of range");
    return TOF_OK;
}

int ready(tof_ctx *ctx, bool need_obs) {
    for (int r = 0; r < ctx->cfg.n_runs; ++r) {
        if (need_obs && !ctx->have_obs[r]) return fail(ctx, TOF_ERR_STATE, "observables not set for run " + std::to_string(r));
        if (ctx->fresh) continue;                           // draws are generated on the device
        if (!ctx->have_z[r][0]) return fail(ctx, TOF_ERR_STATE, "draws (stream 0) not set for run " + std::to_string(r));
        if (ctx->cfg.model == TOF_MODEL_SIMPLE && !ctx->have_z[r][1])
            return fail(ctx, TOF_ERR_STATE, "uniform draws (stream 1) not set");
    }
    return TOF_OK;
}

// Launch the model for n walkers with device pointers.  `out` selects what is produced.
int launch_model(tof_ctx *ctx, const double *d_theta, long long n, int run, ModelOut out, cudaStream_t st) {
    if (n <= 0) return TOF_OK;
    const tof_config &c = ctx->cfg;
    if (out.lnprob && !out.spectra && !out.cells) {          // production calls count NaN results (tof_get_stats)
        if (!ctx->d_nan.p) {
            if (int rc = ensure(ctx, ctx->d_nan, sizeof(unsigned long long))) return rc;
            CU(ctx, cudaMemsetAsync(ctx->d_nan.p, 0, sizeof(unsigned long long), st));
        }
        out.nan_count = static_cast<unsigned long long *>(ctx->d_nan.p);
    }
    if (ctx->stage_timing && ctx->d_stage.p) out.stage_cycles = static_cast<unsigned long long *>(ctx->d_stage.p);
    if (ctx->timing) CU(ctx, cudaEventRecord(ctx->ev0, st));
    // the run as the kernels see it for THIS call: with per-evaluation draws it carries the call's key
    DevRun run0 = ctx->runs[run];
    run0.fresh = ctx->fresh ? 1 : 0;
    run0.fresh_seed = ctx->fresh_seed;
    run0.fresh_epoch = ctx->cur_epoch;
    run0.fresh_walker0 = ctx->cur_walker0;
    if (c.model == TOF_MODEL_ADV) {
        if (c.ode_mode == TOF_ODE_RANGE) {
            // persistent CTAs: one per resident slot, walkers handed out through global counters.
            // d_work = {banded work counter, full-size work counter, queue length}
            const bool prof = out.stage_cycles != nullptr;
            AdvKernel kfull = range_variant(ctx->rng_nt, c.rng_degree, ctx->f32, prof);
            int rc = ensure(ctx, ctx->d_work, 3 * sizeof(unsigned long long));
            if (rc) return rc;
            CU(ctx, cudaMemsetAsync(ctx->d_work.p, 0, 3 * sizeof(unsigned long long), st));
            unsigned long long *cnt = static_cast<unsigned long long *>(ctx->d_work.p);
            const long long slots_full = (long long)ctx->stats.sm_count * std::max(ctx->stats.ctas_per_sm, 1);
            const bool debug = out.spectra != nullptr || out.cells != nullptr;
            out.hcap = c.x_bins * c.e_bins;
            out.rcap = c.rng_n;
            out.lay = ctx->lay_full;
            // few walkers and a big (streamed) draw set: several CTAs per walker
            out.n_split = 1;
            if (!debug && ctx->dm.n_draws >= RANGE_STREAM_MIN) {
                const long long slots = (long long)ctx->stats.sm_count * (ctx->band_enabled ? std::max(ctx->band_ctas, 1) : 1);
                const long long chunks = (ctx->dm.n_draws + 128LL * 16 - 1) / (128LL * 16);   // >= one 128-draw chunk per warp
                // pick the split that minimises the number of rounds of the persistent grid per unit of work,
                // rounds(S)/S with rounds = ceil(n*S/slots), plus a small charge for merging S partial histograms
                // (256 walkers on 296 slots: S = 1 leaves 40 slots idle for a whole walker; S = 15 fills 13 rounds)
                long long S = 1;
                if (n < 2 * slots) {
                    double best = 1e300;
                    for (long long s = 1; s <= std::min<long long>(16, chunks); ++s) {
                        const double rounds = (double)((n * s + slots - 1) / slots);
                        const double cost = rounds / (double)s + (s > 1 ? 0.02 + 0.004 * (double)s : 0.0);
                        if (cost < best - 1e-12) { best = cost; S = s; }
                    }
                }
                if (S > 1) {
                    const size_t stride = (size_t)c.x_bins * c.e_bins;
                    rc = ensure(ctx, ctx->d_split, (size_t)n * S * stride * sizeof(double));
                    if (rc) return rc;
                    if (ctx->d_tickets.bytes < (size_t)n * sizeof(unsigned int)) {
                        rc = ensure(ctx, ctx->d_tickets, (size_t)n * sizeof(unsigned int));
                        if (rc) return rc;
                        CU(ctx, cudaMemsetAsync(ctx->d_tickets.p, 0, ctx->d_tickets.bytes, st));
                    }
                    out.n_split = (int)S;
                    out.split_stride = (int)stride;
                    out.split_scratch = static_cast<double *>(ctx->d_split.p);
                    out.split_tickets = static_cast<unsigned int *>(ctx->d_tickets.p);
                }
            }
            const long long n_work = n * out.n_split;
            if (ctx->zrank_multi && !ctx->fresh && !debug && !prof) {
                // big draw sets: tiles of sorted draws, long runs (one launch); with few walkers n_split CTAs share a walker
                const long long slots = (long long)ctx->stats.sm_count * 2;
                const unsigned grid = (unsigned)std::min<long long>(n_work, slots);
                const size_t stride = (size_t)c.x_bins * c.e_bins;
                rc = ensure(ctx, ctx->d_wide, (size_t)slots * stride * sizeof(double));
                if (rc) return rc;
                ModelOut oz = out;
                oz.work = cnt + 0;
                oz.queue_count = cnt + 2;
                oz.hcap = ctx->zr_hcap;
                oz.rcap = ctx->zr_rcap;
                oz.lay = ctx->lay_zr;
                oz.split_stride = (int)stride;
                oz.wide_scratch = static_cast<double *>(ctx->d_wide.p);
                adv_zrank_multi_kernel<512, 7><<<grid, 512, ctx->zr_smem, st>>>(ctx->dm, run0, d_theta, n, oz);
                ctx->last_model_launches = 1;
            } else if (ctx->zrank && !ctx->fresh && !debug && out.n_split == 1 && ctx->runs[run].zlut) {
                // shipped path: ONE persistent launch, 2 CTAs/SM; walkers whose E-band does not fit shared memory keep
                // their cell sums in this CTA's slice of an L2-resident scratch buffer (cnt[2] counts them)
                const long long slots = (long long)ctx->stats.sm_count * 2;
                const unsigned grid = (unsigned)std::min<long long>(n, slots);
                const size_t stride = (size_t)c.x_bins * c.e_bins;
                rc = ensure(ctx, ctx->d_wide, (size_t)slots * stride * sizeof(double));
                if (rc) return rc;
                ModelOut oz = out;
                oz.work = cnt + 0;
                oz.queue_count = cnt + 2;
                oz.hcap = ctx->zr_hcap;
                oz.rcap = ctx->zr_rcap;
                oz.lay = ctx->lay_zr;
                oz.split_stride = (int)stride;
                oz.wide_scratch = static_cast<double *>(ctx->d_wide.p);
                if (prof) adv_zrank_kernel<512, 7, true><<<grid, 512, ctx->zr_smem, st>>>(ctx->dm, ctx->runs[run], d_theta, n, oz);
                else if (ctx->dm.rank_stride == ZR_TPITCH) adv_zrank_kernel<512, 7, false, ZR_TPITCH><<<grid, 512, ctx->zr_smem, st>>>(ctx->dm, ctx->runs[run], d_theta, n, oz);
                else adv_zrank_kernel<512, 7, false><<<grid, 512, ctx->zr_smem, st>>>(ctx->dm, ctx->runs[run], d_theta, n, oz);
                ctx->last_model_launches = 1;
            } else if (ctx->band_enabled && !debug && (!ctx->fresh || (ctx->planned && out.n_split == 1))) {
                rc = ensure(ctx, ctx->d_queue, (size_t)n * sizeof(int));
                if (rc) return rc;
                // 1) banded launch: every walker whose E-band fits; the others are queued
                ModelOut ob = out;
                ob.work = cnt + 0;
                ob.hcap = ctx->band_hcap;
                ob.rcap = ctx->band_rcap;
                ob.lay = ctx->lay_band;
                ob.queue_out = static_cast<int *>(ctx->d_queue.p);
                ob.queue_count = cnt + 2;
                AdvKernel kband = range_variant(ctx->band_nt, c.rng_degree, ctx->f32, prof);
                // many walkers x one tile of draws, one interval per E-bin: the lean cut of the same kernel
                if (ctx->planned && out.n_split == 1) kband = prof ? adv_planned_kernel<512, 7, true> : adv_planned_kernel<512, 7, false>;
                const long long slots_band = (long long)ctx->stats.sm_count * std::max(ctx->band_ctas, 1);
                kband<<<(unsigned)std::min<long long>(n_work, slots_band), ctx->band_nt, ctx->band_smem, st>>>(ctx->dm, run0, d_theta, n, ob);
                // 2) full-size launch over the queue (exits at once when it is empty)
                out.work = cnt + 1;
                out.queue_in = static_cast<const int *>(ctx->d_queue.p);
                out.queue_count = cnt + 2;
                kfull<<<(unsigned)std::min<long long>(n_work, slots_full), ctx->rng_nt, ctx->adv_smem, st>>>(ctx->dm, run0, d_theta, n, out);
                ctx->stats.kernel_launches += 1;
                ctx->last_model_launches = 2;
            } else {
                out.work = cnt + 1;
                kfull<<<(unsigned)std::min<long long>(n_work, slots_full), ctx->rng_nt, ctx->adv_smem, st>>>(ctx->dm, run0, d_theta, n, out);
                ctx->last_model_launches = 1;
            }
        } else {
            AdvKernel k = adv_variant(ctx->adv_nt, ctx->adv_dpt, c.n_materials);
            k<<<(unsigned)n, ctx->adv_nt, ctx->adv_smem, st>>>(ctx->dm, run0, d_theta, n, out);
        }
        ctx->stats.kernel_launches += 1;
    } else if (c.model == TOF_MODEL_SIMPLE) {
        const int T = c.tof_bins[0];
        int rc = ensure(ctx, ctx->d_counts, (size_t)n * T * sizeof(unsigned long long));
        if (rc) return rc;
        CU(ctx, cudaMemsetAsync(ctx->d_counts.p, 0, (size_t)n * T * sizeof(unsigned long long), st));
        // enough chunks to fill the machine ~4x over, but at least ~4096 draws per CTA
        long long chunks = std::max<long long>(1, std::min<long long>((ctx->stats.sm_count * 8 + n - 1) / n,
                                                                       (ctx->dm.n_draws + 4095) / 4096));
        chunks = std::min<long long>(chunks, 65535);
        dim3 grid((unsigned)n, (unsigned)chunks);          // walkers on grid.x: no 65535 limit on the batch
        simple_hist_kernel<256><<<grid, 256, 0, st>>>(ctx->dm, run0, d_theta, n,
                                                      static_cast<unsigned long long *>(ctx->d_counts.p),
                                                      out.spectra != nullptr);
        simple_finish_kernel<32><<<(unsigned)n, 32, 0, st>>>(ctx->dm, ctx->runs[0], d_theta, n,
                                                             static_cast<unsigned long long *>(ctx->d_counts.p), out);
        ctx->stats.kernel_launches += 2;
    } else if (c.model == TOF_MODEL_ONEBD) {
        DevRunSet rs;
        for (int r = 0; r < TOF_MAX_RUNS; ++r) {
            rs.r[r] = ctx->runs[r];
            rs.r[r].fresh = run0.fresh; rs.r[r].fresh_seed = run0.fresh_seed; rs.r[r].fresh_epoch = run0.fresh_epoch;
            rs.r[r].fresh_walker0 = run0.fresh_walker0;
        }
        const bool debug = out.spectra != nullptr || out.cells != nullptr;
        if (debug) {
            onebd_run_kernel<256><<<(unsigned)n, 256, ctx->onebd_smem, st>>>(ctx->dm, rs, d_theta, n, out, run);
            ctx->stats.kernel_launches += 1;
        } else {
            int rc = ensure(ctx, ctx->d_partial, (size_t)n * c.n_runs * sizeof(double));
            if (rc) return rc;
            ModelOut po{};
            po.lnprob = static_cast<double *>(ctx->d_partial.p);
            onebd_run_kernel<256><<<(unsigned)(n * c.n_runs), 256, ctx->onebd_smem, st>>>(ctx->dm, rs, d_theta, n, po, -1);
            simult_finish_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ctx->dm, d_theta, n, po.lnprob, out.lnprob, out.nan_count);
            ctx->stats.kernel_launches += 2;
        }
    } else if (c.model == TOF_MODEL_SIMULT) {
        DevRunSet rs;
        for (int r = 0; r < TOF_MAX_RUNS; ++r) {
            rs.r[r] = ctx->runs[r];
            rs.r[r].fresh = run0.fresh; rs.r[r].fresh_seed = run0.fresh_seed; rs.r[r].fresh_epoch = run0.fresh_epoch;
            rs.r[r].fresh_walker0 = run0.fresh_walker0;
        }
        const bool debug = out.spectra != nullptr || out.cells != nullptr;
        const bool rng = c.ode_mode == TOF_ODE_RANGE;
        if (debug) {
            // eD_atEachX (ppcTools.py:151-157) is an unweighted histogram of stopped energies: the RK4 kernel produces it in
            // either mode (the range kernel never forms per-sample energies)
            if (rng && !out.unweighted) simult_range_kernel<256, 7><<<(unsigned)n, 256, ctx->simult_smem, st>>>(ctx->dm, rs, d_theta, n, out, run);
            else simult_run_kernel<256><<<(unsigned)n, 256, rng ? ctx->simult_rk4_smem : ctx->simult_smem, st>>>(ctx->dm, rs, d_theta, n, out, run);
            ctx->stats.kernel_launches += 1;
        } else {
            int rc = ensure(ctx, ctx->d_partial, (size_t)n * c.n_runs * sizeof(double));
            if (rc) return rc;
            ModelOut po{};
            po.lnprob = static_cast<double *>(ctx->d_partial.p);
            if (rng) simult_range_kernel<256, 7><<<(unsigned)(n * c.n_runs), 256, ctx->simult_smem, st>>>(ctx->dm, rs, d_theta, n, po, -1);
            else simult_run_kernel<256><<<(unsigned)(n * c.n_runs), 256, ctx->simult_smem, st>>>(ctx->dm, rs, d_theta, n, po, -1);
            simult_finish_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ctx->dm, d_theta, n, po.lnprob, out.lnprob, out.nan_count);
            ctx->stats.kernel_launches += 2;
        }
    } else {
        return fail(ctx, TOF_ERR_INVALID, "model kind not implemented");
    }
    if (ctx->timing) {
        CU(ctx, cudaEventRecord(ctx->ev1, st));
        ctx->timed = true;
    }
    CU(ctx, cudaGetLastError());
    {   // (not while `st` is being captured into a CUDA graph: a captured event cannot be waited on from the host)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap == cudaStreamCaptureStatusNone) {
            CU(ctx, cudaEventRecord(ctx->ev_busy, st));
            ctx->busy = true;
        }
    }
    ctx->stats.evaluations += n;
    return TOF_OK;
}

// Key of the next model call when draws are generated per evaluation: calls made through the batch entry points take
// consecutive epochs; the ensemble entry points pin (epoch, first walker) to (2*step + half, global walker index).
void next_call_key(tof_ctx *ctx) {
    ctx->cur_epoch = ctx->fresh_epoch++;
    ctx->cur_walker0 = 0;
}

template <int NT>
__global__ void fresh_draws_kernel(DevRun run, long long walker, int run_idx, int stream, int sorted, int n, double *out) {
    __shared__ double zs[2 * NT];
    __shared__ double scratch[NT / 32 + 2];
    if (sorted) {
        fresh_sorted_normals<NT>(zs, n, run, walker, run_idx, scratch);
        for (int d = threadIdx.x; d < n; d += NT) out[d] = zs[d];
    } else {
        for (int d = threadIdx.x; d < n; d += NT)
            out[d] = stream == 1 ? fresh_uniform(run, walker, run_idx, d) : fresh_normal(run, walker, run_idx, d, stream);
    }
}

}  // namespace

extern "C" {

int tof_abi_version(void) { return TOF_ABI_VERSION; }
int tof_sizeof_config(void) { return (int)sizeof(tof_config); }

const char *tof_last_error(const tof_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tof_create(const tof_config *cfg, tof_ctx **out) {
    if (!cfg || !out) return fail(nullptr, TOF_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != TOF_ABI_VERSION) return fail(nullptr, TOF_ERR_INVALID, "abi_version mismatch");
    if (cfg->model != TOF_MODEL_SIMPLE && cfg->model != TOF_MODEL_ADV && cfg->model != TOF_MODEL_SIMULT &&
        cfg->model != TOF_MODEL_ONEBD)
        return fail(nullptr, TOF_ERR_INVALID, "unknown model kind");
    if (cfg->ndim < 1 || cfg->ndim > TOF_MAX_DIM) return fail(nullptr, TOF_ERR_INVALID, "ndim out of range");
    if (cfg->n_runs < 1 || cfg->n_runs > TOF_MAX_RUNS) return fail(nullptr, TOF_ERR_INVALID, "n_runs out of range");
    if (cfg->n_loops < 1 || cfg->n_ev_per_loop < 1) return fail(nullptr, TOF_ERR_INVALID, "n_loops / n_ev_per_loop must be >= 1");
    for (int r = 0; r < cfg->n_runs; ++r)
        if (cfg->tof_bins[r] < 1 || !(cfg->tof_max[r] > cfg->tof_min[r]))
            return fail(nullptr, TOF_ERR_INVALID, "bad TOF window");
    const bool cell_model = cfg->model != TOF_MODEL_SIMPLE;
    if (cell_model) {
        if (cfg->x_bins < 1 || cfg->e_bins < 1 || !(cfg->x_max > cfg->x_min) || !(cfg->e_max > cfg->e_min))
            return fail(nullptr, TOF_ERR_INVALID, "bad (x, E) binning");
        if (cfg->model != TOF_MODEL_ONEBD && (cfg->n_materials < 1 || cfg->n_materials > TOF_MAX_MATERIALS))
            return fail(nullptr, TOF_ERR_INVALID, "n_materials out of range");
        if (cfg->n_xs < 4) return fail(nullptr, TOF_ERR_INVALID, "cross-section table too short");
        if (cfg->n_taps < 1) return fail(nullptr, TOF_ERR_INVALID, "n_taps must be >= 1");
        if (cfg->ode_substeps < 1) return fail(nullptr, TOF_ERR_INVALID, "ode_substeps must be >= 1");
        if (!cfg->x_centers || !cfg->e_centers || !cfg->neutron_speed || !cfg->neutron_dist || !cfg->xs_breaks ||
            !cfg->xs_coefs || !cfg->taps)
            return fail(nullptr, TOF_ERR_INVALID, "missing table pointer");
        for (int r = 0; r < cfg->n_runs; ++r)
            if (cfg->tof_bins[r] < cfg->n_taps) return fail(nullptr, TOF_ERR_INVALID, "tof_bins must be >= n_taps");
        if (cfg->model == TOF_MODEL_ADV && cfg->ndim < 2) return fail(nullptr, TOF_ERR_INVALID, "adv model needs ndim >= 2");
    } else {
        if (cfg->ndim != 3) return fail(nullptr, TOF_ERR_INVALID, "simple model has ndim == 3");
        if (cfg->tof_bins[0] > 1024) return fail(nullptr, TOF_ERR_INVALID, "simple model supports at most 1024 TOF bins");
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, TOF_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TOF_ERR_NO_DEVICE, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, TOF_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, TOF_ERR_NO_DEVICE,
                    std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                        "; libtofgpu is built for sm_100a only and has no fallback path");

    tof_ctx *ctx = new tof_ctx();
    ctx->cfg = *cfg;
    auto bail = [&](int rc) {
        g_create_error = ctx->err;
        tof_destroy(ctx);
        return rc;
    };
#define TRY(x)                     \
    do {                           \
        int rc__ = (x);            \
        if (rc__) return bail(rc__); \
    } while (0)
#define CUC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return bail(TOF_ERR_CUDA);                                                             \
        }                                                                                          \
    } while (0)

    CUC(cudaSetDevice(cfg->device));
    CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&ctx->ev0));
    CUC(cudaEventCreate(&ctx->ev1));
    CUC(cudaEventCreateWithFlags(&ctx->ev_busy, cudaEventDisableTiming));
    CUC(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_counters), 2 * sizeof(unsigned long long)));
    ctx->stats.sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;

    DevModel &m = ctx->dm;
    m.model = cfg->model; m.ode_mode = cfg->ode_mode; m.ode_substeps = cfg->ode_substeps;
    m.ode_from_zero = cfg->ode_from_zero; m.prior_strict = cfg->prior_strict; m.nan_to_neginf = cfg->nan_to_neginf;
    m.ndim = cfg->ndim; m.n_runs = cfg->n_runs; m.x_bins = cfg->x_bins; m.e_bins = cfg->e_bins;
    m.n_taps = cfg->n_taps; m.conv_shift = (cfg->n_taps - 1) / 2; m.n_zero_deg = cfg->n_zero_deg;
    m.n_materials = cfg->n_materials; m.n_xs = cfg->n_xs;
    m.n_samples = cfg->n_samples; m.n_ev_per_loop = cfg->n_ev_per_loop; m.n_loops = cfg->n_loops;
    m.n_draws = cfg->n_loops * cfg->n_ev_per_loop;
    m.x_min = cfg->x_min; m.x_max = cfg->x_max; m.e_min = cfg->e_min; m.e_max = cfg->e_max;
    m.c = cfg->speed_of_light; m.m_d = cfg->mass_deuteron; m.m_n = cfg->mass_neutron; m.m_he3 = cfg->mass_he3;
    m.q_ddn = cfg->q_ddn; m.cell_length = cfg->cell_length; m.simple_neutron_base = cfg->simple_neutron_base;
    std::memcpy(m.bethe_A, cfg->bethe_A, sizeof(m.bethe_A));
    std::memcpy(m.bethe_B, cfg->bethe_B, sizeof(m.bethe_B));
    std::memcpy(m.prior_lo, cfg->prior_lo, sizeof(m.prior_lo));
    std::memcpy(m.prior_hi, cfg->prior_hi, sizeof(m.prior_hi));

    for (int r = 0; r < cfg->n_runs; ++r) {
        ctx->runs[r].tof_bins = cfg->tof_bins[r];
        ctx->runs[r].tof_min = cfg->tof_min[r];
        ctx->runs[r].tof_max = cfg->tof_max[r];
    }

    if (cell_model) {
        TRY(upload(ctx, cfg->x_centers, cfg->x_bins, &m.x_centers));
        TRY(upload(ctx, cfg->e_centers, cfg->e_bins, &m.e_centers));
        TRY(upload(ctx, cfg->neutron_speed, cfg->e_bins, &m.neutron_speed));
        {
            std::vector<double> rs(cfg->e_bins);
            for (int j = 0; j < cfg->e_bins; ++j) rs[j] = 1.0 / cfg->neutron_speed[j];   // correctly rounded (div_by_recip)
            TRY(upload(ctx, rs.data(), cfg->e_bins, &m.neutron_rspeed));
        }
        TRY(upload(ctx, cfg->xs_breaks, cfg->n_xs, &m.xs_breaks));
        TRY(upload(ctx, cfg->xs_coefs, (size_t)(cfg->n_xs - 1) * 4, &m.xs_coefs));
        TRY(upload(ctx, cfg->taps, cfg->n_taps, &m.taps));
        for (int r = 0; r < cfg->n_runs; ++r)
            TRY(upload(ctx, cfg->neutron_dist + (size_t)r * cfg->x_bins, cfg->x_bins, &ctx->runs[r].neutron_dist));
        if (cfg->n_zero_deg > 0) {
            if (!cfg->zero_deg_times || !cfg->zero_deg_weights) {
                ctx->err = "n_zero_deg > 0 needs zero_deg_times / zero_deg_weights";
                return bail(TOF_ERR_INVALID);
            }
            TRY(upload(ctx, cfg->zero_deg_times, (size_t)cfg->e_bins * cfg->n_zero_deg, &m.zd_times));
            TRY(upload(ctx, cfg->zero_deg_weights, (size_t)cfg->e_bins * cfg->n_zero_deg, &m.zd_weights));
        }
        // cross-section interval lookup table: uniform cells, entry = interval holding the cell's left edge
        {
            const double *bp = cfg->xs_breaks;
            const int nb = cfg->n_xs;
            double min_gap = bp[1] - bp[0];
            for (int i = 1; i + 1 < nb; ++i) {
                if (!(bp[i + 1] > bp[i])) { ctx->err = "xs_breaks must increase"; return bail(TOF_ERR_INVALID); }
                min_gap = std::min(min_gap, bp[i + 1] - bp[i]);
            }
            const double span = bp[nb - 1] - bp[0];
            int lut_n = (int)std::min<double>(4096.0, std::max<double>(1.0, std::ceil(span / min_gap - 1e-9)));
            std::vector<unsigned char> lut(lut_n);
            if (nb - 1 > 256) { ctx->err = "at most 256 cross-section intervals"; return bail(TOF_ERR_INVALID); }
            int iv = 0;
            for (int cidx = 0; cidx < lut_n; ++cidx) {
                const double left = bp[0] + span * (double)cidx / (double)lut_n;
                while (iv + 2 < nb && left >= bp[iv + 1]) ++iv;
                lut[cidx] = (unsigned char)iv;
            }
            TRY(upload(ctx, lut.data(), lut.size(), &m.xs_lut));
            m.xs_lut_n = lut_n;
            m.xs_lut_lo = bp[0];
            m.xs_lut_inv = (double)lut_n / span;
        }
    }

    const bool use_range = cfg->ode_mode == TOF_ODE_RANGE && (cfg->model == TOF_MODEL_ADV || cfg->model == TOF_MODEL_SIMULT);
    if (cfg->precision != TOF_PRECISION_FP64 && cfg->precision != TOF_PRECISION_FP32) {
        ctx->err = "precision must be TOF_PRECISION_FP64 or TOF_PRECISION_FP32";
        return bail(TOF_ERR_INVALID);
    }
    if (cfg->precision == TOF_PRECISION_FP32 && !(cfg->model == TOF_MODEL_ADV && cfg->ode_mode == TOF_ODE_RANGE)) {
        ctx->err = "TOF_PRECISION_FP32 is built for the adv/intermediate model with TOF_ODE_RANGE only";
        return bail(TOF_ERR_INVALID);
    }
    if (use_range) {
        if (!cfg->t1_coefs || !cfg->rng_breaks || !cfg->rng_bins || !cfg->rng_coefs || !cfg->rng_lut || cfg->rng_n < 1 ||
            cfg->t1_n < 1 || cfg->rng_lut_n < 1 || !(cfg->rng_u_max > 0.0)) {
            ctx->err = "TOF_ODE_RANGE needs the range tables (t1_*, rng_*)";
            return bail(TOF_ERR_INVALID);
        }
        if (cfg->rng_degree != 7) { ctx->err = "rng_degree must be 7 (the built kernels evaluate degree-7 weight polynomials)"; return bail(TOF_ERR_INVALID); }
        if (cfg->ode_substeps < 1) { ctx->err = "ode_substeps must be >= 1"; return bail(TOF_ERR_INVALID); }
        // np.rint counts are accumulated as u32 per TOF bin: bound the total by nSamples/(dE*dx)
        {
            const double de = (cfg->e_max - cfg->e_min) / cfg->e_bins, dx = (cfg->x_max - cfg->x_min) / cfg->x_bins;
            if ((double)cfg->n_samples / (de * dx) * 1.01 + cfg->x_bins * (double)cfg->e_bins > 4.0e9) {
                ctx->err = "n_samples too large for the 32-bit TOF counters of the range kernel; use TOF_ODE_RK4";
                return bail(TOF_ERR_CAPACITY);
            }
        }
        const int P = cfg->rng_degree, RW = P + 3, Mi = cfg->rng_n;
        std::vector<double> rec((size_t)Mi * RW);
        for (int j = 0; j < Mi; ++j) {
            // header: [break that ends the interval (+inf for the last one), E-bin as an int in the low word]
            rec[(size_t)j * RW + 0] = (j + 1 < Mi) ? cfg->rng_breaks[j + 1] : HUGE_VAL;
            {
                const bool shared = (j > 0 && cfg->rng_bins[j - 1] == cfg->rng_bins[j]) ||
                                    (j + 1 < Mi && cfg->rng_bins[j + 1] == cfg->rng_bins[j]);
                long long bits = (long long)(unsigned int)cfg->rng_bins[j];
                if (shared) bits |= (long long)0x8000000000000000ull;   // sign bit: cell needs atomics
                double asd;
                std::memcpy(&asd, &bits, sizeof(asd));
                rec[(size_t)j * RW + 1] = asd;
            }
            for (int k = 0; k <= P; ++k) rec[(size_t)j * RW + 2 + k] = cfg->rng_coefs[(size_t)j * (P + 1) + k];
        }
        TRY(upload(ctx, rec.data(), rec.size(), &m.rng_rec));
        if (cfg->precision == TOF_PRECISION_FP32) {
            // float records: degree-3 interpolant of the degree-7 weight polynomial at the Chebyshev nodes of the
            // interval, monomials in dt; the fit is checked against the FP64 polynomial and refused if it is not
            // single-precision accurate (very wide E-bins)
            std::vector<float> recf((size_t)Mi * RANGE_RWF, 0.0f);
            double worst = 0.0;
            for (int j = 0; j < Mi; ++j) {
                const double *a = cfg->rng_coefs + (size_t)j * (P + 1);
                const double w = ((j + 1 < Mi) ? cfg->rng_breaks[j + 1] : cfg->rng_u_max) - cfg->rng_breaks[j];
                auto p7 = [&](double d) { double v = a[P]; for (int k = P - 1; k >= 0; --k) v = v * d + a[k]; return v; };
                constexpr int N = RANGE_PF + 1;
                long double A[N][N + 1];
                for (int k = 0; k < N; ++k) {
                    const long double sk = 0.5L * (1.0L - cosl((2 * k + 1) * 3.14159265358979323846264338327950288L / (2 * N)));
                    long double pw = 1.0L;
                    for (int c = 0; c < N; ++c) { A[k][c] = pw; pw *= sk; }          // monomials in s = dt / w
                    A[k][N] = (long double)p7((double)(sk * w));
                }
                for (int c = 0; c < N; ++c) {                                          // Gaussian elimination, partial pivoting
                    int piv = c;
                    for (int r = c + 1; r < N; ++r) if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
                    for (int q = 0; q <= N; ++q) std::swap(A[c][q], A[piv][q]);
                    for (int r = 0; r < N; ++r) {
                        if (r == c) continue;
                        const long double f = A[r][c] / A[c][c];
                        for (int q = c; q <= N; ++q) A[r][q] -= f * A[c][q];
                    }
                }
                float cf[N];
                long double wp = 1.0L;
                for (int c = 0; c < N; ++c) { cf[c] = (float)(A[c][N] / A[c][c] / wp); wp *= w; }
                for (int t = 0; t <= 32; ++t) {
                    const double d = w * t / 32.0;
                    double v = cf[N - 1];
                    for (int k = N - 2; k >= 0; --k) v = v * d + (double)cf[k];
                    const double ref = p7(d);
                    if (ref != 0.0) worst = std::max(worst, std::fabs(v / ref - 1.0));
                }
                float *r = recf.data() + (size_t)j * RANGE_RWF;
                for (int c = 0; c < N; ++c) r[c] = cf[c];
                const int32_t bin = cfg->rng_bins[j];
                const int32_t shared = ((j > 0 && cfg->rng_bins[j - 1] == bin) || (j + 1 < Mi && cfg->rng_bins[j + 1] == bin)) ? 1 : 0;
                std::memcpy(r + 4, &bin, 4);
                std::memcpy(r + 5, &shared, 4);
                const double c0 = (double)(A[0][N] / A[0][0]);      // constant term in FP64 (n*c0 is added in FP64)
                std::memcpy(r + 6, &c0, 8);
            }
            if (!(worst <= 3e-7)) {
                ctx->err = "TOF_PRECISION_FP32: the degree-3 weight polynomials miss single precision on this E binning (max rel. error " +
                           std::to_string(worst) + "); use TOF_PRECISION_FP64";
                return bail(TOF_ERR_CAPACITY);
            }
            TRY(upload(ctx, recf.data(), recf.size(), &m.rng_rec_f32));
        }
        TRY(upload(ctx, cfg->t1_coefs, (size_t)cfg->t1_n * 8, &m.t1_coefs));
        TRY(upload(ctx, cfg->rng_lut, (size_t)cfg->rng_lut_n, &m.rng_lut));
        {
            bool ident = Mi == cfg->e_bins;
            for (int j = 0; ident && j < Mi; ++j) ident = cfg->rng_bins[j] == j;
            if (const char *v = std::getenv("TOFGPU_RANGE_PLANNED")) ident = ident && std::atoi(v) != 0;   // tuning / A-B knob
            m.rng_identity = ident ? 1 : 0;
        }
        m.t1_q = cfg->t1_q; m.t1_key_lo = cfg->t1_key_lo; m.t1_n = cfg->t1_n; m.rng_degree = P; m.rng_n = Mi;
        m.rng_lut_n = cfg->rng_lut_n; m.rng_sign = cfg->rng_sign; m.rng_u_max = cfg->rng_u_max;
        m.rng_lut_inv = (double)cfg->rng_lut_n / cfg->rng_u_max; m.e_tab_lo = cfg->e_tab_lo; m.e_tab_hi = cfg->e_tab_hi;
    }
    if (cfg->model == TOF_MODEL_ADV && use_range) {
        const int P = cfg->rng_degree, Mi = cfg->rng_n;
        // the range kernel reuses the (x,E) cell histogram as the density buffer (T doubles) and keeps E-bins and
        // interval indices in 16-bit tables: refuse what does not fit instead of overrunning shared memory
        if ((long long)cfg->tof_bins[0] > (long long)cfg->x_bins * cfg->e_bins) {
            ctx->err = "TOF_ODE_RANGE needs tof_bins <= x_bins*e_bins (the density reuses the cell histogram); use TOF_ODE_RK4";
            return bail(TOF_ERR_CAPACITY);
        }
        if (cfg->e_bins > 65535 || Mi > 65535 || cfg->rng_lut_n > 65535) {
            ctx->err = "TOF_ODE_RANGE keeps E-bins and table intervals in 16 bits: e_bins, rng_n and rng_lut_n must be <= 65535";
            return bail(TOF_ERR_CAPACITY);
        }
        if (const char *v = std::getenv("TOFGPU_RANGE_THREADS")) {
            const int nt = std::atoi(v);
            if (!range_variant(nt, P)) { ctx->err = "TOFGPU_RANGE_THREADS must be 512, 640, 800 or 1024"; return bail(TOF_ERR_INVALID); }
            ctx->rng_nt = nt;
        }
        ctx->lay_full = range_layout(cfg->x_bins, cfg->e_bins, cfg->tof_bins[0], cfg->x_bins * cfg->e_bins, Mi, P, cfg->n_taps,
                                     cfg->rng_lut_n, Mi);
        ctx->adv_smem = ctx->lay_full.total;
        if ((int)ctx->adv_smem > ctx->max_smem_optin) {
            ctx->err = "range kernel needs " + std::to_string(ctx->adv_smem) + " B of shared memory per CTA; device offers " +
                       std::to_string(ctx->max_smem_optin);
            return bail(TOF_ERR_CAPACITY);
        }
        ctx->f32 = cfg->precision == TOF_PRECISION_FP32 && m.n_draws < RANGE_STREAM_MIN;   // big draw sets: FP64 streaming walk
        ctx->stats.fp32_active = ctx->f32 ? 1 : 0;
        if (ctx->f32 && !range_variant(ctx->rng_nt, P, true)) { ctx->err = "TOF_PRECISION_FP32 runs with 512 or 1024 threads"; return bail(TOF_ERR_INVALID); }
        AdvKernel k = range_variant(ctx->rng_nt, P, ctx->f32);
        CUC(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->adv_smem));
        CUC(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, ctx->rng_nt, ctx->adv_smem));
        ctx->stats.smem_bytes = (int)ctx->adv_smem;
        ctx->stats.threads = ctx->rng_nt;
        ctx->stats.ctas_per_sm = occ;
        // banded launch: size the histogram so that two 512-thread CTAs share an SM
        {
            const char *env = std::getenv("TOFGPU_RANGE_BANDED");
            const bool want = !(env && std::atoi(env) == 0);
            if (const char *v = std::getenv("TOFGPU_BAND_THREADS")) {      // tuning knob: 384, 448 or 512
                const int nt = std::atoi(v);
                if (nt > 512 || !range_variant(nt, P)) { ctx->err = "TOFGPU_BAND_THREADS must be 384, 448 or 512"; return bail(TOF_ERR_INVALID); }
                ctx->band_nt = nt;
            }
            if (ctx->f32 && ctx->band_nt != 512) { ctx->err = "TOF_PRECISION_FP32 runs the banded launch with 512 threads"; return bail(TOF_ERR_INVALID); }
            AdvKernel kband = range_variant(ctx->band_nt, P, ctx->f32);
            const int per_cta = (int)prop.sharedMemPerBlockOptin / 2 - 2048;   // two CTAs + the per-CTA reservation
            const int rcap = std::min(Mi, 128);
            const size_t fixed = range_smem_bytes(cfg->x_bins, cfg->e_bins, cfg->tof_bins[0], 0, rcap, P, cfg->n_taps, cfg->rng_lut_n, Mi);
            long long hcap = ((long long)per_cta - (long long)fixed) / 8;
            hcap = std::min<long long>(hcap, (long long)cfg->x_bins * cfg->e_bins - 1);
            hcap &= ~1LL;                                   // the records behind the histogram are read with 16-byte loads
            if (want && kband && hcap >= (long long)cfg->x_bins * 8 && hcap >= cfg->tof_bins[0]) {
                ctx->band_hcap = (int)hcap;
                ctx->band_rcap = rcap;
                ctx->lay_band = range_layout(cfg->x_bins, cfg->e_bins, cfg->tof_bins[0], (int)hcap, rcap, P, cfg->n_taps,
                                             cfg->rng_lut_n, Mi);
                ctx->band_smem = ctx->lay_band.total;
                CUC(cudaFuncSetAttribute(kband, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->band_smem));
                CUC(cudaFuncSetAttribute(kband, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                int occb = 0;
                CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occb, kband, ctx->band_nt, ctx->band_smem));
                ctx->band_ctas = occb;
                ctx->band_enabled = occb >= 2;
                if (ctx->band_enabled && !ctx->f32 && ctx->band_nt == 512 && P == 7 && m.rng_identity && m.n_draws <= RANGE_TILE) {
                    AdvKernel kp = adv_planned_kernel<512, 7, false>;
                    CUC(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->band_smem));
                    CUC(cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    int occp = 0;
                    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, kp, 512, ctx->band_smem));
                    ctx->planned = occp >= 2;
                }
            }
        }
        // ---- adv_zrank_kernel (shipped for single-tile draw sets): rank-hint table + its own shared-memory layout ----
        {
            const char *env = std::getenv("TOFGPU_RANGE_ZRANK");
            const bool want = !(env && std::atoi(env) == 0);
            const bool multi = m.n_draws > RANGE_TILE;       // big draw sets: adv_zrank_multi_kernel (tiles, no hint tables)
            if (want && !ctx->f32 && P == 7 && m.rng_identity && cfg->x_bins >= 1 && (!multi || cfg->x_bins <= 32 * (512 / 32 - 1))) {
                const int per_cta = (int)prop.sharedMemPerBlockOptin / 2 - 2048;
                const int rcap = std::min(Mi, 128);
                const size_t fixed = zrank_layout(cfg->x_bins, cfg->e_bins, cfg->tof_bins[0], 0, rcap, P, cfg->n_taps, Mi).total;
                long long hcap = ((long long)per_cta - (long long)fixed) / 8;
                hcap = std::min<long long>(hcap, (long long)cfg->x_bins * cfg->e_bins);
                hcap &= ~1LL;                               // the records behind the histogram are read with 16-byte loads
                if (hcap >= (long long)cfg->x_bins * 8 && hcap >= cfg->tof_bins[0]) {
                    ctx->zr_hcap = (int)hcap;
                    ctx->zr_rcap = rcap;
                    ctx->lay_zr = zrank_layout(cfg->x_bins, cfg->e_bins, cfg->tof_bins[0], (int)hcap, rcap, P, cfg->n_taps, Mi);
                    ctx->zr_smem = ctx->lay_zr.total;
                    int occz = 0;
                    // (measured on the benchmark shape: 384 threads / 80 registers -9 %, 576 threads / 56 registers -13 %)
                    AdvKernel kz = adv_zrank_kernel<512, 7, false>, kzp = adv_zrank_kernel<512, 7, true>, kz3 = adv_zrank_kernel<512, 7, false, ZR_TPITCH>;
                    for (AdvKernel k2 : {kz, kzp, kz3}) {
                        CUC(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->zr_smem));
                        CUC(cudaFuncSetAttribute(k2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    }
                    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occz, kz, 512, ctx->zr_smem));
                    if (multi) {
                        AdvKernel km = adv_zrank_multi_kernel<512, 7>;
                        CUC(cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->zr_smem));
                        CUC(cudaFuncSetAttribute(km, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                        int occm = 0;
                        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occm, km, 512, ctx->zr_smem));
                        ctx->zrank_multi = occm >= 2;
                    } else if (occz >= 2) {
                        // Theta[j][i] = u^-1(U_j - delta_i): invert the T1 table u(E) on a dense grid (a hint: 1e-5 keV is plenty)
                        auto host_t1 = [&](double E) {
                            uint64_t bits;
                            std::memcpy(&bits, &E, 8);
                            const int hi = (int)(bits >> 32);
                            const int key = hi >> (20 - cfg->t1_q);
                            int idx = key - cfg->t1_key_lo;
                            idx = idx < 0 ? 0 : (idx > cfg->t1_n - 1 ? cfg->t1_n - 1 : idx);
                            const uint64_t mb = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
                            double mant;
                            std::memcpy(&mant, &mb, 8);
                            const double cc = (double)(key & ((1 << cfg->t1_q) - 1));
                            const double t = (mant - 1.0) * (double)(1 << (cfg->t1_q + 1)) - (2.0 * cc + 1.0);
                            const double *k = cfg->t1_coefs + 8 * (size_t)idx;
                            double acc = k[7];
                            for (int q = 6; q >= 0; --q) acc = acc * t + k[q];
                            return acc;
                        };
                        const int NG = 1 << 18;
                        std::vector<double> ug(NG), eg(NG);
                        const double e_a = cfg->e_tab_lo, e_b = cfg->e_tab_hi;
                        double run_max = -1e300;
                        for (int k = 0; k < NG; ++k) {
                            eg[k] = e_a + (e_b - e_a) * ((double)k / (double)NG);
                            double u = host_t1(eg[k]);
                            run_max = u > run_max ? u : run_max;       // monotone up to the fit error: make it so
                            ug[k] = run_max;
                        }
                        const double x_start = cfg->ode_from_zero ? 0.0 : cfg->x_centers[0];
                        // rows j = -1 .. M + 1 (a pair visit prefetches j, j + 1, j + 2 around its window; the pad rows are
                        // read by idle lanes only); up to ZR_TPITCH rows of the cell the pitch is that compile-time constant
                        const int pitch = cfg->x_bins <= ZR_TPITCH ? ZR_TPITCH : cfg->x_bins;
                        std::vector<float> th((size_t)(Mi + 3) * pitch, 1e30f);
                        for (int j = 0; j <= Mi; ++j) {
                            const double U = cfg->rng_breaks[j];
                            for (int i = 0; i < cfg->x_bins; ++i) {
                                const double tq = U - cfg->rng_sign * (cfg->x_centers[i] - x_start);
                                float v;
                                if (!(tq >= ug[0])) v = -1e30f;
                                else if (tq > ug[NG - 1]) v = 1e30f;
                                else {
                                    size_t k = std::upper_bound(ug.begin(), ug.end(), tq) - ug.begin();   // ug[k-1] <= tq < ug[k]
                                    k = k < 1 ? 1 : (k > (size_t)NG - 1 ? (size_t)NG - 1 : k);
                                    const double du = ug[k] - ug[k - 1];
                                    const double fr = du > 0.0 ? (tq - ug[k - 1]) / du : 0.0;
                                    v = (float)(eg[k - 1] + fr * (eg[k] - eg[k - 1]));
                                }
                                th[(size_t)(j + 1) * pitch + i] = v;
                            }
                        }
                        TRY(upload(ctx, th.data(), th.size(), &m.rank_theta));
                        m.rank_theta += pitch;                   // row 0 of the view = interval 0
                        m.rank_stride = pitch;
                        ctx->zrank = true;
                        // what tof_get_stats reports as "the main model kernel" is now this one
                        ctx->stats.smem_bytes = (int)ctx->zr_smem;
                        ctx->stats.threads = ctx->zr_nt;
                        ctx->stats.ctas_per_sm = occz;
                    }
                }
            }
        }
    } else if (cfg->model == TOF_MODEL_ADV) {
        if (const char *v = std::getenv("TOFGPU_ADV_VARIANT")) {
            int nt = 0, dpt = 0;
            if (std::sscanf(v, "%dx%d", &nt, &dpt) == 2 && adv_variant(nt, dpt, 1)) {
                ctx->adv_nt = nt;
                ctx->adv_dpt = dpt;
            } else {
                ctx->err = std::string("TOFGPU_ADV_VARIANT='") + v + "' is not a built variant";
                return bail(TOF_ERR_INVALID);
            }
        }
        ctx->adv_smem = adv_smem_bytes(cfg->x_bins, cfg->e_bins, cfg->tof_bins[0], cfg->n_xs, cfg->n_taps, m.xs_lut_n);
        if ((int)ctx->adv_smem > ctx->max_smem_optin) {
            ctx->err = "model needs " + std::to_string(ctx->adv_smem) + " B of shared memory per CTA; device offers " +
                       std::to_string(ctx->max_smem_optin);
            return bail(TOF_ERR_CAPACITY);
        }
        AdvKernel k = adv_variant(ctx->adv_nt, ctx->adv_dpt, cfg->n_materials);
        CUC(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->adv_smem));
        CUC(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, ctx->adv_nt, ctx->adv_smem));
        ctx->stats.smem_bytes = (int)ctx->adv_smem;
        ctx->stats.threads = ctx->adv_nt;
        ctx->stats.ctas_per_sm = occ;
    } else if (cfg->model == TOF_MODEL_SIMPLE) {
        ctx->stats.threads = 256;
    } else if (cfg->model == TOF_MODEL_ONEBD) {
        if (cfg->ndim != 3 + 2 * cfg->n_runs) { ctx->err = "oneBD model has ndim == 3 + 2*n_runs"; return bail(TOF_ERR_INVALID); }
        if (cfg->stop_n < 4 || !cfg->stop_coefs || !cfg->attenuation || !cfg->taps2 || cfg->n_taps2 < 1 || !(cfg->stop_step > 0.0)) {
            ctx->err = "oneBD model needs stop_coefs / attenuation / taps2";
            return bail(TOF_ERR_INVALID);
        }
        TRY(upload(ctx, cfg->stop_coefs, (size_t)cfg->x_bins * (cfg->stop_n - 1) * 4, &m.stop_coefs));
        TRY(upload(ctx, cfg->attenuation, (size_t)cfg->x_bins, &m.attenuation));
        TRY(upload(ctx, cfg->taps2, (size_t)cfg->n_taps2, &m.taps2));
        m.stop_n = cfg->stop_n; m.n_taps2 = cfg->n_taps2; m.stop_lo = cfg->stop_lo; m.stop_step = cfg->stop_step;
        m.beam_energy = cfg->beam_energy;
        int tmax = 0;
        for (int r = 0; r < cfg->n_runs; ++r) tmax = std::max(tmax, cfg->tof_bins[r]);
        // one private histogram copy per warp when they fit (csi_oneBD's 10 x 100 grid), fewer for the grid of the
        // posterior-predictive variant (ppcTools_oneBD.py: 20 x 400 cells = 64 KB per copy)
        m.onebd_copies = 256 / 32;
        for (;;) {
            ctx->onebd_smem = onebd_smem_bytes(m.onebd_copies, cfg->x_bins, cfg->e_bins, tmax, cfg->n_xs, cfg->n_taps, cfg->n_taps2,
                                               cfg->stop_n, m.xs_lut_n);
            if ((int)ctx->onebd_smem <= ctx->max_smem_optin / 2 || m.onebd_copies == 1) break;
            m.onebd_copies /= 2;
        }
        if ((int)ctx->onebd_smem > ctx->max_smem_optin) {
            ctx->err = "oneBD kernel needs " + std::to_string(ctx->onebd_smem) + " B of shared memory per CTA";
            return bail(TOF_ERR_CAPACITY);
        }
        CUC(cudaFuncSetAttribute(onebd_run_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->onebd_smem));
        CUC(cudaFuncSetAttribute(onebd_run_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, onebd_run_kernel<256>, 256, ctx->onebd_smem));
        ctx->stats.smem_bytes = (int)ctx->onebd_smem;
        ctx->stats.threads = 256;
        ctx->stats.ctas_per_sm = occ;
    } else {
        if (cfg->ndim < 4 + cfg->n_runs) { ctx->err = "simult model needs ndim >= 4 + n_runs"; return bail(TOF_ERR_INVALID); }
        int tmax = 0;
        for (int r = 0; r < cfg->n_runs; ++r) tmax = std::max(tmax, cfg->tof_bins[r]);
        int occ = 0;
        if (use_range) {
            if (cfg->rng_degree != 7) { ctx->err = "the simult range kernel is built for rng_degree 7"; return bail(TOF_ERR_INVALID); }
            ctx->simult_smem = simult_range_smem_bytes(cfg->x_bins, cfg->e_bins, tmax, cfg->rng_n, 7, cfg->n_taps, cfg->rng_lut_n);
        } else {
            ctx->simult_smem = simult_smem_bytes(256, cfg->x_bins, cfg->e_bins, tmax, cfg->n_xs, cfg->n_taps, m.xs_lut_n);
        }
        if ((int)ctx->simult_smem > ctx->max_smem_optin) {
            ctx->err = "simult kernel needs " + std::to_string(ctx->simult_smem) + " B of shared memory per CTA";
            return bail(TOF_ERR_CAPACITY);
        }
        if (use_range) {
            CUC(cudaFuncSetAttribute(simult_range_kernel<256, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->simult_smem));
            CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, simult_range_kernel<256, 7>, 256, ctx->simult_smem));
            ctx->simult_rk4_smem = simult_smem_bytes(256, cfg->x_bins, cfg->e_bins, tmax, cfg->n_xs, cfg->n_taps, m.xs_lut_n);
            if ((int)ctx->simult_rk4_smem <= ctx->max_smem_optin)
                CUC(cudaFuncSetAttribute(simult_run_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->simult_rk4_smem));
            else
                ctx->simult_rk4_smem = 0;
        } else {
            CUC(cudaFuncSetAttribute(simult_run_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->simult_smem));
            CUC(cudaFuncSetAttribute(simult_run_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, simult_run_kernel<256>, 256, ctx->simult_smem));
        }
        ctx->stats.smem_bytes = (int)ctx->simult_smem;
        ctx->stats.threads = 256;
        ctx->stats.ctas_per_sm = occ;
    }
#undef TRY
#undef CUC
    *out = ctx;
    return TOF_OK;
}

void tof_destroy(tof_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->busy && ctx->ev_busy) cudaEventSynchronize(ctx->ev_busy);
    for (DeviceBuf *b : {&ctx->d_theta, &ctx->d_out, &ctx->d_spectra, &ctx->d_cells, &ctx->d_counts, &ctx->d_partial, &ctx->d_work,
                         &ctx->d_queue, &ctx->d_split, &ctx->d_tickets, &ctx->d_ens, &ctx->d_nan, &ctx->d_stage, &ctx->d_wide})
        if (b->p) cudaFree(b->p);
    for (int r = 0; r < TOF_MAX_RUNS; ++r)
        if (ctx->d_zlut[r].p) cudaFree(ctx->d_zlut[r].p);
    for (int r = 0; r < TOF_MAX_RUNS; ++r)
        for (DeviceBuf *b : {&ctx->d_z[r][0], &ctx->d_z[r][1], &ctx->d_obs[r], &ctx->d_obs_idx[r], &ctx->d_obs_val[r]})
            if (b->p) cudaFree(b->p);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_busy) cudaEventDestroy(ctx->ev_busy);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int tof_set_observables(tof_ctx *ctx, int run, const double *counts, int nbins) {
    if (!ctx || !counts) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (int rc = check_run(ctx, run)) return rc;
    if (nbins != ctx->cfg.tof_bins[run]) return fail(ctx, TOF_ERR_INVALID, "observables length != tof_bins[run]");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    std::vector<double> obs(counts, counts + nbins);
    if (ctx->cfg.model == TOF_MODEL_SIMULT || ctx->cfg.model == TOF_MODEL_ONEBD)
        for (double &v : obs)
            if (v == 0.0) v = 1.0;  // simultFit.py:391-392, applied to the private copy
    std::vector<int> idx;
    std::vector<double> val;
    for (int t = 0; t < nbins; ++t)
        if (obs[t] != 0.0) {  // NaN observables are kept so that they poison the sum like np.dot would
            idx.push_back(t);
            val.push_back(obs[t]);
        }
    DevRun &r = ctx->runs[run];
    if (int rc = upload_into(ctx, ctx->d_obs[run], obs.data(), obs.size(), &r.obs)) return rc;
    if (int rc = upload_into(ctx, ctx->d_obs_idx[run], idx.data(), idx.size(), &r.obs_nz_idx)) return rc;
    if (int rc = upload_into(ctx, ctx->d_obs_val[run], val.data(), val.size(), &r.obs_nz_val)) return rc;
    r.n_obs_nz = (int)idx.size();
    ctx->have_obs[run] = true;
    return TOF_OK;
}

int tof_set_draws(tof_ctx *ctx, int run, int stream, const double *values, int64_t n) {
    if (!ctx || (!values && n > 0)) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (int rc = check_run(ctx, run)) return rc;
    if (stream < 0 || stream > 1) return fail(ctx, TOF_ERR_INVALID, "stream must be 0 or 1");
    const long long need = ctx->dm.n_draws;
    if (stream == 0 && n != need)
        return fail(ctx, TOF_ERR_INVALID, "stream 0 needs n_loops*n_ev_per_loop = " + std::to_string(need) + " draws");
    if (stream == 1 && ctx->cfg.model == TOF_MODEL_SIMPLE && n != need)
        return fail(ctx, TOF_ERR_INVALID, "simple model: stream 1 needs as many uniforms as normals");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const double *d = nullptr;
    if (stream == 0 && ctx->cfg.ode_mode == TOF_ODE_RANGE && ctx->cfg.model == TOF_MODEL_SIMULT) {
        // E0 = beamE - (eLoss + scale*exp(s*z)) falls with z: sort every loop's block descending so that E0 ascends
        std::vector<double> sorted(values, values + n);
        const long long per = ctx->cfg.n_ev_per_loop;
        for (long long l = 0; l < ctx->cfg.n_loops; ++l)
            std::sort(sorted.begin() + l * per, sorted.begin() + (l + 1) * per, [](double a, double b) { return a > b || (b != b && a == a); });   // NaNs last: a strict weak order
        if (int rc = upload_into(ctx, ctx->d_z[run][stream], sorted.data(), (size_t)n, &d)) return rc;
    } else if (stream == 0 && ctx->cfg.ode_mode == TOF_ODE_RANGE && ctx->cfg.model == TOF_MODEL_ADV) {
        // the model is a symmetric function of the draws; the range kernel walks them in ascending order
        std::vector<double> sorted(values, values + n);
        std::sort(sorted.begin(), sorted.end(), [](double a, double b) { return a < b || (b != b && a == a); });
        if (int rc = upload_into(ctx, ctx->d_z[run][stream], sorted.data(), (size_t)n, &d)) return rc;
        // adv_zrank_kernel: walker-independent rank lookup over z, zlut[c] = first draw with z >= z_lo + c*(z_hi - z_lo)/ZR_LUT
        DevRun &rz = ctx->runs[run];
        rz.zlut = nullptr;
        if (ctx->zrank && n >= 1 && n <= 65535) {
            long long n_fin = 0;
            while (n_fin < n && std::isfinite(sorted[(size_t)n_fin])) ++n_fin;     // NaNs sort last; +-inf draws: hints off
            const bool all_finite = n_fin == n;
            const double z_lo = sorted[0], z_hi = sorted[(size_t)n - 1];
            if (all_finite && z_hi - z_lo > 1e-3) {                // (a degenerate draw set gets no lookup: the banded pair runs)
                std::vector<unsigned short> lut(ZR_LUT + 1);
                const double inv = (double)ZR_LUT / (z_hi - z_lo);
                size_t dpos = 0;
                for (int cc = 0; cc <= ZR_LUT; ++cc) {
                    const double start = z_lo + (double)cc / inv;
                    while (dpos < (size_t)n && sorted[dpos] < start) ++dpos;
                    lut[cc] = (unsigned short)dpos;
                }
                if (const char *v = std::getenv("TOFGPU_ZR_NOHINT"))   // debugging aid: every search walks from draw 0
                    if (std::atoi(v)) std::fill(lut.begin(), lut.end(), (unsigned short)0);
                // the top entry must be "no draw" for thresholds beyond the last draw, but draws equal to z_hi sit in the
                // last cell: lut[ZR_LUT] is the first draw with z >= z_hi, a low hint as required (never beyond the answer)
                const unsigned short *dl = nullptr;
                if (int rc = upload_into(ctx, ctx->d_zlut[run], lut.data(), lut.size(), &dl)) return rc;
                rz.zlut = dl;
                rz.zlut_lo = z_lo;
                rz.zlut_inv = inv;
            }
        }
    } else {
        if (int rc = upload_into(ctx, ctx->d_z[run][stream], values, (size_t)n, &d)) return rc;
    }
    DevRun &r = ctx->runs[run];
    if (stream == 0) { r.z = d; r.n_z = n; } else { r.z1 = d; r.n_z1 = n; }
    ctx->have_z[run][stream] = true;
    return TOF_OK;
}

int tof_lnprob_batch_device(tof_ctx *ctx, const double *d_theta, int64_t n, double *d_out, void *stream) {
    if (!ctx || (n > 0 && (!d_theta || !d_out))) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (n < 0) return fail(ctx, TOF_ERR_INVALID, "n < 0");
    if (int rc = ready(ctx, true)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    ModelOut o{};
    o.lnprob = d_out;
    next_call_key(ctx);
    return launch_model(ctx, d_theta, n, 0, o, static_cast<cudaStream_t>(stream));
}

int tof_lnprob_batch(tof_ctx *ctx, const double *theta, int64_t n, double *out) {
    if (!ctx || (n > 0 && (!theta || !out))) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (n < 0) return fail(ctx, TOF_ERR_INVALID, "n < 0");
    if (n == 0) return TOF_OK;
    if (int rc = ready(ctx, true)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t tb = (size_t)n * ctx->cfg.ndim * sizeof(double), ob = (size_t)n * sizeof(double);
    if (int rc = ensure(ctx, ctx->d_theta, tb)) return rc;
    if (int rc = ensure(ctx, ctx->d_out, ob)) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->d_theta.p, theta, tb, cudaMemcpyHostToDevice, ctx->stream));
    ModelOut o{};
    o.lnprob = static_cast<double *>(ctx->d_out.p);
    next_call_key(ctx);
    if (int rc = launch_model(ctx, static_cast<const double *>(ctx->d_theta.p), n, 0, o, ctx->stream)) return rc;
    CU(ctx, cudaMemcpyAsync(out, ctx->d_out.p, ob, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return TOF_OK;
}

int tof_model_batch(tof_ctx *ctx, const double *theta, int64_t n, int run, int stage, double *spectra) {
    if (!ctx || (n > 0 && (!theta || !spectra))) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (int rc = check_run(ctx, run)) return rc;
    if (stage < TOF_STAGE_COUNTS || stage > TOF_STAGE_SPREAD) return fail(ctx, TOF_ERR_INVALID, "bad stage");
    if (ctx->cfg.model == TOF_MODEL_SIMPLE && stage == TOF_STAGE_SPREAD)
        return fail(ctx, TOF_ERR_INVALID, "the simple model has no timing-response stage");
    if (n <= 0) return n == 0 ? TOF_OK : fail(ctx, TOF_ERR_INVALID, "n < 0");
    if (int rc = ready(ctx, false)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const int T = ctx->cfg.tof_bins[run];
    const size_t tb = (size_t)n * ctx->cfg.ndim * sizeof(double), sb = (size_t)n * T * sizeof(double);
    if (int rc = ensure(ctx, ctx->d_theta, tb)) return rc;
    if (int rc = ensure(ctx, ctx->d_spectra, sb)) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->d_theta.p, theta, tb, cudaMemcpyHostToDevice, ctx->stream));
    ModelOut o{};
    o.spectra = static_cast<double *>(ctx->d_spectra.p);
    o.stage = stage;
    next_call_key(ctx);
    if (int rc = launch_model(ctx, static_cast<const double *>(ctx->d_theta.p), n, run, o, ctx->stream)) return rc;
    CU(ctx, cudaMemcpyAsync(spectra, ctx->d_spectra.p, sb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return TOF_OK;
}

int tof_cell_counts_batch(tof_ctx *ctx, const double *theta, int64_t n, int run, int64_t *counts) {
    if (!ctx || (n > 0 && (!theta || !counts))) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (int rc = check_run(ctx, run)) return rc;
    if (ctx->cfg.model == TOF_MODEL_SIMPLE) return fail(ctx, TOF_ERR_INVALID, "the simple model has no (x, E) cells");
    if (n <= 0) return n == 0 ? TOF_OK : fail(ctx, TOF_ERR_INVALID, "n < 0");
    if (int rc = ready(ctx, false)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t cells = (size_t)ctx->cfg.x_bins * ctx->cfg.e_bins;
    const size_t tb = (size_t)n * ctx->cfg.ndim * sizeof(double), cb = (size_t)n * cells * sizeof(long long);
    if (int rc = ensure(ctx, ctx->d_theta, tb)) return rc;
    if (int rc = ensure(ctx, ctx->d_cells, cb)) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->d_theta.p, theta, tb, cudaMemcpyHostToDevice, ctx->stream));
    ModelOut o{};
    o.cells = static_cast<long long *>(ctx->d_cells.p);
    next_call_key(ctx);
    if (int rc = launch_model(ctx, static_cast<const double *>(ctx->d_theta.p), n, run, o, ctx->stream)) return rc;
    CU(ctx, cudaMemcpyAsync(counts, ctx->d_cells.p, cb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return TOF_OK;
}

int tof_deuteron_counts_batch(tof_ctx *ctx, const double *theta, int64_t n, int run, int64_t *counts) {
    if (!ctx || (n > 0 && (!theta || !counts))) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (int rc = check_run(ctx, run)) return rc;
    const bool simult_ok = ctx->cfg.model == TOF_MODEL_SIMULT && (ctx->cfg.ode_mode == TOF_ODE_RK4 || ctx->simult_rk4_smem > 0);
    if (!simult_ok && ctx->cfg.model != TOF_MODEL_ONEBD)
        return fail(ctx, TOF_ERR_INVALID, "deuteron counts are built for TOF_MODEL_SIMULT and TOF_MODEL_ONEBD");
    if (n <= 0) return n == 0 ? TOF_OK : fail(ctx, TOF_ERR_INVALID, "n < 0");
    if (int rc = ready(ctx, false)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t cells = (size_t)ctx->cfg.x_bins * ctx->cfg.e_bins;
    const size_t tb = (size_t)n * ctx->cfg.ndim * sizeof(double), cb = (size_t)n * cells * sizeof(long long);
    if (int rc = ensure(ctx, ctx->d_theta, tb)) return rc;
    if (int rc = ensure(ctx, ctx->d_cells, cb)) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->d_theta.p, theta, tb, cudaMemcpyHostToDevice, ctx->stream));
    ModelOut o{};
    o.cells = static_cast<long long *>(ctx->d_cells.p);
    o.unweighted = 1;
    next_call_key(ctx);
    if (int rc = launch_model(ctx, static_cast<const double *>(ctx->d_theta.p), n, run, o, ctx->stream)) return rc;
    CU(ctx, cudaMemcpyAsync(counts, ctx->d_cells.p, cb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return TOF_OK;
}

int tof_stretch_propose(tof_ctx *ctx, const double *d_s, int64_t n, int64_t walker0, const double *d_comp, int64_t n_comp,
                        double a, uint64_t seed, int64_t step, int half, double *d_q, double *d_log_zz, void *stream) {
    if (!ctx || !d_s || !d_comp || !d_q || !d_log_zz) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (n <= 0 || n_comp <= 0 || !(a > 1.0)) return fail(ctx, TOF_ERR_INVALID, "bad stretch-move arguments");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    stretch_propose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_s, n, walker0, d_comp, n_comp, ctx->cfg.ndim, ctx->cfg.ndim, ctx->cfg.ndim, a, seed, step, half, d_q, d_log_zz);
    ctx->stats.kernel_launches += 1;
    CU(ctx, cudaGetLastError());
    return TOF_OK;
}

int tof_stretch_accept(tof_ctx *ctx, double *d_s, double *d_lnprob, int64_t n, int64_t walker0, const double *d_q,
                       const double *d_new_lnprob, const double *d_log_zz, uint64_t seed, int64_t step, int half,
                       int64_t *d_n_accept, void *stream) {
    if (!ctx || !d_s || !d_lnprob || !d_q || !d_new_lnprob || !d_log_zz) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (n <= 0) return fail(ctx, TOF_ERR_INVALID, "n <= 0");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    stretch_accept_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_s, d_lnprob, n, walker0, d_q, d_new_lnprob, d_log_zz, ctx->cfg.ndim, ctx->cfg.ndim, 1, seed, step, half,
        reinterpret_cast<long long *>(d_n_accept));
    ctx->stats.kernel_launches += 1;
    CU(ctx, cudaGetLastError());
    return TOF_OK;
}

int tof_ensemble_step(tof_ctx *ctx, double *d_pos, double *d_lnprob, int64_t n_walkers, int64_t n_steps, double a,
                      uint64_t seed, int64_t step0, int64_t *d_n_accept, void *stream) {
    if (!ctx || !d_pos || !d_lnprob) return fail(ctx, TOF_ERR_INVALID, "null argument");
    const int ndim = ctx->cfg.ndim;
    if (n_walkers < 2 || (n_walkers & 1)) return fail(ctx, TOF_ERR_INVALID, "The number of walkers must be even.");
    if (n_walkers < 2 * (int64_t)ndim)
        return fail(ctx, TOF_ERR_INVALID, "The number of walkers needs to be more than twice the dimension of your parameter space.");
    if (n_steps < 0 || !(a > 1.0)) return fail(ctx, TOF_ERR_INVALID, "bad stretch-move arguments");
    if (int rc = ready(ctx, true)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long h = n_walkers / 2;
    // scratch of one half-step: proposals q[h][ndim], log_zz[h], new_lnprob[h]
    if (int rc = ensure(ctx, ctx->d_ens, (size_t)h * (ndim + 2) * sizeof(double))) return rc;
    double *q = static_cast<double *>(ctx->d_ens.p), *log_zz = q + (size_t)h * ndim, *new_lp = log_zz + h;
    const unsigned grid = (unsigned)((h + 255) / 256);
    for (int64_t s = 0; s < n_steps; ++s) {
        for (int half = 0; half < 2; ++half) {
            double *sl = d_pos + (size_t)half * h * ndim;                 // the half that moves ...
            const double *comp = d_pos + (size_t)(1 - half) * h * ndim;   // ... against the complementary half
            double *lp = d_lnprob + (size_t)half * h;
            const long long walker0 = half * h;
            stretch_propose_kernel<<<grid, 256, 0, st>>>(sl, h, walker0, comp, h, ndim, ndim, ndim, a, seed, step0 + s, half, q, log_zz);
            ModelOut o{};
            o.lnprob = new_lp;
            ctx->cur_epoch = (1ull << 31) | (2 * (uint64_t)(step0 + s) + (uint64_t)half);   // per-evaluation draws: independent of the sharding
            ctx->cur_walker0 = walker0;
            if (int rc = launch_model(ctx, q, h, 0, o, st)) return rc;
            stretch_accept_kernel<<<grid, 256, 0, st>>>(sl, lp, h, walker0, q, new_lp, log_zz, ndim, ndim, 1, seed, step0 + s, half,
                                                        reinterpret_cast<long long *>(d_n_accept ? d_n_accept + walker0 : nullptr));
            ctx->stats.kernel_launches += 2;
        }
    }
    CU(ctx, cudaGetLastError());
    return TOF_OK;
}

int tof_ensemble_half_step(tof_ctx *ctx, double *d_state, int64_t n_walkers, int half, int64_t own0, int64_t n_own, double a,
                           uint64_t seed, int64_t step, int64_t *d_n_accept, void *stream) {
    if (!ctx || !d_state) return fail(ctx, TOF_ERR_INVALID, "null argument");
    const int ndim = ctx->cfg.ndim, ld = ndim + 1;
    if (n_walkers < 2 || (n_walkers & 1)) return fail(ctx, TOF_ERR_INVALID, "The number of walkers must be even.");
    const long long h = n_walkers / 2;
    if (half < 0 || half > 1 || own0 < 0 || n_own < 0 || own0 + n_own > h || !(a > 1.0))
        return fail(ctx, TOF_ERR_INVALID, "bad half-step arguments");
    if (n_own == 0) return TOF_OK;
    if (int rc = ready(ctx, true)) return rc;
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = ensure(ctx, ctx->d_ens, (size_t)n_own * (ndim + 2) * sizeof(double))) return rc;
    double *q = static_cast<double *>(ctx->d_ens.p), *log_zz = q + (size_t)n_own * ndim, *new_lp = log_zz + n_own;
    const long long walker0 = (long long)half * h + own0;
    double *sl = d_state + (size_t)walker0 * ld;                            // own rows of the half that moves
    const double *comp = d_state + (size_t)(1 - half) * h * ld;              // the whole complementary half
    const unsigned grid = (unsigned)((n_own + 255) / 256);
    stretch_propose_kernel<<<grid, 256, 0, st>>>(sl, n_own, walker0, comp, h, ndim, ld, ld, a, seed, step, half, q, log_zz);
    ModelOut o{};
    o.lnprob = new_lp;
    ctx->cur_epoch = (1ull << 31) | (2 * (uint64_t)step + (uint64_t)half);  // per-evaluation draws: independent of the sharding
    ctx->cur_walker0 = walker0;
    if (int rc = launch_model(ctx, q, n_own, 0, o, st)) return rc;
    stretch_accept_kernel<<<grid, 256, 0, st>>>(sl, sl + ndim, n_own, walker0, q, new_lp, log_zz, ndim, ld, ld, seed, step, half,
                                                reinterpret_cast<long long *>(d_n_accept ? d_n_accept + walker0 : nullptr));
    ctx->stats.kernel_launches += 2;
    CU(ctx, cudaGetLastError());
    return TOF_OK;
}

int tof_set_draw_mode(tof_ctx *ctx, int mode, uint64_t seed, uint64_t epoch0) {
    if (!ctx) return TOF_ERR_INVALID;
    if (mode != TOF_DRAWS_BOUND && mode != TOF_DRAWS_PER_EVALUATION) return fail(ctx, TOF_ERR_INVALID, "unknown draw mode");
    if (mode == TOF_DRAWS_PER_EVALUATION) {
        const tof_config &c = ctx->cfg;
        if (c.model == TOF_MODEL_SIMULT && c.ode_mode == TOF_ODE_RANGE)
            return fail(ctx, TOF_ERR_INVALID, "per-evaluation draws for the simultaneous fit need TOF_ODE_RK4 (the range kernel "
                                              "wants every loop's draws sorted)");
        if (c.model == TOF_MODEL_ADV && c.ode_mode == TOF_ODE_RANGE && (ctx->dm.n_draws > RANGE_TILE || ctx->f32))
            return fail(ctx, TOF_ERR_INVALID, "per-evaluation draws with TOF_ODE_RANGE need FP64 and n_draws <= " +
                                              std::to_string(RANGE_TILE) + " (one sorted tile per walker); use TOF_ODE_RK4");
        if (epoch0 >> 31) return fail(ctx, TOF_ERR_INVALID, "epoch0 must be < 2^31 (the upper half is used by the ensemble entry points)");
    }
    ctx->fresh = mode == TOF_DRAWS_PER_EVALUATION;
    ctx->fresh_seed = seed;
    ctx->fresh_epoch = epoch0;
    return TOF_OK;
}

int tof_generate_draws(tof_ctx *ctx, uint64_t epoch, int64_t walker, int run, int stream, int sorted, double *out, int64_t n) {
    if (!ctx || !out) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (n < 1 || (stream != 0 && stream != 1 && stream != 3) || run < 0 || run >= TOF_MAX_RUNS)
        return fail(ctx, TOF_ERR_INVALID, "bad arguments");
    if (sorted && (stream != 0 || n > 2 * 512)) return fail(ctx, TOF_ERR_INVALID, "sorted draws: stream 0, n <= 1024");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    DeviceBuf buf;
    if (int rc = ensure(ctx, buf, (size_t)n * sizeof(double))) return rc;
    DevRun r{};
    r.fresh = 1;
    r.fresh_seed = ctx->fresh_seed;
    r.fresh_epoch = epoch;
    r.fresh_walker0 = 0;
    fresh_draws_kernel<512><<<1, 512, 0, ctx->stream>>>(r, walker, run, stream, sorted, (int)n, static_cast<double *>(buf.p));
    cudaError_t e = cudaMemcpyAsync(out, buf.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(buf.p);
    if (e != cudaSuccess) return fail(ctx, TOF_ERR_CUDA, cudaGetErrorString(e));
    return TOF_OK;
}

int tof_get_stats(const tof_ctx *ctx, tof_stats *out) {
    if (!ctx || !out) return TOF_ERR_INVALID;
    *out = ctx->stats;
    out->band_ctas_per_sm = ctx->band_enabled ? ctx->band_ctas : 0;
    out->band_cells = ctx->band_enabled ? ctx->band_hcap : 0;
    out->band_queued_last = 0;
    out->wide_last = 0;
    // (before the first call: what a production call of this context will launch)
    const bool single = ctx->last_model_launches ? ctx->last_model_launches == 1
                                                 : ((ctx->zrank && ctx->runs[0].zlut != nullptr) || ctx->zrank_multi || !ctx->band_enabled);
    out->model_launches_per_call = (ctx->cfg.model == TOF_MODEL_ADV && ctx->cfg.ode_mode == TOF_ODE_RANGE) ? (single ? 1 : 2) : 0;
    // counters are read on the context's own stream, ordered after its last model launch (ev_busy): the host waits for
    // that launch only -- no device-wide or legacy-stream synchronisation
    const bool want_nan = ctx->d_nan.p != nullptr, want_q = (ctx->band_enabled || single) && ctx->d_work.p != nullptr;
    if ((want_nan || want_q) && ctx->h_counters) {
        cudaSetDevice(ctx->cfg.device);
        bool ok = true;
        if (ctx->busy) ok = cudaStreamWaitEvent(ctx->stream, ctx->ev_busy, 0) == cudaSuccess;
        ctx->h_counters[0] = ctx->h_counters[1] = 0ull;
        if (ok && want_nan)
            ok = cudaMemcpyAsync(ctx->h_counters + 0, ctx->d_nan.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
        if (ok && want_q)
            ok = cudaMemcpyAsync(ctx->h_counters + 1, static_cast<unsigned long long *>(ctx->d_work.p) + 2, sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
        if (ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
            if (want_nan) out->nan_results = (int64_t)ctx->h_counters[0];
            if (want_q) (single ? out->wide_last : out->band_queued_last) = (int64_t)ctx->h_counters[1];
        }
    }
    return TOF_OK;
}

int tof_set_stage_timing(tof_ctx *ctx, int enabled) {
    if (!ctx) return TOF_ERR_INVALID;
    if (enabled && !(ctx->cfg.model == TOF_MODEL_ADV && ctx->cfg.ode_mode == TOF_ODE_RANGE))
        return fail(ctx, TOF_ERR_INVALID, "stage timing is built into the adv/intermediate TOF_ODE_RANGE kernel only");
    if (enabled && (ctx->f32 || !range_variant_prof(ctx->rng_nt, ctx->cfg.rng_degree) || !range_variant_prof(ctx->band_nt, ctx->cfg.rng_degree)))
        return fail(ctx, TOF_ERR_INVALID, "stage timing needs an FP64 context with the default 512/1024-thread launches");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    if (enabled) {      // the instrumented instantiations need the same shared-memory opt-in as the shipped ones
        CU(ctx, cudaFuncSetAttribute(range_variant_prof(ctx->rng_nt, ctx->cfg.rng_degree), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->adv_smem));
        if (ctx->band_enabled)
            CU(ctx, cudaFuncSetAttribute(range_variant_prof(ctx->band_nt, ctx->cfg.rng_degree), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->band_smem));
        if (ctx->planned) {
            AdvKernel kp = adv_planned_kernel<512, 7, true>;
            CU(ctx, cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->band_smem));
        }
    }
    if (enabled && !ctx->d_stage.p) {
        if (int rc = ensure(ctx, ctx->d_stage, (TOF_N_STAGES + 1) * sizeof(unsigned long long))) return rc;
        CU(ctx, cudaMemset(ctx->d_stage.p, 0, (TOF_N_STAGES + 1) * sizeof(unsigned long long)));
    }
    ctx->stage_timing = enabled != 0;
    return TOF_OK;
}

int tof_get_stage_cycles(tof_ctx *ctx, uint64_t cycles[TOF_N_STAGES + 1]) {
    if (!ctx || !cycles) return TOF_ERR_INVALID;
    if (!ctx->d_stage.p) return fail(ctx, TOF_ERR_STATE, "stage timing was never enabled");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    if (ctx->busy) CU(ctx, cudaEventSynchronize(ctx->ev_busy));
    CU(ctx, cudaMemcpy(cycles, ctx->d_stage.p, (TOF_N_STAGES + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemset(ctx->d_stage.p, 0, (TOF_N_STAGES + 1) * sizeof(unsigned long long)));
    return TOF_OK;
}

int tof_set_timing(tof_ctx *ctx, int enabled) {
    if (!ctx) return TOF_ERR_INVALID;
    ctx->timing = enabled != 0;
    ctx->timed = false;
    return TOF_OK;
}

int tof_last_kernel_ms(tof_ctx *ctx, float *ms) {
    if (!ctx || !ms) return fail(ctx, TOF_ERR_INVALID, "null argument");
    if (!ctx->timed) return fail(ctx, TOF_ERR_STATE, "no timed launch recorded (call tof_set_timing(ctx, 1) first)");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    CU(ctx, cudaEventSynchronize(ctx->ev1));
    CU(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return TOF_OK;
}

int tof_measure_fp64_peak(tof_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return fail(ctx, TOF_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    const int blocks = ctx->stats.sm_count * 16, threads = 256, iters = 1 << 16;
    double *d = nullptr;
    CU(ctx, cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t a, b;
    CU(ctx, cudaEventCreate(&a));
    CU(ctx, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(ctx, cudaEventRecord(a, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9);
        CU(ctx, cudaEventRecord(b, ctx->stream));
        CU(ctx, cudaEventSynchronize(b));
        float ms = 0.f;
        CU(ctx, cudaEventElapsedTime(&ms, a, b));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    ctx->stats.kernel_launches += 5;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return TOF_OK;
}

}  // extern "C"
