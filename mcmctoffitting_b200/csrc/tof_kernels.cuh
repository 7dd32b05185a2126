// Model kernels (sm_100a).  One CTA evaluates one walker (one (walker, run) pair for the multi-run models): Monte-Carlo
// draws are spread over the threads, every histogram lives in shared memory, and only theta (in) and lnprob (out)
// touch HBM.
#pragma once
#include "tof_common.cuh"
#include "adv_rk4.cuh"       // adv / intermediate, RK4 per x-interval (the oracle's scheme)
#include "adv_range.cuh"     // adv / intermediate, range-energy tables (shipped)
#include "adv_planned.cuh"   // ... its lean per-phase cut for single-tile walkers (kept as the A/B twin of the next one)
#include "adv_zrank.cuh"     // ... shipped: walker-independent rank hints, fused normalisation, one launch per call
#include "simple_model.cuh"  // config 1
#include "simult_model.cuh"  // config 4 (RK4 and range variants)
#include "onebd_model.cuh"   // csi_oneBD production model
#include "sampler.cuh"       // stretch move, FP64 peak
