// Shared kernel-side declarations: what a model kernel is asked to produce.
#pragma once
#include "tof_device.cuh"

namespace tof {

// Byte offsets of the range kernel's regions inside dynamic shared memory, computed once on the host
// (range_layout, adv_range.cuh) and passed with the launch: inside the kernel every region pointer is "base + one
// constant-bank word" -- cheap to re-materialise under register pressure, where the arithmetic that derives the
// offsets from (T, E, X, hcap, rcap, ...) was being re-executed inside the hot loops (11 % of all instructions).
struct RangeLayout {
    unsigned pa, rec, svd, staps, scratch, sdelta, lut, ulut, srow, hlo, sbrk, sbin, total;
};

// Outputs requested from a model kernel.  Production: only `lnprob`.
struct ModelOut {
    double *lnprob;       // [n] (adv/simple) or [n][n_runs] partials (simult)
    double *spectra;      // optional [n][T] at `stage`
    long long *cells;     // optional [n][X][E] integer cell counts
    int stage;
    int unweighted;       // simult RK4 kernel: cells = unweighted (x,E) histogram of the LAST loop (ppcTools.py:151-157)
    unsigned long long *stage_cycles; // optional [TOF_N_STAGES + 1]: SM clock cycles per stage summed over CTAs, then walkers (adv range kernel)
    unsigned long long *nan_count;   // optional: walkers inside the prior whose log-probability came out NaN (diagnostics)
    unsigned long long *work;  // optional global work counter (persistent CTAs take walkers dynamically)
    // range kernel, banded launch: capacity of the cell histogram (cells) and of the staged T2 records
    int hcap, rcap;
    RangeLayout lay;                 // shared-memory layout of this launch (range kernel)
    int *queue_out;                  // walkers that do not fit the banded layout ...
    unsigned long long *queue_count; // ... and how many
    const int *queue_in;             // full-size launch: process queue_in[0 .. *queue_count)
    // few walkers x big draw sets: n_split CTAs share a walker; partial cell histograms meet in global scratch
    int n_split;
    int split_stride;                // doubles per partial histogram
    double *split_scratch;           // [n][n_split][split_stride]
    unsigned int *split_tickets;     // [n] arrival counters (zero between calls)
    // adv_zrank_kernel: per-CTA scratch histogram in global memory (L2) for walkers whose E-band does not fit the
    // banded shared-memory histogram, split_stride doubles per CTA
    double *wide_scratch;
};

}  // namespace tof
