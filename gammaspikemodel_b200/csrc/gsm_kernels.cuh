// gsm_kernels.cuh — launch interface between the C-ABI layer (gsm_capi.cu) and the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsm_b200.h"

// One resident batch in HBM: structure-of-arrays, tree-major rows of `stride` elements
// (stride % 32 == 0, so every row of every array starts 128-byte aligned).
struct GsmBatch {
    int64_t n_trees;
    int32_t n_nodes;       // max nodes per tree
    int64_t stride;        // row pitch in elements
    const double* height;  // [T][stride]
    const int32_t* parent; // [T][stride], -1 = root
    const int32_t* node_count;  // [T] or nullptr
    const double* rate;    // [T][stride] or nullptr (rates not wired)
    const double* spike;   // [T][stride]
    const int32_t* nstubs; // [T][stride] or nullptr
    const double* params;  // [T][GSM_NPARAM]
    const uint8_t* use_spike;   // [T] or nullptr
    uint32_t flags;
    int32_t stub_mode;
};

// Evaluate the whole batch: one result row per tree, optional per-branch rates / weighted spikes
// (pitch = stride).  Returns the number of kernels launched (>=1) or a negative cudaError_t.
int gsm_launch_eval(const GsmBatch& b, double* d_out_tree, double* d_out_rates, double* d_out_wspikes,
                    cudaStream_t stream);

// Fills the device-side sum_{i=2..n} log(i) table of the current device (idempotent).
cudaError_t gsm_init_tables();

// BranchSpikePrior.getCumulativeProbs for (tree, node) of the resident batch; d_cdf has room for `cap`
// doubles, d_len receives the Java array length.
int gsm_launch_stub_cdf(const GsmBatch& b, int64_t tree, int32_t node, double* d_cdf, int32_t cap, int32_t* d_len,
                        cudaStream_t stream);
