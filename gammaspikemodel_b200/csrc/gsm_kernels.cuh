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
    // EXACT mode (gsm_upload_stubs): stub lists, CSR by tree; nullptr otherwise.  nstubs then holds the derived counts.
    const int64_t* stub_ptr;     // [T + 1]
    const int32_t* stub_branch;  // [S]
    const double* stub_time;     // [S] relative position along the branch
};

// Device scratch of the lean path, owned by whoever owns the stream the evaluation runs on (one per handle slot):
//   rows  n_trees * gsm_scratch_row_bytes() bytes — per-tree sums handed from the lean kernel to the post kernel
//   redo  (n_trees + 4) int32 — [0], [1]: ping-pong counters (zero-initialised once), [4..]: trees the lean kernel
//         hands back to the generic path
//   epoch evaluation counter (selects the live counter)
struct GsmScratch {
    void* rows = nullptr;
    int32_t* redo = nullptr;
    uint32_t epoch = 0;
};
size_t gsm_scratch_row_bytes();

// Per-handle launch context: device facts, probe results and tuning overrides.  Owned by one gsm_handle (calls on a handle
// are serialised by its caller), so nothing here is process-global: two handles on two threads or two devices never share it.
#define GSM_CTX_MAX_KERNELS 48
struct GsmLaunchCtx {
    int device = -1;
    int sm_count = 0;
    size_t smem_per_sm = 0;      // cudaDevAttrMaxSharedMemoryPerMultiprocessor
    size_t smem_optin = 0;       // cudaDevAttrMaxSharedMemoryPerBlockOptin
    int a_dyn[12] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};   // shared-window address of the dynamic segment per (lean config, rates)
    const void* prepared[GSM_CTX_MAX_KERNELS] = {};    // kernels whose attributes were set on this handle's device
    int n_prepared = 0;
    // gsm_set_option (tuning / tests; 0 = automatic)
    int force_generic = 0, block = 0, minb = 0;
    int producer = 0;            // 1: 128 x 3 runs with a producer warp (second warpgroup, setmaxnreg); measured equal to the
                                 // four-warp CTA on C4 and 3 % slower on C3 once its 136-register consumers spill: off
    int warp_kernel = 1;         // trees of up to 256 taxa are evaluated one warp per tree (gsm_lean_warp.cuh); 0: by the CTA kernel
    size_t smem_pad = 0;         // extra dynamic shared memory per CTA of gsm_lean_kernel (occupancy experiments)
    int persistent = 0;          // 1: the persistent lean kernel where it applies (measured slower than one CTA per tree: off)
};
cudaError_t gsm_launch_ctx_init(GsmLaunchCtx& ctx, int device);

// Evaluate the whole batch: one result row per tree, optional per-branch rates / weighted spikes
// (pitch = stride).  Returns the number of kernels launched (>=1) or a negative cudaError_t.
// Without scratch (rows == nullptr) every wiring runs the generic kernel.
int gsm_launch_eval(GsmLaunchCtx& ctx, const GsmBatch& b, GsmScratch& scratch, double* d_out_tree, double* d_out_rates,
                    double* d_out_wspikes, cudaStream_t stream);

// Proposal batch (gsm_eval_dirty): dst rows [p] = src rows [tree_of_prop[p0 + p]] with the CSR edits of proposal p0 + p
// applied; dst has the same n_nodes / stride as src.  Device pointers throughout; a null edit array = unchanged.
struct GsmProposals {
    const int32_t* tree_of_prop;   // [P]
    const int64_t* dirty_ptr;      // [P + 1]
    const int32_t* dirty_nodes;    // [D]
    const double* new_height;      // [D] or nullptr
    const double* new_rate;
    const double* new_spike;
    const int32_t* new_nstubs;
    const double* height_scale;    // [P] or nullptr: Tree.scale factor of the internal non-fake nodes
    const double* new_params;      // [P][GSM_NPARAM] or nullptr
    const uint8_t* new_use_spike;  // [P] or nullptr
};
struct GsmBatchMut {   // writable twin of GsmBatch's arrays
    double* height; int32_t* parent; int32_t* node_count; double* rate; double* spike; int32_t* nstubs; double* params;
    uint8_t* use_spike;
};
// Device scratch of gsm_eval_dirty (all device pointers; owned by the handle):
struct GsmDirtyWork {
    void* children = nullptr;      // [T][stride] int2: the two children of every node (-1 = none)
    void* base_rows = nullptr;     // [T] accumulators of the resident trees (gsm_base_row_bytes() each)
    void* base_K = nullptr;        // [T] per-tree constants (gsm_base_k_bytes() each)
    int32_t* tree_flag = nullptr;  // [T] 1: the tree is not binary
    double* out_base = nullptr;    // [T][GSM_NOUT] result rows of the resident trees
    int32_t* full_list = nullptr;  // [P] proposals that take the full path
    int32_t* counters = nullptr;   // [2] number of entries of full_list, argument-error flag
    double* out_all = nullptr;     // [P][GSM_NOUT] result rows of every proposal
};
size_t gsm_base_row_bytes();
size_t gsm_base_k_bytes();
// children table + accumulators of the resident batch (once per resident batch)
int gsm_launch_dirty_base(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmDirtyWork& w, cudaStream_t stream);
// validates every proposal (counters[1]), evaluates the eligible ones incrementally into out_all and lists the others
int gsm_launch_dirty_delta(const GsmBatch& b, const GsmProposals& pr, int64_t n_props, const GsmDirtyWork& w, cudaStream_t stream);
// full path: dst rows [q] = src rows [tree_of_prop[list[q0 + q]]] with that proposal's edits applied
int gsm_launch_patch(const GsmBatch& src, const GsmProposals& pr, const GsmDirtyWork& w, int64_t q0, int64_t count, const GsmBatchMut& dst,
                     cudaStream_t stream);
int gsm_launch_scatter_rows(const double* d_rows, const int32_t* d_list, int64_t q0, int64_t count, double* d_out_all, cudaStream_t stream);

// EXACT mode: counts[t][i] = number of listed stubs of tree t on branch i (Stubs.getNStubsOnBranch, Stubs:357-364);
// rows of `stride` int32, every entry of a row is written.
int gsm_launch_stub_counts(const GsmBatch& b, int32_t* d_counts, cudaStream_t stream);

// Log-time stub sampling for whole trees (Stubs:389-458, BSP:315-389).  d_trees: the selected trees (nullptr = trees 0 .. n_sel-1);
// items are (selected tree, node) in tree-major order, n_sel * n_nodes of them.
int gsm_launch_stub_len(const GsmBatch& b, int64_t n_sel, const int64_t* d_trees, int64_t* d_lens, double* d_sums, cudaStream_t stream);
int gsm_launch_scan(const int64_t* d_lens, int64_t* d_ptr, int64_t n, cudaStream_t stream);
int gsm_launch_stub_fill(const GsmBatch& b, int64_t n_sel, const int64_t* d_trees, const int64_t* d_ptr, const double* d_sums, double* d_cdf,
                         int64_t cap, cudaStream_t stream);
int gsm_launch_stub_choose(const GsmBatch& b, int64_t n_sel, const int64_t* d_trees, const double* d_u, int32_t* d_out, cudaStream_t stream);

// fp64 pipe micro-benchmark: thread-level DFMA per second of the current device (best of `reps` launches).
int gsm_launch_dfma_peak(const GsmLaunchCtx& ctx, double* d_scratch, int reps, double* fma_per_sec, cudaStream_t stream);

// Fills the device-side sum_{i=2..n} log(i) table of the current device (idempotent).
cudaError_t gsm_init_tables();

// BranchSpikePrior.getCumulativeProbs for (tree, node) of the resident batch; d_cdf has room for `cap`
// doubles, d_len receives the Java array length.
int gsm_launch_stub_cdf(const GsmBatch& b, int64_t tree, int32_t node, double* d_cdf, int32_t cap, int32_t* d_len,
                        cudaStream_t stream);
