/*
 * gsm_b200.h — C ABI of libgsm_b200.so: the B200 (sm_100a) implementation of GammaSpikeModel's
 * per-branch compute, evaluated for batches of tree states at once.
 *
 * This is the drop-in boundary.  The reference is a pure-Java BEAST 2 package with
 * no native interface of its own, so each entry point below names the Java method(s) whose body it
 * replaces; the Java classes keep their public API and call these through Panama (java.lang.foreign)
 * or JNI — see INTEGRATION.md for the binding stub.  Signatures use only int32_t / int64_t /
 * uint32_t / double / pointers: no structs by value, no callbacks, no C++ or torch types.
 *
 * Citations are relative to the reference tree (/root/reference):
 *   PRCM  src/gammaspike/clockmodel/PunctuatedRelaxedClockModel.java
 *   BSP   src/gammaspike/distribution/BranchSpikePrior.java
 *   BRP   src/gammaspike/distribution/BranchRatePrior.java
 *   STP   src/gammaspike/distribution/StumpedTreePrior.java
 *   Stubs src/gammaspike/tree/Stubs.java
 *   SPL   src/gammaspike/logger/SaltativeProportionLogger.java
 *   SpikeSize src/gammaspike/clockmodel/SpikeSize.java
 *
 * Conventions
 *   - A batch is T trees ("states") of at most N nodes each.  Host arrays are tree-major, dense:
 *     element (t, i) at [t * n_nodes + i].  Node numbering is BEAST's (Node.getNr()).
 *   - parent[i] = -1 for the root.  Leaf / sampled-ancestor / "fake" status is derived on the device
 *     from (height, parent) exactly as beast.base.evolution.tree.Node does (isLeaf, isDirectAncestor,
 *     isFake), so topology is one int32 per node.
 *   - Every function returns GSM_OK (0) or a GSM_E_* code; gsm_last_error() gives the text.
 *     An invalid *state* (e.g. lambda <= mu, a negative spike) is NOT an error: the reference returns
 *     Double.NEGATIVE_INFINITY (STP:160-164, BSP:60-75, BRP:36-50) and so do we, in the result row.
 *   - One handle = one chain/batch; calls on one handle are serialised by the caller; different
 *     handles may be used from different threads (one CUDA stream per handle, no global state).
 *   - There is no CPU fallback: if no CUDA device is usable gsm_create fails.
 */
#ifndef GSM_B200_H
#define GSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSM_ABI_VERSION 2

/* ---- status codes ---- */
#define GSM_OK             0
#define GSM_E_ARG          1   /* bad argument */
#define GSM_E_CUDA         2   /* CUDA runtime error (text in gsm_last_error) */
#define GSM_E_STATE        3   /* call order: e.g. eval before upload */
#define GSM_E_UNSUPPORTED  4   /* e.g. n_nodes above GSM_MAX_NODES */
#define GSM_E_NOMEM        5

/* ---- wiring flags: which optional BEAST Inputs are connected (gsm_set_mode) ---- */
#define GSM_F_HAS_INDICATOR        (1u << 0)   /* PRCM.indicator wired (PRCM:34); value per tree = use_spike[t] */
#define GSM_F_NO_SPIKE_DATED_TIPS  (1u << 1)   /* PRCM.noSpikeOnDatedTips = true (PRCM:42, 185-187) */
#define GSM_F_HAS_RATES            (1u << 2)   /* PRCM.rates wired (PRCM:36, 213) */
#define GSM_F_RELAXED_FALSE        (1u << 3)   /* PRCM.relaxed wired and false (PRCM:216-220) */
#define GSM_F_STUBS_TO_CLOCK       (1u << 4)   /* PRCM.stubs wired (PRCM:30, 192-195) */
#define GSM_F_STUBS_TO_SPIKE_PRIOR (1u << 5)   /* BSP.stubs wired (BSP:28, 77) */
#define GSM_F_STUBS_TO_TREE_PRIOR  (1u << 6)   /* STP.stubs wired (STP:46, 219) */
#define GSM_F_IGNORE_TREE_PRIOR    (1u << 7)   /* STP.ignoreTreePrior (STP:49, 174) */
#define GSM_F_IGNORE_STUB_PRIOR    (1u << 8)   /* STP.ignoreStubPrior (STP:50, 219) */
#define GSM_F_COND_SAMPLING        (1u << 9)   /* STP.conditionOnSampling (STP:26, 446) */
#define GSM_F_COND_RHO_SAMPLING    (1u << 10)  /* STP.conditionOnRhoSampling (STP:28, 455) */
#define GSM_F_COND_ROOT            (1u << 11)  /* STP.conditionOnRoot (STP:31, 179) */
#define GSM_F_ORIGIN               (1u << 12)  /* STP.origin wired (STP:24, 179-180); value = params[GSM_P_ORIGIN] */
#define GSM_F_BSP_HAS_INDICATOR    (1u << 13)  /* BSP.indicator wired (BSP:35, 347) — used by gsm_stub_cdf only */

/* ---- Stubs.StubMode (Stubs:36-40) ---- */
#define GSM_STUBS_NOSTUB  0   /* integrate over stub counts: getNStubsOnBranch == -1 (Stubs:348-350) */
#define GSM_STUBS_COUNTS  1   /* stubsPerBranch, dimension N-1 (Stubs:200, 352-355) */
#define GSM_STUBS_EXACT   2   /* reversible-jump placement (Stubs:42-54): the stub list of gsm_upload_stubs defines the
                                 per-branch counts (Stubs:357-364) and the point-process density STP:573-600 / 639-663 */

/* ---- per-tree parameter row: params[t * GSM_NPARAM + k] ---- */
#define GSM_NPARAM 9
#define GSM_P_CLOCK_RATE  0   /* clock.rate / GSMclockRate  (PRCM:228) */
#define GSM_P_SPIKE_MEAN  1   /* GSMspikeMean               (PRCM:204) */
#define GSM_P_SPIKE_SHAPE 2   /* GSMspikeShape              (BSP:58)   */
#define GSM_P_CLOCK_SD    3   /* GSMclockSD                 (BRP:33)   */
#define GSM_P_LAMBDA      4   /* STP.getLambda()            (STP:672-690) */
#define GSM_P_MU          5   /* STP.getMu()                (STP:692-708) */
#define GSM_P_PSI         6   /* STP.getPsi()               (STP:710-728) */
#define GSM_P_RHO         7   /* STP.getRho()               (STP:730-737) */
#define GSM_P_ORIGIN      8   /* STP.origin                 (STP:668-670) */

/* ---- per-tree result row: out_tree[t * GSM_NOUT + k] ---- */
#define GSM_NOUT 8
#define GSM_O_STP_LOGP    0   /* StumpedTreePrior.calculateTreeLogLikelihood (STP:150-284), before the host-side
                                 nExtantTaxa check (STP:196-200; cross-call state kept by the caller) */
#define GSM_O_TREE_LOGP   1   /* its treeLogP component (STP:167) */
#define GSM_O_STUB_LOGP   2   /* its stubLogP component (STP:168) */
#define GSM_O_SPIKE_LOGP  3   /* BranchSpikePrior.calculateLogP (BSP:53-187) */
#define GSM_O_RATE_LOGP   4   /* BranchRatePrior.calculateLogP (BRP:29-65) == template Prior+LogNormal */
#define GSM_O_SUM_BURST   5   /* SaltativeProportionLogger abruptChange (SPL:36, 53-54) */
#define GSM_O_SUM_DIST    6   /* SaltativeProportionLogger totalChange  (SPL:35, 44-49); ProportionOfSaltation = [5]/[6] */
#define GSM_O_N_EXTANT    7   /* number of extant leaves n (STP:190-195) */

#define GSM_MAX_NODES 8192    /* largest tree (nodes) the one-CTA-per-tree kernels stage in shared memory */

typedef struct gsm_handle gsm_handle;

/* ---- lifecycle ---- */
int  gsm_abi_version(void);
int  gsm_device_count(void);                 /* number of visible CUDA devices; 0 => nothing will work */
int  gsm_create(gsm_handle** out, int32_t device_id, int64_t max_trees, int32_t max_nodes);
int  gsm_destroy(gsm_handle* h);
const char* gsm_last_error(const gsm_handle* h);   /* h may be NULL: error of the last failed gsm_create on this thread */

/* Which Inputs are wired and which Stubs mode is used (replaces the `Input.get() != null` tests at
 * PRCM:181,185,192-195,213-216; BSP:77; STP:174,179,219 and Stubs.initAndValidate mode selection Stubs:138-245). */
int  gsm_set_mode(gsm_handle* h, int32_t stub_mode, uint32_t flags);

/* Tuning / test options of one handle; never needed for correct results.  Names: "host_chunk_mb" (MB of inputs per
 * streaming chunk of gsm_eval_host / gsm_eval_dirty, default 256), "force_generic" (1: every wiring runs the
 * statement-by-statement kernel), "block" / "minb" (launch configuration override, 0 = automatic), "producer" (1: the
 * 128 x 3 lean kernel runs with a producer warp), "persistent" (1: the persistent lean kernel where three of its CTAs fit
 * an SM; both variants give bit-identical rows and are off by default: measured no faster), "warp_kernel" (0: trees of up to 256 taxa are evaluated by the CTA kernel instead of one warp per
 * tree), "smem_pad" (extra dynamic
 * shared memory per CTA of the lean kernel, bytes: occupancy experiments).  Stored in the handle: the library reads no
 * environment variables and keeps no process-global mutable state. */
int  gsm_set_option(gsm_handle* h, const char* name, int64_t value);

/* ---- host-buffer path (what the Java plugin calls) ---------------------------------------------
 * All host arrays are caller-owned and are copied before the call returns. */

/* Tree states: Tree.getNodesAsArray() heights and parents.  node_count may be NULL (every tree has
 * n_nodes nodes) or give a per-tree count <= n_nodes (ragged batch; trailing entries ignored). */
int  gsm_upload_trees(gsm_handle* h, int64_t n_trees, int32_t n_nodes,
                      const double* height, const int32_t* parent, const int32_t* node_count);
/* GSMbranchRates (PRCM.rates), GSMspikes (PRCM.spikes / BSP.spikes), stubsPerBranch (Stubs:50).  A NULL
 * pointer leaves that array as it was.  nstubs has n_nodes entries per tree; entry i >= n-1 is ignored
 * in COUNTS mode (Stubs:200, 354). */
int  gsm_upload_branch_params(gsm_handle* h, const double* rate, const double* spike, const int32_t* nstubs);
/* Per-tree scalars (GSM_P_*) and the GSMuseSpikeModel indicator (NULL => true for every tree). */
int  gsm_upload_tree_params(gsm_handle* h, const double* params, const uint8_t* use_spike);

/* Stubs in exact (reversible-jump) mode: the stub list of every tree of the resident batch, CSR by tree (entries
 * [stub_ptr[t], stub_ptr[t+1]) belong to tree t; stub_ptr has n_trees + 1 entries).  stub_branch[j] / stub_time[j] are
 * Stubs.branchNr / Stubs.time (relative position along the branch, Stubs:625-640) of the Java entries 1 .. dim-1; the
 * dummy at index 0, which Stubs.includeStub always rejects (Stubs:285-287), is not passed.  On the device the list
 * replaces the O(S) scans of Stubs.getNStubsOnBranch (Stubs:357-364; used by PRCM:195, BSP:81, STP:232-240) and of
 * getStubLogPForBranchWith/WithoutSampling (STP:573-600, 639-663): per stub log 2 + log lambda + logQt(height)
 * (STP:796-831), per branch -E - log(count!); a stub outside its branch, or on a sampled ancestor when psi > 0 or
 * rho < 1, gives -inf; a stub on the root makes the Java throw (Stubs:631-634, caught at STP:258-265) and the state
 * -inf.  Stubs whose branch is outside [0, n) match no branch in the reference and are ignored.
 * Call after gsm_upload_trees; needed (GSM_E_STATE otherwise) whenever stub_mode is GSM_STUBS_EXACT and the stub prior
 * is evaluated (GSM_F_STUBS_TO_TREE_PRIOR without GSM_F_IGNORE_STUB_PRIOR).  Without a stub list, EXACT mode takes the
 * per-branch counts from the nstubs array of gsm_upload_branch_params (all n entries), which is enough for PRCM / BSP.
 * gsm_eval_host and gsm_eval_dirty do not take stub lists: in EXACT mode with the stub prior evaluated they return
 * GSM_E_UNSUPPORTED (reversible-jump moves change the list itself; upload the proposed state instead). */
int  gsm_upload_stubs(gsm_handle* h, const int64_t* stub_ptr, const int32_t* stub_branch, const double* stub_time);

/* Evaluate the resident batch.  out_tree [T][GSM_NOUT] is required.  out_rates [T][n_nodes] (optional)
 * receives PRCM.getRatesArray() (PRCM:267-276; getRateForBranch PRCM:225-242); out_wspikes [T][n_nodes]
 * (optional) receives GSMweightedSpikes = SpikeSize.getArrayValue(i) with clockModel wired
 * (SpikeSize:48-56 -> PRCM.getBurstSize PRCM:167-208).  Blocks until the outputs are written. */
int  gsm_eval(gsm_handle* h, double* out_tree, double* out_rates, double* out_wspikes);

/* One call = upload + eval + download for a whole host-resident batch, streamed through the device in
 * chunks so that copies overlap the kernels (use pinned host memory, gsm_host_alloc, for full PCIe speed).
 * Arguments as in the gsm_upload_* functions and gsm_eval. */
int  gsm_eval_host(gsm_handle* h, int64_t n_trees, int32_t n_nodes,
                   const double* height, const int32_t* parent, const int32_t* node_count,
                   const double* rate, const double* spike, const int32_t* nstubs,
                   const double* params, const uint8_t* use_spike,
                   double* out_tree, double* out_rates, double* out_wspikes);

/* MCMC proposal batch on the resident batch (BASELINE config 5: chains x proposals on one tree each).  The
 * reference re-evaluates all N branches after every proposal (BSP:294-295, STP:836-842; the proposals are the operators
 * SpikeUpRateDown.java:23-47, SpikeFlipOperator.java:30-67, StumpedTreeConstantDistanceOperator.java:61-183,
 * StumpedTreeScaler.java:63-161, ... and BEAST's scale / bit-flip operators on the hyper-parameters).  Here proposal p
 * is "tree tree_of_prop[p] of the resident batch with the nodes dirty_nodes[dirty_ptr[p] .. dirty_ptr[p+1]) replaced"
 * (CSR; the nodes of one proposal must be distinct): new_height / new_rate / new_spike / new_nstubs hold one value per
 * CSR entry, and a NULL array leaves that quantity as it is.  height_scale [n_props] (may be NULL) is BEAST's
 * Tree.scale(s) as the tree scalers use it (StumpedTreeScaler.java:115; Node.scale): the height of every internal node
 * that is not fake is multiplied by height_scale[p] before the CSR edits are applied (1.0 = no scaling), so a scale move
 * costs 8 bytes instead of N CSR entries.  new_params [n_props][GSM_NPARAM] / new_use_spike
 * [n_props] (either may be NULL) replace the per-tree scalars: hyper-parameter and indicator moves, which dirty every
 * branch.  out_tree [n_props][GSM_NOUT] receives the result row of each proposed state.  The resident batch itself is
 * not modified (an accepted proposal is uploaded by the caller).
 *   Cost is O(dirty): a proposal that edits at most 6 nodes and leaves the per-tree scalars alone is evaluated from the
 * cached accumulators of its resident tree -- only the edited nodes, their children, parents and siblings are
 * re-evaluated (old and new state) and the per-tree epilogue re-run; the cache is built on the first call after the
 * resident batch changed.  All other proposals (height_scale != 1, hyper-parameter / indicator moves, trees whose resident
 * state is invalid) are materialised on the device and evaluated in full by the kernels of gsm_eval.  Results equal
 * gsm_eval on the proposed state to the 1e-10 tolerance of the path (the incremental sums are regrouped, so not bit for
 * bit); tree / node indices are validated on the device (GSM_E_ARG). */
int  gsm_eval_dirty(gsm_handle* h, int64_t n_props, const int32_t* tree_of_prop, const int64_t* dirty_ptr,
                    const int32_t* dirty_nodes, const double* new_height, const double* new_rate,
                    const double* new_spike, const int32_t* new_nstubs, const double* height_scale,
                    const double* new_params, const uint8_t* new_use_spike, double* out_tree);

/* Number of proposals of the last gsm_eval_dirty call that took the full (materialise + re-evaluate) path. */
int64_t gsm_last_dirty_full_count(const gsm_handle* h);

/* BranchSpikePrior.getCumulativeProbs(mu, nodeNr) (BSP:315-389) for one branch of one resident tree, with
 * mu = StumpedTreePrior.getMeanStubNumber(h_node, h_parent) (STP:876-892) as Stubs.sampleNStubsOnBranch
 * computes it (Stubs:424-447).  Writes min(*len, cap) cumulative probabilities; *len = the Java array length. */
int  gsm_stub_cdf(gsm_handle* h, int64_t tree, int32_t node, double* cdf, int32_t cap, int32_t* len);

/* The same for whole trees in one call (a logged state asks for every branch: Stubs.sampleNStubs Stubs:389-401, the tree
 * loggers' Stubs.getArrayValue(i) for every node Stubs:666-674).  trees [n_sel] selects resident trees (NULL = trees
 * 0 .. n_sel-1); items are (selected tree, node) in tree-major order, n_sel * n_nodes of them.  CSR output: the array of
 * item j is cdf[cdf_ptr[j] .. cdf_ptr[j+1]); the root and branches with Poisson mean 0 (no draw in the reference,
 * Stubs:416-434) have empty arrays.  cdf_ptr must hold n_sel * n_nodes + 1 entries; *needed receives the total length; at
 * most `cap` doubles are written (call again with a larger buffer when *needed > cap). */
int  gsm_stub_cdf_trees(gsm_handle* h, int64_t n_sel, const int64_t* trees, int64_t* cdf_ptr, double* cdf, int64_t cap,
                        int64_t* needed);
/* Stubs.sampleNStubsOnBranch for every branch of the selected trees, given the uniforms (Stubs:410-458 with
 * Randomizer.randomChoice(getCumulativeProbs(mean, nodeNr)), Stubs:446-448): nstubs_out[j] = 0 for the root and for mean 0,
 * -1 for every branch when stubs are estimated (stub_mode != GSM_STUBS_NOSTUB, Stubs:412), else the index randomChoice
 * returns for U = u[j] (0 if U <= cf[0]; the first s with cf[s-1] < U <= cf[s]; cf.length if none).  The random numbers
 * stay with the host: BEAST draws one Randomizer.nextDouble() per branch that reaches randomChoice, in node order; u[j] of
 * the other branches is ignored.  Needs GSM_F_STUBS_TO_SPIKE_PRIOR (without a spike prior the reference uses
 * Randomizer.nextPoisson, BEAST's own algorithm: GSM_E_UNSUPPORTED). */
int  gsm_stub_choose(gsm_handle* h, int64_t n_sel, const int64_t* trees, const double* u, int32_t* nstubs_out);

/* ---- the step after the path: the tree-log line of one state (host side, no device work) ----------------
 * StumpedTreeLogger.toNewick (logger/StumpedTreeLogger.java:77-199; the same shape as BEAST's TreeWithMetaDataLogger): the
 * Newick string of the tree with per-node metadata `[&nstubs=..,rate=..,<id>=..]` and branch lengths, every number printed
 * as Java's Double.toString does (JDK 19+ rule: shortest decimal that rounds to the double; decimal notation for
 * 1e-3 <= |d| < 1e7, else d.dddE[-]n; "NaN", "Infinity").  left / right: Node.getLeft / getRight numbers (-1 = none; they fix
 * the print order), parent (-1 = root), height: as uploaded.  nstubs (NULL = no stubs wired; the root prints none,
 * STL:92), rate (NULL = no clock model; values = out_rates of gsm_eval), n_meta metadata functions with ids separated by
 * '\n', values [n_meta][n_nodes] (e.g. out_wspikes for GSMweightedSpikes, GSMbranchRates) and dimensions (a node number
 * >= the dimension prints nothing, STL:159, 187).  Matrix-valued metadata and the reversible-jump stub-location entries
 * (STL:103-123) are not produced.  Returns the length of the string (written if it fits cap - 1, NUL-terminated), -1 on
 * bad arguments.  The caller adds "tree STATE_<n> = " and ";" (STL:59-67). */
int64_t gsm_format_newick(int32_t n_nodes, const int32_t* left, const int32_t* right, const int32_t* parent,
                          const double* height, const int32_t* nstubs, const double* rate, int32_t n_meta,
                          const char* meta_ids, const double* meta_values, const int32_t* meta_dim, char* out, int64_t cap);
/* The inverse reader: one logged tree back into arrays, as util/GetEffectiveRates.java:30-83 reads them through BEAST's
 * tree-log parser.  newick: `(...)[&key=value,...]:length` as gsm_format_newick / StumpedTreeLogger write it (leaf labels are
 * node numbers + 1, STL:87; an optional trailing ';').  Node numbers on return: a leaf keeps the number of its label, internal
 * nodes follow in post-order, the root is last (BEAST's convention).  parent (-1 = root) is required; left / right / length
 * (root: 0) / height (youngest leaf at 0) may be NULL.  keys: n_keys metadata names separated by '\n'; values
 * [n_keys][cap_nodes] receives each key's number per node, NaN where a node does not carry it.  Returns the number of nodes;
 * a value > cap_nodes means the arrays are too small (nothing written); -1: not a binary tree in this format. */
int64_t gsm_parse_newick(const char* newick, int32_t cap_nodes, int32_t* parent, int32_t* left, int32_t* right, double* length,
                         double* height, int32_t n_keys, const char* keys, double* values);
/* GetEffectiveRates.java:55-62: eRate_i = branchRates_i + spikes_i / length_i of the leaves (the per-leaf columns of its
 * output table; rate / spike = the metadata the XML logs under those ids, length = Node.getLength). */
int  gsm_effective_leaf_rates(int32_t n_leaves, const double* rate, const double* spike, const double* length, double* out);
/* Java's Double.toString(d) (as used by every Loggable of the package: SPL:60, SpikeSize:74-78, STL:201-203). */
int64_t gsm_java_double_to_string(double d, char* out, int64_t cap);

/* Pinned host memory for the arrays above (optional; any host pointer is accepted). */
int  gsm_host_alloc(void** out, int64_t bytes);
int  gsm_host_free(void* p);

/* ---- device-resident path (inputs already in HBM; used by the bench and by multi-GPU hosts) ------
 * Row pitch on the device is `stride` elements (>= n_nodes; use gsm_device_stride(n_nodes) so that
 * every row starts 128-byte aligned).  Pointers are device pointers owned by the caller (e.g. torch
 * tensors); they must stay valid until the evaluation that uses them has completed. */
int64_t gsm_device_stride(int32_t n_nodes);
int  gsm_bind_device(gsm_handle* h, int64_t n_trees, int32_t n_nodes, int64_t stride,
                     const double* d_height, const int32_t* d_parent, const int32_t* d_node_count,
                     const double* d_rate, const double* d_spike, const int32_t* d_nstubs,
                     const double* d_params, const uint8_t* d_use_spike);
/* Device-resident stub lists of EXACT mode (see gsm_upload_stubs); call after gsm_bind_device.  The handle allocates the
 * derived per-branch count rows itself.  NULL stub_ptr removes the list. */
int  gsm_bind_device_stubs(gsm_handle* h, const int64_t* d_stub_ptr, const int32_t* d_stub_branch, const double* d_stub_time);
/* Launch on `cuda_stream` (a cudaStream_t; NULL = the handle's own stream) and return without
 * synchronising.  d_out_rates / d_out_wspikes have pitch `stride` and may be NULL. */
int  gsm_eval_device(gsm_handle* h, double* d_out_tree, double* d_out_rates, double* d_out_wspikes,
                     void* cuda_stream);
int  gsm_sync(gsm_handle* h);

/* ---- multi-GPU hosts: one process per GPU, result rows gathered over NVLink by the copy engines -------------
 * Trees are independent, so the only exchange of the path is the gather of the per-tree result rows (64 B per tree).
 * A rank allocates its receive buffer with gsm_device_alloc, exports it (gsm_ipc_export: a 64-byte cudaIpcMemHandle_t
 * that the host passes to its peers through any channel), and every peer opens it (gsm_ipc_open; peer access over
 * NVLink / NVSwitch is enabled on first use) and pushes its rows with gsm_copy_async on a side stream: a plain
 * device-to-device copy, executed by the DMA engines, so that no SM is taken from the kernels it overlaps with.
 * gsm_copy_async also moves host <-> device (cudaMemcpyDefault).  Ordering is the caller's: CUDA events between the
 * kernel stream and the side stream, and a process-level barrier before the receiver reads. */
int  gsm_device_alloc(void** out, int32_t device, int64_t bytes);
int  gsm_device_free(void* p, int32_t device);
int  gsm_ipc_export(const void* dptr, unsigned char* handle64);
int  gsm_ipc_open(int32_t device, const unsigned char* handle64, void** out);
int  gsm_ipc_close(void* p);
int  gsm_copy_async(void* dst, const void* src, int64_t bytes, void* cuda_stream);

/* Measurement helper (bench.py, roofline.fp64): throughput of the device's fp64 pipe in thread-level fused multiply-adds
 * per second, from a DFMA micro-benchmark (eight independent chains per thread, full occupancy, best of five launches). */
int  gsm_measure_fp64_peak(gsm_handle* h, double* fma_per_sec);

/* Number of kernels this handle has launched since creation (bench "gpu_launches"). */
int64_t gsm_launch_count(const gsm_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* GSM_B200_H */
