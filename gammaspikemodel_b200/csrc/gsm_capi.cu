// gsm_capi.cu — the C ABI declared in include/gsm_b200.h: handle management, host<->device staging,
// chunked streaming for host-resident batches, and the launches.  No compute happens on the CPU here
// and there is no fallback: without a usable CUDA device gsm_create fails.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/gsm_b200.h"
#include "gsm_kernels.cuh"

namespace {

thread_local std::string g_create_error;

struct Slot {  // one staging slot of the streaming path
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    double *height = nullptr, *rate = nullptr, *spike = nullptr, *params = nullptr, *out = nullptr;
    double *out_rates = nullptr, *out_wsp = nullptr;
    int32_t *parent = nullptr, *nstubs = nullptr, *node_count = nullptr;
    uint8_t* use_spike = nullptr;
    GsmScratch scratch;
    int64_t scratch_trees = 0;
};

}  // namespace

struct gsm_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t max_trees = 0;
    int32_t max_nodes = 0;
    int64_t max_stride = 0;
    uint32_t flags = 0;
    int32_t stub_mode = GSM_STUBS_COUNTS;
    // handle-owned resident batch
    double *d_height = nullptr, *d_rate = nullptr, *d_spike = nullptr, *d_params = nullptr, *d_out = nullptr;
    double *d_out_rates = nullptr, *d_out_wsp = nullptr;
    int32_t *d_parent = nullptr, *d_nstubs = nullptr, *d_node_count = nullptr;
    uint8_t* d_use = nullptr;
    // EXACT mode stub lists (gsm_upload_stubs / gsm_bind_device_stubs) and the per-branch counts derived from them
    int64_t* d_stub_ptr = nullptr;
    int32_t* d_stub_branch = nullptr;
    double* d_stub_time = nullptr;
    int64_t stub_ptr_cap = 0, stub_cap = 0;
    const int64_t* x_stub_ptr = nullptr;      // device-bound lists (caller-owned)
    const int32_t* x_stub_branch = nullptr;
    const double* x_stub_time = nullptr;
    int32_t* d_stub_counts = nullptr;         // [n_trees][stride] for device-bound batches
    int64_t stub_counts_cap = 0;
    bool have_stubs = false, stub_counts_valid = false;
    double* d_cdf = nullptr;  // scratch for gsm_stub_cdf
    int32_t cdf_cap = 0;
    bool have_trees = false, have_rate = false, have_spike = false, have_nstubs = false, have_params = false;
    bool have_use = false, have_node_count = false, external = false;
    GsmBatch cur{};
    // streaming slots (gsm_eval_host)
    Slot slots[2];
    int64_t slot_trees = 0;
    int64_t slot_stride = 0;
    bool slot_rates = false, slot_wsp = false;
    unsigned char* d_prop = nullptr;   // device copy of a proposal batch (gsm_eval_dirty), grow-only
    int64_t prop_cap = 0;
    GsmScratch scratch;          // lean-path scratch of the handle's own stream
    int64_t scratch_trees = 0;
    // gsm_eval_dirty: base of the resident batch + per-call lists
    GsmDirtyWork dirty;
    unsigned char *dw_children = nullptr, *dw_base_rows = nullptr, *dw_base_K = nullptr;
    int64_t dirty_trees = 0, dirty_cells = 0, dirty_props = 0, last_dirty_full = 0;
    bool base_valid = false;
    uint32_t base_flags = 0;
    int32_t base_mode = 0;
    GsmLaunchCtx ctx;            // device facts, probe results, gsm_set_option overrides (nothing process-global)
    int64_t host_chunk_mb = 256; // streaming chunk of gsm_eval_host / gsm_eval_dirty (gsm_set_option "host_chunk_mb")
    std::string err;
    int64_t launches = 0;
};

namespace {

int fail(gsm_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define GSM_CUDA(h, call)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail((h), GSM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

template <typename T>
cudaError_t dalloc(T** p, int64_t count) {
    if (count <= 0) count = 1;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), static_cast<size_t>(count) * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, static_cast<size_t>(count) * sizeof(T));
    // the memset runs on the legacy default stream; the handle's streams are non-blocking, so finish it here
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}

template <typename T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// lean-path scratch for up to `trees` trees (rows + redo list); grows, never shrinks
cudaError_t ensure_scratch(GsmScratch& sc, int64_t& cap, int64_t trees) {
    if (cap >= trees && sc.rows && sc.redo) return cudaSuccess;
    if (sc.rows) cudaFree(sc.rows);
    if (sc.redo) cudaFree(sc.redo);
    sc = GsmScratch{};
    cap = 0;
    unsigned char* rows = nullptr;
    cudaError_t e = dalloc(&rows, trees * static_cast<int64_t>(gsm_scratch_row_bytes()));
    if (e != cudaSuccess) return e;
    sc.rows = rows;
    e = dalloc(&sc.redo, trees + 4);
    if (e != cudaSuccess) return e;
    cap = trees;
    return cudaSuccess;
}
void free_scratch(GsmScratch& sc, int64_t& cap) {
    if (sc.rows) cudaFree(sc.rows);
    if (sc.redo) cudaFree(sc.redo);
    sc = GsmScratch{};
    cap = 0;
}

// dense host [T][n] -> pitched device [T][stride]
template <typename T>
cudaError_t h2d_rows(T* dst, int64_t stride, const T* src, int64_t n, int64_t rows, cudaStream_t st) {
    if (rows <= 0 || n <= 0) return cudaSuccess;
    return cudaMemcpy2DAsync(dst, static_cast<size_t>(stride) * sizeof(T), src, static_cast<size_t>(n) * sizeof(T),
                             static_cast<size_t>(n) * sizeof(T), static_cast<size_t>(rows), cudaMemcpyHostToDevice, st);
}
template <typename T>
cudaError_t d2h_rows(T* dst, int64_t n, const T* src, int64_t stride, int64_t rows, cudaStream_t st) {
    if (rows <= 0 || n <= 0) return cudaSuccess;
    return cudaMemcpy2DAsync(dst, static_cast<size_t>(n) * sizeof(T), src, static_cast<size_t>(stride) * sizeof(T),
                             static_cast<size_t>(n) * sizeof(T), static_cast<size_t>(rows), cudaMemcpyDeviceToHost, st);
}

void free_slots(gsm_handle* h) {
    for (Slot& s : h->slots) {
        dfree(s.height); dfree(s.rate); dfree(s.spike); dfree(s.params); dfree(s.out);
        dfree(s.out_rates); dfree(s.out_wsp); dfree(s.parent); dfree(s.nstubs); dfree(s.node_count);
        dfree(s.use_spike);
        free_scratch(s.scratch, s.scratch_trees);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = Slot{};
    }
    h->slot_trees = 0;
    h->slot_stride = 0;
}

int check_mode(gsm_handle* h, bool have_rate, bool have_nstubs) {
    if ((h->flags & GSM_F_HAS_RATES) && !have_rate)
        return fail(h, GSM_E_STATE, "GSM_F_HAS_RATES is set but no rate array was provided");
    const bool needs_counts = h->stub_mode != GSM_STUBS_NOSTUB &&
                              (h->flags & (GSM_F_STUBS_TO_CLOCK | GSM_F_STUBS_TO_SPIKE_PRIOR | GSM_F_STUBS_TO_TREE_PRIOR));
    if (needs_counts && !have_nstubs)
        return fail(h, GSM_E_STATE, "stub counts are wired (stub_mode=%d) but no nstubs array was provided", h->stub_mode);
    return GSM_OK;
}

}  // namespace

extern "C" {

int gsm_abi_version(void) { return GSM_ABI_VERSION; }

int gsm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t gsm_device_stride(int32_t n_nodes) { return n_nodes <= 0 ? 32 : round_up(n_nodes, 32); }

const char* gsm_last_error(const gsm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int gsm_create(gsm_handle** out, int32_t device_id, int64_t max_trees, int32_t max_nodes) {
    if (!out) return fail(nullptr, GSM_E_ARG, "gsm_create: out is NULL");
    *out = nullptr;
    if (max_trees < 0 || max_nodes < 0) return fail(nullptr, GSM_E_ARG, "gsm_create: negative capacity");
    if (max_nodes > GSM_MAX_NODES)
        return fail(nullptr, GSM_E_UNSUPPORTED, "gsm_create: max_nodes %d exceeds GSM_MAX_NODES %d", max_nodes, GSM_MAX_NODES);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, GSM_E_CUDA, "gsm_create: no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device_id < 0 || device_id >= ndev) return fail(nullptr, GSM_E_ARG, "gsm_create: device %d of %d", device_id, ndev);
    gsm_handle* h = new (std::nothrow) gsm_handle();
    if (!h) return fail(nullptr, GSM_E_NOMEM, "gsm_create: out of host memory");
    h->device = device_id;
    h->max_trees = max_trees;
    h->max_nodes = max_nodes;
    h->max_stride = gsm_device_stride(max_nodes);
    auto bail = [&](cudaError_t ce, const char* what) {
        int rc = fail(nullptr, ce == cudaErrorMemoryAllocation ? GSM_E_NOMEM : GSM_E_CUDA, "gsm_create: %s: %s", what,
                      cudaGetErrorString(ce));
        gsm_destroy(h);
        return rc;
    };
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = gsm_init_tables()) != cudaSuccess) return bail(e, "table upload");
    if ((e = gsm_launch_ctx_init(h->ctx, device_id)) != cudaSuccess) return bail(e, "device attributes");
    if (max_trees > 0 && max_nodes > 0) {
        const int64_t cells = max_trees * h->max_stride;
        if ((e = dalloc(&h->d_height, cells)) != cudaSuccess) return bail(e, "alloc height");
        if ((e = dalloc(&h->d_rate, cells)) != cudaSuccess) return bail(e, "alloc rate");
        if ((e = dalloc(&h->d_spike, cells)) != cudaSuccess) return bail(e, "alloc spike");
        if ((e = dalloc(&h->d_parent, cells)) != cudaSuccess) return bail(e, "alloc parent");
        if ((e = dalloc(&h->d_nstubs, cells)) != cudaSuccess) return bail(e, "alloc nstubs");
        if ((e = dalloc(&h->d_node_count, max_trees)) != cudaSuccess) return bail(e, "alloc node_count");
        if ((e = dalloc(&h->d_params, max_trees * GSM_NPARAM)) != cudaSuccess) return bail(e, "alloc params");
        if ((e = dalloc(&h->d_use, max_trees)) != cudaSuccess) return bail(e, "alloc use_spike");
        if ((e = dalloc(&h->d_out, max_trees * GSM_NOUT)) != cudaSuccess) return bail(e, "alloc out");
    }
    *out = h;
    return GSM_OK;
}

int gsm_destroy(gsm_handle* h) {
    if (!h) return GSM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_slots(h);
    dfree(h->d_height); dfree(h->d_rate); dfree(h->d_spike); dfree(h->d_parent); dfree(h->d_nstubs);
    dfree(h->d_node_count); dfree(h->d_params); dfree(h->d_use); dfree(h->d_out); dfree(h->d_out_rates);
    dfree(h->d_out_wsp); dfree(h->d_cdf); dfree(h->d_prop);
    dfree(h->d_stub_ptr); dfree(h->d_stub_branch); dfree(h->d_stub_time); dfree(h->d_stub_counts);
    dfree(h->dw_children); dfree(h->dw_base_rows); dfree(h->dw_base_K); dfree(h->dirty.tree_flag); dfree(h->dirty.out_base);
    dfree(h->dirty.full_list); dfree(h->dirty.counters); dfree(h->dirty.out_all);
    free_scratch(h->scratch, h->scratch_trees);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return GSM_OK;
}

int gsm_set_mode(gsm_handle* h, int32_t stub_mode, uint32_t flags) {
    if (!h) return GSM_E_ARG;
    if (stub_mode < GSM_STUBS_NOSTUB || stub_mode > GSM_STUBS_EXACT) return fail(h, GSM_E_ARG, "gsm_set_mode: stub_mode %d", stub_mode);
    if ((flags & GSM_F_COND_SAMPLING) && (flags & GSM_F_COND_RHO_SAMPLING))   // STP:92-94
        return fail(h, GSM_E_ARG, "gsm_set_mode: conditionOnSampling and conditionOnRhoSampling are exclusive");
    if ((flags & GSM_F_ORIGIN) && (flags & GSM_F_COND_ROOT))                  // STP:89-91
        return fail(h, GSM_E_ARG, "gsm_set_mode: origin and conditionOnRoot are exclusive");
    if ((flags & GSM_F_COND_ROOT) && !(flags & (GSM_F_COND_SAMPLING | GSM_F_COND_RHO_SAMPLING)))   // STP:95-98
        return fail(h, GSM_E_ARG, "gsm_set_mode: conditionOnRoot needs conditionOnSampling or conditionOnRhoSampling");
    h->stub_mode = stub_mode;
    h->flags = flags;
    return GSM_OK;
}

int gsm_set_option(gsm_handle* h, const char* name, int64_t value) {
    if (!h || !name) return GSM_E_ARG;
    if (!strcmp(name, "host_chunk_mb")) {
        if (value < 1) return fail(h, GSM_E_ARG, "gsm_set_option: host_chunk_mb must be >= 1");
        h->host_chunk_mb = value;
    } else if (!strcmp(name, "force_generic")) {
        h->ctx.force_generic = value != 0;
    } else if (!strcmp(name, "block")) {
        h->ctx.block = static_cast<int>(value);
    } else if (!strcmp(name, "minb")) {
        h->ctx.minb = static_cast<int>(value);
    } else if (!strcmp(name, "producer")) {
        h->ctx.producer = value != 0;
    } else if (!strcmp(name, "warp_kernel")) {
        h->ctx.warp_kernel = value != 0;
    } else if (!strcmp(name, "smem_pad")) {
        h->ctx.smem_pad = static_cast<size_t>(value);
    } else if (!strcmp(name, "persistent")) {
        h->ctx.persistent = static_cast<int>(value);
    } else {
        return fail(h, GSM_E_ARG, "gsm_set_option: unknown option '%s'", name);
    }
    return GSM_OK;
}

int gsm_upload_trees(gsm_handle* h, int64_t n_trees, int32_t n_nodes, const double* height, const int32_t* parent,
                     const int32_t* node_count) {
    if (!h) return GSM_E_ARG;
    if (n_trees < 0 || n_nodes < 0) return fail(h, GSM_E_ARG, "gsm_upload_trees: negative size");
    if (n_trees > h->max_trees || n_nodes > h->max_nodes)
        return fail(h, GSM_E_ARG, "gsm_upload_trees: %lld x %d exceeds the handle capacity %lld x %d", (long long)n_trees,
                    n_nodes, (long long)h->max_trees, h->max_nodes);
    if (n_trees > 0 && n_nodes > 0 && (!height || !parent)) return fail(h, GSM_E_ARG, "gsm_upload_trees: NULL array");
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t stride = gsm_device_stride(n_nodes);
    GSM_CUDA(h, h2d_rows(h->d_height, stride, height, n_nodes, n_trees, h->stream));
    GSM_CUDA(h, h2d_rows(h->d_parent, stride, parent, n_nodes, n_trees, h->stream));
    if (node_count && n_trees > 0)
        GSM_CUDA(h, cudaMemcpyAsync(h->d_node_count, node_count, n_trees * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->external = false;
    h->have_trees = true;
    h->have_node_count = node_count != nullptr;
    // branch arrays and per-tree scalars belong to a tree upload: a new batch must bring its own
    h->have_rate = h->have_spike = h->have_nstubs = h->have_params = h->have_use = false;
    h->have_stubs = h->stub_counts_valid = false;
    h->base_valid = false;
    h->cur = GsmBatch{};
    h->cur.n_trees = n_trees;
    h->cur.n_nodes = n_nodes;
    h->cur.stride = stride;
    return GSM_OK;
}

int gsm_upload_branch_params(gsm_handle* h, const double* rate, const double* spike, const int32_t* nstubs) {
    if (!h) return GSM_E_ARG;
    if (!h->have_trees || h->external) return fail(h, GSM_E_STATE, "gsm_upload_branch_params: call gsm_upload_trees first");
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t T = h->cur.n_trees, n = h->cur.n_nodes, S = h->cur.stride;
    if (rate) GSM_CUDA(h, h2d_rows(h->d_rate, S, rate, n, T, h->stream));
    if (spike) GSM_CUDA(h, h2d_rows(h->d_spike, S, spike, n, T, h->stream));
    if (nstubs) GSM_CUDA(h, h2d_rows(h->d_nstubs, S, nstubs, n, T, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_rate |= rate != nullptr;
    h->have_spike |= spike != nullptr;
    h->have_nstubs |= nstubs != nullptr;
    if (nstubs) h->stub_counts_valid = false;   // EXACT mode: the derived counts live in the same rows
    h->base_valid = false;
    return GSM_OK;
}

int gsm_upload_tree_params(gsm_handle* h, const double* params, const uint8_t* use_spike) {
    if (!h) return GSM_E_ARG;
    if (!h->have_trees || h->external) return fail(h, GSM_E_STATE, "gsm_upload_tree_params: call gsm_upload_trees first");
    if (!params) return fail(h, GSM_E_ARG, "gsm_upload_tree_params: params is NULL");
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t T = h->cur.n_trees;
    if (T > 0) {
        GSM_CUDA(h, cudaMemcpyAsync(h->d_params, params, T * GSM_NPARAM * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (use_spike) GSM_CUDA(h, cudaMemcpyAsync(h->d_use, use_spike, T, cudaMemcpyHostToDevice, h->stream));
    }
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_params = true;
    h->have_use = use_spike != nullptr;
    h->base_valid = false;
    return GSM_OK;
}

int gsm_upload_stubs(gsm_handle* h, const int64_t* stub_ptr, const int32_t* stub_branch, const double* stub_time) {
    if (!h) return GSM_E_ARG;
    if (!h->have_trees || h->external) return fail(h, GSM_E_STATE, "gsm_upload_stubs: call gsm_upload_trees first");
    if (!stub_ptr) return fail(h, GSM_E_ARG, "gsm_upload_stubs: stub_ptr is NULL");
    const int64_t T = h->cur.n_trees;
    if (stub_ptr[0] != 0) return fail(h, GSM_E_ARG, "gsm_upload_stubs: stub_ptr must start at 0");
    for (int64_t t = 0; t < T; t++)
        if (stub_ptr[t + 1] < stub_ptr[t]) return fail(h, GSM_E_ARG, "gsm_upload_stubs: stub_ptr decreases at tree %lld", (long long)t);
    const int64_t S = stub_ptr[T];
    if (S > 0 && (!stub_branch || !stub_time)) return fail(h, GSM_E_ARG, "gsm_upload_stubs: NULL stub arrays");
    GSM_CUDA(h, cudaSetDevice(h->device));
    if (h->stub_ptr_cap < T + 1) {
        dfree(h->d_stub_ptr);
        h->stub_ptr_cap = 0;
        GSM_CUDA(h, dalloc(&h->d_stub_ptr, T + 1));
        h->stub_ptr_cap = T + 1;
    }
    if (h->stub_cap < S) {
        dfree(h->d_stub_branch); dfree(h->d_stub_time);
        h->stub_cap = 0;
        GSM_CUDA(h, dalloc(&h->d_stub_branch, S));
        GSM_CUDA(h, dalloc(&h->d_stub_time, S));
        h->stub_cap = S;
    }
    GSM_CUDA(h, cudaMemcpyAsync(h->d_stub_ptr, stub_ptr, (T + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    if (S > 0) {
        GSM_CUDA(h, cudaMemcpyAsync(h->d_stub_branch, stub_branch, S * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        GSM_CUDA(h, cudaMemcpyAsync(h->d_stub_time, stub_time, S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_stubs = true;
    h->stub_counts_valid = false;
    h->base_valid = false;
    h->have_nstubs = true;   // the counts are derived from the list (resident_batch)
    return GSM_OK;
}

int gsm_bind_device_stubs(gsm_handle* h, const int64_t* d_stub_ptr, const int32_t* d_stub_branch, const double* d_stub_time) {
    if (!h) return GSM_E_ARG;
    if (!h->have_trees || !h->external) return fail(h, GSM_E_STATE, "gsm_bind_device_stubs: call gsm_bind_device first");
    h->x_stub_ptr = d_stub_ptr;
    h->x_stub_branch = d_stub_branch;
    h->x_stub_time = d_stub_time;
    h->have_stubs = d_stub_ptr != nullptr;
    h->stub_counts_valid = false;
    return GSM_OK;
}

static bool rj_terms_wired(const gsm_handle* h) {   // STP:219 with Stubs.getReversibleJump (STP:573-600, 639-663)
    return h->stub_mode == GSM_STUBS_EXACT && (h->flags & GSM_F_STUBS_TO_TREE_PRIOR) && !(h->flags & GSM_F_IGNORE_STUB_PRIOR);
}

static int resident_batch(gsm_handle* h, GsmBatch* b) {
    if (!h->have_trees) return fail(h, GSM_E_STATE, "no tree batch: call gsm_upload_trees or gsm_bind_device first");
    if (h->external) {
        *b = h->cur;
    } else {
        if (!h->have_params) return fail(h, GSM_E_STATE, "no per-tree parameters: call gsm_upload_tree_params first");
        if (!h->have_spike) return fail(h, GSM_E_STATE, "no spikes: call gsm_upload_branch_params first");
        *b = h->cur;
        b->height = h->d_height;
        b->parent = h->d_parent;
        b->node_count = h->have_node_count ? h->d_node_count : nullptr;
        b->rate = h->have_rate ? h->d_rate : nullptr;
        b->spike = h->d_spike;
        b->nstubs = h->have_nstubs ? h->d_nstubs : nullptr;
        b->params = h->d_params;
        b->use_spike = h->have_use ? h->d_use : nullptr;
    }
    b->flags = h->flags;
    b->stub_mode = h->stub_mode;
    b->stub_ptr = nullptr; b->stub_branch = nullptr; b->stub_time = nullptr;
    if (h->stub_mode == GSM_STUBS_EXACT) {
        if (h->have_stubs) {
            // the list defines the per-branch counts (Stubs:357-364): derived on the device, once per upload
            b->stub_ptr = h->external ? h->x_stub_ptr : h->d_stub_ptr;
            b->stub_branch = h->external ? h->x_stub_branch : h->d_stub_branch;
            b->stub_time = h->external ? h->x_stub_time : h->d_stub_time;
            int32_t* counts = h->d_nstubs;
            if (h->external) {
                const int64_t cells = b->n_trees * b->stride;
                if (h->stub_counts_cap < cells) {
                    dfree(h->d_stub_counts);
                    h->stub_counts_cap = 0;
                    GSM_CUDA(h, cudaSetDevice(h->device));
                    GSM_CUDA(h, dalloc(&h->d_stub_counts, cells));
                    h->stub_counts_cap = cells;
                    h->stub_counts_valid = false;
                }
                counts = h->d_stub_counts;
            }
            if (!counts) return fail(h, GSM_E_STATE, "EXACT mode: no room for the per-branch stub counts (handle created without capacity)");
            if (!h->stub_counts_valid && b->n_trees > 0) {
                GSM_CUDA(h, cudaSetDevice(h->device));
                int nl = gsm_launch_stub_counts(*b, counts, h->stream);
                if (nl < 0) return fail(h, GSM_E_CUDA, "stub count launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
                h->launches += nl;
                GSM_CUDA(h, cudaStreamSynchronize(h->stream));   // evaluations may run on another stream
                h->stub_counts_valid = true;
            }
            b->nstubs = counts;
        } else if (rj_terms_wired(h)) {
            return fail(h, GSM_E_STATE, "stub_mode is GSM_STUBS_EXACT and the stub prior is evaluated: call gsm_upload_stubs "
                                        "(or gsm_bind_device_stubs) first");
        }
    }
    return check_mode(h, b->rate != nullptr, b->nstubs != nullptr);
}

int gsm_eval(gsm_handle* h, double* out_tree, double* out_rates, double* out_wspikes) {
    if (!h) return GSM_E_ARG;
    if (h->external) return fail(h, GSM_E_STATE, "gsm_eval: batch is device-bound; use gsm_eval_device");
    GsmBatch b;
    int rc = resident_batch(h, &b);
    if (rc != GSM_OK) return rc;
    if (b.n_trees == 0) return GSM_OK;
    if (!out_tree) return fail(h, GSM_E_ARG, "gsm_eval: out_tree is NULL");
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t cells = h->max_trees * h->max_stride;
    if (out_rates && !h->d_out_rates) GSM_CUDA(h, dalloc(&h->d_out_rates, cells));
    if (out_wspikes && !h->d_out_wsp) GSM_CUDA(h, dalloc(&h->d_out_wsp, cells));
    GSM_CUDA(h, ensure_scratch(h->scratch, h->scratch_trees, b.n_trees));
    int nl = gsm_launch_eval(h->ctx, b, h->scratch, h->d_out, out_rates ? h->d_out_rates : nullptr, out_wspikes ? h->d_out_wsp : nullptr, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    GSM_CUDA(h, cudaMemcpyAsync(out_tree, h->d_out, b.n_trees * GSM_NOUT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (out_rates) GSM_CUDA(h, d2h_rows(out_rates, (int64_t)b.n_nodes, h->d_out_rates, b.stride, b.n_trees, h->stream));
    if (out_wspikes) GSM_CUDA(h, d2h_rows(out_wspikes, (int64_t)b.n_nodes, h->d_out_wsp, b.stride, b.n_trees, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    return GSM_OK;
}

static int ensure_slots(gsm_handle* h, int64_t chunk, int64_t stride, bool rates, bool wsp) {
    if (h->slot_trees >= chunk && h->slot_stride == stride && (!rates || h->slot_rates) && (!wsp || h->slot_wsp)) return GSM_OK;
    free_slots(h);
    const int64_t cells = chunk * stride;
    for (Slot& s : h->slots) {
        GSM_CUDA(h, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        GSM_CUDA(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        GSM_CUDA(h, dalloc(&s.height, cells));
        GSM_CUDA(h, dalloc(&s.rate, cells));
        GSM_CUDA(h, dalloc(&s.spike, cells));
        GSM_CUDA(h, dalloc(&s.parent, cells));
        GSM_CUDA(h, dalloc(&s.nstubs, cells));
        GSM_CUDA(h, dalloc(&s.node_count, chunk));
        GSM_CUDA(h, dalloc(&s.params, chunk * GSM_NPARAM));
        GSM_CUDA(h, dalloc(&s.use_spike, chunk));
        GSM_CUDA(h, dalloc(&s.out, chunk * GSM_NOUT));
        if (rates) GSM_CUDA(h, dalloc(&s.out_rates, cells));
        if (wsp) GSM_CUDA(h, dalloc(&s.out_wsp, cells));
        GSM_CUDA(h, ensure_scratch(s.scratch, s.scratch_trees, chunk));
    }
    h->slot_trees = chunk;
    h->slot_stride = stride;
    h->slot_rates = rates;
    h->slot_wsp = wsp;
    return GSM_OK;
}

int gsm_eval_host(gsm_handle* h, int64_t n_trees, int32_t n_nodes, const double* height, const int32_t* parent,
                  const int32_t* node_count, const double* rate, const double* spike, const int32_t* nstubs,
                  const double* params, const uint8_t* use_spike, double* out_tree, double* out_rates, double* out_wspikes) {
    if (!h) return GSM_E_ARG;
    if (n_trees < 0 || n_nodes < 0) return fail(h, GSM_E_ARG, "gsm_eval_host: negative size");
    if (n_nodes > GSM_MAX_NODES) return fail(h, GSM_E_UNSUPPORTED, "gsm_eval_host: n_nodes %d exceeds GSM_MAX_NODES", n_nodes);
    if (n_trees == 0 || n_nodes == 0) return GSM_OK;
    if (!height || !parent || !spike || !params || !out_tree) return fail(h, GSM_E_ARG, "gsm_eval_host: NULL required array");
    if (rj_terms_wired(h))
        return fail(h, GSM_E_UNSUPPORTED, "gsm_eval_host: EXACT mode with the stub prior evaluated needs the stub lists; use "
                                          "gsm_upload_trees / gsm_upload_stubs / gsm_eval");
    int rc = check_mode(h, rate != nullptr, nstubs != nullptr);
    if (rc != GSM_OK) return rc;
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t stride = gsm_device_stride(n_nodes);
    // ~256 MB of inputs per slot: long enough transfers for full PCIe rate, short enough to overlap
    const int64_t chunk_mb = h->host_chunk_mb;
    int64_t chunk = std::max<int64_t>(1, (chunk_mb << 20) / (stride * 32));
    chunk = std::min(chunk, n_trees);
    rc = ensure_slots(h, chunk, stride, out_rates != nullptr, out_wspikes != nullptr);
    if (rc != GSM_OK) return rc;
    chunk = std::min(h->slot_trees, n_trees);
    int64_t k = 0;
    for (int64_t t0 = 0; t0 < n_trees; t0 += chunk, ++k) {
        const int64_t T = std::min(chunk, n_trees - t0);
        Slot& s = h->slots[k & 1];
        if (s.busy) GSM_CUDA(h, cudaEventSynchronize(s.done));
        const int64_t off = t0 * n_nodes;
        GSM_CUDA(h, h2d_rows(s.height, stride, height + off, (int64_t)n_nodes, T, s.stream));
        GSM_CUDA(h, h2d_rows(s.parent, stride, parent + off, (int64_t)n_nodes, T, s.stream));
        GSM_CUDA(h, h2d_rows(s.spike, stride, spike + off, (int64_t)n_nodes, T, s.stream));
        if (rate) GSM_CUDA(h, h2d_rows(s.rate, stride, rate + off, (int64_t)n_nodes, T, s.stream));
        if (nstubs) GSM_CUDA(h, h2d_rows(s.nstubs, stride, nstubs + off, (int64_t)n_nodes, T, s.stream));
        if (node_count) GSM_CUDA(h, cudaMemcpyAsync(s.node_count, node_count + t0, T * sizeof(int32_t), cudaMemcpyHostToDevice, s.stream));
        GSM_CUDA(h, cudaMemcpyAsync(s.params, params + t0 * GSM_NPARAM, T * GSM_NPARAM * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        if (use_spike) GSM_CUDA(h, cudaMemcpyAsync(s.use_spike, use_spike + t0, T, cudaMemcpyHostToDevice, s.stream));
        GsmBatch b{};
        b.n_trees = T; b.n_nodes = n_nodes; b.stride = stride;
        b.height = s.height; b.parent = s.parent; b.node_count = node_count ? s.node_count : nullptr;
        b.rate = rate ? s.rate : nullptr; b.spike = s.spike; b.nstubs = nstubs ? s.nstubs : nullptr;
        b.params = s.params; b.use_spike = use_spike ? s.use_spike : nullptr;
        b.flags = h->flags; b.stub_mode = h->stub_mode;
        int nl = gsm_launch_eval(h->ctx, b, s.scratch, s.out, out_rates ? s.out_rates : nullptr, out_wspikes ? s.out_wsp : nullptr, s.stream);
        if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_host launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
        h->launches += nl;
        GSM_CUDA(h, cudaMemcpyAsync(out_tree + t0 * GSM_NOUT, s.out, T * GSM_NOUT * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        if (out_rates) GSM_CUDA(h, d2h_rows(out_rates + off, (int64_t)n_nodes, s.out_rates, stride, T, s.stream));
        if (out_wspikes) GSM_CUDA(h, d2h_rows(out_wspikes + off, (int64_t)n_nodes, s.out_wsp, stride, T, s.stream));
        GSM_CUDA(h, cudaEventRecord(s.done, s.stream));
        s.busy = true;
    }
    for (Slot& s : h->slots) {
        if (s.busy) GSM_CUDA(h, cudaEventSynchronize(s.done));
        s.busy = false;
    }
    return GSM_OK;
}

// scratch of gsm_eval_dirty: base of the resident batch (valid until the batch changes) + per-call lists
static int ensure_dirty_work(gsm_handle* h, const GsmBatch& src, int64_t n_props) {
    const int64_t T = src.n_trees, cells = T * src.stride;
    if (h->dirty_trees < T || h->dirty_cells < cells) {
        dfree(h->dw_children); dfree(h->dw_base_rows); dfree(h->dw_base_K); dfree(h->dirty.tree_flag); dfree(h->dirty.out_base);
        h->dirty_trees = h->dirty_cells = 0;
        h->base_valid = false;
        GSM_CUDA(h, dalloc(&h->dw_children, cells * 8));
        GSM_CUDA(h, dalloc(&h->dw_base_rows, T * static_cast<int64_t>(gsm_base_row_bytes())));
        GSM_CUDA(h, dalloc(&h->dw_base_K, T * static_cast<int64_t>(gsm_base_k_bytes())));
        GSM_CUDA(h, dalloc(&h->dirty.tree_flag, T));
        GSM_CUDA(h, dalloc(&h->dirty.out_base, T * GSM_NOUT));
        h->dirty.children = h->dw_children; h->dirty.base_rows = h->dw_base_rows; h->dirty.base_K = h->dw_base_K;
        h->dirty_trees = T;
        h->dirty_cells = cells;
    }
    if (h->dirty_props < n_props) {
        dfree(h->dirty.full_list); dfree(h->dirty.out_all);
        h->dirty_props = 0;
        GSM_CUDA(h, dalloc(&h->dirty.full_list, n_props));
        GSM_CUDA(h, dalloc(&h->dirty.out_all, n_props * GSM_NOUT));
        h->dirty_props = n_props;
    }
    if (!h->dirty.counters) GSM_CUDA(h, dalloc(&h->dirty.counters, 2));
    return GSM_OK;
}

int gsm_eval_dirty(gsm_handle* h, int64_t n_props, const int32_t* tree_of_prop, const int64_t* dirty_ptr,
                   const int32_t* dirty_nodes, const double* new_height, const double* new_rate, const double* new_spike,
                   const int32_t* new_nstubs, const double* height_scale, const double* new_params, const uint8_t* new_use_spike,
                   double* out_tree) {
    if (!h) return GSM_E_ARG;
    if (n_props < 0) return fail(h, GSM_E_ARG, "gsm_eval_dirty: negative n_props");
    if (n_props == 0) return GSM_OK;
    if (!tree_of_prop || !dirty_ptr || !out_tree) return fail(h, GSM_E_ARG, "gsm_eval_dirty: NULL required array");
    if (rj_terms_wired(h))
        return fail(h, GSM_E_UNSUPPORTED, "gsm_eval_dirty: reversible-jump moves change the stub list itself; upload the "
                                          "proposed state (gsm_upload_stubs) and call gsm_eval");
    GsmBatch src;
    int rc = resident_batch(h, &src);
    if (rc != GSM_OK) return rc;
    const int64_t D = dirty_ptr[n_props];
    if (dirty_ptr[0] != 0 || D < 0) return fail(h, GSM_E_ARG, "gsm_eval_dirty: dirty_ptr must start at 0 and be non-decreasing");
    if (D > 0 && !dirty_nodes) return fail(h, GSM_E_ARG, "gsm_eval_dirty: dirty_nodes is NULL");
    GSM_CUDA(h, cudaSetDevice(h->device));
    // ---- proposal arrays -> one device buffer (16-byte aligned sections) ----
    auto al = [](int64_t v) { return (v + 15) & ~static_cast<int64_t>(15); };
    const int64_t o_tree = 0, o_ptr = al(o_tree + n_props * 4), o_nodes = al(o_ptr + (n_props + 1) * 8), o_h = al(o_nodes + D * 4),
                  o_r = al(o_h + (new_height ? D * 8 : 0)), o_s = al(o_r + (new_rate ? D * 8 : 0)), o_k = al(o_s + (new_spike ? D * 8 : 0)),
                  o_sc = al(o_k + (new_nstubs ? D * 4 : 0)), o_par = al(o_sc + (height_scale ? n_props * 8 : 0)),
                  o_use = al(o_par + (new_params ? n_props * GSM_NPARAM * 8 : 0)),
                  total = al(o_use + (new_use_spike ? n_props : 0)) + 16;
    if (h->prop_cap < total) {
        dfree(h->d_prop);
        h->prop_cap = 0;
        GSM_CUDA(h, dalloc(&h->d_prop, total));
        h->prop_cap = total;
    }
    unsigned char* base = h->d_prop;
    auto up = [&](int64_t off, const void* src_, int64_t bytes) -> cudaError_t {
        if (!src_ || bytes <= 0) return cudaSuccess;
        return cudaMemcpyAsync(base + off, src_, static_cast<size_t>(bytes), cudaMemcpyHostToDevice, h->stream);
    };
    GSM_CUDA(h, up(o_tree, tree_of_prop, n_props * 4));
    GSM_CUDA(h, up(o_ptr, dirty_ptr, (n_props + 1) * 8));
    GSM_CUDA(h, up(o_nodes, dirty_nodes, D * 4));
    GSM_CUDA(h, up(o_h, new_height, D * 8));
    GSM_CUDA(h, up(o_r, new_rate, D * 8));
    GSM_CUDA(h, up(o_s, new_spike, D * 8));
    GSM_CUDA(h, up(o_k, new_nstubs, D * 4));
    GSM_CUDA(h, up(o_sc, height_scale, n_props * 8));
    GSM_CUDA(h, up(o_par, new_params, n_props * GSM_NPARAM * 8));
    GSM_CUDA(h, up(o_use, new_use_spike, n_props));
    GsmProposals pr{};
    pr.tree_of_prop = reinterpret_cast<const int32_t*>(base + o_tree);
    pr.dirty_ptr = reinterpret_cast<const int64_t*>(base + o_ptr);
    pr.dirty_nodes = reinterpret_cast<const int32_t*>(base + o_nodes);
    pr.new_height = new_height ? reinterpret_cast<const double*>(base + o_h) : nullptr;
    pr.new_rate = new_rate ? reinterpret_cast<const double*>(base + o_r) : nullptr;
    pr.new_spike = new_spike ? reinterpret_cast<const double*>(base + o_s) : nullptr;
    pr.new_nstubs = new_nstubs ? reinterpret_cast<const int32_t*>(base + o_k) : nullptr;
    pr.height_scale = height_scale ? reinterpret_cast<const double*>(base + o_sc) : nullptr;
    pr.new_params = new_params ? reinterpret_cast<const double*>(base + o_par) : nullptr;
    pr.new_use_spike = new_use_spike ? reinterpret_cast<const uint8_t*>(base + o_use) : nullptr;
    // ---- base of the resident batch (children table, accumulators): once per resident batch and wiring ----
    rc = ensure_dirty_work(h, src, n_props);
    if (rc != GSM_OK) return rc;
    const bool same_batch = h->base_valid && h->base_flags == h->flags && h->base_mode == h->stub_mode && !h->external;
    if (!same_batch) {
        int nl = gsm_launch_dirty_base(h->ctx, src, h->dirty, h->stream);
        if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_dirty base launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
        h->launches += nl;
        h->base_valid = true;   // device-bound batches: the caller may rewrite the arrays between calls, so never reused
        h->base_flags = h->flags;
        h->base_mode = h->stub_mode;
    }
    // ---- validation, classification and the incremental evaluations: one launch ----
    int nl = gsm_launch_dirty_delta(src, pr, n_props, h->dirty, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_dirty delta launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    int32_t counters[2] = {0, 0};
    GSM_CUDA(h, cudaMemcpyAsync(counters, h->dirty.counters, sizeof counters, cudaMemcpyDeviceToHost, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));   // also: the slots' streams read the proposal buffer
    if (counters[1])
        return fail(h, GSM_E_ARG, "gsm_eval_dirty: a proposal refers to a tree outside the batch, a node outside [0, %d) or "
                                  "a decreasing dirty_ptr", src.n_nodes);
    const int64_t n_full = counters[0];
    h->last_dirty_full = n_full;
    // ---- the other proposals: proposed states materialised chunk by chunk through the two staging slots ----
    if (n_full > 0) {
        const int64_t chunk_mb = h->host_chunk_mb;
        int64_t chunk = std::max<int64_t>(1, (chunk_mb << 20) / (src.stride * 32));
        chunk = std::min(chunk, n_full);
        rc = ensure_slots(h, chunk, src.stride, false, false);
        if (rc != GSM_OK) return rc;
        chunk = std::min(h->slot_trees, n_full);
        int64_t k = 0;
        for (int64_t q0 = 0; q0 < n_full; q0 += chunk, ++k) {
            const int64_t P = std::min(chunk, n_full - q0);
            Slot& s = h->slots[k & 1];
            if (s.busy) GSM_CUDA(h, cudaEventSynchronize(s.done));
            GsmBatchMut dst{s.height, s.parent, s.node_count, s.rate, s.spike, s.nstubs, s.params, s.use_spike};
            nl = gsm_launch_patch(src, pr, h->dirty, q0, P, dst, s.stream);
            if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_dirty patch launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
            h->launches += nl;
            GsmBatch b = src;
            b.n_trees = P;
            b.height = s.height; b.parent = s.parent; b.node_count = s.node_count;
            b.rate = src.rate ? s.rate : nullptr; b.spike = s.spike; b.nstubs = src.nstubs ? s.nstubs : nullptr;
            b.params = s.params; b.use_spike = s.use_spike;
            b.stub_ptr = nullptr; b.stub_branch = nullptr; b.stub_time = nullptr;
            nl = gsm_launch_eval(h->ctx, b, s.scratch, s.out, nullptr, nullptr, s.stream);
            if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_dirty launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
            h->launches += nl;
            nl = gsm_launch_scatter_rows(s.out, h->dirty.full_list, q0, P, h->dirty.out_all, s.stream);
            if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_dirty scatter launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
            h->launches += nl;
            GSM_CUDA(h, cudaEventRecord(s.done, s.stream));
            s.busy = true;
        }
        for (Slot& s : h->slots) {
            if (s.busy) GSM_CUDA(h, cudaEventSynchronize(s.done));
            s.busy = false;
        }
    }
    GSM_CUDA(h, cudaMemcpyAsync(out_tree, h->dirty.out_all, n_props * GSM_NOUT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    return GSM_OK;
}

int64_t gsm_last_dirty_full_count(const gsm_handle* h) { return h ? h->last_dirty_full : 0; }

int gsm_stub_cdf(gsm_handle* h, int64_t tree, int32_t node, double* cdf, int32_t cap, int32_t* len) {
    if (!h) return GSM_E_ARG;
    if (!cdf || !len || cap <= 0) return fail(h, GSM_E_ARG, "gsm_stub_cdf: bad output arguments");
    GsmBatch b;
    int rc = resident_batch(h, &b);
    if (rc != GSM_OK) return rc;
    if (tree < 0 || tree >= b.n_trees || node < 0 || node >= b.n_nodes) return fail(h, GSM_E_ARG, "gsm_stub_cdf: (tree,node) out of range");
    GSM_CUDA(h, cudaSetDevice(h->device));
    if (h->cdf_cap < cap + 1) {
        dfree(h->d_cdf);
        GSM_CUDA(h, dalloc(&h->d_cdf, (int64_t)cap + 1));
        h->cdf_cap = cap + 1;
    }
    int32_t* d_len = reinterpret_cast<int32_t*>(h->d_cdf + cap);
    int nl = gsm_launch_stub_cdf(b, tree, node, h->d_cdf, cap, d_len, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_stub_cdf launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    int32_t n = 0;
    GSM_CUDA(h, cudaMemcpyAsync(&n, d_len, sizeof n, cudaMemcpyDeviceToHost, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    GSM_CUDA(h, cudaMemcpyAsync(cdf, h->d_cdf, sizeof(double) * std::min(n, cap), cudaMemcpyDeviceToHost, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    *len = n;
    return GSM_OK;
}

// scratch of the log-time whole-tree calls: allocated per call (these run once per logged state, not per MCMC step)
namespace {
struct LogScratch {
    int64_t *trees = nullptr, *lens = nullptr, *ptr = nullptr;
    double *sums = nullptr, *cdf = nullptr, *u = nullptr;
    int32_t* out = nullptr;
    ~LogScratch() {
        cudaFree(trees); cudaFree(lens); cudaFree(ptr); cudaFree(sums); cudaFree(cdf); cudaFree(u); cudaFree(out);
    }
};
int check_sel(gsm_handle* h, const GsmBatch& b, int64_t n_sel, const int64_t* trees, const char* who) {
    if (n_sel < 0) return fail(h, GSM_E_ARG, "%s: negative n_sel", who);
    if (!trees && n_sel > b.n_trees) return fail(h, GSM_E_ARG, "%s: n_sel %lld exceeds the batch (%lld trees)", who, (long long)n_sel, (long long)b.n_trees);
    if (trees)
        for (int64_t k = 0; k < n_sel; k++)
            if (trees[k] < 0 || trees[k] >= b.n_trees) return fail(h, GSM_E_ARG, "%s: tree %lld outside the batch", who, (long long)trees[k]);
    return GSM_OK;
}
}  // namespace

int gsm_stub_cdf_trees(gsm_handle* h, int64_t n_sel, const int64_t* trees, int64_t* cdf_ptr, double* cdf, int64_t cap, int64_t* needed) {
    if (!h) return GSM_E_ARG;
    if (!cdf_ptr || !needed || cap < 0 || (cap > 0 && !cdf)) return fail(h, GSM_E_ARG, "gsm_stub_cdf_trees: bad output arguments");
    GsmBatch b;
    int rc = resident_batch(h, &b);
    if (rc != GSM_OK) return rc;
    if ((rc = check_sel(h, b, n_sel, trees, "gsm_stub_cdf_trees")) != GSM_OK) return rc;
    *needed = 0;
    cdf_ptr[0] = 0;
    if (n_sel == 0 || b.n_nodes == 0) return GSM_OK;
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t items = n_sel * b.n_nodes;
    LogScratch sc;
    if (trees) {
        GSM_CUDA(h, dalloc(&sc.trees, n_sel));
        GSM_CUDA(h, cudaMemcpyAsync(sc.trees, trees, n_sel * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    }
    GSM_CUDA(h, dalloc(&sc.lens, items));
    GSM_CUDA(h, dalloc(&sc.sums, items));
    GSM_CUDA(h, dalloc(&sc.ptr, items + 1));
    int nl = gsm_launch_stub_len(b, n_sel, sc.trees, sc.lens, sc.sums, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_stub_cdf_trees: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    nl = gsm_launch_scan(sc.lens, sc.ptr, items, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_stub_cdf_trees: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    GSM_CUDA(h, cudaMemcpyAsync(cdf_ptr, sc.ptr, (items + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    *needed = cdf_ptr[items];
    const int64_t w = std::min(*needed, cap);
    if (w > 0) {
        GSM_CUDA(h, dalloc(&sc.cdf, w));
        nl = gsm_launch_stub_fill(b, n_sel, sc.trees, sc.ptr, sc.sums, sc.cdf, w, h->stream);
        if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_stub_cdf_trees: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
        h->launches += nl;
        GSM_CUDA(h, cudaMemcpyAsync(cdf, sc.cdf, w * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return GSM_OK;
}

int gsm_stub_choose(gsm_handle* h, int64_t n_sel, const int64_t* trees, const double* u, int32_t* nstubs_out) {
    if (!h) return GSM_E_ARG;
    if (!u || !nstubs_out) return fail(h, GSM_E_ARG, "gsm_stub_choose: NULL array");
    if (!(h->flags & GSM_F_STUBS_TO_SPIKE_PRIOR))   // Stubs.spikePrior == null: Randomizer.nextPoisson, BEAST's own algorithm (Stubs:436-441)
        return fail(h, GSM_E_UNSUPPORTED, "gsm_stub_choose: without a BranchSpikePrior wired to the stubs the reference draws with "
                                          "Randomizer.nextPoisson (BEAST core RNG); only the randomChoice branch is provided");
    GsmBatch b;
    int rc = resident_batch(h, &b);
    if (rc != GSM_OK) return rc;
    if ((rc = check_sel(h, b, n_sel, trees, "gsm_stub_choose")) != GSM_OK) return rc;
    if (n_sel == 0 || b.n_nodes == 0) return GSM_OK;
    GSM_CUDA(h, cudaSetDevice(h->device));
    const int64_t items = n_sel * b.n_nodes;
    LogScratch sc;
    if (trees) {
        GSM_CUDA(h, dalloc(&sc.trees, n_sel));
        GSM_CUDA(h, cudaMemcpyAsync(sc.trees, trees, n_sel * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    }
    GSM_CUDA(h, dalloc(&sc.u, items));
    GSM_CUDA(h, dalloc(&sc.out, items));
    GSM_CUDA(h, cudaMemcpyAsync(sc.u, u, items * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int nl = gsm_launch_stub_choose(b, n_sel, sc.trees, sc.u, sc.out, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_stub_choose: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    GSM_CUDA(h, cudaMemcpyAsync(nstubs_out, sc.out, items * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    return GSM_OK;
}

int gsm_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return GSM_E_ARG;
    cudaError_t e = cudaHostAlloc(out, static_cast<size_t>(bytes > 0 ? bytes : 1), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        g_create_error = std::string("gsm_host_alloc: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? GSM_E_NOMEM : GSM_E_CUDA;
    }
    return GSM_OK;
}

int gsm_host_free(void* p) {
    if (!p) return GSM_OK;
    return cudaFreeHost(p) == cudaSuccess ? GSM_OK : GSM_E_CUDA;
}

int gsm_bind_device(gsm_handle* h, int64_t n_trees, int32_t n_nodes, int64_t stride, const double* d_height,
                    const int32_t* d_parent, const int32_t* d_node_count, const double* d_rate, const double* d_spike,
                    const int32_t* d_nstubs, const double* d_params, const uint8_t* d_use_spike) {
    if (!h) return GSM_E_ARG;
    if (n_trees < 0 || n_nodes < 0) return fail(h, GSM_E_ARG, "gsm_bind_device: negative size");
    if (n_nodes > GSM_MAX_NODES) return fail(h, GSM_E_UNSUPPORTED, "gsm_bind_device: n_nodes %d exceeds GSM_MAX_NODES", n_nodes);
    if (stride < n_nodes || (stride % 32) != 0) return fail(h, GSM_E_ARG, "gsm_bind_device: stride must be >= n_nodes and a multiple of 32");
    if (n_trees > 0 && (!d_height || !d_parent || !d_spike || !d_params)) return fail(h, GSM_E_ARG, "gsm_bind_device: NULL required array");
    if ((reinterpret_cast<uintptr_t>(d_height) | reinterpret_cast<uintptr_t>(d_parent) | reinterpret_cast<uintptr_t>(d_rate) |
         reinterpret_cast<uintptr_t>(d_spike) | reinterpret_cast<uintptr_t>(d_nstubs)) & 15u)
        return fail(h, GSM_E_ARG, "gsm_bind_device: height / parent / rate / spike / nstubs must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_params) & 7u) return fail(h, GSM_E_ARG, "gsm_bind_device: params must be 8-byte aligned");
    h->external = true;
    h->have_trees = true;
    h->have_stubs = h->stub_counts_valid = false;
    h->x_stub_ptr = nullptr; h->x_stub_branch = nullptr; h->x_stub_time = nullptr;
    GsmBatch b{};
    b.n_trees = n_trees; b.n_nodes = n_nodes; b.stride = stride;
    b.height = d_height; b.parent = d_parent; b.node_count = d_node_count;
    b.rate = d_rate; b.spike = d_spike; b.nstubs = d_nstubs; b.params = d_params; b.use_spike = d_use_spike;
    h->cur = b;
    return GSM_OK;
}

int gsm_eval_device(gsm_handle* h, double* d_out_tree, double* d_out_rates, double* d_out_wspikes, void* cuda_stream) {
    if (!h) return GSM_E_ARG;
    GsmBatch b;
    int rc = resident_batch(h, &b);
    if (rc != GSM_OK) return rc;
    if (b.n_trees == 0) return GSM_OK;
    if (!d_out_tree) return fail(h, GSM_E_ARG, "gsm_eval_device: d_out_tree is NULL");
    GSM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
    GSM_CUDA(h, ensure_scratch(h->scratch, h->scratch_trees, b.n_trees));
    int nl = gsm_launch_eval(h->ctx, b, h->scratch, d_out_tree, d_out_rates, d_out_wspikes, st);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_eval_device launch: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    return GSM_OK;
}

int gsm_sync(gsm_handle* h) {
    if (!h) return GSM_E_ARG;
    GSM_CUDA(h, cudaSetDevice(h->device));
    GSM_CUDA(h, cudaStreamSynchronize(h->stream));
    return GSM_OK;
}

// ---- device memory shared between the per-GPU processes of one box (result rows gathered by the copy engines) ----
int gsm_device_alloc(void** out, int32_t device, int64_t bytes) {
    if (!out || bytes < 0) return GSM_E_ARG;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return GSM_E_ARG; }
    cudaError_t e = cudaMalloc(out, static_cast<size_t>(bytes > 0 ? bytes : 1));
    if (e != cudaSuccess) {
        g_create_error = std::string("gsm_device_alloc: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? GSM_E_NOMEM : GSM_E_CUDA;
    }
    e = cudaMemset(*out, 0, static_cast<size_t>(bytes > 0 ? bytes : 1));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e == cudaSuccess ? GSM_OK : GSM_E_CUDA;
}
int gsm_device_free(void* p, int32_t device) {
    if (!p) return GSM_OK;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return GSM_E_ARG; }
    return cudaFree(p) == cudaSuccess ? GSM_OK : GSM_E_CUDA;
}
int gsm_ipc_export(const void* dptr, unsigned char* handle64) {
    if (!dptr || !handle64) return GSM_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hnd;
    cudaError_t e = cudaIpcGetMemHandle(&hnd, const_cast<void*>(dptr));
    if (e != cudaSuccess) {
        g_create_error = std::string("gsm_ipc_export: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return GSM_E_CUDA;
    }
    memcpy(handle64, &hnd, 64);
    return GSM_OK;
}
int gsm_ipc_open(int32_t device, const unsigned char* handle64, void** out) {
    if (!handle64 || !out) return GSM_E_ARG;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return GSM_E_ARG; }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(out, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        g_create_error = std::string("gsm_ipc_open: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return GSM_E_CUDA;
    }
    return GSM_OK;
}
int gsm_ipc_close(void* p) {
    if (!p) return GSM_OK;
    return cudaIpcCloseMemHandle(p) == cudaSuccess ? GSM_OK : GSM_E_CUDA;
}
int gsm_copy_async(void* dst, const void* src, int64_t bytes, void* cuda_stream) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return GSM_E_ARG;
    if (bytes == 0) return GSM_OK;
    cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDefault, static_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) {
        g_create_error = std::string("gsm_copy_async: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return GSM_E_CUDA;
    }
    return GSM_OK;
}

int gsm_measure_fp64_peak(gsm_handle* h, double* fma_per_sec) {
    if (!h || !fma_per_sec) return GSM_E_ARG;
    GSM_CUDA(h, cudaSetDevice(h->device));
    if (h->cdf_cap < 2) {
        dfree(h->d_cdf);
        GSM_CUDA(h, dalloc(&h->d_cdf, (int64_t)4097));
        h->cdf_cap = 4097;
    }
    int nl = gsm_launch_dfma_peak(h->ctx, h->d_cdf, 5, fma_per_sec, h->stream);
    if (nl < 0) return fail(h, GSM_E_CUDA, "gsm_measure_fp64_peak: %s", cudaGetErrorString(static_cast<cudaError_t>(-nl)));
    h->launches += nl;
    return GSM_OK;
}

int64_t gsm_launch_count(const gsm_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
