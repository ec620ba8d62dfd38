// gsm_kernels.cu — sm_100a kernels of the GammaSpikeModel per-branch hot path.
//
// gsm_eval_kernel: one CTA per tree state.
//   phase 0  one elected thread issues two TMA bulk copies (cp.async.bulk, mbarrier complete_tx) that bring
//            the tree's height row and parent row into shared memory; meanwhile one warp computes the
//            per-tree constants (c1, c2, logs, lgamma(shape*m) table) and the rest clear the flag bytes.
//   phase 1  every node marks its parent as non-leaf                         (Node.isLeaf)
//   phase 1b every leaf whose parent has the same height marks the parent fake (Node.isDirectAncestor / isFake)
//   phase 2  pass A per node: leaf census, the node's own tree-density term, aux(h) = log G(h) -> shared memory
//   phase 3  pass B per branch: rate/spike/stub-count rows are streamed from HBM with coalesced loads
//            (each byte is read exactly once), parent height/aux/flags are gathered from shared memory
//   phase 4  fixed-order warp-shuffle + cross-warp reduction, per-tree epilogue, one 64-byte result row.
// The math is in gsm_node_math.cuh (shared with the CPU-side logic tests).

#include "gsm_kernels.cuh"
#include "gsm_node_math.cuh"
#include "gsm_lean.cuh"

#include <cstdio>
#include <cstdlib>

__device__ double g_logfact[GSM_LOGFACT_N];
#ifdef GSM_PHASE_TIMING
// debug builds (tools/build_dbg.sh): cycles spent by thread 0 of every CTA of the lean kernel between phase boundaries
__device__ unsigned long long g_phase_cycles[8];
__device__ unsigned long long g_warp_cycles[16];   // pass B duration per warp (lane 0)
extern "C" int gsm_debug_phase_cycles(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    }
    return static_cast<int>(e);
}
extern "C" int gsm_debug_warp_cycles(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_warp_cycles, sizeof(g_warp_cycles));
    if (e == cudaSuccess && reset) {
        unsigned long long z[16] = {};
        e = cudaMemcpyToSymbol(g_warp_cycles, z, sizeof(z));
    }
    return static_cast<int>(e);
}
#define GSM_WARP_T0() const long long tw0_ = clock64()
#define GSM_WARP_T1(stage) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_warp_cycles[(threadIdx.x >> 5) + 8 * (stage)], static_cast<unsigned long long>(clock64() - tw0_)); } while (0)
#define GSM_PHASE_MARK(k)                                                          \
    do {                                                                           \
        if (threadIdx.x == 0) {                                                    \
            const long long now_ = clock64();                                      \
            atomicAdd(&g_phase_cycles[k], static_cast<unsigned long long>(now_ - t_mark_)); \
            t_mark_ = now_;                                                        \
        }                                                                          \
    } while (0)
#define GSM_PHASE_START() long long t_mark_ = clock64()
#else
#define GSM_PHASE_MARK(k) do { } while (0)
#define GSM_PHASE_START() do { } while (0)
#define GSM_WARP_T0() do { } while (0)
#define GSM_WARP_T1(stage) do { } while (0)
#endif
// constant tables of the lean path (stubtab is filled by gsm_init_tables)
__device__ GsmLeanBlob g_lean_blob;   // replicated log / exp tables, filled by gsm_init_tables

// ---------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (global -> shared::cluster, completion on an mbarrier)
// ---------------------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
static __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GSM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GSM_DONE_%=;\n"
        "bra GSM_WAIT_%=;\n"
        "GSM_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned
static __device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

static __device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) global loads: read-only path, do not allocate in L1
static __device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
static __device__ __forceinline__ int2 ld_stream(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
static __device__ __forceinline__ int32_t ld_stream(const int32_t* p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

static __device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }

// shared-memory accesses by 32-bit shared-window address (the hot loops of the lean kernel keep base addresses in
// registers; going through generic pointers costs a window-base computation per access)
static __device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
static __device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
static __device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
static __device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
static __device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
static __device__ __forceinline__ int4 lds_v4s32(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
static __device__ __forceinline__ void sts_f64(uint32_t a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }
static __device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
static __device__ __forceinline__ void sts_v4u32(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
static __device__ __forceinline__ int2 lds_v2s32(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
static __device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(static_cast<uint16_t>(v)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// per-tree scratch row written by the lean kernel and consumed by the post kernel (128 bytes)
// ---------------------------------------------------------------------------------------------------------
struct __align__(16) GsmLeanRow {
    double tree, stub, spike, rate, burst, dist;   // combined factor sums (gsm_lean_combine), reduced over the CTA
    double T;                                      // root height
    int32_t n_nst, n_leafterm, n_plain, n_fake;
    int32_t n_extant, n_extinct, n_da, root_fake;
    int32_t status;                                // 1: valid, 0: tree was sent to the redo list
    int32_t pad[9];
};
static_assert(sizeof(GsmLeanRow) == 128, "GsmLeanRow must be 128 bytes");

// ---------------------------------------------------------------------------------------------------------
// generic evaluation of one tree by one CTA (every wiring, every special value): the statement-by-statement
// path.  Used directly for wirings outside the lean family and for the trees the lean kernel hands back.
// ---------------------------------------------------------------------------------------------------------
// Per-tree state kept for incremental proposal evaluation (gsm_eval_dirty): the reduced accumulators of the generic path
// before the epilogue, the root, and whether the tree qualifies (one root, every parent in range, no invalid-state flag).
struct __align__(16) GsmBaseRow {
    double acc[6];                                   // GsmAcc: tree, stub, spike, rate, burst, dist
    int32_t n_extant, n_leaf, n_da, n_extinct, bad;
    int32_t root, n, ok;
    int32_t pad[8];
};
static_assert(sizeof(GsmBaseRow) == 112, "GsmBaseRow layout");

struct GsmGenericShared {
    GsmTreeConsts K;
    uint64_t mbar;
    int32_t root;
    int32_t nroot, nbadpar;
    double red[6][16];
    int32_t redi[5][16];
};

// Returns true when the tree went through the bulk-copy phase of sh.mbar (the caller flips its parity bit then).
template <int BLOCK, bool WANT_RATES>
static __device__ bool gsm_generic_tree(const GsmBatch& b, int64_t t, uint32_t parity, unsigned char* smem_raw,
                                        GsmGenericShared& sh, double* __restrict__ out_tree, double* __restrict__ out_rates,
                                        double* __restrict__ out_wspikes, GsmBaseRow* __restrict__ base_rows = nullptr,
                                        GsmTreeConsts* __restrict__ base_K = nullptr) {
    constexpr int NWARP = BLOCK / 32;
    const int64_t S = b.stride;
    double* sh_h = reinterpret_cast<double*>(smem_raw);
    double* sh_aux = sh_h + S;
    int32_t* sh_par = reinterpret_cast<int32_t*>(sh_aux + S);
    uint8_t* sh_nonleaf = reinterpret_cast<uint8_t*>(sh_par + S);
    uint8_t* sh_fake = sh_nonleaf + S;
    GsmTreeConsts& K = sh.K;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    double* orow = out_tree + t * GSM_NOUT;
    if (n <= 0) {
        if (tid < GSM_NOUT) orow[tid] = NAN;
        if (base_rows && tid == 0) base_rows[t].ok = 0;
        return false;
    }
    const int64_t row = t * S;

    // ---- phase 0 ----
    if (tid == 0) {
        sh.root = -1;
        sh.nroot = 0;
        sh.nbadpar = 0;
        const uint32_t bytes_h = (static_cast<uint32_t>(n) * 8u + 15u) & ~15u;
        const uint32_t bytes_p = (static_cast<uint32_t>(n) * 4u + 15u) & ~15u;
        mbar_arrive_expect_tx(&sh.mbar, bytes_h + bytes_p);
        tma_bulk_g2s(sh_h, b.height + row, bytes_h, &sh.mbar);
        tma_bulk_g2s(sh_par, b.parent + row, bytes_p, &sh.mbar);
    }
    {
        // clear both flag arrays (2*S bytes, S % 32 == 0) with 32-bit stores
        uint32_t* w = reinterpret_cast<uint32_t*>(sh_nonleaf);
        const int nw = static_cast<int>(S >> 1);
        for (int i = tid; i < nw; i += BLOCK) w[i] = 0u;
    }
    if (warp == NWARP - 1) {
        const double* prow = b.params + t * GSM_NPARAM;
        if (lane == 0) {
            gsm_setup_consts(K, prow, b.use_spike ? static_cast<int32_t>(b.use_spike[t]) : 1, b.flags, b.stub_mode);
        } else if (lane <= GSM_LUT_M + 1) {
            gsm_lut_entry(K, prow[GSM_P_SPIKE_SHAPE], lane - 1);  // m = 0..GSM_LUT_M
        }
    }
    __syncthreads();
    mbar_wait(&sh.mbar, parity);

    // ---- phase 1: non-leaf marks, root, parent sanitising ----
    for (int i = tid; i < n; i += BLOCK) {
        const int32_t p = sh_par[i];
        if (static_cast<uint32_t>(p) < static_cast<uint32_t>(n)) sh_nonleaf[p] = 1;
        else {
            if (p < 0) {
                sh.root = i;
                if (base_rows) atomicAdd(&sh.nroot, 1);
            } else {
                sh_par[i] = -1;  // out-of-range parent: treated as detached (never dereferenced)
                if (base_rows) atomicAdd(&sh.nbadpar, 1);
            }
        }
    }
    __syncthreads();

    // ---- phase 1b: sampled ancestors -> fake parents ----
    for (int i = tid; i < n; i += BLOCK) {
        const int32_t p = sh_par[i];
        if (!sh_nonleaf[i] && p >= 0 && sh_h[p] == sh_h[i]) sh_fake[p] = 1;
    }
    __syncthreads();

    // ---- phase 2: pass A (census, own tree-density term, aux) ----
    GsmAcc A;
    gsm_acc_zero(A);
    for (int i = tid; i < n; i += BLOCK) {
        const double h = sh_h[i];
        const int32_t p = sh_par[i];
        uint32_t nf = 0;
        double hp = h;
        if (sh_nonleaf[i]) nf |= GSM_NF_NONLEAF;
        if (sh_fake[i]) nf |= GSM_NF_FAKE;
        if (p < 0) nf |= GSM_NF_ROOT;
        else {
            hp = sh_h[p];
            if (!(nf & GSM_NF_NONLEAF) && hp == h) nf |= GSM_NF_DA;
        }
        sh_aux[i] = gsm_node_pass_a(K, A, nf, h, hp);
    }
    __syncthreads();

    // ---- phase 3: pass B (branch parameters streamed from HBM) ----
    const double* rate_row = b.rate ? b.rate + row : nullptr;
    const double* spike_row = b.spike + row;
    const int32_t* nst_row = b.nstubs ? b.nstubs + row : nullptr;
    double* orate = (WANT_RATES && out_rates) ? out_rates + row : nullptr;
    double* owsp = (WANT_RATES && out_wspikes) ? out_wspikes + row : nullptr;

    int i = tid;
    double rate_n = 1.0, spike_n = 0.0;
    int32_t nst_n = 0;
    if (i < n) {
        if (rate_row) rate_n = __ldg(rate_row + i);
        spike_n = __ldg(spike_row + i);
        if (nst_row) nst_n = __ldg(nst_row + i);
    }
    while (i < n) {
        const double rate = rate_n, spike = spike_n;
        const int32_t nst = nst_n;
        const int inext = i + BLOCK;
        if (inext < n) {  // software prefetch of the next strip
            if (rate_row) rate_n = __ldg(rate_row + inext);
            spike_n = __ldg(spike_row + inext);
            if (nst_row) nst_n = __ldg(nst_row + inext);
        }
        const double h = sh_h[i];
        const int32_t p = sh_par[i];
        const double aux = sh_aux[i];
        uint32_t nf = 0;
        double hp = h, auxp = aux;
        if (sh_nonleaf[i]) nf |= GSM_NF_NONLEAF;
        if (sh_fake[i]) nf |= GSM_NF_FAKE;
        if (p < 0) {
            nf |= GSM_NF_ROOT;
        } else {
            hp = sh_h[p];
            auxp = sh_aux[p];
            if (sh_fake[p]) nf |= GSM_NF_PARENT_FAKE;
            if (!(nf & GSM_NF_NONLEAF) && hp == h) nf |= GSM_NF_DA;
        }
        gsm_node_pass_b<WANT_RATES>(K, g_logfact, A, i, n, nf, h, hp, aux, auxp, rate, spike, nst,
                                    orate ? orate + i : nullptr, owsp ? owsp + i : nullptr);
        i = inext;
    }

    // ---- phase 3b: EXACT mode, the listed stubs of this tree (STP:573-592 / 639-655) ----
    if (b.stub_ptr != nullptr && K.stub_term && K.stub_mode == GSM_STUBS_EXACT) {
        const int64_t s0 = b.stub_ptr[t], s1 = b.stub_ptr[t + 1];
        for (int64_t j = s0 + tid; j < s1; j += BLOCK) {
            const int32_t br = __ldg(b.stub_branch + j);
            if (static_cast<uint32_t>(br) >= static_cast<uint32_t>(n)) continue;   // matches no branchNr of STP:243-253
            const int32_t p = sh_par[br];
            const double h = sh_h[br];
            const bool root = p < 0;
            const double hp = root ? h : sh_h[p];
            const bool da = !root && !sh_nonleaf[br] && hp == h;
            gsm_stub_term(K, A, root, da, h, hp, __ldg(b.stub_time + j));
        }
    }

    // ---- phase 4: reductions (fixed order) + epilogue ----
    double v0 = warp_sum(A.tree), v1 = warp_sum(A.stub), v2 = warp_sum(A.spike), v3 = warp_sum(A.rate),
           v4 = warp_sum(A.burst), v5 = warp_sum(A.dist);
    int32_t c0 = __reduce_add_sync(0xffffffffu, A.n_extant), c1 = __reduce_add_sync(0xffffffffu, A.n_leaf),
            c2 = __reduce_add_sync(0xffffffffu, A.n_da), c3 = __reduce_add_sync(0xffffffffu, A.n_extinct),
            c4 = static_cast<int32_t>(__reduce_or_sync(0xffffffffu, static_cast<uint32_t>(A.bad)));
    if (lane == 0) {
        sh.red[0][warp] = v0; sh.red[1][warp] = v1; sh.red[2][warp] = v2;
        sh.red[3][warp] = v3; sh.red[4][warp] = v4; sh.red[5][warp] = v5;
        sh.redi[0][warp] = c0; sh.redi[1][warp] = c1; sh.redi[2][warp] = c2; sh.redi[3][warp] = c3; sh.redi[4][warp] = c4;
    }
    __syncthreads();
    if (tid == 0) {
        GsmAcc R;
        gsm_acc_zero(R);
#pragma unroll
        for (int w = 0; w < NWARP; w++) {
            R.tree += sh.red[0][w]; R.stub += sh.red[1][w]; R.spike += sh.red[2][w];
            R.rate += sh.red[3][w]; R.burst += sh.red[4][w]; R.dist += sh.red[5][w];
            R.n_extant += sh.redi[0][w]; R.n_leaf += sh.redi[1][w]; R.n_da += sh.redi[2][w];
            R.n_extinct += sh.redi[3][w]; R.bad |= sh.redi[4][w];
        }
        const int32_t root = sh.root;
        if (base_rows) {
            GsmBaseRow br;
            br.acc[0] = R.tree; br.acc[1] = R.stub; br.acc[2] = R.spike; br.acc[3] = R.rate; br.acc[4] = R.burst; br.acc[5] = R.dist;
            br.n_extant = R.n_extant; br.n_leaf = R.n_leaf; br.n_da = R.n_da; br.n_extinct = R.n_extinct; br.bad = R.bad;
            br.root = root; br.n = n;
            bool fin = true;
#pragma unroll
            for (int k = 0; k < 6; k++) fin = fin && (fabs(br.acc[k]) < INFINITY);
            br.ok = (fin && R.bad == 0 && sh.nroot == 1 && sh.nbadpar == 0 && root >= 0) ? 1 : 0;
#pragma unroll
            for (int k = 0; k < 8; k++) br.pad[k] = 0;
            base_rows[t] = br;
            base_K[t] = K;
        }
        double res[GSM_NOUT];
        if (root < 0) {
#pragma unroll
            for (int k = 0; k < GSM_NOUT; k++) res[k] = NAN;  // no root: not a tree
        } else {
            gsm_finish_consts(K, sh_h[root]);
            gsm_finalize(K, R, n, sh_fake[root] != 0, res);
        }
#pragma unroll
        for (int k = 0; k < GSM_NOUT; k++) orow[k] = res[k];
    }
    return true;
}

// One CTA per tree (wirings outside the lean family).
// MINB = CTAs per SM the register allocation is sized for (__launch_bounds__): BLOCK*MINB resident threads.
template <int BLOCK, int MINB, bool WANT_RATES>
__global__ void __launch_bounds__(BLOCK, MINB) gsm_eval_kernel(const GsmBatch b, double* __restrict__ out_tree,
                                                               double* __restrict__ out_rates,
                                                               double* __restrict__ out_wspikes,
                                                               GsmBaseRow* __restrict__ base_rows,
                                                               GsmTreeConsts* __restrict__ base_K) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ GsmGenericShared sh;
    if (threadIdx.x == 0) {
        mbar_init(&sh.mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    gsm_generic_tree<BLOCK, WANT_RATES>(b, blockIdx.x, 0u, smem_raw, sh, out_tree, out_rates, out_wspikes, base_rows, base_K);
}

// ---------------------------------------------------------------------------------------------------------
// post kernel of the lean path (second and last launch of an evaluation):
//   part 1  one THREAD per tree: per-tree constants with their logarithms, per-node constants, the
//           dispatcher / epilogue of STP:150-284 (gsm_finalize) on the lean kernel's scratch row -> result row
//   part 2  one CTA per tree of the redo list: the generic path
// ---------------------------------------------------------------------------------------------------------
template <int BLOCK, int MINB, bool WANT_RATES>
__global__ void __launch_bounds__(BLOCK, MINB) gsm_post_kernel(const GsmBatch b, const GsmLeanRow* __restrict__ rows,
                                                               const int32_t* __restrict__ redo_count,
                                                               const int32_t* __restrict__ redo_list,
                                                               double* __restrict__ out_tree, double* __restrict__ out_rates,
                                                               double* __restrict__ out_wspikes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ GsmGenericShared sh;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * BLOCK + threadIdx.x; t < b.n_trees;
         t += static_cast<int64_t>(gridDim.x) * BLOCK) {
        const GsmLeanRow r = rows[t];
        if (r.status != 1) continue;
        int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
        if (n > b.n_nodes) n = b.n_nodes;
        GsmTreeConsts K;
        gsm_setup_consts(K, b.params + t * GSM_NPARAM, b.use_spike ? static_cast<int32_t>(b.use_spike[t]) : 1, b.flags,
                         b.stub_mode);
        GsmLeanC c;
        gsm_lean_consts(K, c, n);
        GsmAcc A;
        gsm_acc_zero(A);
        A.n_leaf = c.L;
        A.n_extant = r.n_extant;
        A.n_extinct = r.n_extinct;
        A.n_da = r.n_da;
        GsmLean6 s6;
        s6.tree = r.tree; s6.stub = r.stub; s6.spike = r.spike; s6.rate = r.rate; s6.burst = r.burst; s6.dist = r.dist;
        gsm_lean_finish(K, c, s6, r.n_nst, r.n_leafterm, r.n_plain, r.n_fake, n, A);
        gsm_finish_consts(K, r.T);
        double res[GSM_NOUT];
        gsm_finalize(K, A, n, r.root_fake != 0, res);
        double* orow = out_tree + t * GSM_NOUT;
#pragma unroll
        for (int k = 0; k < GSM_NOUT; k++) orow[k] = res[k];
    }
    const int32_t cnt = *redo_count;
    if (static_cast<int32_t>(blockIdx.x) >= cnt) return;
    if (threadIdx.x == 0) {
        mbar_init(&sh.mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    for (int32_t k = blockIdx.x; k < cnt; k += gridDim.x) {
        if (gsm_generic_tree<BLOCK, WANT_RATES>(b, redo_list[k], parity, smem_raw, sh, out_tree, out_rates, out_wspikes)) parity ^= 1u;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// lean kernel: default wiring family (gsm_lean_wiring_ok), one CTA per tree.  See gsm_lean.cuh.
//
// What stays in shared memory is only what is gathered at random: the heights and aux = log G(h) of the INTERNAL
// nodes (a leaf is nobody's parent), 16 bytes per internal node.  Everything else is read once, in node order, and
// comes through a three-stage ring of chunks of 2 BLOCK nodes filled by TMA bulk copies: rate, spike, stub count
// and parent index of every node, and the heights of the leaves.
//   phase 0  one thread requests the internal heights, the replicated log / exp tables and the first three chunks;
//            one lane computes the arithmetic-only per-tree constants
//   phase 1  64 lanes compute the Gamma tables {log shape - logGamma(shape m)} (commons-math Lanczos with fast
//            reciprocals); trees that may hold sampled ancestors (psi > 0) are scanned for them (leaf height ==
//            parent height -> the parent is fake)
//   phase 2  pass A: aux = log G(h) of the internal nodes (pairs)
//   phase 3  pass B: leaf chunks (aux of a leaf is computed on the spot; a warp whose 64 heights are all zero uses
//            the constant), then internal chunks; one node pair per thread and chunk; CLEAN or GENERAL flavour
//   phase 4  per-thread combination to six sums, warp-shuffle + cross-warp reduction, 128-byte scratch row
// BEAST's numbering (leaves 0..L-1, every parent in [L, n), one root) is validated on the fly (chk_p).
// ---------------------------------------------------------------------------------------------------------
#ifndef GSM_RING_STAGES
#define GSM_RING_STAGES 3
#endif
#ifndef GSM_SCAN_U
#define GSM_SCAN_U 8   // leaves per thread and round of the sampled-ancestor scan
#endif
#ifndef GSM_PA_BATCH
#define GSM_PA_BATCH 64   // pairs per dynamic pass A batch (two per lane and 64)
#endif
#ifndef GSM_PA_DYN
#define GSM_PA_DYN 0   // 1: pass A by dynamic batches also for BLOCK >= 256
#endif
// node pairs per thread and chunk of pass B: the 128 x 2 configuration (few warps, 255 registers) interleaves two pairs
__host__ __device__ constexpr int gsm_lean_pairs(int block, int minb) { return (block == 128 && minb == 2) ? 2 : 1; }
// dynamic shared memory of the lean kernel, as offsets from its start (the shared-window address a_dyn).  The log table
// (8 KB) sits on an 8 KB boundary of the window and the exp table (4 KB) on a 4 KB boundary: right in front of the log
// table when the gap there is large enough, else behind it.  Then internal heights + aux (si x 16 bytes) and the ring.
// The three small per-tree arrays (mtab, ntab, fake flags) fill what is left of the gap, first fit, else follow the ring.
struct GsmLeanLayout {
    uint32_t logtab, exptab, arrays, ring, mtab, ntab, fk, total;
};
__host__ __device__ inline int gsm_lean_si(int n_nodes) { return (((n_nodes + 1) >> 1) + 8 + 1) & ~1; }   // n - L4 + pad, even
// pk (persistent kernel): no mtab
__host__ __device__ inline GsmLeanLayout gsm_lean_layout(uint32_t a_dyn, int si, uint32_t ring_bytes, bool pk = false) {
    GsmLeanLayout l;
    const uint32_t gap = (8192u - (a_dyn & 8191u)) & 8191u;
    uint32_t free0 = 0u, free1 = gap;             // free part of the gap
    l.logtab = gap;
    uint32_t end = gap + GSM_T6_WORDS * 8u;
    if (gap >= GSM_T6E_WORDS * 8u) {
        l.exptab = gap - GSM_T6E_WORDS * 8u;
        free1 = l.exptab;
    } else {
        l.exptab = end;
        end += GSM_T6E_WORDS * 8u;
    }
    l.arrays = end;
    end += static_cast<uint32_t>(si) * 16u;
    l.ring = end;
    end += ring_bytes;
    const uint32_t sz[3] = {pk ? 0u : (GSM_LEAN_MTAB_N + 1) * 16u, (GSM_LEAN_MAX_NST + 1) * 32u,
                            (static_cast<uint32_t>(si) + 15u) & ~15u};
    uint32_t off[3];
    for (int k = 0; k < 3; k++) {
        if (free1 - free0 >= sz[k]) {
            off[k] = free0;
            free0 += sz[k];
        } else {
            off[k] = end;
            end += sz[k];
        }
    }
    l.mtab = off[0];
    l.ntab = off[1];
    l.fk = off[2];
    l.total = end;
    return l;
}
// ring stages (<= GSM_RING_STAGES)
__host__ __device__ constexpr int gsm_lean_stages(int block, int minb) { return (void)block, (void)minb, GSM_RING_STAGES; }
#define GSM_TAIL_SLOTS 5
struct __align__(8) GsmLeanXPair {
    int32_t i0, pad;      // first node of the pair, -1: none
    int32_t i[2][2];      // parent, stub count
    double d[2][3];       // height, rate, spike
};
template <int NWARP>
struct GsmLeanShared {
    GsmTreeConsts K;
    uint64_t mbar;                     // internal heights + tables
    uint64_t full[GSM_RING_STAGES];    // ring: chunk landed
    uint64_t empty[GSM_RING_STAGES];   // ring: every warp holds its pairs of the chunk in registers
    int32_t root, nroot, bad, root_last;
    int32_t pa_next;         // pass A: next batch of pairs
    // tail nodes: slots 0-2 = L-1, L (odd L) and n-1, rows fetched by asynchronous copies in phase 0; slots 3-4 = the pair of
    // internal nodes that holds the root when the root is not node n-1 (xp, handed over by the thread that met it in pass B)
    double tail_d[3][3];                  // height, rate, spike
    int32_t tail_i[3][2];                 // parent, stub count
    GsmLeanXPair xp;
    double tail_s6[GSM_TAIL_SLOTS][6];    // their combined sums ...
    int32_t tail_c[GSM_TAIL_SLOTS][6];    // ... counts (leafterm, plain, fake, extant, extinct, da) ...
    uint32_t tail_u[GSM_TAIL_SLOTS][4];   // ... and check words (chk, chk_n, hmax, chk_p)
    double red[6][NWARP];
    int32_t redi[7][NWARP];
    uint32_t redu[5][NWARP];
};

static __device__ __forceinline__ void lean_redo(int32_t* __restrict__ redo_count, int32_t* __restrict__ redo_list,
                                                 GsmLeanRow* __restrict__ rows, int64_t t) {
    const int32_t k = atomicAdd(redo_count, 1);
    redo_list[k] = static_cast<int32_t>(t);
    rows[t].status = 0;
}

// 1 / x for a positive normal x: hardware approximation + two Newton steps (relative error ~1e-16), no special cases
static __device__ __forceinline__ double lean_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return fma(fma(-x, r, 1.0), r, r);
}
// commons-math Lanczos logGamma (same terms, same summation order as gsm_log_gamma / the Java loop) with reciprocals
// instead of divisions and the kernel's table-driven logarithm: agrees with the Java value to ~1e-15 relative.
static __device__ __forceinline__ double lean_log_gamma(const GsmT6& t6, double x) {
    double q[15];
#pragma unroll
    for (int i = 14; i > 0; --i) q[i] = gsm_lanczos_d[i] * lean_rcp(x + i);
    double sum = 0.0;
#pragma unroll
    for (int i = 14; i > 0; --i) sum = sum + q[i];
    sum = sum + gsm_lanczos_d[0];
    const double tmp = x + 607.0 / 128.0 + .5;
    return ((x + .5) * gsm_log6_hi(t6, tmp)) - tmp + 0.91893853320467274178 + gsm_log6_hi(t6, sum * lean_rcp(x));
}

#ifndef GSM_PK_TMA_DOORBELL
#define GSM_PK_TMA_DOORBELL 0
#endif
// named barriers of the persistent kernel (gsm_lean_pk.cuh): 0 __syncthreads, 2 consumers, then
#define GSM_PK_BAR_STAGE 3    // + stage: a consumer warp has taken its share of the chunk in this stage (wakes the copy warp)
#define GSM_PK_BAR_SCAL 6     // + buffer: the consumers are through with a tree (wakes the scalar warp, two trees ahead)
#define GSM_PK_BAR_GAMMA 8    // + buffer: ... (wakes the Gamma warps, one tree ahead)

// ---- the ring -------------------------------------------------------------------------------------------
// Virtual chunk v: v < nlc is leaf chunk v (nodes [v * 2 BLOCK, ...) up to L), else internal chunk v - nlc (nodes from
// L4 = L & ~3 on, up to n - 1).  Stage layout (BLOCK = 256: 16 KB):
//   [rate: 2 BLOCK x 8][spike: 2 BLOCK x 8][height: 2 BLOCK x 8, leaf chunks only][stub count: 2 BLOCK x 4][parent: 2 BLOCK x 4]
// full mbarrier: arrive.expect_tx by the requester, complete_tx by the copies.  A warp that holds its pairs of a chunk in
// registers arrives on the stage's empty mbarrier.
struct LeanRing {
    uint32_t a_ring;            // shared-window address of stage 0
    uint32_t a_full, a_empty;   // ... of full[0], empty[0]
    const double* rate;         // row pointers
    const double* spike;
    const double* height;
    const int32_t* nst;         // nullptr: no stub counts
    const int32_t* parent;
    int L, L4, nlc, nvc;        // leaves, L & ~3, leaf chunks, all chunks
    int end_leaf, end_int;      // node ranges to copy: leaf chunks stop at end_leaf (L rounded up to 4), internal at end_int
};

static __device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GSM_WAITA_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GSM_DONEA_%=;\n"
        "bra GSM_WAITA_%=;\n"
        "GSM_DONEA_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
static __device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ void tma_bulk_g2s_a(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src_gmem),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

static __device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
static __device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// one thread: request virtual chunk v into stage s.  CH = nodes per chunk (PAIRS node pairs per thread).
template <int CH>
static __device__ __forceinline__ void ring_issue(const LeanRing& rg, int v, uint32_t s) {
    constexpr uint32_t STAGE = CH * 32u;
    constexpr int BLOCK = CH / 2;   // layout unit below: CH nodes = 2 BLOCK
    const bool leafc = v < rg.nlc;
    const int o = leafc ? v * 2 * BLOCK : rg.L4 + (v - rg.nlc) * 2 * BLOCK;
    int cnt = (leafc ? rg.end_leaf : rg.end_int) - o;
    if (cnt > 2 * BLOCK) cnt = 2 * BLOCK;
    const uint32_t b8 = static_cast<uint32_t>(cnt) * 8u, b4 = static_cast<uint32_t>(cnt) * 4u;
    const uint32_t bar = rg.a_full + s * 8u, st = rg.a_ring + s * STAGE;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"(2u * b8 + (leafc ? b8 : 0u) + (rg.nst ? 2u * b4 : b4))
                 : "memory");
    tma_bulk_g2s_a(st, rg.rate + o, b8, bar);
    tma_bulk_g2s_a(st + BLOCK * 16u, rg.spike + o, b8, bar);
    if (leafc) tma_bulk_g2s_a(st + BLOCK * 32u, rg.height + o, b8, bar);
    if (rg.nst) tma_bulk_g2s_a(st + BLOCK * 48u, rg.nst + o, b4, bar);
    tma_bulk_g2s_a(st + BLOCK * 56u, rg.parent + o, b4, bar);
}

// per-thread state of pass B that is not a sum
struct LeanBCtx {
    uint32_t a_hs, a_as;   // shared-window addresses of the internal height / aux arrays, biased: node i is at a_hs + i * 8
    uint32_t a_fk;         // ... of the fake flags (bytes), biased the same way: node i at a_fk + i
    uint32_t a_hL, a_aL, a_fL;   // the same arrays addressed by d = i - L
    uint32_t chk_p;        // max over the nodes of (parent - L) as unsigned: >= n - L means outside [L, n)
    uint32_t nL;           // n - L
    double auxz, ez, Gz;   // aux, e, G of height zero
};

// one node pair (i0, i0 + 1), both leaves (LEAF = 1) or both internal (LEAF = 0)
// XPD: distance in shared memory from the ring's first full barrier to the GsmLeanXPair hand-over slots (both members of the
// kernel's shared struct; 0: the kernel has none); a_full: address of that barrier (live in the chunk loop anyway)
template <bool GENERAL, bool WANT_RATES, int LEAF, bool STDW, int LT, bool MIX, int XPD>
static __device__ __forceinline__ void lean_b_pair(const GsmLeanC& c, GsmLeanAccA& SA, GsmLeanAccB& SB, LeanBCtx& x, uint32_t a_full, int i0, bool on,
                                                   const double2 rr, const double2 ss, double2 h2, const int2 kk, int2 pp,
                                                   int32_t* sh_root_nroot, double2* __restrict__ orate2, double2* __restrict__ owsp2) {
    double2 a2;
    double e0 = x.ez, G0 = x.Gz, e1 = x.ez, G1 = x.Gz;
    if (LEAF == 1) {
        // aux of the two leaves: the constant if every height of the warp is zero
        a2 = make_double2(x.auxz, x.auxz);
        const bool zero = (__double2hiint(h2.x) | __double2loint(h2.x) | __double2hiint(h2.y) | __double2loint(h2.y)) == 0;
        if (!__all_sync(0xffffffffu, zero || !on)) {
            if (!on) h2 = make_double2(0.0, 0.0);
            a2.x = gsm_lean_aux(c, SA, h2.x, e0, G0);
            a2.y = gsm_lean_aux(c, SA, h2.y, e1, G1);
        }
        if (!on) return;
    } else {
        if (!on) return;
        h2 = lds_v2f64(x.a_hs + i0 * 8);
        a2 = lds_v2f64(x.a_as + i0 * 8);
        if (!GENERAL && (pp.x | pp.y) < 0) {
            // An ordinary tree whose root is not node n-1 (BEAST's operators leave the root anywhere): the pair that holds it
            // goes to two tail lanes, which evaluate single nodes with the GENERAL flavour after pass B; every other pair of
            // the tree keeps the CLEAN body.  (A second such pair, or no hand-over slots: not a tree this path finishes.)
            if (XPD == 0) {
                SB.chk = 0xffffffffu;
                return;
            }
            const uint32_t a_xpair = a_full + static_cast<uint32_t>(XPD);
            uint32_t old;
            asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a_xpair), "r"(0xffffffffu), "r"(static_cast<uint32_t>(i0)) : "memory");
            if (old != 0xffffffffu) {
                SB.chk = 0xffffffffu;
                return;
            }
            const uint32_t a_xi = a_xpair + 8u, a_xd = a_xpair + 24u;   // GsmLeanXPair::i, ::d
            sts_f64(a_xd + 0, h2.x); sts_f64(a_xd + 8, rr.x); sts_f64(a_xd + 16, ss.x);
            sts_f64(a_xd + 24, h2.y); sts_f64(a_xd + 32, rr.y); sts_f64(a_xd + 40, ss.y);
            sts_u32(a_xi + 0, static_cast<uint32_t>(pp.x)); sts_u32(a_xi + 4, static_cast<uint32_t>(kk.x));
            sts_u32(a_xi + 8, static_cast<uint32_t>(pp.y)); sts_u32(a_xi + 12, static_cast<uint32_t>(kk.y));
            return;
        }
    }
    bool root0 = false, root1 = false;
    if (GENERAL) {   // a root has no parent: it points to itself
        root0 = pp.x < 0;
        root1 = pp.y < 0;
        if (root0 || root1) {
            atomicAdd(sh_root_nroot + 1, (root0 ? 1 : 0) + (root1 ? 1 : 0));
            sh_root_nroot[0] = root0 ? i0 : i0 + 1;
            if (LEAF == 1) SB.chk = 0xffffffffu;   // a leaf cannot be the root
            if (root0) pp.x = LEAF == 1 ? c.L : i0;
            if (root1) pp.y = LEAF == 1 ? c.L : i0 + 1;
        }
    }
    // parent slots: d = parent - L must be in [0, n - L); a malformed tree is flagged (chk_p) and its gathers clamped
    const uint32_t d0 = static_cast<uint32_t>(pp.x - c.L), d1 = static_cast<uint32_t>(pp.y - c.L);
    x.chk_p = gsm_umax(x.chk_p, gsm_umax(d0, d1));
    const uint32_t q0 = umin(d0, x.nL - 1u), q1 = umin(d1, x.nL - 1u);
    double hp0 = lds_f64(x.a_hL + q0 * 8), hp1 = lds_f64(x.a_hL + q1 * 8);
    double ap0 = lds_f64(x.a_aL + q0 * 8), ap1 = lds_f64(x.a_aL + q1 * 8);
    bool pf0 = false, pf1 = false, fs0 = false, fs1 = false;
    if (GENERAL) {
        pf0 = lds_u8(x.a_fL + q0) != 0;
        pf1 = lds_u8(x.a_fL + q1) != 0;
        if (LEAF == 0) {
            fs0 = lds_u8(x.a_fk + i0) != 0;
            fs1 = lds_u8(x.a_fk + i0 + 1) != 0;
        }
        if (root0) { hp0 = h2.x; ap0 = a2.x; }
        if (root1) { hp1 = h2.y; ap1 = a2.y; }
    }
    if (LEAF == 1 && LT != 0) {
        const bool da0 = GENERAL && hp0 == h2.x && !root0, da1 = GENERAL && hp1 == h2.y && !root1;
        gsm_lean_leaf_term<LT>(c, SA, !da0, h2.x, a2.x, e0, G0);
        gsm_lean_leaf_term<LT>(c, SA, !da1, h2.y, a2.y, e1, G1);
    }
    double er[2], ew[2];
    gsm_lean_b_node<GENERAL, WANT_RATES, LEAF, STDW, MIX>(c, SB, i0, root0, pf0, fs0, h2.x, hp0, a2.x, ap0, rr.x, ss.x, kk.x, &er[0], &ew[0]);
    gsm_lean_b_node<GENERAL, WANT_RATES, LEAF, STDW, MIX>(c, SB, i0 + 1, root1, pf1, fs1, h2.y, hp1, a2.y, ap1, rr.y, ss.y, kk.y, &er[1],
                                                          &ew[1]);
    if (WANT_RATES) {
        if (orate2) orate2[i0 >> 1] = make_double2(er[0], er[1]);
        if (owsp2) owsp2[i0 >> 1] = make_double2(ew[0], ew[1]);
    }
}

// virtual chunks [v0, v1) of pass B, all of one LEAF class; `first` = first node of chunk v0, `lo`/`hi`: a thread's pair
// (i0, i0 + 1) is processed when lo <= i0 and i0 + 1 < hi.  PAIRS node pairs per thread and chunk (pair tid + p BLOCK).
template <bool GENERAL, bool WANT_RATES, int LEAF, bool STDW, int LT, int BLOCK, int PAIRS, int NST, bool MIX, int PROD, int XPD>
static __device__ __forceinline__ void lean_b_chunks(const GsmLeanC& c, GsmLeanAccA& SA, GsmLeanAccB& SB, LeanBCtx& x,
                                                     const LeanRing& rg, int v0, int v1, int first, int lo, int hi, int tid,
                                                     int32_t* sh_root_nroot, double2* __restrict__ orate2,
                                                     double2* __restrict__ owsp2, uint32_t& s, uint32_t& ph) {
    // s, ph: ring stage of chunk v0 and the parity of its full barrier; on return those of chunk v1 (one tree per CTA:
    // v0 % NST, (v0 / NST) & 1; the persistent kernel's ring runs on across trees)
    constexpr int HB = BLOCK * PAIRS;            // pairs per chunk
    constexpr uint32_t STAGE = HB * 64u;
    const bool counts = STDW || rg.nst != nullptr;
    int i0 = first + 2 * tid;
    // this thread's slot of stage 0 and the first mbarrier: kept in registers (opaque to the compiler, which otherwise
    // re-derives them from the kernel parameters in every iteration)
    uint32_t st0 = rg.a_ring + static_cast<uint32_t>(tid) * 16u, a_full = rg.a_full;
    uint32_t st1 = st0 - static_cast<uint32_t>(tid) * 8u;   // the 4-byte rows: pair q at + q * 8
    asm volatile("" : "+r"(st0), "+r"(a_full), "+r"(st1));
    const uint32_t a_empty = a_full + static_cast<uint32_t>(rg.a_empty - rg.a_full);
#pragma unroll 1
    for (int v = v0; v < v1; ++v) {
#ifdef GSM_PHASE_TIMING
        const long long tw_ = clock64();
#endif
        // (testing the barrier of chunk v + 1 before the body of chunk v, so that the latency of the test passes under the
        // body, measured 2.5 % slower: profiles/r02_tuning.md)
        mbar_wait_a(a_full + s * 8u, ph);
#ifdef GSM_PHASE_TIMING
        if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[7], static_cast<unsigned long long>(clock64() - tw_));   // waiting for a chunk
#endif
        double2 rr[PAIRS], ss[PAIRS], h2[PAIRS];
        int2 kk[PAIRS], pp[PAIRS];
#pragma unroll
        for (int p = 0; p < PAIRS; p++) {
            const uint32_t st = st0 + s * STAGE + static_cast<uint32_t>(p * BLOCK) * 16u;
            rr[p] = lds_v2f64(st);
            ss[p] = lds_v2f64(st + HB * 16u);
            h2[p] = make_double2(0.0, 0.0);
            if (LEAF == 1) h2[p] = lds_v2f64(st + HB * 32u);
            kk[p] = make_int2(-1, -1);    // without stub counts getNStubsOnBranch == -1 (Stubs:348-350)
            const uint32_t sq = st1 + s * STAGE + static_cast<uint32_t>(p * BLOCK) * 8u;
            if (counts) kk[p] = lds_v2s32(sq + HB * 48u);
            pp[p] = lds_v2s32(sq + HB * 56u);
        }
        __syncwarp();
        // PROD == 2 (persistent kernel): the copy warp sleeps on a named barrier per stage (a polling warp takes issue slots
        // from the consumers); the mbarrier carries the ordering, the named barrier only wakes the sleeper
#if GSM_PK_TMA_DOORBELL
        if (PROD == 2) asm volatile("bar.arrive %0, %1;" ::"r"(GSM_PK_BAR_STAGE + s), "n"(BLOCK + 32) : "memory");
#endif
        if ((tid & 31) == 0) {
            mbar_arrive_a(a_empty + s * 8u);
            // Chunk v + STAGES - 1 goes into the stage chunk v - 1 came from, once every warp has taken its pairs of v - 1.
            // PROD: the producer warp of the CTA does that (lean_producer).  Else the requesting duty rotates over the warps,
            // so that no warp falls behind the others because of it.
            if (!PROD && (tid >> 5) == (v & (BLOCK / 32 - 1)) && v >= 1 && v + NST - 1 < rg.nvc) {
                const uint32_t sp = s == 0u ? NST - 1u : s - 1u;
                mbar_wait_a(a_empty + sp * 8u, (s == 0u ? ph ^ 1u : ph));
                ring_issue<2 * HB>(rg, v + NST - 1, sp);
            }
        }
#pragma unroll
        for (int p = 0; p < PAIRS; p++) {
            const int ip = i0 + 2 * p * BLOCK;
            const bool on = ip >= lo && ip + 1 < hi;
            lean_b_pair<GENERAL, WANT_RATES, LEAF, STDW, LT, MIX, XPD>(c, SA, SB, x, a_full, ip, on, rr[p], ss[p], h2[p], kk[p], pp[p],
                                                                       sh_root_nroot, orate2, owsp2);
        }
        i0 += 2 * HB;
        if (++s == NST) {
            s = 0u;
            ph ^= 1u;
        }
    }
}

// main loop of pass B: leaf chunks (pairs below L), internal chunks (pairs from L on, up to n - 1).  The first three
// chunks were requested in phase 0.
template <bool GENERAL, bool WANT_RATES, bool STDW, int LT, int BLOCK, int PAIRS, int NST, bool MIX, int PROD, int XPD>
static __device__ __forceinline__ void lean_pass_b(const GsmLeanC& c, GsmLeanAccA& SA, GsmLeanAccB& SB, LeanBCtx& x, const LeanRing& rg,
                                                   int n, int tid, int32_t* sh_root_nroot, double2* __restrict__ orate2,
                                                   double2* __restrict__ owsp2, uint32_t& s, uint32_t& ph) {
    lean_b_chunks<GENERAL, WANT_RATES, 1, STDW, LT, BLOCK, PAIRS, NST, MIX, PROD, XPD>(c, SA, SB, x, rg, 0, rg.nlc, 0, 0, rg.L, tid, sh_root_nroot, orate2,
                                                                                  owsp2, s, ph);
    lean_b_chunks<GENERAL, WANT_RATES, 0, STDW, LT, BLOCK, PAIRS, NST, MIX, PROD, XPD>(c, SA, SB, x, rg, rg.nlc, rg.nvc, rg.L4, rg.L, n - 1, tid,
                                                                                  sh_root_nroot, orate2, owsp2, s, ph);
}

// MIXK: the mixture-mode spike prior (BSP:102-172) is a kernel of its own, so that its body does not weigh on the register
// allocation of the known-n kernel.
// PROD: a second warpgroup of four warps holds the producer warp, which requests every chunk after the first NST ones, so
// that the four consumer warps never run request code.  setmaxnreg works on whole warpgroups (measured: it never returns
// for a warpgroup of one warp, and register allocation per warp has a granularity that makes 5 warps x 136 registers
// take the room of 5 x 144): the kernel starts at 80 registers (3 CTAs x 256 threads), the producer warpgroup gives
// 56 x 128 back and the consumer warpgroup takes them.
#define GSM_PROD_CONSUMER_REGS 136
#define GSM_PROD_PRODUCER_REGS 24
#define GSM_PROD_THREADS 128
#define GSM_LEAN_BOUNDS(B, M, P) __launch_bounds__((B) + ((P) ? GSM_PROD_THREADS : 0), M)
template <int BLOCK, int MINB, bool WANT_RATES, bool MIXK, bool PROD = false>
__global__ void GSM_LEAN_BOUNDS(BLOCK, MINB, PROD) gsm_lean_kernel(const GsmBatch b, GsmLeanRow* __restrict__ rows,
                                                               int32_t* __restrict__ redo_count,
                                                               int32_t* __restrict__ redo_count_next,
                                                               int32_t* __restrict__ redo_list, double* __restrict__ out_rates,
                                                               double* __restrict__ out_wspikes) {
    constexpr int NWARP = BLOCK / 32;
    constexpr int PAIRS = gsm_lean_pairs(BLOCK, MINB);   // node pairs per thread and chunk
    constexpr int HB = BLOCK * PAIRS;                    // pairs per chunk
    constexpr int NST = gsm_lean_stages(BLOCK, MINB);    // ring stages
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ GsmLeanShared<NWARP> sh;
    // The two replicated tables sit on 8 KB boundaries of the shared window (entry offsets are OR-ed into their addresses).
    // The gap in front of them holds the small per-tree arrays (Gamma tables, fake flags) when it is large enough; the host
    // sizes the dynamic segment from the same rule (gsm_lean_layout), knowing where it starts (probe launch: rows == nullptr).
    const uint32_t a_dyn = smem_u32(smem_raw);
    if (rows == nullptr) {
        if (blockIdx.x == 0 && threadIdx.x == 0) redo_count[0] = static_cast<int32_t>(a_dyn);
        return;
    }
    const int S = static_cast<int>(b.stride);
    const int SI = gsm_lean_si(b.n_nodes);     // slots of the internal arrays: n - L4 <= (n + 1) / 2 + 3, + the pair padding
    const GsmLeanLayout lay = gsm_lean_layout(a_dyn, SI, NST * HB * 64);
    uint64_t* sh_logtab = reinterpret_cast<uint64_t*>(smem_raw + lay.logtab);
    uint64_t* sh_exptab = reinterpret_cast<uint64_t*>(smem_raw + lay.exptab);
    double* sh_hs = reinterpret_cast<double*>(smem_raw + lay.arrays);
    double* sh_as = sh_hs + SI;
    unsigned char* sh_ring = smem_raw + lay.ring;   // 16-byte aligned
    GsmLogEntry* sh_mtab = reinterpret_cast<GsmLogEntry*>(smem_raw + lay.mtab);
    GsmLeanN* sh_ntab = reinterpret_cast<GsmLeanN*>(smem_raw + lay.ntab);
    uint8_t* sh_fk = smem_raw + lay.fk;
    GsmTreeConsts& K = sh.K;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t t = blockIdx.x;
    // barriers of the consumer warps once the producer warp has gone its own way (named barrier 2)
    auto csync = [&]() {
        if (PROD) asm volatile("bar.sync 2, %0;" ::"n"(BLOCK) : "memory");
        else __syncthreads();
    };
    auto csync_or = [&](bool pred) -> bool {
        if (PROD) {
            uint32_t r;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, 2, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(r)
                         : "r"(pred ? 1u : 0u), "n"(BLOCK)
                         : "memory");
            return r != 0u;
        }
        return __syncthreads_or(pred ? 1 : 0) != 0;
    };
    GSM_PHASE_START();
    if (t == 0 && tid == 0) *redo_count_next = 0;   // counter of the next evaluation on this handle (ping-pong)
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    uint32_t dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    if (lay.total > dyn_bytes) n = 0;   // the launch did not provide this layout (cannot happen with launch_lean): generic path
    if (!((n & 1) && n >= 3 && n <= GSM_LEAN_MAX_NODES)) {   // not a binary BEAST tree the lean path covers (incl. n <= 0)
        if (tid == 0) lean_redo(redo_count, redo_list, rows, t);
        return;
    }
    const int64_t row = t * S;
    const int L = (n + 1) >> 1;
    const int L4 = L & ~3;

    LeanRing rg;
    rg.a_ring = smem_u32(sh_ring);
    rg.a_full = smem_u32(&sh.full[0]);
    rg.a_empty = smem_u32(&sh.empty[0]);
    rg.rate = b.rate + row;
    rg.spike = b.spike + row;
    rg.height = b.height + row;
    rg.nst = b.stub_mode != GSM_STUBS_NOSTUB ? b.nstubs + row : nullptr;
    rg.parent = b.parent + row;
    rg.L = L;
    rg.L4 = L4;
    rg.nlc = ((L >> 1) + HB - 1) / HB;                       // pairs below L: L / 2
    rg.nvc = rg.nlc + ((n - 1 - L4 + 1) / 2 + HB - 1) / HB;  // pairs (i0, i0 + 1), i0 = L4, L4 + 2, ... with i0 + 1 < n - 1
    rg.end_leaf = (L + 3) & ~3;
    rg.end_int = (n - 1 + 3) & ~3;
    const int nvc = rg.nvc;

    if (PROD) {
        if (warp >= NWARP) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GSM_PROD_PRODUCER_REGS));
            __syncthreads();   // barriers initialised
            __syncthreads();   // per-tree scalars in shared memory, first NST chunks requested
            // the producer: chunk v goes into stage v % NST once every consumer warp has taken its pairs of chunk v - NST
            // (the consumers leave without a pass B when the per-tree constants send the tree to the generic path)
            if (warp == NWARP && lane == 0 && gsm_lean_consts_ok(K) && ((K.spike_known_n == 0) == MIXK)) {
                uint32_t s = 0u, ph = 0u;   // stage and parity of its previous use
#pragma unroll 1
                for (int v = NST; v < nvc; ++v) {
                    mbar_wait_a(rg.a_empty + s * 8u, ph);
                    ring_issue<2 * HB>(rg, v, s);
                    if (++s == NST) {
                        s = 0u;
                        ph ^= 1u;
                    }
                }
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GSM_PROD_CONSUMER_REGS));
    }
    // global loads whose latency would otherwise sit between the first two barriers: issued before anything else
    int32_t parent_last = 0;
    double prm_v[GSM_NPARAM];
    int32_t use_v = 1;
    if (tid == 1) parent_last = __ldg(b.parent + row + n - 1);
    if (warp == NWARP - 1 && lane == 0) {
#pragma unroll
        for (int k = 0; k < GSM_NPARAM; k++) prm_v[k] = __ldg(b.params + t * GSM_NPARAM + k);
        if (b.use_spike) use_v = static_cast<int32_t>(__ldg(b.use_spike + t));
    }
    // ---- phase 0 ----
    if (tid == 0) {
        mbar_init(&sh.mbar, 1);
#pragma unroll
        for (int k = 0; k < NST; k++) {
            mbar_init(&sh.full[k], 1);
            mbar_init(&sh.empty[k], NWARP);
        }
        fence_mbar_init();
        sh.root = -1;
        sh.nroot = 0;
        sh.bad = 0;
        sh.pa_next = 0;
        sh.xp.i0 = -1;
    }
    __syncthreads();
    // requests spread over the warps (a bulk copy costs its issuing thread a few hundred cycles)
    if (lane == 0) {
        if (warp == 0) {
            const uint32_t bytes_h = (static_cast<uint32_t>(n - L4) * 8u + 15u) & ~15u;
            mbar_arrive_expect_tx(&sh.mbar, bytes_h + static_cast<uint32_t>(sizeof(GsmLeanBlob)));
            tma_bulk_g2s(sh_exptab, g_lean_blob.exptab, GSM_T6E_WORDS * 8u, &sh.mbar);
            tma_bulk_g2s(sh_hs, b.height + row + L4, bytes_h, &sh.mbar);
            tma_bulk_g2s(sh_logtab, g_lean_blob.logtab, GSM_T6_WORDS * 8u, &sh.mbar);
        } else if (warp <= NST && warp != NWARP - 1) {
            if (warp - 1 < nvc) ring_issue<2 * HB>(rg, warp - 1, static_cast<uint32_t>(warp - 1));
        }
    } else if (tid == 1) {
        sh.root_last = parent_last < 0 ? 1 : 0;
    }
    // The tail nodes (node n-1 and, when L is odd, L-1 and L) are evaluated by three lanes after pass B from rows read
    // straight from global memory: requested now, as asynchronous copies, so that their latency is long gone by then.
    const bool tail_lane = warp == (NWARP > 1 ? 1 : 0) && lane >= 29;
    // (lanes 27-28 of the same warp take the root's pair when it is handed over: slots 3-4; nothing of that is kept in
    // registers across pass B)
#define GSM_TSLOT(lane_) ((lane_) >= 29 ? (lane_) - 29 : (lane_) - 24)
    int tail_node = -1;
    if (tail_lane) {
        if (lane == 31) tail_node = n - 1;
        else if (L & 1) tail_node = lane == 29 ? L - 1 : L;
        if (tail_node >= 0) {
            const int k = GSM_TSLOT(lane);
            cp_async_8(smem_u32(&sh.tail_d[k][0]), b.height + row + tail_node);
            cp_async_8(smem_u32(&sh.tail_d[k][1]), b.rate + row + tail_node);
            cp_async_8(smem_u32(&sh.tail_d[k][2]), b.spike + row + tail_node);
            cp_async_4(smem_u32(&sh.tail_i[k][0]), b.parent + row + tail_node);
            if (b.stub_mode != GSM_STUBS_NOSTUB) cp_async_4(smem_u32(&sh.tail_i[k][1]), b.nstubs + row + tail_node);
        }
    }
    // before any exit from here on: the bulk copies target this CTA's shared memory and must have landed
    auto drain = [&]() {
#pragma unroll
        for (int k = 0; k < NST; k++)
            if (k < nvc) mbar_wait_a(rg.a_full + k * 8u, 0u);
        asm volatile("cp.async.wait_all;" ::: "memory");
    };
    // The lane that computes the per-tree scalars does that first: when its warp also has a chunk to request (four-warp
    // CTAs), the request follows the barrier (that chunk is the last of the first three to be consumed).
    if (warp == NWARP - 1 && lane == 0) gsm_setup_scalars(K, prm_v, use_v, b.flags, b.stub_mode);
    __syncthreads();
    if (warp == NWARP - 1 && lane == 0 && warp <= NST && warp - 1 < nvc)
        ring_issue<2 * HB>(rg, warp - 1, static_cast<uint32_t>(warp - 1));
    GSM_PHASE_MARK(0);   // barrier init, TMA requests, per-tree scalars
    mbar_wait(&sh.mbar, 0);
    GSM_PHASE_MARK(1);   // internal heights + tables landed
    if (!gsm_lean_consts_ok(K)) {   // invalid state (lambda <= mu, shape <= 0, ...): the generic path reproduces the reference
        drain();
        if (tid == 0) lean_redo(redo_count, redo_list, rows, t);
        return;
    }
    GsmLeanC c;
    gsm_lean_consts(K, c, n);
    if (c.mix != MIXK) {   // cannot happen with launch_lean: the kernel flavour follows the wiring
        drain();
        if (tid == 0) lean_redo(redo_count, redo_list, rows, t);
        return;
    }
    c.t6.log_ar = smem_u32(sh_logtab) | (static_cast<uint32_t>(lane & 15) * 8u);
    c.t6.exp_ar = smem_u32(sh_exptab) | (static_cast<uint32_t>(lane & 7) * 8u);
    c.mtab = sh_mtab;
    c.ntab = sh_ntab;
    c.mixtab = reinterpret_cast<const GsmLogEntry*>(sh_ntab);   // mixture mode has no stub counts: the two tables share the space
    // shared-window addresses of the internal arrays, biased so that node i is at + i * 8
    LeanBCtx x;
    x.a_hs = smem_u32(sh_hs) - static_cast<uint32_t>(L4) * 8u;
    x.a_as = smem_u32(sh_as) - static_cast<uint32_t>(L4) * 8u;
    x.a_fk = smem_u32(sh_fk) - static_cast<uint32_t>(L4);
    x.a_hL = x.a_hs + static_cast<uint32_t>(L) * 8u;
    x.a_aL = x.a_as + static_cast<uint32_t>(L) * 8u;
    x.a_fL = x.a_fk + static_cast<uint32_t>(L);
    x.chk_p = 0u;
    x.nL = static_cast<uint32_t>(n - L);
    static_assert(offsetof(GsmLeanXPair, i) == 8 && offsetof(GsmLeanXPair, d) == 24, "GsmLeanXPair layout (lean_b_pair)");

    // ---- phase 1: Gamma tables (64 lanes), fake flags ----
    // Sampled ancestors are expected where the tree prior generates them (psi > 0, origin / root conditioning): those trees
    // are scanned.  Elsewhere a zero-length leaf is caught by the length check of the CLEAN bodies or, in a GENERAL body
    // (root not last, tail nodes), by c.sa_unscanned, and the tree goes to the generic path.
    const bool detect = !(c.leaf_term == 0 && K.psi == 0.0);
    c.sa_unscanned = !detect;
    if (tid < GSM_LEAN_MTAB_N - 1) {
        // mtab[m] = {shape m - 1, log shape - logGamma(shape m)}, ntab[m - 1] = {log shape - logGamma(shape m), -log((m - 1)!)}
        // (BSP:97-99 with the Lanczos logGamma of commons-math; STP:567)
        const int m = tid + 1;
        const double shape = K.shape;
        GsmLogEntry e;
        GsmLeanN f;
        e.invc = shape * m - 1;
        e.logc = gsm_log6_hi(c.t6, shape) - lean_log_gamma(c.t6, shape * m);
        sh_mtab[m] = e;
        if (!MIXK) {
            f.gcon = e.logc;
            f.nlf = -g_logfact[m - 1];
            f.w = e.invc;
            f.nd = static_cast<double>(m - 1);
            sh_ntab[m - 1] = f;
        } else {
            // g[m] = Gamma(shape m) / Gamma(shape (m + 1)); [m - 1] = {g[m] / m, 1 / m}, [64 + m] = {g[m] / (m + 1), 1 / (m + 1)}
            const double lg_m = gsm_log6_hi(c.t6, shape) - e.logc;   // logGamma(shape m)
            const double g = gsm_exp6(c.t6, lg_m - lean_log_gamma(c.t6, shape * (m + 1)));
            GsmLogEntry* mt = reinterpret_cast<GsmLogEntry*>(sh_ntab);
            GsmLogEntry r;
            r.logc = lean_rcp(static_cast<double>(m));
            r.invc = g * r.logc;
            mt[m - 1] = r;
            if (m < GSM_LEAN_MAX_NST) {
                r.logc = lean_rcp(static_cast<double>(m + 1));
                r.invc = g * r.logc;
                mt[GSM_LEAN_MAX_NST + 1 + m] = r;
            }
            if (tid == 0) {   // [64 + 0]: only its 1 / (k + 1) is used (the k = 0 term of a fake node's child is zero)
                r.logc = 1.0;
                r.invc = 0.0;
                mt[GSM_LEAN_MAX_NST + 1] = r;
            }
        }
        if (tid == 0) {
            e.invc = -1.0;
            e.logc = 0.0;
            sh_mtab[0] = e;
        }
    }
    bool anyda = false;
    if (detect) {
        // Node.isDirectAncestor: a leaf at its parent's height -> the parent isFake.  Leaf heights and parents straight from
        // global memory (the ring will read them again from L2).
        for (int k = tid; k < (SI + 3) >> 2; k += BLOCK) reinterpret_cast<uint32_t*>(sh_fk)[k] = 0u;
        csync();
        const double* hrow = b.height + row;
        const int32_t* prow = b.parent + row;
        // GSM_SCAN_U leaves per thread and round: all their global loads are issued before the first one is used
#pragma unroll 1
        for (int i0 = tid; i0 < L; i0 += GSM_SCAN_U * BLOCK) {
            int32_t p[GSM_SCAN_U];
            double h[GSM_SCAN_U];
#pragma unroll
            for (int u = 0; u < GSM_SCAN_U; u++) {
                const int i = i0 + u * BLOCK;
                p[u] = i < L ? __ldg(prow + i) : -1;
                h[u] = i < L ? __ldg(hrow + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < GSM_SCAN_U; u++) {
                if (static_cast<uint32_t>(p[u] - L) < x.nL && h[u] == lds_f64(x.a_hs + p[u] * 8)) {
                    reinterpret_cast<uint8_t*>(sh_fk)[p[u] - L4] = 1;
                    anyda = true;
                }
            }
        }
    }
    // One tail node, evaluated by its lane with the GENERAL flavour; the combined sums go to shared memory.
    //   early  (right here, for a tree that is certainly ordinary: no sampled-ancestor scan, root = node n-1): the lane
    //          computes the aux values it needs itself, while the other warps run pass A
    //   late   (after pass B): aux and the fake flags come from shared memory
    auto tail_eval = [&](bool early, bool gen) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        const int i = tail_node;
        const int k = GSM_TSLOT(lane);
        const double* td = k < 3 ? sh.tail_d[k] : sh.xp.d[k - 3];
        const int32_t* ti = k < 3 ? sh.tail_i[k] : sh.xp.i[k - 3];
        GsmLeanAccB ST;
        gsm_lean_acc_b_zero(ST);
        GsmLeanAccA STA;
        gsm_lean_acc_a_zero(STA);
        uint32_t chk_p = 0u;
        const double h = td[0];
        int32_t p = ti[0];
        const bool root = p < 0;
        if (root) {
            atomicAdd(&sh.nroot, 1);
            sh.root = i;
            if (i < L) ST.chk = 0xffffffffu;
            p = i >= L ? i : L;
        }
        chk_p = static_cast<uint32_t>(p - L);
        if (!(static_cast<uint32_t>(p - L) < x.nL)) p = L;
        double e, G, aux, ep, Gp;
        if (i < L || early) aux = gsm_lean_aux(c, STA, h, e, G);
        else aux = sh_as[i - L4];
        double hp = sh_hs[p - L4], auxp;
        if (root) { hp = h; auxp = aux; }
        else if (early) auxp = gsm_lean_aux(c, STA, hp, ep, Gp);
        else auxp = sh_as[p - L4];
        const bool pf = gen && sh_fk[p - L4] != 0, fs = gen && i >= L && sh_fk[i - L4] != 0;
        if (i < L) {
            const bool da = hp == h && !root;
            if (c.leaf_term == 1) gsm_lean_leaf_term<1>(c, STA, !da, h, aux, e, G);
            else if (c.leaf_term == 2) gsm_lean_leaf_term<2>(c, STA, !da, h, aux, e, G);
        }
        int32_t nst = -1;
        if (c.use_counts) nst = (i < n - 1 || b.stub_mode == GSM_STUBS_EXACT) ? ti[1] : 0;
        double er, ew;
        gsm_lean_b_node<true, WANT_RATES, 2, false, MIXK>(c, ST, i, root, pf, fs, h, hp, aux, auxp, td[1], td[2], nst, &er,
                                                          &ew);
        if (WANT_RATES) {
            if (out_rates) out_rates[row + i] = er;
            if (out_wspikes) out_wspikes[row + i] = ew;
        }
        GsmLean6 t6;
        gsm_lean_combine<true>(K, c, STA, ST, t6);
        double* o = sh.tail_s6[k];
        o[0] = t6.tree; o[1] = t6.stub; o[2] = t6.spike; o[3] = t6.rate; o[4] = t6.burst; o[5] = t6.dist;
        int32_t* oc = sh.tail_c[k];
        oc[0] = STA.n_leafterm; oc[1] = STA.n_plain; oc[2] = ST.n_fake; oc[3] = ST.n_extant; oc[4] = ST.n_extinct; oc[5] = ST.n_da;
        uint32_t* ou = sh.tail_u[k];
        ou[0] = gsm_umax(STA.chk, ST.chk); ou[1] = ST.chk_n; ou[2] = STA.hmax; ou[3] = chk_p;
    };
    const bool early_tail = !detect && sh.root_last != 0;
    // The early tail lanes (warp 1) read Gamma table entries written by warp 0: the two Gamma warps meet on a named
    // barrier of their own first (they finish their entries at the same time; the pass A warps are not held up).
    if (early_tail && NWARP > 1 && warp < 2) asm volatile("bar.sync 1, 64;" ::: "memory");
    if (tail_lane && tail_node >= 0 && early_tail) tail_eval(true, false);
    GSM_PHASE_MARK(2);   // Gamma tables, sampled-ancestor scan, tail nodes

    // ---- phase 2: pass A: aux of the internal nodes ----
    GsmLeanAccA SA;
    gsm_lean_acc_a_zero(SA);
    {
        const int npa = (n - L4 + 1) >> 1;
        const uint32_t a_h0 = smem_u32(sh_hs), a_a0 = smem_u32(sh_as);
        auto pair_aux = [&](int j) {
            double2 h2 = lds_v2f64(a_h0 + j * 16);
            if (L4 + 2 * j + 1 >= n) h2.y = h2.x;   // padding slot of the last pair
            double e0, G0, e1, G1;
            const double a0 = gsm_lean_aux(c, SA, h2.x, e0, G0);
            const double a1 = gsm_lean_aux(c, SA, h2.y, e1, G1);
            sts_v2f64(a_a0 + j * 16, a0, a1);
        };
        auto pair_aux2 = [&](int j, int j2) {   // two pairs = four nodes, staged (gsm_lean_aux4)
            const double2 ha = lds_v2f64(a_h0 + j * 16);
            double2 hb = lds_v2f64(a_h0 + j2 * 16);
            if (L4 + 2 * j2 + 1 >= n) hb.y = hb.x;   // padding slot of the last pair
            const double h4[4] = {ha.x, ha.y, hb.x, hb.y};
            double a4[4];
            gsm_lean_aux4(c, SA, h4, a4);
            sts_v2f64(a_a0 + j * 16, a4[0], a4[1]);
            sts_v2f64(a_a0 + j2 * 16, a4[2], a4[3]);
        };
        if (GSM_PA_DYN || BLOCK < 256) {
            // Warps take batches of 64 pairs (two per lane) from a shared counter: the two Gamma warps join when their table
            // entries are done.
#pragma unroll 1
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(&sh.pa_next, GSM_PA_BATCH);
                j = __shfl_sync(0xffffffffu, j, 0);
                if (j >= npa) break;
                j += lane;
#pragma unroll
                for (int u = 0; u < GSM_PA_BATCH / 64; u++, j += 64) {
                    if (j + 32 < npa) pair_aux2(j, j + 32);
                    else if (j < npa) pair_aux(j);
                }
            }
        } else {
            // the warps that did not compute Gamma table entries take all of pass A, two pairs (four nodes) at a time
            constexpr int NA = BLOCK - 64;
#pragma unroll 1
            for (int j = tid - 64; j < npa && j >= 0; j += 2 * NA) {
                const int j2 = j + NA;
                if (j2 < npa) pair_aux2(j, j2);
                else pair_aux(j);
            }
        }
    }
    {
        GsmLeanAccA S0;
        gsm_lean_acc_a_zero(S0);
        x.auxz = gsm_lean_aux(c, S0, 0.0, x.ez, x.Gz);
    }
    GSM_PHASE_MARK(3);   // pass A (thread 0's share)
    const bool any = csync_or(anyda);
    GSM_PHASE_MARK(4);   // waiting for the other warps
    // GENERAL bodies for trees with sampled ancestors; an ordinary tree whose root is not node n-1 keeps the CLEAN bodies and
    // hands the pair with the root to the tail lanes (lean_b_pair)
    const bool general = any;

    // ---- phase 3: pass B ----
    GSM_WARP_T0();
    GsmLeanAccB SB;
    gsm_lean_acc_b_zero(SB);
    double2* orate2 = (WANT_RATES && out_rates) ? reinterpret_cast<double2*>(out_rates + row) : nullptr;
    double2* owsp2 = (WANT_RATES && out_wspikes) ? reinterpret_cast<double2*>(out_wspikes + row) : nullptr;
    const bool stdw = gsm_lean_std_wiring(c);
    int32_t* rn = &sh.root;   // {root, nroot}
    uint32_t ring_s = 0u, ring_ph = 0u;
    constexpr int XPD = static_cast<int>(offsetof(GsmLeanShared<NWARP>, xp)) - static_cast<int>(offsetof(GsmLeanShared<NWARP>, full));
    static_assert(XPD != 0, "hand-over slots");
#define GSM_PASS_B(G, W, LT, M) lean_pass_b<G, WANT_RATES, W, LT, BLOCK, PAIRS, NST, M, PROD, XPD>(c, SA, SB, x, rg, n, tid, rn, orate2, owsp2, ring_s, ring_ph)
#define GSM_PASS_B3(G, LT) { if constexpr (MIXK) GSM_PASS_B(G, false, LT, true); else { if (stdw) GSM_PASS_B(G, true, LT, false); else GSM_PASS_B(G, false, LT, false); } }
    if (general) {
        if (!detect) for (int k = tid; k < (SI + 3) >> 2; k += BLOCK) reinterpret_cast<uint32_t*>(sh_fk)[k] = 0u;
        if (!detect) csync();
        if (c.leaf_term == 0) GSM_PASS_B3(true, 0)
        else if (c.leaf_term == 1) GSM_PASS_B3(true, 1)
        else GSM_PASS_B3(true, 2)
    } else {
        if (c.leaf_term == 0) GSM_PASS_B3(false, 0)
        else if (c.leaf_term == 1) GSM_PASS_B3(false, 1)
        else GSM_PASS_B3(false, 2)
    }
#undef GSM_PASS_B3
#undef GSM_PASS_B
    GSM_WARP_T1(0);
    // ---- tail: node n-1 (no stub count in COUNTS mode, Stubs:354; the root of an ordinary tree) and, when L is odd, the
    // pair (L-1, L) that straddles the leaf / internal boundary ----
    const bool xpair_lane = warp == (NWARP > 1 ? 1 : 0) && (lane == 27 || lane == 28);
    if (!general && !sh.root_last) {   // the root was met in pass B
        csync();   // the hand-over slots were written by whichever thread met it
        if (xpair_lane && sh.xp.i0 >= 0) tail_node = sh.xp.i0 + (lane - 27);
    }
    const bool tail_on = (tail_lane || xpair_lane) && tail_node >= 0;
    if (tail_on && !early_tail) tail_eval(false, general);
    GsmLean6 s6;
    if (general) {
        gsm_lean_combine<true>(K, c, SA, SB, s6);
    } else {
        gsm_lean_combine<false>(K, c, SA, SB, s6);
    }
    int32_t t_leafterm = 0, t_plain = 0, t_fake = 0, t_extant = 0, t_extinct = 0, t_da = 0;
    uint32_t t_chk = 0u, t_chk_n = 0u, t_hmax = 0u;
    if (tail_on) {
        const int k = GSM_TSLOT(lane);
        const double* o = sh.tail_s6[k];
        s6.tree += o[0]; s6.stub += o[1]; s6.spike += o[2]; s6.rate += o[3]; s6.burst += o[4]; s6.dist += o[5];
        const int32_t* oc = sh.tail_c[k];
        t_leafterm = oc[0]; t_plain = oc[1]; t_fake = oc[2]; t_extant = oc[3]; t_extinct = oc[4]; t_da = oc[5];
        const uint32_t* ou = sh.tail_u[k];
        t_chk = ou[0]; t_chk_n = ou[1]; t_hmax = ou[2];
        x.chk_p = gsm_umax(x.chk_p, ou[3]);
    }

    // ---- phase 4: reductions (fixed order) ----
    GSM_WARP_T1(1);
    GSM_PHASE_MARK(5);   // pass B (thread 0's share)
    const double v0 = warp_sum(s6.tree), v1 = warp_sum(s6.stub), v2 = warp_sum(s6.spike), v3 = warp_sum(s6.rate),
                 v4 = warp_sum(s6.burst), v5 = warp_sum(s6.dist);
    const int32_t c0 = 0, c1 = __reduce_add_sync(0xffffffffu, SA.n_leafterm + t_leafterm),
                  c2 = __reduce_add_sync(0xffffffffu, SA.n_plain + t_plain), c3 = __reduce_add_sync(0xffffffffu, SB.n_fake + t_fake),
                  c4 = __reduce_add_sync(0xffffffffu, SB.n_extant + t_extant),
                  c5 = __reduce_add_sync(0xffffffffu, SB.n_extinct + t_extinct), c6 = __reduce_add_sync(0xffffffffu, SB.n_da + t_da);
    const uint32_t u0 = warp_max_u32(gsm_umax(gsm_umax(SA.chk, SB.chk), t_chk)),
                   u1 = warp_max_u32(gsm_umax(SB.chk_n, t_chk_n)), u2 = warp_max_u32(SB.chk_len),
                   u3 = warp_max_u32(gsm_umax(SA.hmax, t_hmax)), u4 = warp_max_u32(x.chk_p);
    if (lane == 0) {
        sh.red[0][warp] = v0; sh.red[1][warp] = v1; sh.red[2][warp] = v2;
        sh.red[3][warp] = v3; sh.red[4][warp] = v4; sh.red[5][warp] = v5;
        sh.redi[0][warp] = c0; sh.redi[1][warp] = c1; sh.redi[2][warp] = c2; sh.redi[3][warp] = c3;
        sh.redi[4][warp] = c4; sh.redi[5][warp] = c5; sh.redi[6][warp] = c6;
        sh.redu[0][warp] = u0; sh.redu[1][warp] = u1; sh.redu[2][warp] = u2; sh.redu[3][warp] = u3; sh.redu[4][warp] = u4;
    }
    csync();
    GSM_PHASE_MARK(6);   // waiting for the other warps (pass B)
    if (warp == 0) {
        // lanes 0..5 sum one double each over the warps (fixed order); lanes 8..14 the integers; lanes 16..20 the maxima
        double dv = 0.0;
        int32_t iv = 0;
        uint32_t uv = 0u;
        if (lane < 6) {
#pragma unroll
            for (int w = 0; w < NWARP; w++) dv += sh.red[lane][w];
        } else if (lane >= 8 && lane < 15) {
#pragma unroll
            for (int w = 0; w < NWARP; w++) iv += sh.redi[lane - 8][w];
        } else if (lane >= 16 && lane < 21) {
#pragma unroll
            for (int w = 0; w < NWARP; w++) uv = gsm_umax(uv, sh.redu[lane - 16][w]);
        }
        const uint32_t chk = __shfl_sync(0xffffffffu, uv, 16), chk_n = __shfl_sync(0xffffffffu, uv, 17),
                       chk_len = __shfl_sync(0xffffffffu, uv, 18), hmax = __shfl_sync(0xffffffffu, uv, 19),
                       chk_p = __shfl_sync(0xffffffffu, uv, 20);
        const int root = sh.root;
        bool bad = chk >= GSM_CHK_LIMIT || (c.use_counts && chk_n > static_cast<uint32_t>(GSM_LEAN_MAX_NST)) ||
                   (!general && chk_len >= GSM_CHK_LIMIT) || hmax >= 0x7ff00000u || chk_p >= x.nL || sh.nroot != 1 || root < L;
        if (!bad) bad = !(K.c1 * __hiloint2double(static_cast<int>(hmax) + 1, 0) < 700.0);   // exp(-c1 h) stays in range
        if (bad) {
            if (lane == 0) lean_redo(redo_count, redo_list, rows, t);
        } else {
            GsmLeanRow* r = rows + t;
            double* rd = reinterpret_cast<double*>(r);
            int32_t* ri = reinterpret_cast<int32_t*>(rd + 7);
            if (lane < 6) rd[lane] = dv;
            else if (lane == 6) rd[6] = sh_hs[root - L4];
            else if (lane >= 8 && lane < 15) {
                // redi order: n_nst, n_leafterm, n_plain, n_fake, n_extant, n_extinct, n_da  == row order
                ri[lane - 8] = iv;
            } else if (lane == 15) {
                ri[7] = (general && sh_fk[root - L4] != 0) ? 1 : 0;
                ri[8] = 1;
            }
        }
    }
}

#include "gsm_lean_pk.cuh"
#include "gsm_lean_warp.cuh"

// ---------------------------------------------------------------------------------------------------------
// BranchSpikePrior.getCumulativeProbs (BSP:315-389) for one branch; log-time only, one thread.
// ---------------------------------------------------------------------------------------------------------
__global__ void gsm_stub_cdf_kernel(const GsmBatch b, int64_t t, int32_t node, double* __restrict__ cdf, int32_t cap,
                                    int32_t* __restrict__ len_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    GsmTreeConsts K;
    gsm_setup_consts(K, b.params + t * GSM_NPARAM, b.use_spike ? static_cast<int32_t>(b.use_spike[t]) : 1, b.flags,
                     b.stub_mode);
    for (int m = 0; m <= GSM_LUT_M; m++) gsm_lut_entry(K, K.shape, m);
    const int64_t row = t * b.stride;
    const double h = b.height[row + node];
    const int32_t p = b.parent[row + node];
    const double hp = p < 0 ? h : b.height[row + p];
    double mu = 0.0;                                                        // STP:876-878
    if (hp > h) {                                                           // STP:749-769 through aux = log G
        const double a1 = gsm_log(fma(K.omc2, gsm_exp(-K.c1 * h), K.opc2));
        const double a0 = gsm_log(fma(K.omc2, gsm_exp(-K.c1 * hp), K.opc2));
        mu = fma(hp - h, K.kE, 2 * (a1 - a0));
    }
    const double x = b.spike[row + node];                                   // BSP:328
    const bool use_gamma = (b.flags & GSM_F_BSP_HAS_INDICATOR) && K.use_spike;   // BSP:347
    const double log_mu = log(mu);
    const double xs = x * K.shape, lxs = log(xs);
    int32_t k = 0, len = 0;
    double pcs = 0.0, psum = 0.0;
    while (pcs < GSM_MAX_CUM_SUM && k < GSM_K_CAP) {                        // BSP:334
        double branchP = 0.0;
        double pl = -mu + k * log_mu;                                       // BSP:339
        pl -= gsm_logfact(g_logfact, k);                                    // BSP:340
        const double pReal = exp(pl);
        if (use_gamma) {
            const double g = gsm_gamma_logdens(K, k + 1, x, xs, lxs);       // BSP:344-350
            if (g == -INFINITY || g != g) {
                if (pcs > 0) break;                                         // BSP:352
            } else {
                branchP += exp(pl + g);                                     // BSP:355
            }
        } else {
            branchP += pReal;                                               // BSP:361
        }
        pcs += pReal;
        psum += branchP;
        if (len < cap) cdf[len] = branchP;
        len++;
        k++;
    }
    const int32_t w = len < cap ? len : cap;
    double cum = 0.0;
    for (int32_t i = 0; i < w; i++) {                                       // BSP:374-387
        const double pr = cdf[i] / psum;
        cdf[i] = pr + cum;
        cum += pr;
    }
    *len_out = len;
}

// ---------------------------------------------------------------------------------------------------------
// Log-time stub sampling for whole trees (Stubs.sampleNStubs / sampleNStubsOnBranch, Stubs:389-458; the tree loggers call
// Stubs.getArrayValue for every node, Stubs:666-674): one CTA per selected tree, one thread per branch at a time.
//   gsm_stub_walk      the loop of BranchSpikePrior.getCumulativeProbs (BSP:334-370) for one branch
//   gsm_stub_len       pass 1: array length and normalising sum of every branch
//   gsm_stub_fill      pass 2: the normalised cumulative arrays (BSP:374-387), CSR
//   gsm_stub_choose    Randomizer.randomChoice(cf) for a caller-supplied uniform per branch (the RNG itself is BEAST core)
// ---------------------------------------------------------------------------------------------------------
struct GsmStubBranch {
    double mu, x;
    bool draws;   // false: root or Poisson mean 0 -> 0 stubs without a draw (Stubs:416-434)
};
static __device__ __forceinline__ GsmStubBranch gsm_stub_branch(const GsmBatch& b, const GsmTreeConsts& K, int64_t t, int32_t node) {
    GsmStubBranch r;
    const int64_t row = t * b.stride;
    const double h = b.height[row + node];
    const int32_t p = b.parent[row + node];
    const double hp = p < 0 ? h : b.height[row + p];
    r.mu = 0.0;                                                              // STP:876-878
    if (p >= 0 && hp > h) {                                                  // STP:749-769 through aux = log G
        const double a1 = gsm_log(fma(K.omc2, gsm_exp(-K.c1 * h), K.opc2));
        const double a0 = gsm_log(fma(K.omc2, gsm_exp(-K.c1 * hp), K.opc2));
        r.mu = fma(hp - h, K.kE, 2 * (a1 - a0));
    }
    r.x = b.spike[row + node];                                               // BSP:328
    r.draws = p >= 0 && r.mu != 0.0;
    return r;
}
// calls emit(k, branchP) for every entry of the Java list `probs`; returns its length, *psum = branchPSum
template <typename F>
static __device__ __forceinline__ int32_t gsm_stub_walk(const GsmTreeConsts& K, bool use_gamma, double mu, double x, double* psum, F&& emit) {
    const double log_mu = log(mu);
    const double xs = x * K.shape, lxs = log(xs);
    int32_t k = 0;
    double pcs = 0.0, sum = 0.0;
    while (pcs < GSM_MAX_CUM_SUM && k < GSM_K_CAP) {                         // BSP:334
        double branchP = 0.0;
        double pl = -mu + k * log_mu;                                        // BSP:339
        pl -= gsm_logfact(g_logfact, k);                                     // BSP:340
        const double pReal = exp(pl);
        if (use_gamma) {
            const double g = gsm_gamma_logdens(K, k + 1, x, xs, lxs);        // BSP:344-350
            if (g == -INFINITY || g != g) {
                if (pcs > 0) break;                                          // BSP:352
            } else {
                branchP += exp(pl + g);                                      // BSP:355
            }
        } else {
            branchP += pReal;                                                // BSP:361
        }
        pcs += pReal;
        sum += branchP;
        emit(k, branchP);
        k++;
    }
    *psum = sum;
    return k;
}
static __device__ __forceinline__ void gsm_stub_tree_consts(const GsmBatch& b, int64_t t, GsmTreeConsts& K) {
    if (threadIdx.x == 0) {
        gsm_setup_consts(K, b.params + t * GSM_NPARAM, b.use_spike ? static_cast<int32_t>(b.use_spike[t]) : 1, b.flags, b.stub_mode);
        for (int m = 0; m <= GSM_LUT_M; m++) gsm_lut_entry(K, K.shape, m);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) gsm_stub_len_kernel(const GsmBatch b, const int64_t* __restrict__ trees, int64_t* __restrict__ lens,
                                                           double* __restrict__ sums) {
    __shared__ GsmTreeConsts K;
    const int64_t t = trees ? trees[blockIdx.x] : blockIdx.x;
    gsm_stub_tree_consts(b, t, K);
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    const bool use_gamma = (b.flags & GSM_F_BSP_HAS_INDICATOR) && K.use_spike;   // BSP:347
    for (int32_t i = threadIdx.x; i < b.n_nodes; i += blockDim.x) {
        int64_t len = 0;
        double sum = 0.0;
        if (i < n) {
            const GsmStubBranch br = gsm_stub_branch(b, K, t, i);
            if (br.draws) len = gsm_stub_walk(K, use_gamma, br.mu, br.x, &sum, [](int32_t, double) {});
        }
        lens[static_cast<int64_t>(blockIdx.x) * b.n_nodes + i] = len;
        sums[static_cast<int64_t>(blockIdx.x) * b.n_nodes + i] = sum;
    }
}

__global__ void __launch_bounds__(128) gsm_stub_fill_kernel(const GsmBatch b, const int64_t* __restrict__ trees, const int64_t* __restrict__ ptr,
                                                            const double* __restrict__ sums, double* __restrict__ cdf, int64_t cap) {
    __shared__ GsmTreeConsts K;
    const int64_t t = trees ? trees[blockIdx.x] : blockIdx.x;
    gsm_stub_tree_consts(b, t, K);
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    const bool use_gamma = (b.flags & GSM_F_BSP_HAS_INDICATOR) && K.use_spike;
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t item = static_cast<int64_t>(blockIdx.x) * b.n_nodes + i;
        const int64_t o = ptr[item];
        if (ptr[item + 1] == o) continue;
        const GsmStubBranch br = gsm_stub_branch(b, K, t, i);
        const double total = sums[item];
        double cum = 0.0, dummy;
        gsm_stub_walk(K, use_gamma, br.mu, br.x, &dummy, [&](int32_t k, double branchP) {
            const double pr = branchP / total;                               // BSP:374-377
            if (o + k < cap) cdf[o + k] = pr + cum;                          // BSP:381-387
            cum += pr;
        });
    }
}

// Randomizer.randomChoice(cf) (BEAST.base 2.7.7, recalled): s = 0 if U <= cf[0]; else the first s >= 1 with
// cf[s-1] < U <= cf[s]; cf.length when no entry qualifies (e.g. a 0/0 array).
__global__ void __launch_bounds__(128) gsm_stub_choose_kernel(const GsmBatch b, const int64_t* __restrict__ trees, const double* __restrict__ u,
                                                              int32_t* __restrict__ out) {
    __shared__ GsmTreeConsts K;
    const int64_t t = trees ? trees[blockIdx.x] : blockIdx.x;
    gsm_stub_tree_consts(b, t, K);
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    const bool use_gamma = (b.flags & GSM_F_BSP_HAS_INDICATOR) && K.use_spike;
    for (int32_t i = threadIdx.x; i < b.n_nodes; i += blockDim.x) {
        const int64_t item = static_cast<int64_t>(blockIdx.x) * b.n_nodes + i;
        int32_t pick = 0;
        if (b.stub_mode != GSM_STUBS_NOSTUB) pick = -1;                      // Stubs:412
        else if (i < n) {
            const GsmStubBranch br = gsm_stub_branch(b, K, t, i);
            if (br.draws) {
                double total = 0.0, dummy;
                const int32_t len = gsm_stub_walk(K, use_gamma, br.mu, br.x, &total, [](int32_t, double) {});
                const double U = u[item];
                double cum = 0.0, prev = 0.0;
                int32_t s = len;
                gsm_stub_walk(K, use_gamma, br.mu, br.x, &dummy, [&](int32_t k, double branchP) {
                    const double pr = branchP / total;
                    const double cf = pr + cum;
                    cum += pr;
                    if (s == len && U <= cf && (k == 0 || U > prev)) s = k;
                    prev = cf;
                });
                pick = s;
            }
        }
        out[item] = pick;
    }
}

int gsm_launch_stub_len(const GsmBatch& b, int64_t n_sel, const int64_t* d_trees, int64_t* d_lens, double* d_sums, cudaStream_t stream) {
    if (n_sel <= 0) return 0;
    gsm_stub_len_kernel<<<static_cast<unsigned>(n_sel), 128, 0, stream>>>(b, d_trees, d_lens, d_sums);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -static_cast<int>(e);
}
int gsm_launch_stub_fill(const GsmBatch& b, int64_t n_sel, const int64_t* d_trees, const int64_t* d_ptr, const double* d_sums, double* d_cdf,
                         int64_t cap, cudaStream_t stream) {
    if (n_sel <= 0) return 0;
    gsm_stub_fill_kernel<<<static_cast<unsigned>(n_sel), 128, 0, stream>>>(b, d_trees, d_ptr, d_sums, d_cdf, cap);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -static_cast<int>(e);
}
int gsm_launch_stub_choose(const GsmBatch& b, int64_t n_sel, const int64_t* d_trees, const double* d_u, int32_t* d_out, cudaStream_t stream) {
    if (n_sel <= 0) return 0;
    gsm_stub_choose_kernel<<<static_cast<unsigned>(n_sel), 128, 0, stream>>>(b, d_trees, d_u, d_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -static_cast<int>(e);
}
// exclusive prefix sum of the array lengths -> CSR pointers (n + 1 entries); single CTA, log-time path
__global__ void __launch_bounds__(1024) gsm_scan_kernel(const int64_t* __restrict__ lens, int64_t* __restrict__ ptr, int64_t n) {
    __shared__ int64_t part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (n + 1023) / 1024, lo = tid * per, hi = lo + per < n ? lo + per : n;
    int64_t s = 0;
    for (int64_t i = lo; i < hi; i++) s += lens[i];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int64_t run = 0;
        for (int k = 0; k < 1024; k++) { const int64_t v = part[k]; part[k] = run; run += v; }
        ptr[n] = run;
    }
    __syncthreads();
    int64_t run = part[tid];
    for (int64_t i = lo; i < hi; i++) { ptr[i] = run; run += lens[i]; }
}
int gsm_launch_scan(const int64_t* d_lens, int64_t* d_ptr, int64_t n, cudaStream_t stream) {
    gsm_scan_kernel<<<1, 1024, 0, stream>>>(d_lens, d_ptr, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// ---------------------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------------------
cudaError_t gsm_init_tables() {
    double h[GSM_LOGFACT_N];
    h[0] = 0.0;
    h[1] = 0.0;
    double s = 0.0;
    for (int i = 2; i < GSM_LOGFACT_N; i++) {  // same ascending order as the Java loops (STP:567, BSP:121)
        s += log(static_cast<double>(i));
        h[i] = s;
    }
    cudaError_t e = cudaMemcpyToSymbol(g_logfact, h, sizeof(h));
    if (e != cudaSuccess) return e;
    static GsmLeanBlob blob;
    gsm_lean_blob_fill(blob);
    return cudaMemcpyToSymbol(g_lean_blob, &blob, sizeof(blob));
}

size_t gsm_scratch_row_bytes() { return sizeof(GsmLeanRow); }

static size_t generic_smem_bytes(int64_t stride) { return static_cast<size_t>(stride) * 22u; }
static size_t lean_smem_bytes(int32_t n_nodes, int a_dyn, int pairs_per_chunk, int stages, bool pk = false) {
    return gsm_lean_layout(static_cast<uint32_t>(a_dyn), gsm_lean_si(n_nodes), static_cast<uint32_t>(stages) * pairs_per_chunk * 64u, pk).total;
}

// Kernel attributes are set once per handle and kernel, to values that do not depend on the batch (the device's opt-in
// maximum): concurrent handles that race here write the same values.
template <typename KERN>
static cudaError_t prep(GsmLaunchCtx& ctx, KERN kern, size_t smem) {
    if (smem > ctx.smem_optin) return cudaErrorInvalidConfiguration;
    const void* key = reinterpret_cast<const void*>(kern);
    for (int k = 0; k < ctx.n_prepared; k++)
        if (ctx.prepared[k] == key) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    if (smem + fa.sharedSizeBytes > ctx.smem_optin) return cudaErrorInvalidConfiguration;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx.smem_optin - fa.sharedSizeBytes));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (ctx.n_prepared < GSM_CTX_MAX_KERNELS) ctx.prepared[ctx.n_prepared++] = key;
    return cudaSuccess;
}

cudaError_t gsm_launch_ctx_init(GsmLaunchCtx& ctx, int device) {
    ctx = GsmLaunchCtx{};
    ctx.device = device;
    int v = 0;
    cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    ctx.sm_count = v > 0 ? v : 148;
    if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device)) != cudaSuccess) return e;
    ctx.smem_per_sm = v > 0 ? static_cast<size_t>(v) : 233472u;
    if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return e;
    ctx.smem_optin = v > 0 ? static_cast<size_t>(v) : 232448u;
    return cudaSuccess;
}

template <int BLOCK, int MINB>
static cudaError_t launch_generic(GsmLaunchCtx& ctx, const GsmBatch& b, double* o, double* r, double* w, cudaStream_t st,
                                  GsmBaseRow* base_rows = nullptr, GsmTreeConsts* base_K = nullptr) {
    const size_t smem = generic_smem_bytes(b.stride);
    const dim3 grid(static_cast<unsigned>(b.n_trees));
    cudaError_t e;
    if (r || w) {
        auto kern = gsm_eval_kernel<BLOCK, MINB, true>;
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, o, r, w, base_rows, base_K);
    } else {
        auto kern = gsm_eval_kernel<BLOCK, MINB, false>;
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, o, r, w, base_rows, base_K);
    }
    return cudaGetLastError();
}

// shared-window address at which a kernel's dynamic shared memory starts (probe launch, once per kernel and process)
template <typename KERN>
static cudaError_t lean_probe(KERN kern, int block, const GsmBatch& b, const GsmScratch& sc, cudaStream_t st, int* a_dyn) {
    kern<<<1, block, 0, st>>>(b, nullptr, sc.redo + 2, nullptr, nullptr, nullptr, nullptr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int32_t v = 0;
    if ((e = cudaMemcpyAsync(&v, sc.redo + 2, sizeof v, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    *a_dyn = v;
    return cudaSuccess;
}

// dynamic + static + driver-reserved shared memory one CTA of the lean kernel <BLOCK, MINB> takes for this batch
// index of a compiled-in lean configuration in GsmLaunchCtx::a_dyn
constexpr int gsm_lean_cfg_index(int block, int minb, bool prod) { return prod ? 4 : (block == 128 ? (minb == 4 ? 0 : (minb == 3 ? 1 : 3)) : 2); }
template <int BLOCK, int MINB, bool PROD>
static cudaError_t lean_cta_bytes(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmScratch& sc, bool rates, cudaStream_t st, size_t* dyn,
                                  size_t* all) {
    int* a_dyn = ctx.a_dyn + 2 * gsm_lean_cfg_index(BLOCK, MINB, PROD);
    constexpr int HB = BLOCK * gsm_lean_pairs(BLOCK, MINB), NST = gsm_lean_stages(BLOCK, MINB);
    const int k = rates ? 1 : 0;
    if (a_dyn[k] < 0) {
        cudaError_t e;
        // (the mixture kernels have the same static shared memory: one probe serves both flavours)
        if (rates) e = lean_probe(gsm_lean_kernel<BLOCK, MINB, true, false, PROD>, BLOCK + (PROD ? GSM_PROD_THREADS : 0), b, sc, st, &a_dyn[k]);
        else e = lean_probe(gsm_lean_kernel<BLOCK, MINB, false, false, PROD>, BLOCK + (PROD ? GSM_PROD_THREADS : 0), b, sc, st, &a_dyn[k]);
        if (e != cudaSuccess) return e;
    }
    *dyn = lean_smem_bytes(b.n_nodes, a_dyn[k], HB, NST);
    // the dynamic segment starts at a_dyn = reserved + static (rounded): a_dyn + dyn is the CTA's whole footprint
    *all = static_cast<size_t>(a_dyn[k]) + *dyn;
    return cudaSuccess;
}

template <int BLOCK, int MINB, bool PROD>
static cudaError_t launch_lean(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmScratch& sc, double* r, double* w, cudaStream_t st) {
    const dim3 grid(static_cast<unsigned>(b.n_trees));
    GsmLeanRow* rows = reinterpret_cast<GsmLeanRow*>(sc.rows);
    int32_t* cnt = sc.redo + (sc.epoch & 1);
    int32_t* cnt_next = sc.redo + ((sc.epoch + 1) & 1);
    int32_t* list = sc.redo + 4;
    cudaError_t e;
    size_t smem = 0, all = 0;
    if ((e = lean_cta_bytes<BLOCK, MINB, PROD>(ctx, b, sc, r || w, st, &smem, &all)) != cudaSuccess) return e;
    const bool mixk = gsm_lean_mixture(b.flags, b.stub_mode);
#define GSM_LAUNCH_LEAN(R, M)                                                                        \
    {                                                                                                \
        auto kern = gsm_lean_kernel<BLOCK, MINB, R, M, PROD>;                                        \
        if ((e = prep(ctx, kern, smem + ctx.smem_pad)) != cudaSuccess) return e;                     \
        kern<<<grid, BLOCK + (PROD ? GSM_PROD_THREADS : 0), smem + ctx.smem_pad, st>>>(b, rows, cnt, cnt_next, list, r, w); \
    }
    if (r || w) {
        if (mixk) GSM_LAUNCH_LEAN(true, true) else GSM_LAUNCH_LEAN(true, false)
    } else {
        if (mixk) GSM_LAUNCH_LEAN(false, true) else GSM_LAUNCH_LEAN(false, false)
    }
#undef GSM_LAUNCH_LEAN
    return cudaGetLastError();
}

// persistent lean kernel (gsm_lean_pk.cuh): 3 CTAs per SM, each walking over trees blockIdx.x, blockIdx.x + gridDim.x, ...
static cudaError_t lean_pk_cta_bytes(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmScratch& sc, bool rates, cudaStream_t st, size_t* dyn,
                                     size_t* all) {
    int* a_dyn = ctx.a_dyn + 2 * 5;
    const int k = rates ? 1 : 0;
    if (a_dyn[k] < 0) {
        cudaError_t e;
        if (rates) e = lean_probe(gsm_lean_pk_kernel<true>, GSM_PK_THREADS, b, sc, st, &a_dyn[k]);
        else e = lean_probe(gsm_lean_pk_kernel<false>, GSM_PK_THREADS, b, sc, st, &a_dyn[k]);
        if (e != cudaSuccess) return e;
    }
    *dyn = lean_smem_bytes(b.n_nodes, a_dyn[k], GSM_PK_CONSUMERS, GSM_RING_STAGES, true);
    *all = static_cast<size_t>(a_dyn[k]) + *dyn;
    return cudaSuccess;
}
static cudaError_t launch_lean_pk(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmScratch& sc, double* r, double* w, cudaStream_t st) {
    int64_t g = static_cast<int64_t>(ctx.sm_count) * 3;
    if (g > b.n_trees) g = b.n_trees;
    const dim3 grid(static_cast<unsigned>(g));
    GsmLeanRow* rows = reinterpret_cast<GsmLeanRow*>(sc.rows);
    int32_t* cnt = sc.redo + (sc.epoch & 1);
    int32_t* cnt_next = sc.redo + ((sc.epoch + 1) & 1);
    int32_t* list = sc.redo + 4;
    cudaError_t e;
    size_t smem = 0, all = 0;
    if ((e = lean_pk_cta_bytes(ctx, b, sc, r || w, st, &smem, &all)) != cudaSuccess) return e;
    if (r || w) {
        auto kern = gsm_lean_pk_kernel<true>;
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;
        kern<<<grid, GSM_PK_THREADS, smem, st>>>(b, rows, cnt, cnt_next, list, r, w);
    } else {
        auto kern = gsm_lean_pk_kernel<false>;
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;
        kern<<<grid, GSM_PK_THREADS, smem, st>>>(b, rows, cnt, cnt_next, list, r, w);
    }
    return cudaGetLastError();
}

// small trees: one warp per tree (gsm_lean_warp.cuh), one resident CTA per SM
static cudaError_t launch_lean_warp(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmScratch& sc, double* r, double* w, cudaStream_t st) {
    const int si = gsm_lean_si(b.n_nodes);
    int warps = GSM_WARP_WARPS;
    while (warps > 1 && gsm_warp_cta_bytes(si, warps) + 64u > ctx.smem_optin) --warps;
    const size_t smem = gsm_warp_cta_bytes(si, warps);
    int64_t g = (b.n_trees + warps - 1) / warps;
    if (g > ctx.sm_count) g = ctx.sm_count;
    const dim3 grid(static_cast<unsigned>(g));
    GsmLeanRow* rows = reinterpret_cast<GsmLeanRow*>(sc.rows);
    int32_t* cnt = sc.redo + (sc.epoch & 1);
    int32_t* cnt_next = sc.redo + ((sc.epoch + 1) & 1);
    int32_t* list = sc.redo + 4;
    cudaError_t e;
    const bool mixk = gsm_lean_mixture(b.flags, b.stub_mode);
#define GSM_LAUNCH_WARP(R, M)                                                          \
    {                                                                                  \
        auto kern = gsm_lean_warp_kernel<R, M>;                                        \
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;                      \
        kern<<<grid, 32 * warps, smem, st>>>(b, rows, cnt, cnt_next, list, r, w);      \
    }
    if (r || w) {
        if (mixk) GSM_LAUNCH_WARP(true, true) else GSM_LAUNCH_WARP(true, false)
    } else {
        if (mixk) GSM_LAUNCH_WARP(false, true) else GSM_LAUNCH_WARP(false, false)
    }
#undef GSM_LAUNCH_WARP
    return cudaGetLastError();
}

template <int BLOCK, int MINB>
static cudaError_t launch_post(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmScratch& sc, double* o, double* r, double* w,
                               cudaStream_t st) {
    const int sm_count = ctx.sm_count;
    const size_t smem = generic_smem_bytes(b.stride);
    int64_t g = (b.n_trees + BLOCK - 1) / BLOCK;
    const int64_t cap = static_cast<int64_t>(sm_count) * MINB;
    if (g < cap) g = b.n_trees < cap ? b.n_trees : cap;   // enough CTAs for the redo list when the batch is small
    if (g > 4 * cap) g = 4 * cap;
    if (g < 1) g = 1;
    const dim3 grid(static_cast<unsigned>(g));
    const GsmLeanRow* rows = reinterpret_cast<const GsmLeanRow*>(sc.rows);
    const int32_t* cnt = sc.redo + (sc.epoch & 1);
    const int32_t* list = sc.redo + 4;
    cudaError_t e;
    if (r || w) {
        auto kern = gsm_post_kernel<BLOCK, MINB, true>;
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, rows, cnt, list, o, r, w);
    } else {
        auto kern = gsm_post_kernel<BLOCK, MINB, false>;
        if ((e = prep(ctx, kern, smem)) != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, rows, cnt, list, o, r, w);
    }
    return cudaGetLastError();
}

// launch configurations of the lean kernel that are compiled in
#ifdef GSM_DEV_BUILD
#define GSM_LEAN_CFGS GSM_CFG(128, 3)
#else
#define GSM_LEAN_CFGS GSM_CFG(128, 4) GSM_CFG(128, 3) GSM_CFG(256, 2)
#endif

// Launch configuration: one CTA per tree.  The shared-memory footprint per node (18 B lean, 22 B generic) bounds
// the number of resident trees per SM; BLOCK and the register budget (MINB) are picked from the measured sweeps in
// profiles/.  gsm_set_option("block" / "minb" / "force_generic") overrides the choice for tuning runs.
int gsm_launch_eval(GsmLaunchCtx& ctx, const GsmBatch& b, GsmScratch& sc, double* d_out_tree, double* d_out_rates,
                    double* d_out_wspikes, cudaStream_t stream) {
    if (b.n_trees <= 0) return 0;
    if (ctx.sm_count <= 0) return -static_cast<int>(cudaErrorInvalidValue);
    // the lean kernel moves rows with 16-byte bulk copies and writes rates / weighted spikes as double2
    const bool aligned = ((reinterpret_cast<uintptr_t>(b.rate) | reinterpret_cast<uintptr_t>(b.spike) |
                           reinterpret_cast<uintptr_t>(b.nstubs) | reinterpret_cast<uintptr_t>(b.height) |
                           reinterpret_cast<uintptr_t>(b.parent) | reinterpret_cast<uintptr_t>(d_out_rates) |
                           reinterpret_cast<uintptr_t>(d_out_wspikes)) & 15u) == 0 && (b.stride % 32) == 0;
    const bool lean = gsm_lean_wiring_ok(b.flags, b.stub_mode) && b.rate != nullptr && sc.rows != nullptr && sc.redo != nullptr &&
                      aligned && (b.stub_mode == GSM_STUBS_NOSTUB || b.nstubs != nullptr) && !ctx.force_generic;
    cudaError_t e = cudaErrorInvalidConfiguration;
    int launches = 0;
    if (lean) {
        // measured (profiles/r01_tuning.md): small trees want many small CTAs per SM (the per-tree prologue of one CTA
        // overlaps the main loops of the others); from ~2,000 nodes on shared memory allows two CTAs per SM only
        // 128 x 4 up to 512 nodes; 128 x 3 as long as three of its CTAs fit the SM's shared memory (up to ~4,300 nodes);
        // 256 x 2 above
        int block = 128, minb = 4;
        if (b.n_nodes > 512) {
            size_t dyn = 0, all = 0;
            cudaError_t pe = lean_cta_bytes<128, 3, false>(ctx, b, sc, d_out_rates || d_out_wspikes, stream, &dyn, &all);
            if (pe != cudaSuccess) return -static_cast<int>(pe);
            if (3 * ((all + 1023) & ~static_cast<size_t>(1023)) <= ctx.smem_per_sm) {
                minb = 3;
            } else {
                block = 256;
                minb = 2;
            }
        }
        if (ctx.block > 0) block = ctx.block;
        if (ctx.minb > 0) minb = ctx.minb;
        // the persistent kernel (known-n spike prior) whenever three of its CTAs fit an SM, unless gsm_set_option("persistent", 0)
        bool pk = ctx.persistent != 0 && ctx.block <= 0 && ctx.minb <= 0 && !gsm_lean_mixture(b.flags, b.stub_mode);
        if (ctx.persistent == 2) pk = !gsm_lean_mixture(b.flags, b.stub_mode);   // forced (tuning runs)
        if (pk) {
            size_t dyn = 0, all = 0;
            cudaError_t pe = lean_pk_cta_bytes(ctx, b, sc, d_out_rates || d_out_wspikes, stream, &dyn, &all);
            if (pe != cudaSuccess) return -static_cast<int>(pe);
            if (3 * ((all + 1023) & ~static_cast<size_t>(1023)) > ctx.smem_per_sm) pk = false;
        }
        // small trees (up to 256 taxa): one warp per tree unless gsm_set_option("warp_kernel", 0)
        const bool wk = ctx.warp_kernel != 0 && ctx.block <= 0 && ctx.minb <= 0 && ctx.persistent == 0 && b.n_nodes <= GSM_WARP_MAX_NODES;
        if (wk) {
            e = launch_lean_warp(ctx, b, sc, d_out_rates, d_out_wspikes, stream);
        } else if (pk) {
            e = launch_lean_pk(ctx, b, sc, d_out_rates, d_out_wspikes, stream);
        } else if (block == 128 && minb == 3 && ctx.producer != 0 && !gsm_lean_mixture(b.flags, b.stub_mode)) {
            // gsm_set_option("producer", 1): 128 x 3 with a producer warp (the mixture flavour keeps the 168 registers of the
            // four-warp CTA: 1.59e10 against 1.79e10 branch evals/s with 136)
            e = launch_lean<128, 3, true>(ctx, b, sc, d_out_rates, d_out_wspikes, stream);
        } else {
#define GSM_CFG(B, M) \
    if (block == B && minb == M) e = launch_lean<B, M, false>(ctx, b, sc, d_out_rates, d_out_wspikes, stream);
            GSM_LEAN_CFGS
#undef GSM_CFG
        }
        if (e != cudaSuccess) return -static_cast<int>(e);
        if (b.n_nodes <= 2048) e = launch_post<256, 3>(ctx, b, sc, d_out_tree, d_out_rates, d_out_wspikes, stream);
        else e = launch_post<384, 2>(ctx, b, sc, d_out_tree, d_out_rates, d_out_wspikes, stream);
        if (e != cudaSuccess) return -static_cast<int>(e);
        sc.epoch++;
        launches = 2;
    } else {
        int block = b.n_nodes <= 512 ? 128 : (b.n_nodes <= 2048 ? 256 : 384);
        int minb = b.n_nodes <= 512 ? 6 : (b.n_nodes <= 2048 ? 3 : 2);
        if (ctx.block > 0) block = ctx.block;
        if (ctx.minb > 0) minb = ctx.minb;
#define GSM_CFG(B, M) \
    if (block == B && minb == M) e = launch_generic<B, M>(ctx, b, d_out_tree, d_out_rates, d_out_wspikes, stream);
        GSM_CFG(128, 8) GSM_CFG(128, 6) GSM_CFG(128, 4)
        GSM_CFG(256, 4) GSM_CFG(256, 3) GSM_CFG(256, 2)
        GSM_CFG(384, 2) GSM_CFG(512, 2) GSM_CFG(512, 1)
#undef GSM_CFG
        if (e != cudaSuccess) return -static_cast<int>(e);
        launches = 1;
    }
    return launches;
}

// ---------------------------------------------------------------------------------------------------------
// proposal batches (gsm_eval_dirty, BASELINE config 5)
//
// The reference re-evaluates every branch after every proposal (BSP:294-295, STP:836-842).  Here a proposal that edits a
// few nodes and leaves the per-tree scalars alone is evaluated INCREMENTALLY, O(dirty): with the reduced accumulators of
// the resident state ("base", captured once per resident batch by the generic kernel),
//     acc(proposed) = acc(base) - sum_{j in closure} term_j(resident) + sum_{j in closure} term_j(proposed),
// where the closure of an edited node d is d itself and, when its height changes, its two children, its parent and its
// sibling (branch lengths, expected stub counts, sampled-ancestor / fake status: Node.isDirectAncestor / isFake look one
// level up and down).  Both term sets are evaluated by the statement-by-statement node functions the base was built with
// (gsm_node_pass_a / gsm_node_pass_b), so the resident terms cancel to rounding; the epilogue (gsm_finalize) then runs on
// the updated accumulators with the proposed root height.  One warp per proposal, one lane per closure node, fixed order.
// Everything else -- hyper-parameter / indicator moves, proposals with more than GSM_DELTA_MAX_DIRTY edited nodes (tree
// scalers), trees whose base is not finite or not a proper tree -- is materialised and evaluated in full (gsm_patch_kernel +
// the kernels above), so the reference's -inf / NaN behaviour is inherited there.
// ---------------------------------------------------------------------------------------------------------
#define GSM_DELTA_MAX_DIRTY 6    // 5 closure candidates per edited node <= 30 lanes

// children of every internal node (BEAST trees are binary; a third child marks the tree as not eligible)
__global__ void __launch_bounds__(256) gsm_children_kernel(const GsmBatch b, int2* __restrict__ children, int32_t* __restrict__ tree_flag) {
    const int64_t t = blockIdx.x;
    int2* row = children + t * b.stride;
    for (int64_t i = threadIdx.x; i < b.stride; i += blockDim.x) row[i] = make_int2(-1, -1);
    if (threadIdx.x == 0) tree_flag[t] = 0;
    __syncthreads();
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    const int32_t* par = b.parent + t * b.stride;
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int32_t p = par[i];
        if (static_cast<uint32_t>(p) >= static_cast<uint32_t>(n)) continue;
        int* slot = reinterpret_cast<int*>(row + p);
        if (atomicCAS(slot, -1, i) != -1 && atomicCAS(slot + 1, -1, i) != -1) tree_flag[t] = 1;
    }
}

struct GsmDirtyDev {   // device view of GsmDirtyWork
    const int2* children;
    const GsmBaseRow* base_rows;
    const GsmTreeConsts* base_K;
    const int32_t* tree_flag;
    int32_t* full_list;
    int32_t* counters;     // [0] proposals for the full path, [1] argument error
    double* out_all;
};

// the two children of j in a fixed order (smaller node number first); -1 = none
static __device__ __forceinline__ int2 delta_children(const int2* __restrict__ ch, int32_t j) {
    int2 c = ch[j];
    if (c.y >= 0 && c.y < c.x) { const int32_t x = c.x; c.x = c.y; c.y = x; }
    return c;
}

// one warp per proposal
__global__ void __launch_bounds__(128) gsm_delta_kernel(const GsmBatch b, const GsmProposals pr, int64_t n_props, const GsmDirtyDev w) {
    __shared__ GsmTreeConsts sK[4];
    __shared__ int32_t s_node[4][GSM_DELTA_MAX_DIRTY];
    __shared__ double s_h[4][GSM_DELTA_MAX_DIRTY], s_r[4][GSM_DELTA_MAX_DIRTY], s_s[4][GSM_DELTA_MAX_DIRTY];
    __shared__ int32_t s_k[4][GSM_DELTA_MAX_DIRTY];
    __shared__ int32_t s_cand[4][32];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * 4 + wp;
    if (p >= n_props) return;
    const int64_t t64 = pr.tree_of_prop[p];
    const int64_t d0 = pr.dirty_ptr[p], d1 = pr.dirty_ptr[p + 1];
    // ---- validation (was a host loop over every CSR entry) and classification ----
    bool arg_bad = t64 < 0 || t64 >= b.n_trees || d1 < d0;
    const int64_t t = arg_bad ? 0 : t64;
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    if (!arg_bad) {
        bool bad = false;
        for (int64_t d = d0 + lane; d < d1; d += 32) {
            const int32_t node = pr.dirty_nodes[d];
            bad = bad || node < 0 || node >= b.n_nodes;
        }
        arg_bad = __any_sync(0xffffffffu, bad);
    }
    if (arg_bad) {
        if (lane == 0) w.counters[1] = 1;
        return;
    }
    const int64_t nd = d1 - d0;
    const GsmBaseRow base = w.base_rows[t];
    bool full = nd > GSM_DELTA_MAX_DIRTY || base.ok == 0 || w.tree_flag[t] != 0 || (pr.height_scale && !(pr.height_scale[p] == 1.0));
    if (!full && pr.new_params) {   // per-tree scalars replaced: every branch term changes
        bool diff = false;
        if (lane < GSM_NPARAM) {
            const double a = pr.new_params[p * GSM_NPARAM + lane], c = b.params[t * GSM_NPARAM + lane];
            diff = !(a == c);
        }
        full = __any_sync(0xffffffffu, diff);
    }
    if (!full && pr.new_use_spike) {
        const uint8_t cur = b.use_spike ? b.use_spike[t] : static_cast<uint8_t>(1);
        full = (pr.new_use_spike[p] != 0) != (cur != 0);
    }
    const int64_t row = t * b.stride;
    if (!full) {   // an edited node beyond the tree's own node count cannot be expressed as a delta
        bool out = false;
        if (lane < nd) out = pr.dirty_nodes[d0 + lane] >= n;
        full = __any_sync(0xffffffffu, out);
    }
    if (full) {
        if (lane == 0) w.full_list[atomicAdd(w.counters, 1)] = static_cast<int32_t>(p);
        return;
    }
    // ---- the edits of this proposal -> shared memory ----
    if (lane < nd) {
        const int64_t d = d0 + lane;
        const int32_t node = pr.dirty_nodes[d];
        s_node[wp][lane] = node;
        s_h[wp][lane] = pr.new_height ? pr.new_height[d] : b.height[row + node];
        s_r[wp][lane] = (pr.new_rate && b.rate) ? pr.new_rate[d] : (b.rate ? b.rate[row + node] : 1.0);
        s_s[wp][lane] = pr.new_spike ? pr.new_spike[d] : b.spike[row + node];
        s_k[wp][lane] = (pr.new_nstubs && b.nstubs) ? pr.new_nstubs[d] : (b.nstubs ? b.nstubs[row + node] : 0);
    }
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(w.base_K + t);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sK[wp]);
        for (int k = lane; k < static_cast<int>(sizeof(GsmTreeConsts) / 4); k += 32) dst[k] = src[k];
    }
    __syncwarp();
    const GsmTreeConsts& K = sK[wp];
    const int2* ch = w.children + row;
    const int32_t* par = b.parent + row;
    const int ndi = static_cast<int>(nd);
    // patched reads of the proposed state (NEW) or plain reads of the resident state
    auto find = [&](int32_t j) { int f = -1; for (int k = 0; k < ndi; k++) if (s_node[wp][k] == j) f = k; return f; };
    auto H = [&](int32_t j, bool nw) { const int f = nw ? find(j) : -1; return f >= 0 ? s_h[wp][f] : b.height[row + j]; };
    // ---- closure candidates: lane c -> edited node c / 5, role c % 5 (self, children, parent, sibling) ----
    int32_t cand = -1;
    if (lane < 5 * ndi) {
        const int k = lane / 5, role = lane % 5;
        const int32_t d = s_node[wp][k];
        const bool moved = !(__double_as_longlong(s_h[wp][k]) == __double_as_longlong(b.height[row + d]));
        if (role == 0) cand = d;
        else if (moved) {
            const int32_t pd = par[d];
            if (role == 1) cand = delta_children(ch, d).x;
            else if (role == 2) cand = delta_children(ch, d).y;
            else if (role == 3) cand = pd >= 0 ? pd : -1;
            else if (pd >= 0) {
                const int2 c = delta_children(ch, pd);
                cand = c.x == d ? c.y : c.x;
            }
        }
    }
    s_cand[wp][lane] = cand;
    __syncwarp();
    for (int k = 0; k < lane; k++)
        if (s_cand[wp][k] == cand) cand = -1;   // each node once
    // ---- terms of this lane's node in the resident and in the proposed state ----
    GsmAcc A[2];
    gsm_acc_zero(A[0]);
    gsm_acc_zero(A[1]);
    if (cand >= 0) {
        const int32_t j = cand;
        const int32_t pj = par[j];
        const bool root = pj < 0;
        const int2 cj = delta_children(ch, j);
        const bool nonleaf = cj.x >= 0;
        int2 cp = make_int2(-1, -1);
        if (!root) cp = delta_children(ch, pj);
#pragma unroll 1
        for (int v = 0; v < 2; v++) {
            const bool nw = v == 1;
            const double h = H(j, nw);
            const double hp = root ? h : H(pj, nw);
            uint32_t nf = 0;
            if (nonleaf) nf |= GSM_NF_NONLEAF;
            if (root) nf |= GSM_NF_ROOT;
            if (!nonleaf && !root && hp == h) nf |= GSM_NF_DA;
            // Node.isFake: a non-leaf with a direct-ancestor child (a leaf child at the same height)
            auto leaf_at = [&](int32_t c, double hh) { return c >= 0 && ch[c].x < 0 && H(c, nw) == hh; };
            if (nonleaf && (leaf_at(cj.x, h) || leaf_at(cj.y, h))) nf |= GSM_NF_FAKE;
            if (!root && (leaf_at(cp.x, hp) || leaf_at(cp.y, hp))) nf |= GSM_NF_PARENT_FAKE;
            const double aux = gsm_node_pass_a(K, A[v], nf, h, hp);
            double auxp = aux;
            if (!root) auxp = K.need_aux ? gsm_log(fma(K.omc2, gsm_exp(-K.c1 * hp), K.opc2)) : 0.0;   // == gsm_node_pass_a of the parent
            const int f = nw ? find(j) : -1;
            const double rate = f >= 0 ? s_r[wp][f] : (b.rate ? b.rate[row + j] : 1.0);
            const double spike = f >= 0 ? s_s[wp][f] : b.spike[row + j];
            const int32_t nst = f >= 0 ? s_k[wp][f] : (b.nstubs ? b.nstubs[row + j] : 0);
            gsm_node_pass_b<false>(K, g_logfact, A[v], j, n, nf, h, hp, aux, auxp, rate, spike, nst, nullptr, nullptr);
        }
    }
    // ---- fixed-order reduction of (proposed - resident), update of the base accumulators, epilogue ----
    double dv[6] = {A[1].tree - A[0].tree, A[1].stub - A[0].stub, A[1].spike - A[0].spike,
                    A[1].rate - A[0].rate, A[1].burst - A[0].burst, A[1].dist - A[0].dist};
    // a term that is -inf / NaN in the proposed state must survive the subtraction of a finite resident term
#pragma unroll
    for (int k = 0; k < 6; k++) dv[k] = warp_sum(dv[k]);
    const int32_t c_ext = __reduce_add_sync(0xffffffffu, A[1].n_extant - A[0].n_extant),
                  c_da = __reduce_add_sync(0xffffffffu, A[1].n_da - A[0].n_da),
                  c_exc = __reduce_add_sync(0xffffffffu, A[1].n_extinct - A[0].n_extinct),
                  c_bad = static_cast<int32_t>(__reduce_or_sync(0xffffffffu, static_cast<uint32_t>(A[1].bad)));
    if (lane == 0) {
        GsmAcc R;
        R.tree = base.acc[0] + dv[0]; R.stub = base.acc[1] + dv[1]; R.spike = base.acc[2] + dv[2];
        R.rate = base.acc[3] + dv[3]; R.burst = base.acc[4] + dv[4]; R.dist = base.acc[5] + dv[5];
        R.n_extant = base.n_extant + c_ext; R.n_leaf = base.n_leaf; R.n_da = base.n_da + c_da;
        R.n_extinct = base.n_extinct + c_exc; R.bad = c_bad;
        const int32_t root = base.root;
        const double T = H(root, true);
        const int2 cr = delta_children(ch, root);
        auto leaf_at_new = [&](int32_t c) { return c >= 0 && ch[c].x < 0 && H(c, true) == T; };
        const bool root_fake = leaf_at_new(cr.x) || leaf_at_new(cr.y);
        GsmTreeConsts K2 = K;
        gsm_finish_consts(K2, T);
        double res[GSM_NOUT];
        gsm_finalize(K2, R, n, root_fake, res);
        double* orow = w.out_all + p * GSM_NOUT;
#pragma unroll
        for (int k = 0; k < GSM_NOUT; k++) orow[k] = res[k];
    }
}

// one CTA per proposal of the full path copies the rows of its tree (16-byte accesses) and then applies the proposal's edits
__global__ void __launch_bounds__(256) gsm_patch_kernel(const GsmBatch src, const GsmProposals pr, const int32_t* __restrict__ list,
                                                        const int2* __restrict__ children, int64_t q0, const GsmBatchMut dst) {
    const int64_t q = blockIdx.x;            // row of the destination batch
    const int64_t p = list[q0 + q];          // proposal
    const int64_t t = pr.tree_of_prop[p];
    const int64_t S = src.stride;
    const int64_t so = t * S, doff = q * S;
    const int tid = threadIdx.x;
    // rows are 128-byte aligned and S % 32 == 0: whole rows as 16-byte words
    {
        const double2* h = reinterpret_cast<const double2*>(src.height + so);
        const double2* s_ = reinterpret_cast<const double2*>(src.spike + so);
        double2* dh = reinterpret_cast<double2*>(dst.height + doff);
        double2* ds = reinterpret_cast<double2*>(dst.spike + doff);
        for (int64_t i = tid; i < S / 2; i += blockDim.x) {
            dh[i] = h[i];
            ds[i] = s_[i];
        }
        if (src.rate) {
            const double2* r = reinterpret_cast<const double2*>(src.rate + so);
            double2* dr = reinterpret_cast<double2*>(dst.rate + doff);
            for (int64_t i = tid; i < S / 2; i += blockDim.x) dr[i] = r[i];
        }
        const int4* pa = reinterpret_cast<const int4*>(src.parent + so);
        int4* dp = reinterpret_cast<int4*>(dst.parent + doff);
        for (int64_t i = tid; i < S / 4; i += blockDim.x) dp[i] = pa[i];
        if (src.nstubs) {
            const int4* k = reinterpret_cast<const int4*>(src.nstubs + so);
            int4* dk = reinterpret_cast<int4*>(dst.nstubs + doff);
            for (int64_t i = tid; i < S / 4; i += blockDim.x) dk[i] = k[i];
        }
    }
    if (tid < GSM_NPARAM) dst.params[q * GSM_NPARAM + tid] = pr.new_params ? pr.new_params[p * GSM_NPARAM + tid] : src.params[t * GSM_NPARAM + tid];
    if (tid == 32) dst.use_spike[q] = pr.new_use_spike ? pr.new_use_spike[p] : (src.use_spike ? src.use_spike[t] : static_cast<uint8_t>(1));
    if (tid == 33) dst.node_count[q] = src.node_count ? src.node_count[t] : src.n_nodes;
    const double scale = pr.height_scale ? pr.height_scale[p] : 1.0;
    if (!(scale == 1.0)) {
        __syncthreads();   // (uniform branch) the scaled heights overwrite row entries other threads have just copied
        // Tree.scale(s) -> Node.scale: every internal node that is not fake (no leaf child at its own height)
        int32_t n = src.node_count ? src.node_count[t] : src.n_nodes;
        if (n > src.n_nodes) n = src.n_nodes;
        const int2* ch = children + so;
        for (int32_t i = tid; i < n; i += blockDim.x) {
            const int2 c = ch[i];
            if (c.x < 0) continue;
            const double h = src.height[so + i];
            const bool fake = (ch[c.x].x < 0 && src.height[so + c.x] == h) || (c.y >= 0 && ch[c.y].x < 0 && src.height[so + c.y] == h);
            if (!fake) dst.height[doff + i] = h * scale;
        }
    }
    __syncthreads();   // the edits overwrite what this CTA has just copied
    const int64_t d0 = pr.dirty_ptr[p], d1 = pr.dirty_ptr[p + 1];
    for (int64_t d = d0 + tid; d < d1; d += blockDim.x) {
        const int32_t node = pr.dirty_nodes[d];
        if (node < 0 || node >= src.n_nodes) continue;   // flagged by gsm_delta_kernel as an argument error
        if (pr.new_height) dst.height[doff + node] = pr.new_height[d];
        if (pr.new_rate && src.rate) dst.rate[doff + node] = pr.new_rate[d];
        if (pr.new_spike) dst.spike[doff + node] = pr.new_spike[d];
        if (pr.new_nstubs && src.nstubs) dst.nstubs[doff + node] = pr.new_nstubs[d];
    }
}

// result rows of a chunk of the full path -> their proposals' rows
__global__ void gsm_scatter_rows_kernel(const double* __restrict__ rows, const int32_t* __restrict__ list, int64_t q0, int64_t count,
                                        double* __restrict__ out_all) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count * GSM_NOUT) return;
    const int64_t q = i / GSM_NOUT, k = i % GSM_NOUT;
    out_all[static_cast<int64_t>(list[q0 + q]) * GSM_NOUT + k] = rows[i];
}

size_t gsm_base_row_bytes() { return sizeof(GsmBaseRow); }
size_t gsm_base_k_bytes() { return sizeof(GsmTreeConsts); }

static GsmDirtyDev dirty_dev(const GsmDirtyWork& w) {
    GsmDirtyDev d;
    d.children = reinterpret_cast<const int2*>(w.children);
    d.base_rows = reinterpret_cast<const GsmBaseRow*>(w.base_rows);
    d.base_K = reinterpret_cast<const GsmTreeConsts*>(w.base_K);
    d.tree_flag = w.tree_flag;
    d.full_list = w.full_list;
    d.counters = w.counters;
    d.out_all = w.out_all;
    return d;
}

// base of the resident batch: children table + accumulators of every tree (generic kernel, result rows to w.out_base)
int gsm_launch_dirty_base(GsmLaunchCtx& ctx, const GsmBatch& b, const GsmDirtyWork& w, cudaStream_t stream) {
    if (b.n_trees <= 0) return 0;
    gsm_children_kernel<<<static_cast<unsigned>(b.n_trees), 256, 0, stream>>>(b, reinterpret_cast<int2*>(w.children), w.tree_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    GsmBaseRow* rows = reinterpret_cast<GsmBaseRow*>(w.base_rows);
    GsmTreeConsts* Ks = reinterpret_cast<GsmTreeConsts*>(w.base_K);
    if (b.n_nodes <= 512) e = launch_generic<128, 6>(ctx, b, w.out_base, nullptr, nullptr, stream, rows, Ks);
    else if (b.n_nodes <= 2048) e = launch_generic<256, 3>(ctx, b, w.out_base, nullptr, nullptr, stream, rows, Ks);
    else e = launch_generic<384, 2>(ctx, b, w.out_base, nullptr, nullptr, stream, rows, Ks);
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 2;
}

// validation + classification + the incremental evaluation of every eligible proposal
int gsm_launch_dirty_delta(const GsmBatch& b, const GsmProposals& pr, int64_t n_props, const GsmDirtyWork& w, cudaStream_t stream) {
    if (n_props <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(w.counters, 0, 2 * sizeof(int32_t), stream);
    if (e != cudaSuccess) return -static_cast<int>(e);
    gsm_delta_kernel<<<static_cast<unsigned>((n_props + 3) / 4), 128, 0, stream>>>(b, pr, n_props, dirty_dev(w));
    e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}

int gsm_launch_patch(const GsmBatch& src, const GsmProposals& pr, const GsmDirtyWork& w, int64_t q0, int64_t count, const GsmBatchMut& dst,
                     cudaStream_t stream) {
    if (count <= 0) return 0;
    gsm_patch_kernel<<<static_cast<unsigned>(count), 256, 0, stream>>>(src, pr, w.full_list, reinterpret_cast<const int2*>(w.children), q0, dst);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}

int gsm_launch_scatter_rows(const double* d_rows, const int32_t* d_list, int64_t q0, int64_t count, double* d_out_all, cudaStream_t stream) {
    if (count <= 0) return 0;
    const int64_t n = count * GSM_NOUT;
    gsm_scatter_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(d_rows, d_list, q0, count, d_out_all);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}

// EXACT mode: per-branch stub counts of every tree from its stub list (Stubs.getNStubsOnBranch, Stubs:357-364)
__global__ void __launch_bounds__(256) gsm_stub_counts_kernel(const GsmBatch b, int32_t* __restrict__ counts) {
    const int64_t t = blockIdx.x;
    int32_t* row = counts + t * b.stride;
    for (int64_t i = threadIdx.x; i < b.stride; i += blockDim.x) row[i] = 0;
    __syncthreads();
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    const int64_t s0 = b.stub_ptr[t], s1 = b.stub_ptr[t + 1];
    for (int64_t j = s0 + threadIdx.x; j < s1; j += blockDim.x) {
        const int32_t br = b.stub_branch[j];
        if (static_cast<uint32_t>(br) < static_cast<uint32_t>(n)) atomicAdd(row + br, 1);
    }
}

int gsm_launch_stub_counts(const GsmBatch& b, int32_t* d_counts, cudaStream_t stream) {
    if (b.n_trees <= 0) return 0;
    gsm_stub_counts_kernel<<<static_cast<unsigned>(b.n_trees), 256, 0, stream>>>(b, d_counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------
// fp64 pipe micro-benchmark (bench.py: the denominator of roofline.fp64): eight independent DFMA chains per thread,
// 2,048 resident threads per SM.  Returns thread-level fused multiply-adds per second (best of `reps` launches).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8) gsm_dfma_peak_kernel(double* __restrict__ out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123456.789) out[0] = s;   // keeps the chains alive; practically never true
}

int gsm_launch_dfma_peak(const GsmLaunchCtx& ctx, double* d_scratch, int reps, double* fma_per_sec, cudaStream_t stream) {
    const int iters = 2000;
    const dim3 grid(static_cast<unsigned>(ctx.sm_count * 8));
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return -static_cast<int>(e);
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return -static_cast<int>(e); }
    double best_ms = 1e30;
    int launches = 0;
    for (int r = 0; r < reps + 1; r++) {   // first launch = warm-up
        cudaEventRecord(e0, stream);
        gsm_dfma_peak_kernel<<<grid, 256, 0, stream>>>(d_scratch, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, stream);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        launches++;
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return -static_cast<int>(e);
    *fma_per_sec = static_cast<double>(grid.x) * 256.0 * iters * 128.0 / (best_ms * 1e-3);
    return launches;
}

int gsm_launch_stub_cdf(const GsmBatch& b, int64_t tree, int32_t node, double* d_cdf, int32_t cap, int32_t* d_len,
                        cudaStream_t stream) {
    gsm_stub_cdf_kernel<<<1, 32, 0, stream>>>(b, tree, node, d_cdf, cap, d_len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}
