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

#include <cstdio>
#include <cstdlib>

__device__ double g_logfact[GSM_LOGFACT_N];

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
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
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

// ---------------------------------------------------------------------------------------------------------
// generic kernel (every wiring)
// ---------------------------------------------------------------------------------------------------------
// MINB = CTAs per SM the register allocation is sized for (__launch_bounds__): BLOCK*MINB resident threads.
template <int BLOCK, int MINB, bool WANT_RATES>
__global__ void __launch_bounds__(BLOCK, MINB) gsm_eval_kernel(const GsmBatch b, double* __restrict__ out_tree,
                                                         double* __restrict__ out_rates,
                                                         double* __restrict__ out_wspikes) {
    constexpr int NWARP = BLOCK / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int64_t S = b.stride;
    double* sh_h = reinterpret_cast<double*>(smem_raw);
    double* sh_aux = sh_h + S;
    int32_t* sh_par = reinterpret_cast<int32_t*>(sh_aux + S);
    uint8_t* sh_nonleaf = reinterpret_cast<uint8_t*>(sh_par + S);
    uint8_t* sh_fake = sh_nonleaf + S;

    __shared__ GsmTreeConsts K;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int32_t sh_root;
    __shared__ double sh_red[6][NWARP];
    __shared__ int32_t sh_redi[5][NWARP];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t t = blockIdx.x;
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    double* orow = out_tree + t * GSM_NOUT;
    if (n <= 0) {
        if (tid < GSM_NOUT) orow[tid] = NAN;
        return;
    }
    const int64_t row = t * S;

    // ---- phase 0 ----
    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        sh_root = -1;
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes_h = (static_cast<uint32_t>(n) * 8u + 15u) & ~15u;
        const uint32_t bytes_p = (static_cast<uint32_t>(n) * 4u + 15u) & ~15u;
        mbar_arrive_expect_tx(&mbar, bytes_h + bytes_p);
        tma_bulk_g2s(sh_h, b.height + row, bytes_h, &mbar);
        tma_bulk_g2s(sh_par, b.parent + row, bytes_p, &mbar);
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
    mbar_wait(&mbar, 0);

    // ---- phase 1: non-leaf marks, root, parent sanitising ----
    for (int i = tid; i < n; i += BLOCK) {
        const int32_t p = sh_par[i];
        if (static_cast<uint32_t>(p) < static_cast<uint32_t>(n)) sh_nonleaf[p] = 1;
        else {
            if (p < 0) sh_root = i;
            else sh_par[i] = -1;  // out-of-range parent: treated as detached (never dereferenced)
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

    // ---- phase 4: reductions (fixed order) + epilogue ----
    double v0 = warp_sum(A.tree), v1 = warp_sum(A.stub), v2 = warp_sum(A.spike), v3 = warp_sum(A.rate),
           v4 = warp_sum(A.burst), v5 = warp_sum(A.dist);
    int32_t c0 = __reduce_add_sync(0xffffffffu, A.n_extant), c1 = __reduce_add_sync(0xffffffffu, A.n_leaf),
            c2 = __reduce_add_sync(0xffffffffu, A.n_da), c3 = __reduce_add_sync(0xffffffffu, A.n_extinct),
            c4 = static_cast<int32_t>(__reduce_or_sync(0xffffffffu, static_cast<uint32_t>(A.bad)));
    if (lane == 0) {
        sh_red[0][warp] = v0; sh_red[1][warp] = v1; sh_red[2][warp] = v2;
        sh_red[3][warp] = v3; sh_red[4][warp] = v4; sh_red[5][warp] = v5;
        sh_redi[0][warp] = c0; sh_redi[1][warp] = c1; sh_redi[2][warp] = c2; sh_redi[3][warp] = c3; sh_redi[4][warp] = c4;
    }
    __syncthreads();
    if (tid == 0) {
        GsmAcc R;
        gsm_acc_zero(R);
#pragma unroll
        for (int w = 0; w < NWARP; w++) {
            R.tree += sh_red[0][w]; R.stub += sh_red[1][w]; R.spike += sh_red[2][w];
            R.rate += sh_red[3][w]; R.burst += sh_red[4][w]; R.dist += sh_red[5][w];
            R.n_extant += sh_redi[0][w]; R.n_leaf += sh_redi[1][w]; R.n_da += sh_redi[2][w];
            R.n_extinct += sh_redi[3][w]; R.bad |= sh_redi[4][w];
        }
        const int32_t root = sh_root;
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
}

// ---------------------------------------------------------------------------------------------------------
// fast kernel: default wiring family (gsm_fast_path_ok), one CTA per tree.
//   shared memory per node: height 8 B + aux 8 B + packed word (parent index + NL/FAKE/DA/PF bits) 2 or 4 B
//   -> 18 B/node for trees of <= 4096 nodes (72 KB at 3,999 nodes: three trees resident per SM).
//   phase 0  TMA bulk copy of the height row; parent row read with coalesced loads and packed into words;
//            last warp computes the per-tree constants
//   phase 1  children mark parents non-leaf (idempotent 16/32-bit read-modify-write, no atomics needed)
//   phase 1b zero-length leaves mark parents fake
//   phase 2  pass A (per-variant loop): census, own tree term, aux -> smem, DA / parent-fake bits -> word
//   phase 3  pass B, two adjacent nodes per thread, 128-bit global and shared loads, one-strip prefetch
//   phase 4  fixed-order reduction + epilogue
// ---------------------------------------------------------------------------------------------------------
template <typename WORD, int VARIANT, int BLOCK>
static __device__ __forceinline__ void fast_pass_a_loop(const GsmFastA& fa, const GsmTabs& T, GsmAcc& A, int n, int tid,
                                                        const double* sh_h, double* sh_aux, WORD* sh_w) {
    using W = GsmWord<WORD>;
    const int nfull = n >> 1;   // full pairs; an odd last node is handled below
    for (int j = tid; j < nfull; j += BLOCK) {
        const int i0 = 2 * j;
        const double2 h2 = *reinterpret_cast<const double2*>(sh_h + i0);
        const double h[2] = {h2.x, h2.y};
        uint32_t w[2];
        double hp[2];
        bool leaf[2], fake_self[2], da[2], root[2], pf[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            w[k] = sh_w[i0 + k];
            const int p = static_cast<int>(w[k] & W::IDX);
            hp[k] = sh_h[p];
            const uint32_t wp = sh_w[p];
            root[k] = (p == i0 + k);
            leaf[k] = !(w[k] & W::NL);
            fake_self[k] = (w[k] & W::FAKE) != 0;
            da[k] = leaf[k] && !root[k] && hp[k] == h[k];
            pf[k] = !root[k] && (wp & W::FAKE);
        }
        double aux[2];
        const bool sp = !fa.safe || gsm_fast_a_special(fa, h[0]) || gsm_fast_a_special(fa, h[1]);
        if (__any_sync(__activemask(), sp)) {
#pragma unroll
            for (int k = 0; k < 2; k++)
                aux[k] = gsm_fast_pass_a_checked<VARIANT>(fa, T, A, leaf[k], fake_self[k], da[k], root[k], h[k], hp[k]);
        } else {
            gsm_fast_pass_a_pair<VARIANT>(fa, T, A, leaf, fake_self, da, root, h, hp, aux);
        }
        *reinterpret_cast<double2*>(sh_aux + i0) = make_double2(aux[0], aux[1]);
#pragma unroll
        for (int k = 0; k < 2; k++)
            sh_w[i0 + k] = static_cast<WORD>(w[k] | (da[k] ? W::DA : 0u) | (pf[k] ? W::PF : 0u));
    }
    if ((n & 1) && tid == (nfull % BLOCK)) {
        const int i = n - 1;
        const uint32_t w = sh_w[i];
        const int p = static_cast<int>(w & W::IDX);
        const double h = sh_h[i], hp = sh_h[p];
        const bool root = (p == i), leaf = !(w & W::NL), fake_self = (w & W::FAKE) != 0;
        const bool da = leaf && !root && hp == h;
        const bool pf = !root && (sh_w[p] & W::FAKE);
        sh_aux[i] = gsm_fast_pass_a_checked<VARIANT>(fa, T, A, leaf, fake_self, da, root, h, hp);
        sh_w[i] = static_cast<WORD>(w | (da ? W::DA : 0u) | (pf ? W::PF : 0u));
    }
}

template <typename WORD, int BLOCK, int MINB, bool WANT_RATES>
__global__ void __launch_bounds__(BLOCK, MINB) gsm_eval_fast_kernel(const GsmBatch b, double* __restrict__ out_tree,
                                                                    double* __restrict__ out_rates,
                                                                    double* __restrict__ out_wspikes) {
    using W = GsmWord<WORD>;
    constexpr int NWARP = BLOCK / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int64_t S = b.stride;
    double* sh_h = reinterpret_cast<double*>(smem_raw);
    double* sh_aux = sh_h + S;
    WORD* sh_w = reinterpret_cast<WORD*>(sh_aux + S);

    __shared__ GsmLogEntry sh_logtab[GSM_LOG_TAB_N];
    __shared__ uint64_t sh_exptab[GSM_EXP_TAB_N];
    __shared__ GsmTreeConsts K;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int32_t sh_root;
    __shared__ int32_t sh_anyzero;
    __shared__ double sh_red[6][NWARP];
    __shared__ int32_t sh_redi[5][NWARP];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t t = blockIdx.x;
    int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
    if (n > b.n_nodes) n = b.n_nodes;
    double* orow = out_tree + t * GSM_NOUT;
    if (n <= 0) {
        if (tid < GSM_NOUT) orow[tid] = NAN;
        return;
    }
    const int64_t row = t * S;

    // ---- phase 0 ----
    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        sh_root = -1;
        sh_anyzero = 0;
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes_h = (static_cast<uint32_t>(n) * 8u + 15u) & ~15u;
        mbar_arrive_expect_tx(&mbar, bytes_h);
        tma_bulk_g2s(sh_h, b.height + row, bytes_h, &mbar);
    }
    for (int i = tid; i < GSM_LOG_TAB_N; i += BLOCK) {   // log / exp tables -> shared memory (3 KB, L2-resident source)
        const double2 e = gsm_log_tab_d[i];
        sh_logtab[i].invc = e.x;
        sh_logtab[i].logc = e.y;
        sh_exptab[i] = gsm_exp_tab_d[i];
    }
    {
        const int32_t* prow = b.parent + row;
        for (int i = tid; i < n; i += BLOCK) {
            int32_t p = ld_stream(prow + i);
            if (static_cast<uint32_t>(p) >= static_cast<uint32_t>(n)) {   // root (-1) or out of range: points to itself
                if (p < 0) sh_root = i;
                p = i;
            }
            sh_w[i] = static_cast<WORD>(p);
        }
    }
    if (warp == NWARP - 1) {
        const double* prm = b.params + t * GSM_NPARAM;
        if (lane == 0) {
            gsm_setup_consts(K, prm, b.use_spike ? static_cast<int32_t>(b.use_spike[t]) : 1, b.flags, b.stub_mode);
        } else if (lane <= GSM_LUT_M + 1) {
            gsm_lut_entry(K, prm[GSM_P_SPIKE_SHAPE], lane - 1);
        }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);

    // ---- phase 1: non-leaf marks; note whether any branch has zero length (only then can there be sampled ancestors) ----
    for (int i = tid; i < n; i += BLOCK) {
        const int p = static_cast<int>(sh_w[i] & W::IDX);
        if (p != i) {
            sh_w[p] = static_cast<WORD>(sh_w[p] | W::NL);
            if (sh_h[p] == sh_h[i]) sh_anyzero = 1;
        }
    }
    __syncthreads();
    // ---- phase 1b: sampled ancestors -> fake parents ----
    if (sh_anyzero) {
        for (int i = tid; i < n; i += BLOCK) {
            const uint32_t w = sh_w[i];
            const int p = static_cast<int>(w & W::IDX);
            if (!(w & W::NL) && p != i && sh_h[p] == sh_h[i]) sh_w[p] = static_cast<WORD>(sh_w[p] | W::FAKE);
        }
        __syncthreads();
    }

    // ---- phase 2: pass A ----
    GsmAcc A;
    gsm_acc_zero(A);
    GsmFastA fa;
    GsmFastB fb;
    gsm_fast_consts(K, fa, fb, n);
    GsmTabs T;
    T.logtab = sh_logtab;
    T.exptab = sh_exptab;
    switch (fa.variant) {
        case GSM_TV_COMPLETE_RHO: fast_pass_a_loop<WORD, GSM_TV_COMPLETE_RHO, BLOCK>(fa, T, A, n, tid, sh_h, sh_aux, sh_w); break;
        case GSM_TV_NO_PSI: fast_pass_a_loop<WORD, GSM_TV_NO_PSI, BLOCK>(fa, T, A, n, tid, sh_h, sh_aux, sh_w); break;
        case GSM_TV_PSI: fast_pass_a_loop<WORD, GSM_TV_PSI, BLOCK>(fa, T, A, n, tid, sh_h, sh_aux, sh_w); break;
        case GSM_TV_CONDITIONED: fast_pass_a_loop<WORD, GSM_TV_CONDITIONED, BLOCK>(fa, T, A, n, tid, sh_h, sh_aux, sh_w); break;
        default: fast_pass_a_loop<WORD, GSM_TV_NONE, BLOCK>(fa, T, A, n, tid, sh_h, sh_aux, sh_w); break;
    }
    __syncthreads();

    // ---- phase 3: pass B, two adjacent nodes per thread ----
    {
        const double* rate_row = b.rate + row;
        const double* spike_row = b.spike + row;
        const int32_t* nst_row = b.nstubs ? b.nstubs + row : nullptr;
        const double2* rate2 = reinterpret_cast<const double2*>(rate_row);
        const double2* spike2 = reinterpret_cast<const double2*>(spike_row);
        const int2* nst2 = reinterpret_cast<const int2*>(nst_row);
        double* orate = (WANT_RATES && out_rates) ? out_rates + row : nullptr;
        double* owsp = (WANT_RATES && out_wspikes) ? out_wspikes + row : nullptr;
        double dummy[2];
        const int nfull = n >> 1;
        int j = tid;
        double2 r_n = make_double2(1.0, 1.0), s_n = make_double2(1.0, 1.0);
        int2 k_n = make_int2(0, 0);
        if (j < nfull) {
            r_n = ld_stream(rate2 + j);
            s_n = ld_stream(spike2 + j);
            if (nst2) k_n = ld_stream(nst2 + j);
        }
        while (j < nfull) {
            const double rate[2] = {r_n.x, r_n.y}, spike[2] = {s_n.x, s_n.y};
            const int32_t nst_raw[2] = {k_n.x, k_n.y};
            const int jn = j + BLOCK;
            if (jn < nfull) {  // prefetch the next strip
                r_n = ld_stream(rate2 + jn);
                s_n = ld_stream(spike2 + jn);
                if (nst2) k_n = ld_stream(nst2 + jn);
            }
            const int i0 = 2 * j;
            const double2 h2 = *reinterpret_cast<const double2*>(sh_h + i0);
            const double2 a2 = *reinterpret_cast<const double2*>(sh_aux + i0);
            const double h[2] = {h2.x, h2.y}, aux[2] = {a2.x, a2.y};
            bool da[2], pf[2], root[2];
            double hp[2], auxp[2], len[2], dA[2], E[2];
            bool sp = false;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t w = sh_w[i0 + k];
                const int p = static_cast<int>(w & W::IDX);
                da[k] = (w & W::DA) != 0;
                pf[k] = (w & W::PF) != 0;
                root[k] = (p == i0 + k);
                hp[k] = sh_h[p];
                auxp[k] = sh_aux[p];
                len[k] = hp[k] - h[k];                                      // a root points to itself: len == 0
                dA[k] = aux[k] - auxp[k];
                E[k] = fma(len[k], fb.kE, 2 * dA[k]);                       // STP:749-769 via aux
                sp = sp || gsm_fast_b_special(fb, i0 + k, da[k], pf[k], root[k], len[k], E[k], rate[k], spike[k], nst_raw[k]);
            }
            if (__any_sync(__activemask(), sp)) {
#pragma unroll
                for (int k = 0; k < 2; k++)
                    gsm_fast_pass_b_checked<WANT_RATES>(fb, K, T, g_logfact, A, i0 + k, da[k], pf[k], root[k], h[k], hp[k], aux[k],
                                                        auxp[k], rate[k], spike[k], nst_raw[k], orate ? orate + i0 + k : dummy,
                                                        owsp ? owsp + i0 + k : dummy);
            } else {
                gsm_fast_pass_b_pair<WANT_RATES>(fb, K, T, g_logfact, A, i0, da, pf, root, len, E, dA, rate, spike, nst_raw,
                                                 orate ? orate + i0 : dummy, owsp ? owsp + i0 : dummy);
            }
            j = jn;
        }
        if ((n & 1) && tid == (nfull % BLOCK)) {   // odd last node
            const int i = n - 1;
            const uint32_t w = sh_w[i];
            const int p = static_cast<int>(w & W::IDX);
            gsm_fast_pass_b_checked<WANT_RATES>(fb, K, T, g_logfact, A, i, (w & W::DA) != 0, (w & W::PF) != 0, p == i, sh_h[i], sh_h[p],
                                                sh_aux[i], sh_aux[p], __ldg(rate_row + i), __ldg(spike_row + i),
                                                nst_row ? __ldg(nst_row + i) : 0, orate ? orate + i : dummy, owsp ? owsp + i : dummy);
        }
    }

    // ---- phase 4: reductions (fixed order) + epilogue ----
    double v0 = warp_sum(A.tree), v1 = warp_sum(A.stub), v2 = warp_sum(A.spike), v3 = warp_sum(A.rate),
           v4 = warp_sum(A.burst), v5 = warp_sum(A.dist);
    int32_t c0 = __reduce_add_sync(0xffffffffu, A.n_extant), c1 = __reduce_add_sync(0xffffffffu, A.n_leaf),
            c2 = __reduce_add_sync(0xffffffffu, A.n_da), c3 = __reduce_add_sync(0xffffffffu, A.n_extinct),
            c4 = static_cast<int32_t>(__reduce_or_sync(0xffffffffu, static_cast<uint32_t>(A.bad)));
    if (lane == 0) {
        sh_red[0][warp] = v0; sh_red[1][warp] = v1; sh_red[2][warp] = v2;
        sh_red[3][warp] = v3; sh_red[4][warp] = v4; sh_red[5][warp] = v5;
        sh_redi[0][warp] = c0; sh_redi[1][warp] = c1; sh_redi[2][warp] = c2; sh_redi[3][warp] = c3; sh_redi[4][warp] = c4;
    }
    __syncthreads();
    if (tid == 0) {
        GsmAcc R;
        gsm_acc_zero(R);
#pragma unroll
        for (int w = 0; w < NWARP; w++) {
            R.tree += sh_red[0][w]; R.stub += sh_red[1][w]; R.spike += sh_red[2][w];
            R.rate += sh_red[3][w]; R.burst += sh_red[4][w]; R.dist += sh_red[5][w];
            R.n_extant += sh_redi[0][w]; R.n_leaf += sh_redi[1][w]; R.n_da += sh_redi[2][w];
            R.n_extinct += sh_redi[3][w]; R.bad |= sh_redi[4][w];
        }
        const int32_t root = sh_root;
        double res[GSM_NOUT];
        if (root < 0) {
#pragma unroll
            for (int k = 0; k < GSM_NOUT; k++) res[k] = NAN;
        } else {
            gsm_fast_fix_acc(K, R, n);
            gsm_finish_consts(K, sh_h[root]);
            gsm_finalize(K, R, n, (sh_w[root] & W::FAKE) != 0, res);
        }
#pragma unroll
        for (int k = 0; k < GSM_NOUT; k++) orow[k] = res[k];
    }
}

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
    return cudaMemcpyToSymbol(g_logfact, h, sizeof(h));
}

static size_t eval_smem_bytes(int64_t stride) { return static_cast<size_t>(stride) * 22u; }

template <int BLOCK, int MINB>
static cudaError_t launch_block(const GsmBatch& b, double* o, double* r, double* w, cudaStream_t st) {
    const size_t smem = eval_smem_bytes(b.stride);
    const dim3 grid(static_cast<unsigned>(b.n_trees));
    cudaError_t e;
    if (r || w) {
        auto kern = gsm_eval_kernel<BLOCK, MINB, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, o, r, w);
    } else {
        auto kern = gsm_eval_kernel<BLOCK, MINB, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, o, r, w);
    }
    return cudaGetLastError();
}

template <typename WORD, int BLOCK, int MINB>
static cudaError_t launch_fast(const GsmBatch& b, double* o, double* r, double* w, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(b.stride) * (16u + sizeof(WORD));
    const dim3 grid(static_cast<unsigned>(b.n_trees));
    cudaError_t e;
    if (r || w) {
        auto kern = gsm_eval_fast_kernel<WORD, BLOCK, MINB, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, o, r, w);
    } else {
        auto kern = gsm_eval_fast_kernel<WORD, BLOCK, MINB, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, st>>>(b, o, r, w);
    }
    return cudaGetLastError();
}

// Launch configuration: one CTA per tree.  The shared-memory footprint per node (18-22 B) bounds the number of
// resident trees per SM; BLOCK and the register budget (MINB) are picked from the measured sweeps in
// profiles/.  GSM_BLOCK / GSM_MINB / GSM_GENERIC override the choice for tuning runs.
int gsm_launch_eval(const GsmBatch& b, double* d_out_tree, double* d_out_rates, double* d_out_wspikes,
                    cudaStream_t stream) {
    if (b.n_trees <= 0) return 0;
    const bool fast = gsm_fast_path_ok(b.flags, b.stub_mode) && b.rate != nullptr && !getenv("GSM_GENERIC");
    int block, minb;
    if (fast) {
        block = b.n_nodes <= 512 ? 128 : 256;
        minb = b.n_nodes <= 512 ? 6 : 3;
    } else {
        block = b.n_nodes <= 512 ? 128 : (b.n_nodes <= 2048 ? 256 : 384);
        minb = b.n_nodes <= 512 ? 6 : (b.n_nodes <= 2048 ? 3 : 2);
    }
    if (const char* env = getenv("GSM_BLOCK")) block = atoi(env);
    if (const char* env = getenv("GSM_MINB")) minb = atoi(env);
    cudaError_t e = cudaErrorInvalidConfiguration;
    if (fast) {
        const bool w16 = b.n_nodes <= 4096;
#define GSM_CFG(B, M)                                                                                      \
    if (block == B && minb == M)                                                                           \
        e = w16 ? launch_fast<uint16_t, B, M>(b, d_out_tree, d_out_rates, d_out_wspikes, stream)           \
                : launch_fast<uint32_t, B, M>(b, d_out_tree, d_out_rates, d_out_wspikes, stream);
        GSM_CFG(128, 8) GSM_CFG(128, 6) GSM_CFG(128, 4)
        GSM_CFG(256, 4) GSM_CFG(256, 3) GSM_CFG(256, 2)
        GSM_CFG(384, 2) GSM_CFG(512, 2) GSM_CFG(512, 1)
#undef GSM_CFG
    } else {
#define GSM_CFG(B, M) \
    if (block == B && minb == M) e = launch_block<B, M>(b, d_out_tree, d_out_rates, d_out_wspikes, stream);
        GSM_CFG(128, 8) GSM_CFG(128, 6) GSM_CFG(128, 4)
        GSM_CFG(256, 4) GSM_CFG(256, 3) GSM_CFG(256, 2)
        GSM_CFG(384, 2) GSM_CFG(512, 2) GSM_CFG(512, 1)
#undef GSM_CFG
    }
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}

int gsm_launch_stub_cdf(const GsmBatch& b, int64_t tree, int32_t node, double* d_cdf, int32_t cap, int32_t* d_len,
                        cudaStream_t stream) {
    gsm_stub_cdf_kernel<<<1, 32, 0, stream>>>(b, tree, node, d_cdf, cap, d_len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    return 1;
}
