// gsm_lean_pk.cuh — persistent variant of the lean kernel (included by gsm_kernels.cu, after gsm_lean_kernel whose
// helpers it uses).
//
// gsm_lean_kernel runs one CTA per tree: a CTA spends ~3,000 of its ~30,000 cycles before the first node is evaluated
// (barriers, table copies, per-tree scalars, the latency of the first rows), its four warps take turns at running the
// request code of the ring, and the Gamma tables take two of the four warps for ~1,900 cycles.  Here a CTA stays on its SM
// and walks over trees blockIdx.x, blockIdx.x + gridDim.x, ...; it has two warpgroups:
//
//   consumers (warps 0-3, 136 registers)   pass A, pass B, reduction of one tree after the other; nothing else
//   front end (warps 4-7, 24 registers)    warp 4: every bulk copy of the CTA (the ring runs on across trees: the internal
//                                          heights of tree k + 1 are requested into the stages the last chunks of tree k
//                                          leave behind);  warp 5: per-tree scalars, tail-node rows, root position of
//                                          tree k + 1;  warps 6-7: its Gamma table
//
// The front end works one tree ahead: per-tree scalars into double-buffered shared memory (GsmPkFront; mbarrier kready hands
// them to the consumers, kfree hands the buffer back), Gamma table entries into the registers of their lanes, stored to the
// one ntab when the consumers are through with the previous tree (tready; the consumers need it from pass B on).  The log /
// exp tables are copied once per CTA.
// Same arithmetic, same summation order as gsm_lean_kernel<128, 3>: results are bit-identical to it.
//
// STATUS: measured SLOWER than one CTA per tree and therefore OFF by default (gsm_set_option("persistent", 1) selects it; the
// GPU parity tests run it).  C4 60,000 x 2,000 taxa: 2.33 ms against 2.13 ms; C3 2.11 against 1.74 ms.  What the phase counters
// and the ncu capture show (profiles/r02_tuning.md): the consumers' idle set-up time does go away (3,200 -> 900 cycles per tree)
// and pass A halves, but pass B gets 27 % slower and the end of a tree 2,000 cycles longer: with three always-busy consumer
// groups per SM the per-warp issue rate of pass B drops (1 CTA per SM: 13,000 cycles per pass B, 2: 15,000, 3: 17,300 -- the
// idle phases of one CTA per tree are what lets the other two run faster), the front-end warps at 24 registers spill their
// scalar / Lanczos code to local memory and end up on the critical path, and their polling is 45 % of the instructions issued.
#pragma once

#define GSM_PK_CONSUMERS 128
#define GSM_PK_THREADS 256
#define GSM_PK_CONSUMER_REGS 136
#define GSM_PK_FRONT_REGS 24


// wait of a front-end warp: it is ahead of the consumers most of the time, and a polling warp takes issue slots from them.
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes or the time is over.
template <int NS>
static __device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GSM_WAITH_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra GSM_DONEH_%=;\n"
        "bra GSM_WAITH_%=;\n"
        "GSM_DONEH_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(NS)
        : "memory");
}
#ifndef GSM_PK_TMA_DOORBELL
#define GSM_PK_TMA_DOORBELL 0
#endif
// How a front-end warp waits (measured on C4, 60,000 x 2,000 taxa, profiles/r02_tuning.md): plain polling 2.33 ms; named-barrier
// doorbells (GSM_PK_TMA_DOORBELL / GSM_PK_FE_DOORBELL) 2.72 - 3.03 ms; try_wait with a suspend-time hint (GSM_PK_HINT_*, ns)
// 2.94 ms whatever the hint: a parked warp wakes up too late for a ring that is two chunks deep.  Polling is the default.
#ifndef GSM_PK_FE_DOORBELL
#define GSM_PK_FE_DOORBELL 0
#endif
#ifndef GSM_PK_HINT_FAST
#define GSM_PK_HINT_FAST 0
#endif
#ifndef GSM_PK_HINT_SLOW
#define GSM_PK_HINT_SLOW 0
#endif
template <int NS>
static __device__ __forceinline__ void mbar_wait_fe(uint32_t bar, uint32_t parity) {
    if (NS > 0) mbar_wait_hint<NS>(bar, parity);
    else mbar_wait_a(bar, parity);
}

struct GsmPkFront {
    GsmTreeConsts K;
    int32_t n, skip, root_last, pad;
    double tail_d[3][3];    // height, rate, spike of the up to three tail nodes
    int32_t tail_i[3][2];   // parent, stub count
};
struct GsmPkShared {
    GsmPkFront fr[2];
    uint64_t mbar;                      // log / exp tables
    uint64_t full[GSM_RING_STAGES];     // ring: chunk landed
    uint64_t empty[GSM_RING_STAGES];    // ring: every consumer warp holds its share of the chunk in registers
    uint64_t kready[2], tready[2], kfree[2];
    int32_t root, nroot;
    double tail_s6[3][6];
    int32_t tail_c[3][6];
    uint32_t tail_u[3][4];
    double red[6][GSM_PK_CONSUMERS / 32];
    int32_t redi[7][GSM_PK_CONSUMERS / 32];
    uint32_t redu[5][GSM_PK_CONSUMERS / 32];
};

template <bool WANT_RATES>
__global__ void __launch_bounds__(GSM_PK_THREADS, 3) gsm_lean_pk_kernel(const GsmBatch b, GsmLeanRow* __restrict__ rows,
                                                                        int32_t* __restrict__ redo_count,
                                                                        int32_t* __restrict__ redo_count_next,
                                                                        int32_t* __restrict__ redo_list, double* __restrict__ out_rates,
                                                                        double* __restrict__ out_wspikes) {
    constexpr int BLOCK = GSM_PK_CONSUMERS;
    constexpr int NWARP = BLOCK / 32;
    constexpr int HB = BLOCK;                 // pairs per chunk
    constexpr int NST = GSM_RING_STAGES;
    constexpr uint32_t STAGE = HB * 64u;      // bytes per ring stage
    constexpr int HST = STAGE / 8;            // internal heights per stage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ GsmPkShared sh;
    const uint32_t a_dyn = smem_u32(smem_raw);
    if (rows == nullptr) {   // probe launch: where the dynamic segment starts
        if (blockIdx.x == 0 && threadIdx.x == 0) redo_count[0] = static_cast<int32_t>(a_dyn);
        return;
    }
    const int S = static_cast<int>(b.stride);
    const int SI = gsm_lean_si(b.n_nodes);
    const GsmLeanLayout lay = gsm_lean_layout(a_dyn, SI, NST * STAGE, true);
    uint64_t* sh_logtab = reinterpret_cast<uint64_t*>(smem_raw + lay.logtab);
    uint64_t* sh_exptab = reinterpret_cast<uint64_t*>(smem_raw + lay.exptab);
    double* sh_hs = reinterpret_cast<double*>(smem_raw + lay.arrays);
    double* sh_as = sh_hs + SI;
    unsigned char* sh_ring = smem_raw + lay.ring;
    GsmLeanN* sh_ntab = reinterpret_cast<GsmLeanN*>(smem_raw + lay.ntab);   // [64]
    uint8_t* sh_fk = smem_raw + lay.fk;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t T = b.n_trees;
    const int64_t tstep = gridDim.x;
    uint32_t dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    if (lay.total > dyn_bytes) {   // the launch did not provide this layout (cannot happen with launch_lean_pk): generic path
        if (tid == 0)
            for (int64_t t = blockIdx.x; t < T; t += tstep) lean_redo(redo_count, redo_list, rows, t);
        return;
    }
    const uint32_t a_ring = smem_u32(sh_ring), a_full = smem_u32(&sh.full[0]), a_empty = smem_u32(&sh.empty[0]);
    const uint32_t a_kready = smem_u32(&sh.kready[0]), a_tready = smem_u32(&sh.tready[0]), a_kfree = smem_u32(&sh.kfree[0]);
    if (tid == 0) {
        if (blockIdx.x == 0) *redo_count_next = 0;   // counter of the next evaluation on this handle (ping-pong)
        mbar_init(&sh.mbar, 1);
#pragma unroll
        for (int k = 0; k < NST; k++) {
            mbar_init(&sh.full[k], 1);
            mbar_init(&sh.empty[k], NWARP);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            mbar_init(&sh.kready[k], 1);
            mbar_init(&sh.tready[k], 2);
            mbar_init(&sh.kfree[k], NWARP);
        }
        fence_mbar_init();
        sh.root = -1;
        sh.nroot = 0;
        mbar_arrive_expect_tx(&sh.mbar, static_cast<uint32_t>(sizeof(GsmLeanBlob)));
        tma_bulk_g2s(sh_exptab, g_lean_blob.exptab, GSM_T6E_WORDS * 8u, &sh.mbar);
        tma_bulk_g2s(sh_logtab, g_lean_blob.logtab, GSM_T6_WORDS * 8u, &sh.mbar);
    }
    __syncthreads();

    // =====================================================================================================
    // front end
    // =====================================================================================================
    if (warp >= NWARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GSM_PK_FRONT_REGS));
        if (warp == NWARP) {
            // ---- every bulk copy of the CTA: per tree the internal heights (as many stages as they take), then its chunks ----
            {
                uint32_t s = 0u, ph = 0u, g = 0u;   // stage, parity of its previous use, chunks requested so far
                int k = 0;
#pragma unroll 1
                for (int64_t t = blockIdx.x; t < T; t += tstep, ++k) {
                    const uint32_t bf = static_cast<uint32_t>(k) & 1u;
                    mbar_wait_a(a_kready + bf * 8u, static_cast<uint32_t>(k >> 1) & 1u);
                    const GsmPkFront& f = sh.fr[bf];
                    if (f.skip) continue;
                    const int n = f.n;
                    const int64_t row = t * S;
                    const int L = (n + 1) >> 1, L4 = L & ~3;
                    LeanRing rg;
                    rg.a_ring = a_ring;
                    rg.a_full = a_full;
                    rg.a_empty = a_empty;
                    rg.rate = b.rate + row;
                    rg.spike = b.spike + row;
                    rg.height = b.height + row;
                    rg.nst = b.stub_mode != GSM_STUBS_NOSTUB ? b.nstubs + row : nullptr;
                    rg.parent = b.parent + row;
                    rg.L = L;
                    rg.L4 = L4;
                    rg.nlc = ((L >> 1) + HB - 1) / HB;
                    rg.nvc = rg.nlc + ((n - 1 - L4 + 1) / 2 + HB - 1) / HB;
                    rg.end_leaf = (L + 3) & ~3;
                    rg.end_int = (n - 1 + 3) & ~3;
                    const uint32_t bytes_h = (static_cast<uint32_t>(n - L4) * 8u + 15u) & ~15u;
                    const int nh = static_cast<int>((bytes_h + STAGE - 1u) / STAGE);
                    const int nv = nh + rg.nvc;
#pragma unroll 1
                    for (int j = 0; j < nv; ++j) {
                        if (g >= static_cast<uint32_t>(NST)) {
                            // sleep until every consumer warp has rung the bell of this stage, then the mbarrier (ordering)
#if GSM_PK_TMA_DOORBELL
                            asm volatile("bar.sync %0, %1;" ::"r"(GSM_PK_BAR_STAGE + s), "n"(BLOCK + 32) : "memory");
                            if (lane == 0) mbar_wait_a(a_empty + s * 8u, ph);
#else
                            if (lane == 0) {
                                mbar_wait_fe<GSM_PK_HINT_FAST>(a_empty + s * 8u, ph);
                            }
#endif
                        }
                        if (lane != 0) {
                        } else if (j < nh) {
                            const uint32_t o = static_cast<uint32_t>(j) * STAGE;
                            const uint32_t bytes = bytes_h - o < STAGE ? bytes_h - o : STAGE;
                            const uint32_t bar = a_full + s * 8u;
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                            tma_bulk_g2s_a(a_ring + s * STAGE, b.height + row + L4 + j * HST, bytes, bar);
                        } else {
                            ring_issue<2 * HB>(rg, j - nh, s);
                        }
                        ++g;
                        if (++s == static_cast<uint32_t>(NST)) {
                            s = 0u;
                            if (g > static_cast<uint32_t>(NST)) ph ^= 1u;
                        }
                    }
                }
            }
        } else if (warp == NWARP + 1) {
            // ---- per-tree scalars, tail-node rows, root position ----
            int k = 0;
#pragma unroll 1
            for (int64_t t = blockIdx.x; t < T; t += tstep, ++k) {
                const uint32_t bf = static_cast<uint32_t>(k) & 1u;
                if (k >= 2) {
#if GSM_PK_FE_DOORBELL
                    asm volatile("bar.sync %0, %1;" ::"r"(GSM_PK_BAR_SCAL + bf), "n"(BLOCK + 32) : "memory");
                    mbar_wait_a(a_kfree + bf * 8u, static_cast<uint32_t>((k >> 1) - 1) & 1u);
#else
                    mbar_wait_fe<GSM_PK_HINT_SLOW>(a_kfree + bf * 8u, static_cast<uint32_t>((k >> 1) - 1) & 1u);
#endif
                }
                GsmPkFront& f = sh.fr[bf];
                int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
                if (n > b.n_nodes) n = b.n_nodes;
                const bool valid = (n & 1) && n >= 3 && n <= GSM_LEAN_MAX_NODES;   // a binary BEAST tree the lean path covers
                const int64_t row = t * S;
                if (lane == 0) {
                    double prm_v[GSM_NPARAM];
#pragma unroll
                    for (int q = 0; q < GSM_NPARAM; q++) prm_v[q] = __ldg(b.params + t * GSM_NPARAM + q);
                    const int32_t use_v = b.use_spike ? static_cast<int32_t>(__ldg(b.use_spike + t)) : 1;
                    gsm_setup_scalars(f.K, prm_v, use_v, b.flags, b.stub_mode);
                    // invalid states (lambda <= mu, shape <= 0, ...) and the mixture-mode wiring go to the generic path
                    const bool ok = valid && gsm_lean_consts_ok(f.K) && f.K.spike_known_n != 0;
                    f.n = n;
                    f.skip = ok ? 0 : 1;
                    f.root_last = (valid && __ldg(b.parent + row + n - 1) < 0) ? 1 : 0;
                } else if (valid && lane >= 29) {
                    const int L = (n + 1) >> 1;
                    int node = -1;
                    if (lane == 31) node = n - 1;
                    else if (L & 1) node = lane == 29 ? L - 1 : L;
                    if (node >= 0) {
                        const int q = lane - 29;
                        f.tail_d[q][0] = __ldg(b.height + row + node);
                        f.tail_d[q][1] = __ldg(b.rate + row + node);
                        f.tail_d[q][2] = __ldg(b.spike + row + node);
                        f.tail_i[q][0] = __ldg(b.parent + row + node);
                        f.tail_i[q][1] = b.stub_mode != GSM_STUBS_NOSTUB ? __ldg(b.nstubs + row + node) : 0;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_kready + bf * 8u);
            }
        } else {
            // ---- Gamma table of the tree: ntab[n] = {log shape - logGamma(shape (n + 1)), -log n!, shape (n + 1) - 1, n} ----
            // (BSP:97-99 with the Lanczos logGamma of commons-math; STP:567)
            mbar_wait(&sh.mbar, 0);
            GsmT6 t6;
            t6.log_ar = smem_u32(sh_logtab) | (static_cast<uint32_t>(lane & 15) * 8u);
            t6.exp_ar = smem_u32(sh_exptab) | (static_cast<uint32_t>(lane & 7) * 8u);
            const int m = (warp - (NWARP + 2)) * 32 + lane + 1;   // 1..64
            int k = 0;
#pragma unroll 1
            for (int64_t t = blockIdx.x; t < T; t += tstep, ++k) {
                const uint32_t bf = static_cast<uint32_t>(k) & 1u;
                mbar_wait_a(a_kready + bf * 8u, static_cast<uint32_t>(k >> 1) & 1u);
                const GsmPkFront& f = sh.fr[bf];
                const bool skip = f.skip != 0;
                GsmLeanN q;
                if (!skip) {
                    const double shape = f.K.shape;
                    q.w = shape * m - 1;
                    q.gcon = gsm_log6_hi(t6, shape) - lean_log_gamma(t6, shape * m);
                    q.nlf = -g_logfact[m - 1];
                    q.nd = static_cast<double>(m - 1);
                }
                // the consumers are through with the table of the previous tree
                if (k >= 1) {
#if GSM_PK_FE_DOORBELL
                    asm volatile("bar.sync %0, %1;" ::"r"(GSM_PK_BAR_GAMMA + (bf ^ 1u)), "n"(BLOCK + 64) : "memory");
                    mbar_wait_a(a_kfree + (bf ^ 1u) * 8u, static_cast<uint32_t>((k - 1) >> 1) & 1u);
#else
                    mbar_wait_fe<GSM_PK_HINT_SLOW>(a_kfree + (bf ^ 1u) * 8u, static_cast<uint32_t>((k - 1) >> 1) & 1u);
#endif
                }
                if (!skip) sh_ntab[m - 1] = q;
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_tready + bf * 8u);
            }
        }
        return;
    }

    // =====================================================================================================
    // consumers
    // =====================================================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GSM_PK_CONSUMER_REGS));
    auto csync = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(BLOCK) : "memory"); };
    auto csync_or = [&](bool pred) -> bool {
        uint32_t r;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, 2, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(r)
                     : "r"(pred ? 1u : 0u), "n"(BLOCK)
                     : "memory");
        return r != 0u;
    };
    mbar_wait(&sh.mbar, 0);   // log / exp tables
#ifdef GSM_PK_STAGGER
    {   // the CTAs of an SM start their first trees a third of a tree apart instead of in lockstep
        const long long t_start = clock64();
        const long long wait_c = static_cast<long long>((blockIdx.x / 148u) % 3u) * GSM_PK_STAGGER;
        while (clock64() - t_start < wait_c) __nanosleep(200);
    }
#endif
    GsmT6 t6;
    t6.log_ar = smem_u32(sh_logtab) | (static_cast<uint32_t>(lane & 15) * 8u);
    t6.exp_ar = smem_u32(sh_exptab) | (static_cast<uint32_t>(lane & 7) * 8u);
    uint32_t ring_s = 0u, ring_ph = 0u;   // stage of the next chunk of the ring and the parity of its full barrier
    const bool tail_lane = warp == 1 && lane >= 29;
    int kt = 0;
    GSM_PHASE_START();
#pragma unroll 1
    for (int64_t t = blockIdx.x; t < T; t += tstep, ++kt) {
        const uint32_t bf = static_cast<uint32_t>(kt) & 1u;
        mbar_wait_a(a_kready + bf * 8u, static_cast<uint32_t>(kt >> 1) & 1u);
        GSM_PHASE_MARK(0);   // waiting for the front end
        const GsmPkFront& f = sh.fr[bf];
        const GsmTreeConsts& K = f.K;
        if (f.skip) {
            if (tid == 0) lean_redo(redo_count, redo_list, rows, t);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_kfree + bf * 8u);
#if GSM_PK_FE_DOORBELL
            if (t + 2 * tstep < T) asm volatile("bar.arrive %0, %1;" ::"r"(GSM_PK_BAR_SCAL + bf), "n"(BLOCK + 32) : "memory");
            if (t + tstep < T) asm volatile("bar.arrive %0, %1;" ::"r"(GSM_PK_BAR_GAMMA + bf), "n"(BLOCK + 64) : "memory");
#endif
            continue;
        }
        const int n = f.n;
        const int64_t row = t * S;
        const int L = (n + 1) >> 1;
        const int L4 = L & ~3;
        LeanRing rg;
        rg.a_ring = a_ring;
        rg.a_full = a_full;
        rg.a_empty = a_empty;
        rg.rate = nullptr;
        rg.spike = nullptr;
        rg.height = nullptr;
        rg.nst = b.stub_mode != GSM_STUBS_NOSTUB ? b.nstubs : nullptr;   // only "are there stub counts" is used here
        rg.parent = nullptr;
        rg.L = L;
        rg.L4 = L4;
        rg.nlc = ((L >> 1) + HB - 1) / HB;
        rg.nvc = rg.nlc + ((n - 1 - L4 + 1) / 2 + HB - 1) / HB;
        rg.end_leaf = (L + 3) & ~3;
        rg.end_int = (n - 1 + 3) & ~3;

        GsmLeanC c;
        gsm_lean_consts(K, c, n);
        c.t6 = t6;
        c.mtab = nullptr;
        c.ntab = sh_ntab;
        c.mixtab = nullptr;
        LeanBCtx x;
        x.a_hs = smem_u32(sh_hs) - static_cast<uint32_t>(L4) * 8u;
        x.a_as = smem_u32(sh_as) - static_cast<uint32_t>(L4) * 8u;
        x.a_fk = smem_u32(sh_fk) - static_cast<uint32_t>(L4);
        x.a_hL = x.a_hs + static_cast<uint32_t>(L) * 8u;
        x.a_aL = x.a_as + static_cast<uint32_t>(L) * 8u;
        x.a_fL = x.a_fk + static_cast<uint32_t>(L);
        x.chk_p = 0u;
        x.nL = static_cast<uint32_t>(n - L);
        // Sampled ancestors are expected where the tree prior generates them (psi > 0, origin / root conditioning): those trees
        // are scanned.  Elsewhere a zero-length leaf is caught by the length check of the CLEAN bodies or, in a GENERAL body
        // (root not last, tail nodes), by c.sa_unscanned, and the tree goes to the generic path.
        const bool detect = !(c.leaf_term == 0 && K.psi == 0.0);
        c.sa_unscanned = !detect;
        const bool root_last = f.root_last != 0;
        if (detect || !root_last)
            for (int q = tid; q < (SI + 3) >> 2; q += BLOCK) reinterpret_cast<uint32_t*>(sh_fk)[q] = 0u;

        GSM_PHASE_MARK(1);   // per-tree set-up of the consumers
        // ---- pass A: aux of the internal nodes; their heights arrive through the ring ----
        GsmLeanAccA SA;
        gsm_lean_acc_a_zero(SA);
        {
            const int npa = (n - L4 + 1) >> 1;
            const uint32_t a_h0 = smem_u32(sh_hs), a_a0 = smem_u32(sh_as);
#pragma unroll 1
            for (int j0 = 0; j0 < npa; j0 += HST / 2) {
                mbar_wait_a(a_full + ring_s * 8u, ring_ph);
                const uint32_t a_st = a_ring + ring_s * STAGE;
                const int cnt = npa - j0 < HST / 2 ? npa - j0 : HST / 2;
#pragma unroll 1
                for (int q = tid; q < cnt; q += 2 * BLOCK) {
                    const int q2 = q + BLOCK;
                    if (q2 < cnt) {   // two pairs = four nodes, staged (gsm_lean_aux4)
                        const double2 ha = lds_v2f64(a_st + q * 16);
                        double2 hb = lds_v2f64(a_st + q2 * 16);
                        if (L4 + 2 * (j0 + q2) + 1 >= n) hb.y = hb.x;   // padding slot of the last pair
                        const double h4[4] = {ha.x, ha.y, hb.x, hb.y};
                        double a4[4];
                        gsm_lean_aux4(c, SA, h4, a4);
                        sts_v2f64(a_h0 + (j0 + q) * 16, ha.x, ha.y);
                        sts_v2f64(a_h0 + (j0 + q2) * 16, hb.x, hb.y);
                        sts_v2f64(a_a0 + (j0 + q) * 16, a4[0], a4[1]);
                        sts_v2f64(a_a0 + (j0 + q2) * 16, a4[2], a4[3]);
                    } else {
                        double2 h2 = lds_v2f64(a_st + q * 16);
                        if (L4 + 2 * (j0 + q) + 1 >= n) h2.y = h2.x;
                        double e0, G0, e1, G1;
                        const double a0 = gsm_lean_aux(c, SA, h2.x, e0, G0);
                        const double a1 = gsm_lean_aux(c, SA, h2.y, e1, G1);
                        sts_v2f64(a_h0 + (j0 + q) * 16, h2.x, h2.y);
                        sts_v2f64(a_a0 + (j0 + q) * 16, a0, a1);
                    }
                }
                __syncwarp();
#if GSM_PK_TMA_DOORBELL
                asm volatile("bar.arrive %0, %1;" ::"r"(GSM_PK_BAR_STAGE + ring_s), "n"(BLOCK + 32) : "memory");
#endif
                if (lane == 0) mbar_arrive_a(a_empty + ring_s * 8u);
                if (++ring_s == static_cast<uint32_t>(NST)) {
                    ring_s = 0u;
                    ring_ph ^= 1u;
                }
            }
        }
        {
            GsmLeanAccA S0;
            gsm_lean_acc_a_zero(S0);
            x.auxz = gsm_lean_aux(c, S0, 0.0, x.ez, x.Gz);
        }
        GSM_PHASE_MARK(2);   // pass A (thread 0's share, waits for the heights included)
        csync();   // heights and aux of every internal node are in shared memory (and the fake flags are cleared)
        bool general = !root_last;
        if (detect) {
            // Node.isDirectAncestor: a leaf at its parent's height -> the parent isFake.  Leaf heights and parents straight from
            // global memory (the ring will read them again from L2).
            bool anyda = false;
            const double* hrow = b.height + row;
            const int32_t* prow = b.parent + row;
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
            if (csync_or(anyda)) general = true;
        }

        // One tail node, evaluated by its lane with the GENERAL flavour after pass B; the combined sums go to shared memory.
        int tail_node = -1;
        if (tail_lane) {
            if (lane == 31) tail_node = n - 1;
            else if (L & 1) tail_node = lane == 29 ? L - 1 : L;
        }
        auto tail_eval = [&](bool gen) {
            const int i = tail_node;
            const int q = lane - 29;
            GsmLeanAccB ST;
            gsm_lean_acc_b_zero(ST);
            GsmLeanAccA STA;
            gsm_lean_acc_a_zero(STA);
            uint32_t chk_p = 0u;
            const double h = f.tail_d[q][0];
            int32_t p = f.tail_i[q][0];
            const bool root = p < 0;
            if (root) {
                atomicAdd(&sh.nroot, 1);
                sh.root = i;
                if (i < L) ST.chk = 0xffffffffu;
                p = i >= L ? i : L;
            }
            chk_p = static_cast<uint32_t>(p - L);
            if (!(static_cast<uint32_t>(p - L) < x.nL)) p = L;
            double e = 0.0, G = 0.0, aux;
            if (i < L) aux = gsm_lean_aux(c, STA, h, e, G);
            else aux = sh_as[i - L4];
            double hp = sh_hs[p - L4], auxp;
            if (root) { hp = h; auxp = aux; }
            else auxp = sh_as[p - L4];
            const bool pf = gen && sh_fk[p - L4] != 0, fs = gen && i >= L && sh_fk[i - L4] != 0;
            if (i < L) {
                const bool da = hp == h && !root;
                if (c.leaf_term == 1) gsm_lean_leaf_term<1>(c, STA, !da, h, aux, e, G);
                else if (c.leaf_term == 2) gsm_lean_leaf_term<2>(c, STA, !da, h, aux, e, G);
            }
            int32_t nst = -1;
            if (c.use_counts) nst = (i < n - 1 || b.stub_mode == GSM_STUBS_EXACT) ? f.tail_i[q][1] : 0;
            double er, ew;
            gsm_lean_b_node<true, WANT_RATES, 2, false, false>(c, ST, i, root, pf, fs, h, hp, aux, auxp, f.tail_d[q][1], f.tail_d[q][2], nst, &er,
                                                               &ew);
            if (WANT_RATES) {
                if (out_rates) out_rates[row + i] = er;
                if (out_wspikes) out_wspikes[row + i] = ew;
            }
            GsmLean6 t6r;
            gsm_lean_combine<true>(K, c, STA, ST, t6r);
            double* o = sh.tail_s6[q];
            o[0] = t6r.tree; o[1] = t6r.stub; o[2] = t6r.spike; o[3] = t6r.rate; o[4] = t6r.burst; o[5] = t6r.dist;
            int32_t* oc = sh.tail_c[q];
            oc[0] = STA.n_leafterm; oc[1] = STA.n_plain; oc[2] = ST.n_fake; oc[3] = ST.n_extant; oc[4] = ST.n_extinct; oc[5] = ST.n_da;
            uint32_t* ou = sh.tail_u[q];
            ou[0] = gsm_umax(STA.chk, ST.chk); ou[1] = ST.chk_n; ou[2] = STA.hmax; ou[3] = chk_p;
        };

        GSM_PHASE_MARK(3);   // waiting for the other warps, sampled-ancestor scan
        mbar_wait_a(a_tready + bf * 8u, static_cast<uint32_t>(kt >> 1) & 1u);   // the Gamma table of this tree
        // ---- pass B ----
        GsmLeanAccB SB;
        gsm_lean_acc_b_zero(SB);
        double2* orate2 = (WANT_RATES && out_rates) ? reinterpret_cast<double2*>(out_rates + row) : nullptr;
        double2* owsp2 = (WANT_RATES && out_wspikes) ? reinterpret_cast<double2*>(out_wspikes + row) : nullptr;
        const bool stdw = gsm_lean_std_wiring(c);
        int32_t* rn = &sh.root;   // {root, nroot}
#define GSM_PASS_B(G, W, LT) lean_pass_b<G, WANT_RATES, W, LT, BLOCK, 1, NST, false, 2, 0>(c, SA, SB, x, rg, n, tid, rn, orate2, owsp2, ring_s, ring_ph)
#define GSM_PASS_B3(G, LT) { if (stdw) GSM_PASS_B(G, true, LT); else GSM_PASS_B(G, false, LT); }
        if (general) {
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
        // ---- tail: node n-1 (no stub count in COUNTS mode, Stubs:354; the root of an ordinary tree) and, when L is odd, the
        // pair (L-1, L) that straddles the leaf / internal boundary ----
        GSM_PHASE_MARK(4);   // pass B (thread 0's share)
        const bool tail_on = tail_lane && tail_node >= 0;
        if (tail_on) tail_eval(general);
        GsmLean6 s6;
        if (general) gsm_lean_combine<true>(K, c, SA, SB, s6);
        else gsm_lean_combine<false>(K, c, SA, SB, s6);
        int32_t t_leafterm = 0, t_plain = 0, t_fake = 0, t_extant = 0, t_extinct = 0, t_da = 0;
        uint32_t t_chk = 0u, t_chk_n = 0u, t_hmax = 0u;
        if (tail_on) {
            const int q = lane - 29;
            const double* o = sh.tail_s6[q];
            s6.tree += o[0]; s6.stub += o[1]; s6.spike += o[2]; s6.rate += o[3]; s6.burst += o[4]; s6.dist += o[5];
            const int32_t* oc = sh.tail_c[q];
            t_leafterm = oc[0]; t_plain = oc[1]; t_fake = oc[2]; t_extant = oc[3]; t_extinct = oc[4]; t_da = oc[5];
            const uint32_t* ou = sh.tail_u[q];
            t_chk = ou[0]; t_chk_n = ou[1]; t_hmax = ou[2];
            x.chk_p = gsm_umax(x.chk_p, ou[3]);
        }

        // ---- reductions (fixed order) ----
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
        GSM_PHASE_MARK(5);   // combine + warp reductions
        csync();   // every warp is through with pass B: the arrays may be overwritten once warp 0 has read the root's slots
        GSM_PHASE_MARK(6);   // waiting for the other warps (pass B, tail nodes)
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
            const int nroot = sh.nroot;
            bool bad = chk >= GSM_CHK_LIMIT || (c.use_counts && chk_n > static_cast<uint32_t>(GSM_LEAN_MAX_NST)) ||
                       (!general && chk_len >= GSM_CHK_LIMIT) || hmax >= 0x7ff00000u || chk_p >= x.nL || nroot != 1 || root < L;
            if (!bad) bad = !(K.c1 * __hiloint2double(static_cast<int>(hmax) + 1, 0) < 700.0);   // exp(-c1 h) stays in range
            // the root's slots of the arrays (an ordinary tree's root is the tail node n - 1)
            double root_h = 0.0;
            int32_t root_fake = 0;
            if (!bad) {
                root_h = general ? sh_hs[root - L4] : f.tail_d[2][0];
                root_fake = (general && sh_fk[root - L4] != 0) ? 1 : 0;
            }
            __syncwarp();
            if (lane == 0) {
                sh.root = -1;
                sh.nroot = 0;
            }
            if (bad) {
                if (lane == 0) lean_redo(redo_count, redo_list, rows, t);
            } else {
                GsmLeanRow* r = rows + t;
                double* rd = reinterpret_cast<double*>(r);
                int32_t* ri = reinterpret_cast<int32_t*>(rd + 7);
                if (lane < 6) rd[lane] = dv;
                else if (lane == 6) rd[6] = root_h;
                else if (lane >= 8 && lane < 15) {
                    // redi order: n_nst, n_leafterm, n_plain, n_fake, n_extant, n_extinct, n_da  == row order
                    ri[lane - 8] = iv;
                } else if (lane == 15) {
                    ri[7] = root_fake;
                    ri[8] = 1;
                }
            }
        }
        if (general) csync();   // (warp 0 has read the root's slots of the arrays)
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_kfree + bf * 8u);
        // wake the front-end warps that wait for this tree to be over (they exist when the trees they prepare do)
#if GSM_PK_FE_DOORBELL
        if (t + 2 * tstep < T) asm volatile("bar.arrive %0, %1;" ::"r"(GSM_PK_BAR_SCAL + bf), "n"(BLOCK + 32) : "memory");
        if (t + tstep < T) asm volatile("bar.arrive %0, %1;" ::"r"(GSM_PK_BAR_GAMMA + bf), "n"(BLOCK + 64) : "memory");
#endif
    }
}
