// gsm_lean_warp.cuh — the lean path for SMALL trees: one WARP per tree (included by gsm_kernels.cu after gsm_lean_kernel,
// whose node functions it uses).
//
// gsm_lean_kernel gives a tree to a CTA of four warps: for a 200-taxon tree (399 nodes, BASELINE config 2) the two chunks of
// pass B are ~2,500 of the ~15,000 cycles the CTA is resident -- the rest is per-tree set-up, the 64-entry Gamma table, four
// CTA barriers, the cross-warp reduction and the launch of the CTA itself, with four CTAs per SM to overlap them.  Here a
// resident CTA of twelve warps (one per SM, 168 registers) walks over the trees with one tree per warp: no CTA barrier after
// the tables are loaded, the log / exp tables are copied once per CTA, a tree's streamed rows are read straight from global
// memory / L2 with 128-bit loads one iteration ahead (a warp's 32 node pairs are 512 contiguous bytes per array), only what
// is gathered at random (internal heights, aux, fake flags) and the per-tree tables live in the warp's 8 KB of shared
// memory, and the Gamma table is built only as far as the tree's largest stub count needs it.
// Both spike-prior flavours (known stub count; MIXK: the Poisson mixture of the templates' wiring), as in the CTA kernel.
// Same node arithmetic as gsm_lean_kernel (lean_b_pair / gsm_lean_b_node / gsm_lean_combine); the partition of the nodes over
// threads differs, so rows agree with the CTA kernel to rounding (1e-13), not bit for bit.
#pragma once

#ifndef GSM_WARP_MAX_NODES
#define GSM_WARP_MAX_NODES 511      // trees of up to 256 taxa
#endif
#ifndef GSM_WARP_WARPS
#define GSM_WARP_WARPS 12           // warps per CTA (one CTA per SM: 12 x 32 x 168 registers)
#endif

struct __align__(16) GsmWarpTree {  // a warp's shared memory, fixed part; the height / aux / fake-flag arrays follow
    GsmTreeConsts K;
    GsmLeanXPair xp;
    int32_t root, nroot;
    int32_t pad[2];
    GsmLogEntry mtab[2];            // mixture mode: {shape m - 1, log shape - logGamma(shape m)} for m = 0, 1 (the only ones it uses)
    GsmLeanN ntab[GSM_LEAN_MAX_NST + 1];   // mixture mode: overlaid by the Gamma-ratio tables (128 x 16 bytes), as in the CTA kernel
};
__host__ __device__ inline uint32_t gsm_warp_fixed_bytes() { return (static_cast<uint32_t>(sizeof(GsmWarpTree)) + 15u) & ~15u; }
__host__ __device__ inline uint32_t gsm_warp_bytes(int si) {
    return gsm_warp_fixed_bytes() + static_cast<uint32_t>(si) * 16u + ((static_cast<uint32_t>(si) + 15u) & ~15u);
}
// dynamic shared memory of a CTA: room for the 8 KB alignment of the log table + both tables + the warps
__host__ __device__ inline uint32_t gsm_warp_cta_bytes(int si, int warps) {
    return 8192u + (GSM_T6_WORDS + GSM_T6E_WORDS) * 8u + static_cast<uint32_t>(warps) * gsm_warp_bytes(si);
}

// one iteration's streamed values of a node pair
struct LeanWarpRegs {
    double2 rr, ss, h2;
    int2 kk, pp;
};
template <int LEAF>
static __device__ __forceinline__ void lean_warp_load(LeanWarpRegs& r, bool in, const double* rrow, const double* srow, const double* hrow,
                                                      const int32_t* nrow, const int32_t* prow, int i0) {
    r.rr = make_double2(1.0, 1.0);
    r.ss = make_double2(0.0, 0.0);
    r.h2 = make_double2(0.0, 0.0);
    r.kk = make_int2(-1, -1);   // without stub counts getNStubsOnBranch == -1 (Stubs:348-350)
    r.pp = make_int2(0, 0);
    if (in) {
        r.rr = ld_stream(reinterpret_cast<const double2*>(rrow + i0));
        r.ss = ld_stream(reinterpret_cast<const double2*>(srow + i0));
        if (LEAF == 1) r.h2 = ld_stream(reinterpret_cast<const double2*>(hrow + i0));
        if (nrow) r.kk = ld_stream(reinterpret_cast<const int2*>(nrow + i0));
        r.pp = ld_stream(reinterpret_cast<const int2*>(prow + i0));
    }
}

// pass B of one tree by one warp: leaf pairs (i0, i0 + 1 < L), internal pairs from L4 on (L <= i0, i0 + 1 < n - 1); the loads of
// the next 32 pairs are issued before the bodies of these
template <bool GENERAL, bool WANT_RATES, bool STDW, int LT, int XPD, bool MIX>
static __device__ __forceinline__ void lean_warp_pass_b(const GsmLeanC& c, GsmLeanAccA& SA, GsmLeanAccB& SB, LeanBCtx& x, uint32_t a_base, int n,
                                                        int L, int L4, int lane, const double* rrow, const double* srow, const double* hrow,
                                                        const int32_t* nrow, const int32_t* prow, int32_t* root_nroot,
                                                        double2* __restrict__ orate2, double2* __restrict__ owsp2) {
    {
        LeanWarpRegs cur, nxt;
        int i0 = 2 * lane;
        lean_warp_load<1>(cur, i0 + 1 < L, rrow, srow, hrow, nrow, prow, i0);
#pragma unroll 1
        for (int base = 0; base < L; base += 64) {
            const int i1 = i0 + 64;
            lean_warp_load<1>(nxt, base + 64 < L && i1 + 1 < L, rrow, srow, hrow, nrow, prow, i1);
            lean_b_pair<GENERAL, WANT_RATES, 1, STDW, LT, MIX, XPD>(c, SA, SB, x, a_base, i0, i0 + 1 < L, cur.rr, cur.ss, cur.h2, cur.kk, cur.pp,
                                                                      root_nroot, orate2, owsp2);
            cur = nxt;
            i0 = i1;
        }
    }
    {
        LeanWarpRegs cur, nxt;
        const int hi = n - 1;
        int i0 = L4 + 2 * lane;
        lean_warp_load<0>(cur, i0 >= L && i0 + 1 < hi, rrow, srow, hrow, nrow, prow, i0);
#pragma unroll 1
        for (int base = L4; base + 1 < hi; base += 64) {
            const int i1 = i0 + 64;
            lean_warp_load<0>(nxt, base + 65 < hi && i1 + 1 < hi, rrow, srow, hrow, nrow, prow, i1);
            lean_b_pair<GENERAL, WANT_RATES, 0, STDW, LT, MIX, XPD>(c, SA, SB, x, a_base, i0, i0 >= L && i0 + 1 < hi, cur.rr, cur.ss, cur.h2, cur.kk,
                                                                      cur.pp, root_nroot, orate2, owsp2);
            cur = nxt;
            i0 = i1;
        }
    }
}

// MIXK: the mixture-mode spike prior (BSP:102-172, Stubs nostub mode) is a kernel of its own, as for the CTA kernel
template <bool WANT_RATES, bool MIXK>
__global__ void __launch_bounds__(32 * GSM_WARP_WARPS, 1) gsm_lean_warp_kernel(const GsmBatch b, GsmLeanRow* __restrict__ rows,
                                                                               int32_t* __restrict__ redo_count,
                                                                               int32_t* __restrict__ redo_count_next,
                                                                               int32_t* __restrict__ redo_list, double* __restrict__ out_rates,
                                                                               double* __restrict__ out_wspikes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarp = static_cast<int>(blockDim.x >> 5);
    const int S = static_cast<int>(b.stride);
    const int SI = gsm_lean_si(b.n_nodes);
    const uint32_t a_dyn = smem_u32(smem_raw);
    const uint32_t gap = (8192u - (a_dyn & 8191u)) & 8191u;   // the log table sits on an 8 KB boundary of the shared window
    uint64_t* sh_logtab = reinterpret_cast<uint64_t*>(smem_raw + gap);
    uint64_t* sh_exptab = sh_logtab + GSM_T6_WORDS;
    unsigned char* wbase = smem_raw + gap + (GSM_T6_WORDS + GSM_T6E_WORDS) * 8u + static_cast<uint32_t>(warp) * gsm_warp_bytes(SI);
    GsmWarpTree& W = *reinterpret_cast<GsmWarpTree*>(wbase);
    double* sh_hs = reinterpret_cast<double*>(wbase + gsm_warp_fixed_bytes());
    double* sh_as = sh_hs + SI;
    uint8_t* sh_fk = reinterpret_cast<uint8_t*>(sh_as + SI);
    uint32_t dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    const int64_t T = b.n_trees;
    const int64_t tstep = static_cast<int64_t>(gridDim.x) * nwarp;
    if (gsm_warp_cta_bytes(SI, nwarp) > dyn_bytes || b.n_nodes > GSM_WARP_MAX_NODES) {   // cannot happen with launch_lean_warp
        if (lane == 0)
            for (int64_t t = static_cast<int64_t>(blockIdx.x) * nwarp + warp; t < T; t += tstep) lean_redo(redo_count, redo_list, rows, t);
        return;
    }
    if (tid == 0) {
        if (blockIdx.x == 0) *redo_count_next = 0;   // counter of the next evaluation on this handle (ping-pong)
        mbar_init(&mbar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&mbar, static_cast<uint32_t>(sizeof(GsmLeanBlob)));
        tma_bulk_g2s(sh_logtab, g_lean_blob.logtab, GSM_T6_WORDS * 8u, &mbar);
        tma_bulk_g2s(sh_exptab, g_lean_blob.exptab, GSM_T6E_WORDS * 8u, &mbar);
    }
    __syncthreads();
    mbar_wait(&mbar, 0);
    GsmT6 t6;
    t6.log_ar = smem_u32(sh_logtab) | (static_cast<uint32_t>(lane & 15) * 8u);
    t6.exp_ar = smem_u32(sh_exptab) | (static_cast<uint32_t>(lane & 7) * 8u);
    const uint32_t a_w = smem_u32(wbase);
    constexpr int XPD = static_cast<int>(offsetof(GsmWarpTree, xp));
    static_assert(XPD != 0, "hand-over slots");

#pragma unroll 1
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * nwarp + warp; t < T; t += tstep) {
        __syncwarp();
        int32_t n = b.node_count ? b.node_count[t] : b.n_nodes;
        if (n > b.n_nodes) n = b.n_nodes;
        const bool valid = (n & 1) && n >= 3 && n <= GSM_WARP_MAX_NODES;   // a binary BEAST tree this path covers
        const int64_t row = t * S;
        GsmTreeConsts& K = W.K;
        if (lane == 0 && valid) {
            double prm_v[GSM_NPARAM];
#pragma unroll
            for (int q = 0; q < GSM_NPARAM; q++) prm_v[q] = __ldg(b.params + t * GSM_NPARAM + q);
            const int32_t use_v = b.use_spike ? static_cast<int32_t>(__ldg(b.use_spike + t)) : 1;
            gsm_setup_scalars(K, prm_v, use_v, b.flags, b.stub_mode);
            W.root = -1;
            W.nroot = 0;
            W.xp.i0 = -1;
        }
        __syncwarp();
        // invalid states (lambda <= mu, shape <= 0, ...) go to the generic path (the kernel flavour follows the wiring)
        if (!(valid && gsm_lean_consts_ok(K) && (K.spike_known_n == 0) == MIXK)) {
            if (lane == 0) lean_redo(redo_count, redo_list, rows, t);
            continue;
        }
        const int L = (n + 1) >> 1;
        const int L4 = L & ~3;
        const double* hrow = b.height + row;
        const double* rrow = b.rate + row;
        const double* srow = b.spike + row;
        const int32_t* prow = b.parent + row;
        const int32_t* nrow = b.stub_mode != GSM_STUBS_NOSTUB ? b.nstubs + row : nullptr;
        GsmLeanC c;
        gsm_lean_consts(K, c, n);
        c.t6 = t6;
        c.mtab = MIXK ? W.mtab : nullptr;
        c.ntab = W.ntab;
        c.mixtab = reinterpret_cast<const GsmLogEntry*>(W.ntab);   // mixture mode has no stub counts: the two tables share the space
        LeanBCtx x;
        x.a_hs = smem_u32(sh_hs) - static_cast<uint32_t>(L4) * 8u;
        x.a_as = smem_u32(sh_as) - static_cast<uint32_t>(L4) * 8u;
        x.a_fk = smem_u32(sh_fk) - static_cast<uint32_t>(L4);
        x.a_hL = x.a_hs + static_cast<uint32_t>(L) * 8u;
        x.a_aL = x.a_as + static_cast<uint32_t>(L) * 8u;
        x.a_fL = x.a_fk + static_cast<uint32_t>(L);
        x.chk_p = 0u;
        x.nL = static_cast<uint32_t>(n - L);
        const bool detect = !(c.leaf_term == 0 && K.psi == 0.0);   // sampled ancestors are looked for where the tree prior generates them
        c.sa_unscanned = !detect;

        // The heights of the internal nodes (pass A) are requested before the Gamma table is built: their latency passes under it.
        constexpr int NPA_U = (GSM_WARP_MAX_NODES + 1) / 2 / 2 / 32 + 2;   // pair slots per lane (n - L4 <= (n + 1) / 2 + 3)
        double2 hq[NPA_U];
        const int npa = (n - L4 + 1) >> 1;
        {
            const double2* h2row = reinterpret_cast<const double2*>(hrow + L4);
#pragma unroll
            for (int u = 0; u < NPA_U; u++) {
                const int j = lane + 32 * u;
                hq[u] = j < npa ? __ldg(h2row + j) : make_double2(0.0, 0.0);
                if (j < npa && L4 + 2 * j + 1 >= n) hq[u].y = hq[u].x;   // padding slot of the last pair
            }
        }
        // ---- Gamma table, as far as the largest stub count of the tree needs it:
        //      ntab[k] = {log shape - logGamma(shape (k + 1)), -log k!, shape (k + 1) - 1, k}   (BSP:97-99, STP:567)
        if (MIXK) {
            // g[m] = Gamma(shape m) / Gamma(shape (m + 1)); [m - 1] = {g[m] / m, 1 / m}, [64 + m] = {g[m] / (m + 1), 1 / (m + 1)}
            const double shape = K.shape;
            GsmLogEntry* mt = reinterpret_cast<GsmLogEntry*>(W.ntab);
            for (int m = lane + 1; m <= GSM_LEAN_MTAB_N - 1; m += 32) {
                const double logc = gsm_log6_hi(t6, shape) - lean_log_gamma(t6, shape * m);
                if (m == 1) {
                    W.mtab[1].invc = shape * m - 1;
                    W.mtab[1].logc = logc;
                    W.mtab[0].invc = -1.0;
                    W.mtab[0].logc = 0.0;
                }
                const double lg_m = gsm_log6_hi(t6, shape) - logc;   // logGamma(shape m)
                const double g = gsm_exp6(t6, lg_m - lean_log_gamma(t6, shape * (m + 1)));
                GsmLogEntry r;
                r.logc = lean_rcp(static_cast<double>(m));
                r.invc = g * r.logc;
                mt[m - 1] = r;
                if (m < GSM_LEAN_MAX_NST) {
                    r.logc = lean_rcp(static_cast<double>(m + 1));
                    r.invc = g * r.logc;
                    mt[GSM_LEAN_MAX_NST + 1 + m] = r;
                }
                if (m == 1) {   // [64 + 0]: only its 1 / (k + 1) is used (the k = 0 term of a fake node's child is zero)
                    r.logc = 1.0;
                    r.invc = 0.0;
                    mt[GSM_LEAN_MAX_NST + 1] = r;
                }
            }
        } else {
            uint32_t nmax = static_cast<uint32_t>(GSM_LEAN_MAX_NST);
            if (nrow) {
                // (all loads of a lane are issued before the first is used)
                constexpr int NU = (GSM_WARP_MAX_NODES + 31) / 32;
                uint32_t v[NU];
#pragma unroll
                for (int u = 0; u < NU; u++) {
                    const int i = lane + 32 * u;
                    v[u] = i < n ? static_cast<uint32_t>(__ldg(nrow + i)) : 0u;
                }
                nmax = 0u;
#pragma unroll
                for (int u = 0; u < NU; u++) nmax = gsm_umax(nmax, v[u]);
                nmax = warp_max_u32(nmax);
                if (nmax > static_cast<uint32_t>(GSM_LEAN_MAX_NST)) nmax = static_cast<uint32_t>(GSM_LEAN_MAX_NST);   // (such a tree is handed back below)
            }
            const int mtop = static_cast<int>(nmax) + 2 < GSM_LEAN_MAX_NST + 1 ? static_cast<int>(nmax) + 2 : GSM_LEAN_MAX_NST + 1;
            const double shape = K.shape;
            for (int m = lane + 1; m <= mtop; m += 32) {
                GsmLeanN q;
                q.w = shape * m - 1;
                q.gcon = gsm_log6_hi(t6, shape) - lean_log_gamma(t6, shape * m);
                q.nlf = -g_logfact[m - 1];
                q.nd = static_cast<double>(m - 1);
                W.ntab[m - 1] = q;
            }
        }
        // ---- pass A: internal heights into shared memory, aux of the internal nodes; fake flags cleared ----
        GsmLeanAccA SA;
        gsm_lean_acc_a_zero(SA);
        {
            const uint32_t a_h0 = smem_u32(sh_hs), a_a0 = smem_u32(sh_as);
#pragma unroll
            for (int u = 0; u < NPA_U; u += 2) {
                const int j = lane + 32 * u, j2 = j + 32;
                if (u + 1 < NPA_U && j2 < npa) {   // two pairs = four nodes, staged (gsm_lean_aux4)
                    const double h4[4] = {hq[u].x, hq[u].y, hq[u + 1 < NPA_U ? u + 1 : u].x, hq[u + 1 < NPA_U ? u + 1 : u].y};
                    double a4[4];
                    gsm_lean_aux4(c, SA, h4, a4);
                    sts_v2f64(a_h0 + j * 16, h4[0], h4[1]);
                    sts_v2f64(a_h0 + j2 * 16, h4[2], h4[3]);
                    sts_v2f64(a_a0 + j * 16, a4[0], a4[1]);
                    sts_v2f64(a_a0 + j2 * 16, a4[2], a4[3]);
                } else if (j < npa) {
                    double e0, G0, e1, G1;
                    const double a0 = gsm_lean_aux(c, SA, hq[u].x, e0, G0);
                    const double a1 = gsm_lean_aux(c, SA, hq[u].y, e1, G1);
                    sts_v2f64(a_h0 + j * 16, hq[u].x, hq[u].y);
                    sts_v2f64(a_a0 + j * 16, a0, a1);
                }
            }
            for (int q = lane; q < (SI + 3) >> 2; q += 32) reinterpret_cast<uint32_t*>(sh_fk)[q] = 0u;
        }
        {
            GsmLeanAccA S0;
            gsm_lean_acc_a_zero(S0);
            x.auxz = gsm_lean_aux(c, S0, 0.0, x.ez, x.Gz);
        }
        __syncwarp();
        // ---- sampled-ancestor scan: a leaf at its parent's height -> the parent isFake (Node.isDirectAncestor) ----
        bool general = false;
        if (detect) {
            bool anyda = false;
            for (int i = lane; i < L; i += 32) {
                const int32_t p = __ldg(prow + i);
                const double h = __ldg(hrow + i);
                if (static_cast<uint32_t>(p - L) < x.nL && h == lds_f64(x.a_hs + p * 8)) {
                    sh_fk[p - L4] = 1;
                    anyda = true;
                }
            }
            general = __any_sync(0xffffffffu, anyda) != 0;
            __syncwarp();
        }

        // ---- pass B ----
        GsmLeanAccB SB;
        gsm_lean_acc_b_zero(SB);
        double2* orate2 = (WANT_RATES && out_rates) ? reinterpret_cast<double2*>(out_rates + row) : nullptr;
        double2* owsp2 = (WANT_RATES && out_wspikes) ? reinterpret_cast<double2*>(out_wspikes + row) : nullptr;
        const bool stdw = gsm_lean_std_wiring(c);
        int32_t* rn = &W.root;   // {root, nroot}
#define GSM_WPASS_B(G, W_, LT, M_) lean_warp_pass_b<G, WANT_RATES, W_, LT, XPD, M_>(c, SA, SB, x, a_w, n, L, L4, lane, rrow, srow, hrow, nrow, prow, rn, orate2, owsp2)
#define GSM_WPASS_B3(G, LT) { if constexpr (MIXK) GSM_WPASS_B(G, false, LT, true); else { if (stdw) GSM_WPASS_B(G, true, LT, false); else GSM_WPASS_B(G, false, LT, false); } }
        if (general) {
            if (c.leaf_term == 0) GSM_WPASS_B3(true, 0)
            else if (c.leaf_term == 1) GSM_WPASS_B3(true, 1)
            else GSM_WPASS_B3(true, 2)
        } else {
            if (c.leaf_term == 0) GSM_WPASS_B3(false, 0)
            else if (c.leaf_term == 1) GSM_WPASS_B3(false, 1)
            else GSM_WPASS_B3(false, 2)
        }
#undef GSM_WPASS_B3
#undef GSM_WPASS_B
        __syncwarp();
        // ---- the nodes the pair loops leave out, one lane each, with the single-node GENERAL function: node n-1 (lane 0; no stub
        // count in COUNTS mode, Stubs:354), the pair (L-1, L) that straddles the leaf / internal boundary when L is odd (lanes 1, 2),
        // the pair that holds the root when an ordinary tree's root is not node n-1 (lanes 3, 4) ----
        int tail_node = -1;
        if (lane == 0) tail_node = n - 1;
        else if (lane <= 2) { if (L & 1) tail_node = L - 2 + lane; }
        else if (lane <= 4) { if (W.xp.i0 >= 0) tail_node = W.xp.i0 + (lane - 3); }
        GsmLean6 s6;
        if (general) gsm_lean_combine<true>(K, c, SA, SB, s6);
        else gsm_lean_combine<false>(K, c, SA, SB, s6);
        int32_t t_leafterm = 0, t_plain = 0, t_fake = 0, t_extant = 0, t_extinct = 0, t_da = 0;
        uint32_t t_chk = 0u, t_chk_n = 0u, t_hmax = 0u;
        if (tail_node >= 0) {
            const int i = tail_node;
            GsmLeanAccB ST;
            gsm_lean_acc_b_zero(ST);
            GsmLeanAccA STA;
            gsm_lean_acc_a_zero(STA);
            const double h = __ldg(hrow + i);
            int32_t p = __ldg(prow + i);
            const bool root = p < 0;
            if (root) {
                atomicAdd(&W.nroot, 1);
                W.root = i;
                if (i < L) ST.chk = 0xffffffffu;
                p = i >= L ? i : L;
            }
            const uint32_t chk_p = static_cast<uint32_t>(p - L);
            if (!(static_cast<uint32_t>(p - L) < x.nL)) p = L;
            double e = 0.0, G = 0.0, aux;
            if (i < L) aux = gsm_lean_aux(c, STA, h, e, G);
            else aux = sh_as[i - L4];
            double hp = sh_hs[p - L4], auxp;
            if (root) { hp = h; auxp = aux; }
            else auxp = sh_as[p - L4];
            const bool pf = general && sh_fk[p - L4] != 0, fs = general && i >= L && sh_fk[i - L4] != 0;
            if (i < L) {
                const bool da = hp == h && !root;
                if (c.leaf_term == 1) gsm_lean_leaf_term<1>(c, STA, !da, h, aux, e, G);
                else if (c.leaf_term == 2) gsm_lean_leaf_term<2>(c, STA, !da, h, aux, e, G);
            }
            int32_t nst = -1;
            if (c.use_counts) nst = (i < n - 1 || b.stub_mode == GSM_STUBS_EXACT) ? __ldg(nrow + i) : 0;
            double er, ew;
            gsm_lean_b_node<true, WANT_RATES, 2, false, MIXK>(c, ST, i, root, pf, fs, h, hp, aux, auxp, __ldg(rrow + i), __ldg(srow + i), nst, &er, &ew);
            if (WANT_RATES) {
                if (out_rates) out_rates[row + i] = er;
                if (out_wspikes) out_wspikes[row + i] = ew;
            }
            GsmLean6 t6r;
            gsm_lean_combine<true>(K, c, STA, ST, t6r);
            s6.tree += t6r.tree; s6.stub += t6r.stub; s6.spike += t6r.spike; s6.rate += t6r.rate; s6.burst += t6r.burst; s6.dist += t6r.dist;
            t_leafterm = STA.n_leafterm; t_plain = STA.n_plain; t_fake = ST.n_fake; t_extant = ST.n_extant; t_extinct = ST.n_extinct; t_da = ST.n_da;
            t_chk = gsm_umax(STA.chk, ST.chk); t_chk_n = ST.chk_n; t_hmax = STA.hmax;
            x.chk_p = gsm_umax(x.chk_p, chk_p);
        }
        __syncwarp();
        // ---- reductions over the warp (fixed order) and the scratch row ----
        const double v0 = warp_sum(s6.tree), v1 = warp_sum(s6.stub), v2 = warp_sum(s6.spike), v3 = warp_sum(s6.rate),
                     v4 = warp_sum(s6.burst), v5 = warp_sum(s6.dist);
        const int32_t c1 = __reduce_add_sync(0xffffffffu, SA.n_leafterm + t_leafterm),
                      c2 = __reduce_add_sync(0xffffffffu, SA.n_plain + t_plain), c3 = __reduce_add_sync(0xffffffffu, SB.n_fake + t_fake),
                      c4 = __reduce_add_sync(0xffffffffu, SB.n_extant + t_extant),
                      c5 = __reduce_add_sync(0xffffffffu, SB.n_extinct + t_extinct), c6 = __reduce_add_sync(0xffffffffu, SB.n_da + t_da);
        const uint32_t chk = warp_max_u32(gsm_umax(gsm_umax(SA.chk, SB.chk), t_chk)), chk_n = warp_max_u32(gsm_umax(SB.chk_n, t_chk_n)),
                       chk_len = warp_max_u32(SB.chk_len), hmax = warp_max_u32(gsm_umax(SA.hmax, t_hmax)), chk_p = warp_max_u32(x.chk_p);
        if (lane == 0) {
            const int root = W.root;
            bool bad = chk >= GSM_CHK_LIMIT || (c.use_counts && chk_n > static_cast<uint32_t>(GSM_LEAN_MAX_NST)) ||
                       (!general && chk_len >= GSM_CHK_LIMIT) || hmax >= 0x7ff00000u || chk_p >= x.nL || W.nroot != 1 || root < L;
            if (!bad) bad = !(K.c1 * __hiloint2double(static_cast<int>(hmax) + 1, 0) < 700.0);   // exp(-c1 h) stays in range
            if (bad) {
                lean_redo(redo_count, redo_list, rows, t);
            } else {
                GsmLeanRow* r = rows + t;
                r->tree = v0; r->stub = v1; r->spike = v2; r->rate = v3; r->burst = v4; r->dist = v5;
                r->T = sh_hs[root - L4];
                r->n_nst = 0; r->n_leafterm = c1; r->n_plain = c2; r->n_fake = c3;
                r->n_extant = c4; r->n_extinct = c5; r->n_da = c6;
                r->root_fake = (general && sh_fk[root - L4] != 0) ? 1 : 0;
                r->status = 1;
            }
        }
    }
}
