// gsm_emu.cpp — TEST INFRASTRUCTURE.  Compiles the product's per-node device math
// (gammaspikemodel_b200/csrc/gsm_node_math.cuh and gsm_lean.cuh, GSM_HD = plain inline under g++) and walks one tree at a
// time through the same phases as gsm_eval_kernel, sequentially.  It lets the CPU-only test tier check the
// algebra and the flag/epilogue logic of the kernel against the oracle without a GPU.  It is never loaded
// by the product package and is not a fallback.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../gammaspikemodel_b200/csrc/gsm_node_math.cuh"
#include "../../gammaspikemodel_b200/csrc/gsm_lean.cuh"

static double g_logfact_host[GSM_LOGFACT_N];
static bool g_init = false;

static void init_tables() {
    if (g_init) return;
    g_logfact_host[0] = g_logfact_host[1] = 0.0;
    double s = 0.0;
    for (int i = 2; i < GSM_LOGFACT_N; i++) {
        s += std::log((double)i);
        g_logfact_host[i] = s;
    }
    g_init = true;
}

template <bool WANT>
static void eval_tree(uint32_t flags, int32_t stub_mode, int32_t n, const double* h, const int32_t* par_in,
                      const double* rate, const double* spike, const int32_t* nstubs_in, const double* params,
                      int32_t use_spike, double* out, double* out_rates, double* out_wsp, int64_t n_stub = 0,
                      const int32_t* stub_branch = nullptr, const double* stub_time = nullptr) {
    std::vector<int32_t> counts;
    const int32_t* nstubs = nstubs_in;
    if (stub_mode == GSM_STUBS_EXACT && stub_branch) {   // gsm_stub_counts_kernel
        counts.assign(n > 0 ? n : 1, 0);
        for (int64_t j = 0; j < n_stub; j++)
            if ((uint32_t)stub_branch[j] < (uint32_t)n) counts[stub_branch[j]]++;
        nstubs = counts.data();
    }
    if (n <= 0) {
        for (int k = 0; k < GSM_NOUT; k++) out[k] = NAN;
        return;
    }
    GsmTreeConsts K;
    gsm_setup_consts(K, params, use_spike, flags, stub_mode);
    for (int m = 0; m <= GSM_LUT_M; m++) gsm_lut_entry(K, params[GSM_P_SPIKE_SHAPE], m);
    std::vector<int32_t> par(par_in, par_in + n);
    std::vector<uint8_t> nonleaf(n, 0), fake(n, 0);
    std::vector<double> aux(n, 0.0);
    int32_t root = -1;
    for (int i = 0; i < n; i++) {  // phase 1
        int32_t p = par[i];
        if ((uint32_t)p < (uint32_t)n) nonleaf[p] = 1;
        else if (p < 0) root = i;
        else par[i] = -1;
    }
    for (int i = 0; i < n; i++) {  // phase 1b
        int32_t p = par[i];
        if (!nonleaf[i] && p >= 0 && h[p] == h[i]) fake[p] = 1;
    }
    GsmAcc A;
    gsm_acc_zero(A);
    for (int i = 0; i < n; i++) {  // phase 2: pass A
        int32_t p = par[i];
        uint32_t nf = 0;
        double hp = h[i];
        if (nonleaf[i]) nf |= GSM_NF_NONLEAF;
        if (fake[i]) nf |= GSM_NF_FAKE;
        if (p < 0) nf |= GSM_NF_ROOT;
        else {
            hp = h[p];
            if (!(nf & GSM_NF_NONLEAF) && hp == h[i]) nf |= GSM_NF_DA;
        }
        aux[i] = gsm_node_pass_a(K, A, nf, h[i], hp);
    }
    for (int i = 0; i < n; i++) {  // phase 3: pass B
        int32_t p = par[i];
        uint32_t nf = 0;
        double hp = h[i], auxp = aux[i];
        if (nonleaf[i]) nf |= GSM_NF_NONLEAF;
        if (fake[i]) nf |= GSM_NF_FAKE;
        if (p < 0) {
            nf |= GSM_NF_ROOT;
        } else {
            hp = h[p];
            auxp = aux[p];
            if (fake[p]) nf |= GSM_NF_PARENT_FAKE;
            if (!(nf & GSM_NF_NONLEAF) && hp == h[i]) nf |= GSM_NF_DA;
        }
        gsm_node_pass_b<WANT>(K, g_logfact_host, A, i, n, nf, h[i], hp, aux[i], auxp, rate ? rate[i] : 1.0, spike[i],
                              nstubs ? nstubs[i] : 0, out_rates ? out_rates + i : nullptr, out_wsp ? out_wsp + i : nullptr);
    }
    if (stub_branch && K.stub_term && K.stub_mode == GSM_STUBS_EXACT) {   // phase 3b
        for (int64_t j = 0; j < n_stub; j++) {
            const int32_t br = stub_branch[j];
            if ((uint32_t)br >= (uint32_t)n) continue;
            const int32_t p = par[br];
            const bool rt = p < 0;
            const double hp = rt ? h[br] : h[p];
            const bool da = !rt && !nonleaf[br] && hp == h[br];
            gsm_stub_term(K, A, rt, da, h[br], hp, stub_time[j]);
        }
    }
    if (root < 0) {
        for (int k = 0; k < GSM_NOUT; k++) out[k] = NAN;
        return;
    }
    gsm_finish_consts(K, h[root]);
    gsm_finalize(K, A, n, fake[root] != 0, out);
}

// EXACT mode with stub lists (CSR by tree): the generic walk, as gsm_eval_kernel + gsm_stub_counts_kernel do it
extern "C" int gsm_emu_eval_batch_stubs(uint32_t flags, int32_t stub_mode, int64_t n_trees, int32_t n_nodes, int64_t stride,
                                        const double* height, const int32_t* parent, const double* rate, const double* spike,
                                        const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                                        const int32_t* node_count, const int64_t* stub_ptr, const int32_t* stub_branch,
                                        const double* stub_time, double* out_tree, double* out_rates, double* out_wsp) {
    init_tables();
    for (int64_t t = 0; t < n_trees; t++) {
        int32_t n = node_count ? node_count[t] : n_nodes;
        const int64_t r = t * stride;
        const int64_t s0 = stub_ptr ? stub_ptr[t] : 0, s1 = stub_ptr ? stub_ptr[t + 1] : 0;
        eval_tree<true>(flags, stub_mode, n, height + r, parent + r, rate ? rate + r : nullptr, spike + r,
                        nstubs ? nstubs + r : nullptr, params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1,
                        out_tree + t * GSM_NOUT, out_rates ? out_rates + r : nullptr, out_wsp ? out_wsp + r : nullptr, s1 - s0,
                        stub_ptr ? stub_branch + s0 : nullptr, stub_ptr ? stub_time + s0 : nullptr);
    }
    return 0;
}

extern "C" int gsm_emu_eval_batch(uint32_t flags, int32_t stub_mode, int64_t n_trees, int32_t n_nodes, int64_t stride,
                                  const double* height, const int32_t* parent, const double* rate, const double* spike,
                                  const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                                  const int32_t* node_count, double* out_tree, double* out_rates, double* out_wsp) {
    init_tables();
    for (int64_t t = 0; t < n_trees; t++) {
        int32_t n = node_count ? node_count[t] : n_nodes;
        const int64_t r = t * stride;
        if (out_rates || out_wsp)
            eval_tree<true>(flags, stub_mode, n, height + r, parent + r, rate ? rate + r : nullptr, spike + r,
                            nstubs ? nstubs + r : nullptr, params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1,
                            out_tree + t * GSM_NOUT, out_rates ? out_rates + r : nullptr, out_wsp ? out_wsp + r : nullptr);
        else
            eval_tree<false>(flags, stub_mode, n, height + r, parent + r, rate ? rate + r : nullptr, spike + r,
                             nstubs ? nstubs + r : nullptr, params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1,
                             out_tree + t * GSM_NOUT, nullptr, nullptr);
    }
    return 0;
}

// ---- the lean path (gsm_lean_kernel + gsm_post_kernel), walked sequentially by one "thread".  A tree the lean
// path hands back (redo list on the device) is evaluated by the generic walk above, exactly as the post kernel does. ----
static GsmLeanBlob g_blob;
static bool g_blob_init = false;
static void init_blob() {
    if (g_blob_init) return;
    gsm_lean_blob_fill(g_blob);
    init_tables();
    g_blob_init = true;
}

// pass B as the kernel splits it: leaf pairs (aux of a leaf computed on the spot), internal pairs from L4 = L & ~3 on, the
// tail (node n-1 and, when L is odd, the pair that straddles L)
template <bool GENERAL, bool WANT, bool STDW, bool MIX>
static void lean_b_all(const GsmLeanC& c, GsmLeanAccA& SA, GsmLeanAccB& SB, GsmLeanAccA& STA, GsmLeanAccB& ST, int n, int32_t stub_mode,
                       const double* h, const double* aux_int /* [i - L4] */, const int32_t* par, const uint8_t* fk /* [i - L4] */,
                       const double* rate, const double* spike, const int32_t* nstubs, int* root_out, int* nroot, bool* bad,
                       double* out_rates, double* out_wsp) {
    double dr, dw;
    const int L = c.L, L4 = L & ~3;
    const uint32_t nL = (uint32_t)(n - L);
    auto node = [&](int i, int leafclass, GsmLeanAccA& A, GsmLeanAccB& B, bool tail) {
        int32_t p = par[i];
        const bool root = p < 0;
        if (root && (GENERAL || tail)) {
            ++*nroot;
            *root_out = i;
            if (i < L) *bad = true;
            p = i >= L ? i : L;
        }
        if (!((uint32_t)(p - L) < nL)) { *bad = true; p = L; }   // chk_p
        double e = 0, G = 0, aux;
        if (i < L) aux = gsm_lean_aux(c, A, h[i], e, G);
        else aux = aux_int[i - L4];
        double hp = h[p], auxp = aux_int[p - L4];
        if (root && (GENERAL || tail)) { hp = h[i]; auxp = aux; }
        const bool gen = GENERAL || tail;
        const bool pf = gen && fk[p - L4] != 0, fs = gen && i >= L && fk[i - L4] != 0;
        if (i < L) {
            const bool da = gen && hp == h[i] && !root;
            if (c.leaf_term == 1) gsm_lean_leaf_term<1>(c, A, !da, h[i], aux, e, G);
            else if (c.leaf_term == 2) gsm_lean_leaf_term<2>(c, A, !da, h[i], aux, e, G);
        }
        int32_t k = nstubs ? nstubs[i] : -1;
        if (tail) {
            k = -1;
            if (c.use_counts) k = (i < n - 1 || stub_mode == GSM_STUBS_EXACT) ? (nstubs ? nstubs[i] : 0) : 0;
        }
        double* o_r = out_rates ? out_rates + i : &dr;
        double* o_w = out_wsp ? out_wsp + i : &dw;
        const bool rt = root && gen;
        if (tail) gsm_lean_b_node<true, WANT, 2, false, MIX>(c, B, i, rt, pf, fs, h[i], hp, aux, auxp, rate[i], spike[i], k, o_r, o_w);
        else if (leafclass == 1) gsm_lean_b_node<GENERAL, WANT, 1, STDW, MIX>(c, B, i, rt, pf, fs, h[i], hp, aux, auxp, rate[i], spike[i], k, o_r, o_w);
        else gsm_lean_b_node<GENERAL, WANT, 0, STDW, MIX>(c, B, i, rt, pf, fs, h[i], hp, aux, auxp, rate[i], spike[i], k, o_r, o_w);
    };
    for (int i0 = 0; i0 + 1 < L; i0 += 2) { node(i0, 1, SA, SB, false); node(i0 + 1, 1, SA, SB, false); }
    int xpair = -1;   // the pair that holds the root of an ordinary tree whose root is not node n-1: handed to the tail lanes
    for (int i0 = L4; i0 + 1 < n - 1; i0 += 2) {
        if (i0 < L) continue;
        if (!GENERAL && (par[i0] | par[i0 + 1]) < 0) {
            if (xpair >= 0) *bad = true;   // a second pair with a root
            else xpair = i0;
            continue;
        }
        node(i0, 0, SA, SB, false);
        node(i0 + 1, 0, SA, SB, false);
    }
    if (xpair >= 0) { node(xpair, 2, STA, ST, true); node(xpair + 1, 2, STA, ST, true); }
    if (L & 1) { node(L - 1, 2, STA, ST, true); node(L, 2, STA, ST, true); }
    node(n - 1, 2, STA, ST, true);
}

// returns false when the tree is handed back to the generic path
template <bool WANT>
static bool eval_tree_lean(uint32_t flags, int32_t stub_mode, int32_t n, const double* h, const int32_t* par_in,
                           const double* rate, const double* spike, const int32_t* nstubs, const double* params,
                           int32_t use_spike, double* out, double* out_rates, double* out_wsp) {
    if (!((n & 1) && n >= 3 && n <= GSM_LEAN_MAX_NODES)) return false;
    const int L = (n + 1) >> 1, L4 = L & ~3;
    GsmTreeConsts K;
    memset(&K, 0, sizeof K);
    gsm_setup_scalars(K, params, use_spike, flags, stub_mode);
    GsmLogEntry mtab[GSM_LEAN_MTAB_N];
    GsmLeanN ntab[GSM_LEAN_MAX_NST + 1];
    mtab[0].invc = -1.0;
    mtab[0].logc = 0.0;
    GsmLogEntry* mixtab = reinterpret_cast<GsmLogEntry*>(ntab);   // mixture mode: overlays ntab, as in the kernel
    memset(ntab, 0, sizeof ntab);
    for (int m = 1; m < GSM_LEAN_MTAB_N; m++) {
        mtab[m].invc = K.shape * m - 1;
        mtab[m].logc = gsm_log(K.shape) - gsm_log_gamma(K.shape * m);
        if (K.spike_known_n) {
            ntab[m - 1].gcon = mtab[m].logc;
            ntab[m - 1].nlf = -g_logfact_host[m - 1];
            ntab[m - 1].w = mtab[m].invc;
            ntab[m - 1].nd = (double)(m - 1);
        } else {
            const double g = gsm_exp(gsm_log_gamma(K.shape * m) - gsm_log_gamma(K.shape * (m + 1)));
            mixtab[m - 1].logc = 1.0 / m;
            mixtab[m - 1].invc = g / m;
            if (m < GSM_LEAN_MAX_NST) {
                mixtab[GSM_LEAN_MAX_NST + 1 + m].logc = 1.0 / (m + 1);
                mixtab[GSM_LEAN_MAX_NST + 1 + m].invc = g / (m + 1);
            }
        }
    }
    if (!K.spike_known_n) {
        mixtab[GSM_LEAN_MAX_NST + 1].logc = 1.0;
        mixtab[GSM_LEAN_MAX_NST + 1].invc = 0.0;
    }
    if (!gsm_lean_consts_ok(K)) return false;
    GsmLeanC c;
    gsm_lean_consts(K, c, n);
    c.t6.log_t = g_blob.logtab; c.t6.exp_t = g_blob.exptab; c.mtab = mtab; c.ntab = ntab; c.mixtab = mixtab;
    const uint32_t nL = (uint32_t)(n - L);
    // phase 1: sampled-ancestor scan (only where they can occur)
    const bool detect = !(c.leaf_term == 0 && K.psi == 0.0);
    c.sa_unscanned = !detect;
    std::vector<uint8_t> fk(n - L4 + 8, 0);
    bool anyda = false;
    if (detect)
        for (int i = 0; i < L; i++) {
            const int32_t p = par_in[i];
            if ((uint32_t)(p - L) < nL && h[i] == h[p]) { fk[p - L4] = 1; anyda = true; }
        }
    // phase 2: aux of the internal nodes (and of the up to three leaves from L4 on)
    GsmLeanAccA SA, STA;
    gsm_lean_acc_a_zero(SA);
    gsm_lean_acc_a_zero(STA);
    std::vector<double> aux_int(n - L4 + 8, 0.0);
    for (int i = L4; i < n; i++) {
        double e, G;
        aux_int[i - L4] = gsm_lean_aux(c, SA, h[i], e, G);
    }
    const bool general = anyda;   // (a root that is not node n-1 alone keeps the CLEAN bodies)
    GsmLeanAccB SB, ST;
    gsm_lean_acc_b_zero(SB);
    gsm_lean_acc_b_zero(ST);
    const int32_t* ns = c.use_counts ? nstubs : nullptr;
    const bool stdw = gsm_lean_std_wiring(c);
    int root = -1, nroot = 0;
    bool bad = false;
#define EMU_B(G, W, M) lean_b_all<G, WANT, W, M>(c, SA, SB, STA, ST, n, stub_mode, h, aux_int.data(), par_in, fk.data(), rate, spike, ns, &root, &nroot, &bad, out_rates, out_wsp)
    if (general) { if (stdw) EMU_B(true, true, false); else if (c.mix) EMU_B(true, false, true); else EMU_B(true, false, false); }
    else { if (stdw) EMU_B(false, true, false); else if (c.mix) EMU_B(false, false, true); else EMU_B(false, false, false); }
#undef EMU_B
    GsmLean6 s6, t6;
    if (general) gsm_lean_combine<true>(K, c, SA, SB, s6);
    else gsm_lean_combine<false>(K, c, SA, SB, s6);
    gsm_lean_combine<true>(K, c, STA, ST, t6);
    s6.tree += t6.tree; s6.stub += t6.stub; s6.spike += t6.spike; s6.rate += t6.rate; s6.burst += t6.burst; s6.dist += t6.dist;
    const uint32_t chk = gsm_umax(gsm_umax(gsm_umax(SA.chk, STA.chk), SB.chk), ST.chk), chk_n = gsm_umax(SB.chk_n, ST.chk_n);
    const uint32_t hmax = gsm_umax(SA.hmax, STA.hmax);
    bool redo = chk >= GSM_CHK_LIMIT || (c.use_counts && chk_n > (uint32_t)GSM_LEAN_MAX_NST) ||
                (!general && SB.chk_len >= GSM_CHK_LIMIT) || hmax >= 0x7ff00000u || bad || nroot != 1 || root < L;
    if (!redo) redo = !(K.c1 * gsm_hilo((int32_t)hmax + 1, 0) < 700.0);
    if (redo) return false;
    // ---- gsm_post_kernel part 1 ----
    GsmTreeConsts K2;
    memset(&K2, 0, sizeof K2);
    gsm_setup_consts(K2, params, use_spike, flags, stub_mode);
    GsmLeanC c2;
    gsm_lean_consts(K2, c2, n);
    GsmAcc A;
    gsm_acc_zero(A);
    A.n_leaf = c2.L;
    A.n_extant = SB.n_extant + ST.n_extant;
    A.n_extinct = SB.n_extinct + ST.n_extinct;
    A.n_da = SB.n_da + ST.n_da;
    gsm_lean_finish(K2, c2, s6, 0, SA.n_leafterm + STA.n_leafterm, SA.n_plain + STA.n_plain, SB.n_fake + ST.n_fake, n, A);
    gsm_finish_consts(K2, h[root]);
    gsm_finalize(K2, A, n, general && fk[root - L4] != 0, out);
    return true;
}

// returns 1 when the wiring is outside the lean family (caller uses gsm_emu_eval_batch); *n_lean = trees finished
// by the lean path (the rest went through the generic walk, as on the device)
extern "C" int gsm_emu_eval_batch_lean(uint32_t flags, int32_t stub_mode, int64_t n_trees, int32_t n_nodes, int64_t stride,
                                       const double* height, const int32_t* parent, const double* rate, const double* spike,
                                       const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                                       const int32_t* node_count, double* out_tree, double* out_rates, double* out_wsp,
                                       int64_t* n_lean) {
    init_blob();
    if (n_lean) *n_lean = 0;
    if (!gsm_lean_wiring_ok(flags, stub_mode) || !rate) return 1;
    for (int64_t t = 0; t < n_trees; t++) {
        int32_t n = node_count ? node_count[t] : n_nodes;
        const int64_t r = t * stride;
        const bool want = out_rates || out_wsp;
        bool ok;
        if (want)
            ok = eval_tree_lean<true>(flags, stub_mode, n, height + r, parent + r, rate + r, spike + r, nstubs ? nstubs + r : nullptr,
                                      params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1, out_tree + t * GSM_NOUT,
                                      out_rates ? out_rates + r : nullptr, out_wsp ? out_wsp + r : nullptr);
        else
            ok = eval_tree_lean<false>(flags, stub_mode, n, height + r, parent + r, rate + r, spike + r, nstubs ? nstubs + r : nullptr,
                                       params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1, out_tree + t * GSM_NOUT, nullptr, nullptr);
        if (ok) {
            if (n_lean) ++*n_lean;
            continue;
        }
        if (want)
            eval_tree<true>(flags, stub_mode, n, height + r, parent + r, rate + r, spike + r, nstubs ? nstubs + r : nullptr,
                            params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1, out_tree + t * GSM_NOUT,
                            out_rates ? out_rates + r : nullptr, out_wsp ? out_wsp + r : nullptr);
        else
            eval_tree<false>(flags, stub_mode, n, height + r, parent + r, rate + r, spike + r, nstubs ? nstubs + r : nullptr,
                             params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1, out_tree + t * GSM_NOUT, nullptr, nullptr);
    }
    return 0;
}

// which: 0 degree-5 log (per-branch sums), 1 degree-6 log (aux)
extern "C" void gsm_emu_log6_vec(const double* x, double* y, int64_t n, int32_t which) {
    init_blob();
    GsmT6 t;
    t.log_t = g_blob.logtab; t.exp_t = g_blob.exptab;
    for (int64_t i = 0; i < n; i++) y[i] = which ? gsm_log6_hi(t, x[i]) : gsm_log6(t, x[i]);
}
extern "C" void gsm_emu_exp6_vec(const double* x, double* y, int64_t n) {
    init_blob();
    GsmT6 t;
    t.log_t = g_blob.logtab; t.exp_t = g_blob.exptab;
    for (int64_t i = 0; i < n; i++) y[i] = gsm_exp6(t, x[i]);
}

// ---- table-driven log / exp of the kernels, exposed for accuracy tests ----
extern "C" void gsm_emu_log_vec(const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = gsm_log(x[i]);
}
extern "C" void gsm_emu_exp_vec(const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = gsm_exp(x[i]);
}
