// gsm_emu.cpp — TEST INFRASTRUCTURE.  Compiles the product's per-node device math
// (gammaspikemodel_b200/csrc/gsm_node_math.cuh, GSM_HD = plain inline under g++) and walks one tree at a
// time through the same phases as gsm_eval_kernel, sequentially.  It lets the CPU-only test tier check the
// algebra and the flag/epilogue logic of the kernel against the oracle without a GPU.  It is never loaded
// by the product package and is not a fallback.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../gammaspikemodel_b200/csrc/gsm_node_math.cuh"

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
                      const double* rate, const double* spike, const int32_t* nstubs, const double* params,
                      int32_t use_spike, double* out, double* out_rates, double* out_wsp) {
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
    if (root < 0) {
        for (int k = 0; k < GSM_NOUT; k++) out[k] = NAN;
        return;
    }
    gsm_finish_consts(K, h[root]);
    gsm_finalize(K, A, n, fake[root] != 0, out);
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

// ---- the fast path of gsm_eval_fast_kernel, walked sequentially in node pairs (32-bit packed words).  A "warp" is
// one pair here: the pair takes the checked path when either node is special, exactly as a voting warp would. ----
template <int VARIANT>
static void fast_a(const GsmFastA& fa, const GsmTabs& T, GsmAcc& A, int n, const double* h, double* aux, uint32_t* w) {
    using W = GsmWord<uint32_t>;
    std::vector<uint32_t> wnew(w, w + n);
    auto flags = [&](int i, bool& leaf, bool& fake_self, bool& da, bool& root, bool& pf, double& hp) {
        const int p = (int)(w[i] & W::IDX);
        hp = h[p];
        root = p == i;
        leaf = !(w[i] & W::NL);
        fake_self = (w[i] & W::FAKE) != 0;
        da = leaf && !root && hp == h[i];
        pf = !root && (w[p] & W::FAKE);
    };
    const int nfull = n >> 1;
    for (int j = 0; j < nfull; j++) {
        const int i0 = 2 * j;
        bool leaf[2], fake_self[2], da[2], root[2], pf[2];
        double hh[2] = {h[i0], h[i0 + 1]}, hp[2], a[2];
        for (int k = 0; k < 2; k++) flags(i0 + k, leaf[k], fake_self[k], da[k], root[k], pf[k], hp[k]);
        const bool sp = !fa.safe || gsm_fast_a_special(fa, hh[0]) || gsm_fast_a_special(fa, hh[1]);
        if (sp) for (int k = 0; k < 2; k++) a[k] = gsm_fast_pass_a_checked<VARIANT>(fa, T, A, leaf[k], fake_self[k], da[k], root[k], hh[k], hp[k]);
        else gsm_fast_pass_a_pair<VARIANT>(fa, T, A, leaf, fake_self, da, root, hh, hp, a);
        for (int k = 0; k < 2; k++) {
            aux[i0 + k] = a[k];
            wnew[i0 + k] = w[i0 + k] | (da[k] ? W::DA : 0u) | (pf[k] ? W::PF : 0u);
        }
    }
    if (n & 1) {
        const int i = n - 1;
        bool leaf, fake_self, da, root, pf;
        double hp;
        flags(i, leaf, fake_self, da, root, pf, hp);
        aux[i] = gsm_fast_pass_a_checked<VARIANT>(fa, T, A, leaf, fake_self, da, root, h[i], hp);
        wnew[i] = w[i] | (da ? W::DA : 0u) | (pf ? W::PF : 0u);
    }
    for (int i = 0; i < n; i++) w[i] = wnew[i];
}

template <bool WANT>
static void eval_tree_fast(uint32_t flags, int32_t stub_mode, int32_t n, const double* h, const int32_t* par_in,
                           const double* rate, const double* spike, const int32_t* nstubs, const double* params,
                           int32_t use_spike, double* out, double* out_rates, double* out_wsp) {
    using W = GsmWord<uint32_t>;
    if (n <= 0) {
        for (int k = 0; k < GSM_NOUT; k++) out[k] = NAN;
        return;
    }
    GsmTreeConsts K;
    gsm_setup_consts(K, params, use_spike, flags, stub_mode);
    for (int m = 0; m <= GSM_LUT_M; m++) gsm_lut_entry(K, params[GSM_P_SPIKE_SHAPE], m);
    GsmTabs T;
    T.logtab = gsm_log_tab_h;
    T.exptab = gsm_exp_tab_h;
    std::vector<uint32_t> w(n);
    std::vector<double> aux(n, 0.0);
    int32_t root = -1;
    bool anyzero = false;
    for (int i = 0; i < n; i++) {
        int32_t p = par_in[i];
        if ((uint32_t)p >= (uint32_t)n) { if (p < 0) root = i; p = i; }
        w[i] = (uint32_t)p;
    }
    for (int i = 0; i < n; i++) {
        int p = (int)(w[i] & W::IDX);
        if (p != i) { w[p] |= W::NL; if (h[p] == h[i]) anyzero = true; }
    }
    if (anyzero)
        for (int i = 0; i < n; i++) {
            int p = (int)(w[i] & W::IDX);
            if (!(w[i] & W::NL) && p != i && h[p] == h[i]) w[p] |= W::FAKE;
        }
    GsmAcc A;
    gsm_acc_zero(A);
    GsmFastA fa;
    GsmFastB fb;
    gsm_fast_consts(K, fa, fb, n);
    switch (fa.variant) {
        case GSM_TV_COMPLETE_RHO: fast_a<GSM_TV_COMPLETE_RHO>(fa, T, A, n, h, aux.data(), w.data()); break;
        case GSM_TV_NO_PSI: fast_a<GSM_TV_NO_PSI>(fa, T, A, n, h, aux.data(), w.data()); break;
        case GSM_TV_PSI: fast_a<GSM_TV_PSI>(fa, T, A, n, h, aux.data(), w.data()); break;
        case GSM_TV_CONDITIONED: fast_a<GSM_TV_CONDITIONED>(fa, T, A, n, h, aux.data(), w.data()); break;
        default: fast_a<GSM_TV_NONE>(fa, T, A, n, h, aux.data(), w.data()); break;
    }
    double dummy[2];
    const int nfull = n >> 1;
    for (int j = 0; j <= nfull; j++) {
        const int i0 = 2 * j;
        const int cnt = (j < nfull) ? 2 : (n & 1);
        if (cnt == 0) break;
        bool da[2], pf[2], rt[2], sp = false;
        double hp[2], auxp[2], len[2], dA[2], E[2], rr[2], ss[2];
        int32_t nr[2];
        for (int k = 0; k < cnt; k++) {
            const int i = i0 + k;
            const int p = (int)(w[i] & W::IDX);
            da[k] = (w[i] & W::DA) != 0; pf[k] = (w[i] & W::PF) != 0; rt[k] = p == i;
            hp[k] = h[p]; auxp[k] = aux[p]; len[k] = hp[k] - h[i]; dA[k] = aux[i] - auxp[k];
            E[k] = fma(len[k], fb.kE, 2 * dA[k]);
            rr[k] = rate[i]; ss[k] = spike[i]; nr[k] = nstubs ? nstubs[i] : 0;
            sp = sp || gsm_fast_b_special(fb, i, da[k], pf[k], rt[k], len[k], E[k], rr[k], ss[k], nr[k]);
        }
        if (sp || cnt == 1) {
            for (int k = 0; k < cnt; k++)
                gsm_fast_pass_b_checked<WANT>(fb, K, T, g_logfact_host, A, i0 + k, da[k], pf[k], rt[k], h[i0 + k], hp[k], aux[i0 + k],
                                              auxp[k], rr[k], ss[k], nr[k], out_rates ? out_rates + i0 + k : dummy,
                                              out_wsp ? out_wsp + i0 + k : dummy);
        } else {
            gsm_fast_pass_b_pair<WANT>(fb, K, T, g_logfact_host, A, i0, da, pf, rt, len, E, dA, rr, ss, nr,
                                       out_rates ? out_rates + i0 : dummy, out_wsp ? out_wsp + i0 : dummy);
        }
    }
    if (root < 0) {
        for (int k = 0; k < GSM_NOUT; k++) out[k] = NAN;
        return;
    }
    gsm_fast_fix_acc(K, A, n);
    gsm_finish_consts(K, h[root]);
    gsm_finalize(K, A, n, (w[root] & W::FAKE) != 0, out);
}

// returns 1 when the wiring is outside the fast-path family (caller falls back to gsm_emu_eval_batch)
extern "C" int gsm_emu_eval_batch_fast(uint32_t flags, int32_t stub_mode, int64_t n_trees, int32_t n_nodes, int64_t stride,
                                       const double* height, const int32_t* parent, const double* rate, const double* spike,
                                       const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                                       const int32_t* node_count, double* out_tree, double* out_rates, double* out_wsp) {
    init_tables();
    if (!gsm_fast_path_ok(flags, stub_mode) || !rate) return 1;
    for (int64_t t = 0; t < n_trees; t++) {
        int32_t n = node_count ? node_count[t] : n_nodes;
        const int64_t r = t * stride;
        if (out_rates || out_wsp)
            eval_tree_fast<true>(flags, stub_mode, n, height + r, parent + r, rate + r, spike + r, nstubs ? nstubs + r : nullptr,
                                 params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1, out_tree + t * GSM_NOUT,
                                 out_rates ? out_rates + r : nullptr, out_wsp ? out_wsp + r : nullptr);
        else
            eval_tree_fast<false>(flags, stub_mode, n, height + r, parent + r, rate + r, spike + r, nstubs ? nstubs + r : nullptr,
                                  params + t * GSM_NPARAM, use_spike ? use_spike[t] : 1, out_tree + t * GSM_NOUT, nullptr, nullptr);
    }
    return 0;
}

// ---- table-driven log / exp of the kernels, exposed for accuracy tests ----
extern "C" void gsm_emu_log_vec(const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = gsm_log(x[i]);
}
extern "C" void gsm_emu_exp_vec(const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = gsm_exp(x[i]);
}
