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

// ---- table-driven log / exp of the kernels, exposed for accuracy tests ----
extern "C" void gsm_emu_log_vec(const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = gsm_log(x[i]);
}
extern "C" void gsm_emu_exp_vec(const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = gsm_exp(x[i]);
}
