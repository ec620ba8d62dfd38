// gsm_node_math.cuh — per-tree constants and per-node (per-branch) terms of the GammaSpikeModel hot path.
//
// Everything here is GSM_HD (__host__ __device__) so that the exact code the sm_100a kernels run can
// also be compiled by g++ for the CPU-side logic tests (tests/emu_*.cpp).  The product only ever runs
// it inside the kernels of gsm_kernels.cu.
//
// Citations (relative to the reference tree):
//   PRCM  src/gammaspike/clockmodel/PunctuatedRelaxedClockModel.java
//   BSP   src/gammaspike/distribution/BranchSpikePrior.java
//   BRP   src/gammaspike/distribution/BranchRatePrior.java
//   STP   src/gammaspike/distribution/StumpedTreePrior.java
//   Stubs src/gammaspike/tree/Stubs.java
//   SPL   src/gammaspike/logger/SaltativeProportionLogger.java
//
// Design notes (DESIGN.md §4 has the derivations):
//   * per-tree constants (c1, c2, logs of rates, lgamma(shape*m) table) are computed once per tree
//     and the per-node code only pays for the transcendentals that depend on the node's own inputs;
//   * one "aux" value per node (exp(-c1*h) for the sampling variants, log(lambda*exp((lambda-mu)h)-mu)
//     for psi=0,rho=1) is staged in shared memory so a branch reads its parent's value instead of
//     recomputing it;
//   * sums that the Java accumulates term by term are accumulated factor-wise where that only
//     changes rounding at the 1e-16 level (tolerance of the path is 1e-10 relative).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/gsm_b200.h"

#if defined(__CUDACC__)
#define GSM_HD __host__ __device__ __forceinline__
#else
#define GSM_HD inline
#endif

#define GSM_LUT_M 16           // lgamma(shape*m) table, m = 1..GSM_LUT_M
#define GSM_LOGFACT_N 128      // sum_{i=2..n} log(i) table
#define GSM_MAX_CUM_SUM 0.999  // BSP:39
#define GSM_LEAF_EPS 1e-10     // STP:57 HEIGHT_THRESHOLD_OF_LEAF
#define GSM_K_CAP 100000       // safety cap for the mixture truncation loop (the Java has none)

// tree-density variants (STP:179-214)
enum { GSM_TV_NONE = 0,        // ignoreTreePrior
       GSM_TV_CONDITIONED = 1, // origin / conditionOnRoot            STP:428-487
       GSM_TV_COMPLETE_RHO = 2,// psi == 0 && rho == 1                STP:383-426
       GSM_TV_NO_PSI = 3,      // psi == 0, rho < 1                   STP:351-379
       GSM_TV_PSI = 4 };       // psi > 0                             STP:290-347
// what the per-node aux value holds
enum { GSM_AUX_NONE = 0, GSM_AUX_EXP_C1 = 1 /* exp(-c1*h) */, GSM_AUX_LB = 2 /* log(lambda*e^{(lambda-mu)h}-mu) */ };

// node flag bits as assembled by the kernel
#define GSM_NF_NONLEAF     1u
#define GSM_NF_DA          2u   // Node.isDirectAncestor
#define GSM_NF_FAKE        4u   // Node.isFake
#define GSM_NF_PARENT_FAKE 8u
#define GSM_NF_ROOT        16u

struct GsmTreeConsts {
    // wiring
    uint32_t flags;
    int32_t stub_mode;
    int32_t tree_variant;
    int32_t aux_kind;
    int32_t stp_early;        // 1: STP returns -inf before looking at the tree (STP:160-164, 176-178)
    int32_t spike_invalid;    // shape <= 0 (BSP:60-63)
    int32_t rate_invalid;     // sigma <= 0 (BRP:36-39)
    int32_t use_spike;
    int32_t spike_known_n;    // BSP:77 known-n branch (else mixture)
    int32_t stub_term;        // STP:219 stub density is evaluated
    int32_t stub_sampling;    // !(psi==0 && rho==1)  STP:243-253
    int32_t stub_counts;      // stubs.estimateStubs()
    // scalars
    double base, spike_mean, shape, scale, log_scale, sigma;
    double m_ln, inv_2s2, log_s_sqrt2pi;
    double lambda, mu, psi, rho, origin;
    double lmm;               // lambda - mu
    double c1, c2, kE;        // kE = lambda + mu + psi - c1
    double omc2, opc2;        // 1 - c2, 1 + c2
    double log_lambda, log_psi, log_rho, log_lmm, log_4rho, log_2;
    double np_a, np_b;        // NO_PSI: rho*lambda, lambda*(1-rho)-mu
    double np_const;          // NO_PSI per-internal-node constant  log l + log rho + 2 log(l-m)
    // root-dependent (gsm_finish_consts)
    double T;                 // root height
    double lg[GSM_LUT_M + 1]; // lg[m] = logGamma(shape*m), m >= 1
};

struct GsmAcc {
    double tree, stub, spike, rate, burst, dist;
    int32_t n_extant, n_leaf, n_da, n_extinct, bad;
};
// GsmAcc.bad bits
#define GSM_BAD_NEG_SPIKE 1   // BSP:72-75
#define GSM_BAD_NEG_RATE  2   // BRP:47-50
#define GSM_BAD_SA_STUBS  4   // STP:232-240

GSM_HD double gsm_neg_inf() { return -INFINITY; }

// ---- commons-math Gamma.logGamma (Lanczos, g = 607/128), as BEAST 2.7.7 ships it ----
GSM_HD double gsm_log_gamma(double x) {
    const double L[15] = {0.99999999999999709182,    57.156235665862923517,     -59.597960355475491248,
                          14.136097974741747174,     -0.49191381609762019978,   .33994649984811888699e-4,
                          .46523628927048575665e-4,  -.98374475304879564677e-4, .15808870322491248884e-3,
                          -.21026444172410488319e-3, .21743961811521264320e-3,  -.16431810653676389022e-3,
                          .84418223983852743293e-4,  -.26190838401581408670e-4, .36899182659531622704e-5};
    if (!(x > 0.0)) return NAN;
    double sum = 0.0;
#pragma unroll
    for (int i = 14; i > 0; --i) sum = sum + (L[i] / (x + i));
    sum = sum + L[0];
    double tmp = x + 607.0 / 128.0 + .5;
    return ((x + .5) * log(tmp)) - tmp + 0.91893853320467274178 /* 0.5*log(2*pi) */ + log(sum / x);
}

GSM_HD double gsm_lgamma_shape_m(const GsmTreeConsts& K, int32_t m) {
    return (m <= GSM_LUT_M) ? K.lg[m] : gsm_log_gamma(K.shape * m);
}

// sum_{i=2..n} log(i): table for small n (the table is filled in the Java's summation order), loop beyond
GSM_HD double gsm_logfact(const double* __restrict__ logfact, int32_t n) {
    if (n < GSM_LOGFACT_N) return n < 2 ? 0.0 : logfact[n];
    double s = logfact[GSM_LOGFACT_N - 1];
    for (int32_t i = GSM_LOGFACT_N; i <= n; i++) s += log((double)i);
    return s;
}

// GammaDistributionImpl(alpha = shape*m, beta = 1/shape).logDensity(x), BSP:97-99
GSM_HD double gsm_gamma_logdens(const GsmTreeConsts& K, int32_t m, double x) {
    if (x < 0) return gsm_neg_inf();
    double alpha = K.shape * m;
    return log(x / K.scale) * (alpha - 1) - K.log_scale - x / K.scale - gsm_lgamma_shape_m(K, m);
}

// ---- per-tree constants that do not depend on the tree's heights ----
GSM_HD void gsm_setup_consts(GsmTreeConsts& K, const double* __restrict__ p, int32_t use_spike, uint32_t flags,
                             int32_t stub_mode) {
    K.flags = flags;
    K.stub_mode = stub_mode;
    K.use_spike = use_spike;
    K.base = p[GSM_P_CLOCK_RATE];
    K.spike_mean = p[GSM_P_SPIKE_MEAN];
    K.shape = p[GSM_P_SPIKE_SHAPE];
    K.sigma = p[GSM_P_CLOCK_SD];
    K.lambda = p[GSM_P_LAMBDA];
    K.mu = p[GSM_P_MU];
    K.psi = p[GSM_P_PSI];
    K.rho = p[GSM_P_RHO];
    K.origin = p[GSM_P_ORIGIN];
    K.T = 0.0;

    // BranchSpikePrior BSP:58-64
    K.spike_invalid = (K.shape <= 0) ? 1 : 0;
    K.scale = 1.0 / K.shape;
    K.log_scale = log(K.scale);
    const bool stubs_spike = (flags & GSM_F_STUBS_TO_SPIKE_PRIOR) != 0;
    K.stub_counts = (stub_mode != GSM_STUBS_NOSTUB) ? 1 : 0;                 // Stubs:648-650
    K.spike_known_n = (!stubs_spike || K.stub_counts) ? 1 : 0;               // BSP:77

    // BranchRatePrior BRP:33-41
    K.rate_invalid = (K.sigma <= 0) ? 1 : 0;
    K.m_ln = 0.0 - (0.5 * K.sigma * K.sigma);                                // log(1) - 0.5 s^2
    K.inv_2s2 = 1.0 / (2 * K.sigma * K.sigma);
    K.log_s_sqrt2pi = log(K.sigma * 2.50662827463100050242 /* sqrt(2*pi) */);

    // StumpedTreePrior dispatcher STP:150-214
    K.lmm = K.lambda - K.mu;
    K.stp_early = (K.lambda <= K.mu) ? 1 : 0;                                // STP:160-164
    K.tree_variant = GSM_TV_NONE;
    if (!(flags & GSM_F_IGNORE_TREE_PRIOR)) {
        if (exp(K.psi) == INFINITY || exp(-K.psi) == 0) K.stp_early = 1;     // STP:176-178
        if (flags & (GSM_F_ORIGIN | GSM_F_COND_ROOT)) K.tree_variant = GSM_TV_CONDITIONED;
        else if (K.psi == 0 && K.rho == 1) K.tree_variant = GSM_TV_COMPLETE_RHO;
        else if (K.psi == 0) K.tree_variant = GSM_TV_NO_PSI;
        else K.tree_variant = GSM_TV_PSI;
    }
    K.stub_term = (!(flags & GSM_F_IGNORE_STUB_PRIOR) && (flags & GSM_F_STUBS_TO_TREE_PRIOR)) ? 1 : 0;  // STP:219
    K.stub_sampling = !(K.psi == 0 && K.rho == 1) ? 1 : 0;

    // STP:496-502
    double d = K.lambda - K.mu - K.psi;
    K.c1 = sqrt(fabs(d * d + 4 * K.lambda * K.psi));
    K.c2 = -(K.lambda - K.mu - 2 * K.lambda * K.rho - K.psi) / K.c1;
    K.kE = K.lambda + K.mu + K.psi - K.c1;
    K.omc2 = 1 - K.c2;
    K.opc2 = 1 + K.c2;
    K.log_lambda = log(K.lambda);
    K.log_psi = log(K.psi);
    K.log_rho = log(K.rho);
    K.log_lmm = log(K.lmm);
    K.log_4rho = log(4 * K.rho);
    K.log_2 = 0.69314718055994530942;
    K.np_a = K.rho * K.lambda;
    K.np_b = K.lambda * (1 - K.rho) - K.mu;
    K.np_const = K.log_lambda + K.log_rho + 2 * K.log_lmm;

    // which per-node value children need from their parent
    const bool need_E = (K.stub_term && K.stub_counts) || (!K.spike_known_n);   // Poisson mean of the stub count
    K.aux_kind = GSM_AUX_NONE;
    if (need_E) K.aux_kind = K.stub_sampling ? GSM_AUX_EXP_C1 : GSM_AUX_LB;
    if (K.tree_variant == GSM_TV_COMPLETE_RHO && K.aux_kind == GSM_AUX_NONE) K.aux_kind = GSM_AUX_LB;
}

// lgamma(shape*m) table entry (each entry is independent: the kernel spreads m over the lanes of a warp)
GSM_HD double gsm_lut_entry(double shape, int32_t m) { return m >= 1 ? gsm_log_gamma(shape * m) : 0.0; }

GSM_HD void gsm_finish_consts(GsmTreeConsts& K, double root_height) { K.T = root_height; }

// ---- aux value of a node (see GSM_AUX_*) ----
GSM_HD double gsm_aux(const GsmTreeConsts& K, double h) {
    if (K.aux_kind == GSM_AUX_EXP_C1) return exp(-K.c1 * h);
    if (K.aux_kind == GSM_AUX_LB) return log(K.lambda * exp(K.lmm * h) - K.mu);
    return 0.0;
}

// Expected stub count of the branch (h1 = node height, h0 = parent height), STP:749-769.
//   sampling:    E = (h0-h1)(l+m+psi-c1) + 2 log( ((c2-1)e^{-c1 h1} - c2 - 1) / ((c2-1)e^{-c1 h0} - c2 - 1) )   STP:752-753
//   no sampling: E = I(h1) - I(h0), I(h) = 2( log(l e^{(l-m)h} - m) + l (T-h) )                                     STP:758-769
//                  = 2( Lb(h1) - Lb(h0) + l (h0 - h1) )     [the l*T terms cancel exactly in exact arithmetic]
GSM_HD double gsm_expected_stubs(const GsmTreeConsts& K, double h1, double h0, double aux1, double aux0) {
    if (K.stub_sampling) {
        double cm1 = K.c2 - 1;
        double f = (cm1 * aux1 - K.c2 - 1) / (cm1 * aux0 - K.c2 - 1);
        return (h0 - h1) * K.kE + 2 * log(f);
    }
    return 2 * ((aux1 - aux0) + K.lambda * (h0 - h1));
}

// STP:489-491, with e = exp(-c1 t) supplied and exp(c1 t) = 1/e
GSM_HD double gsm_q_from_e(const GsmTreeConsts& K, double e_neg, double e_pos) {
    return e_pos * K.opc2 * K.opc2 + e_neg * K.omc2 * K.omc2 + 2 * (1 - K.c2 * K.c2);
}
// STP:506-507
GSM_HD double gsm_p0_from_e(const GsmTreeConsts& K, double e_neg) {
    return (K.lambda + K.mu + K.psi - K.c1 * (K.opc2 - e_neg * K.omc2) / (K.opc2 + e_neg * K.omc2)) / (2 * K.lambda);
}

GSM_HD void gsm_acc_zero(GsmAcc& a) {
    a.tree = a.stub = a.spike = a.rate = a.burst = a.dist = 0.0;
    a.n_extant = a.n_leaf = a.n_da = a.n_extinct = a.bad = 0;
}

// Number of spikes on a branch, BSP:197-213
GSM_HD int32_t gsm_n_spikes(uint32_t nf, int32_t nstubs) {
    if (nf & GSM_NF_ROOT) return nstubs + 1;
    if (nf & GSM_NF_DA) return 0;
    if (nf & GSM_NF_PARENT_FAKE) return nstubs;
    return nstubs + 1;
}

// Stubs.getNStubsOnBranch Stubs:344-370
GSM_HD int32_t gsm_nstubs_on_branch(const GsmTreeConsts& K, int32_t i, int32_t n, int32_t raw) {
    if (K.stub_mode == GSM_STUBS_NOSTUB) return -1;
    if (K.stub_mode == GSM_STUBS_COUNTS) return (i >= n - 1) ? 0 : raw;
    return raw;
}

// PRCM.getBurstSize PRCM:178-208
GSM_HD double gsm_burst(const GsmTreeConsts& K, uint32_t nf, double h, double spike, int32_t nst_branch) {
    if ((K.flags & GSM_F_HAS_INDICATOR) && !K.use_spike) return 0.0;                     // PRCM:181-183
    if ((K.flags & GSM_F_NO_SPIKE_DATED_TIPS) && !(nf & GSM_NF_NONLEAF) && h > 0) return 0.0;  // PRCM:185-187
    if (nf & GSM_NF_PARENT_FAKE) {                                                       // PRCM:192-202
        if (!(K.flags & GSM_F_STUBS_TO_CLOCK)) return 0.0;
        if (nst_branch == 0) return 0.0;
    }
    return spike * K.spike_mean;                                                         // PRCM:204-206
}

// Mixture-mode spike density of one branch (BSP:102-172): log sum_k Poisson(k; mu) * Gamma(x; shape*m(k), 1/shape)
GSM_HD double gsm_spike_mixture(const GsmTreeConsts& K, const double* __restrict__ logfact, uint32_t nf, double x,
                                double mu) {
    if (mu > 0) {                                                                        // BSP:112
        double branchP = 0;
        double cumsum = 0;
        const double log_mu = log(mu);
        const double log_x_over_scale = log(x / K.scale);
        const double x_over_scale = x / K.scale;
        int32_t k = 0;
        while (cumsum < GSM_MAX_CUM_SUM && k < GSM_K_CAP) {                              // BSP:117
            double p = -mu + k * log_mu;                                                 // BSP:120
            p -= gsm_logfact(logfact, k);                                                // BSP:121
            double pReal = exp(p);
            cumsum += pReal;
            int32_t m = gsm_n_spikes(nf, k);                                             // BSP:127
            if (m == 0) {
                if (!(x != 0)) branchP += pReal;                                         // BSP:128-135
            } else {
                double alpha = K.shape * m;
                double g = (x < 0) ? gsm_neg_inf()
                                   : log_x_over_scale * (alpha - 1) - K.log_scale - x_over_scale - gsm_lgamma_shape_m(K, m);
                if (!(x == 0 || g == gsm_neg_inf() || g != g)) branchP += exp(p + g);    // BSP:140-144
            }
            k++;
        }
        return log(branchP);                                                             // BSP:151
    }
    int32_t m = gsm_n_spikes(nf, 0);                                                     // BSP:157
    if (m == 0) return (x != 0) ? gsm_neg_inf() : 0.0;                                   // BSP:158-164
    return gsm_gamma_logdens(K, 1, x);                                                   // BSP:166-167
}

// ---- all per-node terms -----------------------------------------------------------------------------------
// i: node number, n: node count of this tree, nf: GSM_NF_* flags, h: height, hp: parent height (h for root),
// aux / auxp: aux value of node / parent, rate / spike / nst: the branch's parameters.
// WANT_RATES: also produce getRateForBranch (PRCM:225-242) and getBurstSize for all nodes (SpikeSize:48-56).
template <bool WANT_RATES>
GSM_HD void gsm_node_terms(const GsmTreeConsts& K, const double* __restrict__ logfact, GsmAcc& A, int32_t i, int32_t n,
                           uint32_t nf, double h, double hp, double aux, double auxp, double rate, double spike,
                           int32_t nst_raw, double* out_rate, double* out_wspike) {
    const bool root = (nf & GSM_NF_ROOT) != 0;
    const bool leaf = !(nf & GSM_NF_NONLEAF);
    const bool da = (nf & GSM_NF_DA) != 0;
    const double len = root ? 0.0 : hp - h;                                  // Node.getLength
    const int32_t nst = gsm_nstubs_on_branch(K, i, n, nst_raw);

    // ---- leaf census (STP:190-195, 291-312, 352-359) ----
    if (leaf) {
        A.n_leaf++;
        if (da) A.n_da++;
        else if (h <= GSM_LEAF_EPS) A.n_extant++;
        else if (len > GSM_LEAF_EPS) A.n_extinct++;
    }

    // ---- clock model: burst, SPL sums, optional per-branch outputs ----
    const double burst = gsm_burst(K, nf, h, spike, nst);
    const bool eligible = !(len <= 0 || da || root);                         // PRCM:229, SPL:41
    double r = 1.0;                                                          // PRCM:210-222
    if ((K.flags & GSM_F_HAS_RATES) && !(K.flags & GSM_F_RELAXED_FALSE)) r = rate;
    if (WANT_RATES) {
        double eff = K.base;
        if (eligible) {
            double dist = len * K.base * r + burst;                          // PRCM:233
            eff = dist / len;                                                // PRCM:236
            A.dist += len * eff;                                             // SPL:44-49
            A.burst += burst;                                                // SPL:53-54
        }
        if (out_rate) *out_rate = eff;
        if (out_wspike) *out_wspike = burst;
    } else if (eligible) {
        // len * ((len*base*r + burst)/len) differs from the numerator by <= 1 ulp
        A.dist += len * K.base * r + burst;
        A.burst += burst;
    }

    // ---- BranchRatePrior BRP:43-60 (all N entries incl. root) ----
    if (K.flags & GSM_F_HAS_RATES) {
        if (rate < 0) A.bad |= GSM_BAD_NEG_RATE;
        if (rate <= 0) {
            A.rate += gsm_neg_inf();                                         // BRP:168-170
        } else {
            double lr = log(rate);
            double x0 = lr - K.m_ln;
            A.rate += -(x0 * x0) * K.inv_2s2 - K.log_s_sqrt2pi - lr;         // BRP:171
        }
    }

    // expected number of stubs on this branch, where some term needs it
    double E = 0.0;
    const bool have_len = hp > h && !root;                                   // STP:558 / 878 (h0 <= h1 -> no stubs)
    if (K.aux_kind != GSM_AUX_NONE && have_len && ((K.stub_term && K.stub_counts) || !K.spike_known_n))
        E = gsm_expected_stubs(K, h, hp, aux, auxp);

    // ---- BranchSpikePrior BSP:69-174 (all N entries incl. root) ----
    if (spike < 0) A.bad |= GSM_BAD_NEG_SPIKE;                               // BSP:72-75
    if (K.spike_known_n) {
        int32_t ns = (K.flags & GSM_F_STUBS_TO_SPIKE_PRIOR) ? nst : 0;       // BSP:81
        int32_t m = gsm_n_spikes(nf, ns);                                    // BSP:85
        if (m == 0) {
            if (spike != 0) A.spike += gsm_neg_inf();                        // BSP:86-95
        } else {
            A.spike += gsm_gamma_logdens(K, m, spike);                       // BSP:97-99
        }
    } else {
        double mu_st = have_len ? E : 0.0;                                   // BSP:106-108, STP:876-892
        A.spike += gsm_spike_mixture(K, logfact, nf, spike, mu_st);
    }

    // ---- stub density, count mode: STP:555-571 / 621-637 ----
    if (K.stub_term) {
        if (da && nst > 0) A.bad |= GSM_BAD_SA_STUBS;                        // STP:232-240
        if (K.stub_counts && !root) {
            if (!have_len) {
                if (nst != 0) A.stub += gsm_neg_inf();                       // STP:558-561
            } else {
                double lp = nst * log(E) - E;                                // STP:566
                lp -= gsm_logfact(logfact, nst);                             // STP:567
                A.stub += lp;
            }
        }
    }

    // ---- tree density, per-node part ----
    switch (K.tree_variant) {
        case GSM_TV_COMPLETE_RHO: {
            // STP:398-422: sum_i [B(h_i) - B(h_parent)], B(h) = tlm - log(l e^{lT} - m e^{tlm + mT}) = -mT - Lb(h)
            // so each branch contributes Lb(h_parent) - Lb(h_i) to blueIntegral (root: 0), i.e. +(aux - auxp) to treeLogP.
            if (!root) A.tree += aux - auxp;
            if (!leaf) {
                // STP:423-426, 826-831: log l + log(1 - Qt), 1 - Qt = (l-m) e^{(l-m)h} / (l e^{(l-m)h} - m)
                A.tree += K.log_lambda + (K.log_lmm + K.lmm * h - aux);
            }
        } break;
        case GSM_TV_NO_PSI: {
            if (!leaf && !(nf & GSM_NF_FAKE)) {                              // STP:371-377
                A.tree += K.np_const - K.lmm * h;
                A.tree -= 2 * log(K.np_a + K.np_b * exp(-K.lmm * h));
            }
        } break;
        case GSM_TV_PSI: {
            if (!leaf) {
                if (!(nf & GSM_NF_FAKE)) {                                   // STP:329-334
                    double e_neg = (K.aux_kind == GSM_AUX_EXP_C1) ? aux : exp(-K.c1 * h);
                    double qv = gsm_q_from_e(K, e_neg, exp(K.c1 * h));       // p1 = 4 rho / q  (STP:523-524 vs 489-491)
                    A.tree += K.log_lambda + (K.log_4rho - log(qv));
                }
            } else if (!(h <= GSM_LEAF_EPS) && !da) {                        // STP:337-343
                double e_neg = (K.aux_kind == GSM_AUX_EXP_C1) ? aux : exp(-K.c1 * h);
                double qv = gsm_q_from_e(K, e_neg, exp(K.c1 * h));
                A.tree += log(gsm_p0_from_e(K, e_neg)) - (K.log_4rho - log(qv));
            }
        } break;
        case GSM_TV_CONDITIONED: {                                           // STP:467-483
            if (leaf) {
                if (!da) {
                    if (h > 0.000000000005 || K.rho == 0.) {
                        double e_neg = (K.aux_kind == GSM_AUX_EXP_C1) ? aux : exp(-K.c1 * h);
                        double qv = gsm_q_from_e(K, e_neg, exp(K.c1 * h));
                        A.tree += K.log_psi + log(qv) + log(gsm_p0_from_e(K, e_neg));
                    } else {
                        A.tree += K.log_4rho;
                    }
                }
            } else if (nf & GSM_NF_FAKE) {
                A.tree += K.log_psi;
            } else {
                double e_neg = (K.aux_kind == GSM_AUX_EXP_C1) ? aux : exp(-K.c1 * h);
                double qv = gsm_q_from_e(K, e_neg, exp(K.c1 * h));
                A.tree += K.log_lambda - log(qv);
            }
        } break;
        default: break;
    }
}

// ---- per-tree epilogue: combine reduced sums into the result row (GSM_O_*) ----
GSM_HD void gsm_finalize(const GsmTreeConsts& K, const GsmAcc& A, int32_t n, bool root_fake, double* __restrict__ out) {
    // BranchSpikePrior
    double sp = A.spike;
    if (K.spike_invalid || (A.bad & GSM_BAD_NEG_SPIKE)) sp = gsm_neg_inf();                 // BSP:60-63, 72-75
    else if (sp == INFINITY) sp = gsm_neg_inf();                                            // BSP:177-179
    out[GSM_O_SPIKE_LOGP] = sp;
    // BranchRatePrior
    double rp = 0.0;
    if (K.flags & GSM_F_HAS_RATES) {
        rp = A.rate;
        if (K.rate_invalid || (A.bad & GSM_BAD_NEG_RATE)) rp = gsm_neg_inf();               // BRP:36-39, 47-50
    }
    out[GSM_O_RATE_LOGP] = rp;
    out[GSM_O_SUM_BURST] = A.burst;
    out[GSM_O_SUM_DIST] = A.dist;
    out[GSM_O_N_EXTANT] = (double)A.n_extant;

    // StumpedTreePrior STP:150-284
    double tree = 0.0, stub = 0.0, stp;
    bool early = K.stp_early != 0;
    if (!early && K.tree_variant != GSM_TV_NONE) {
        const double T = K.T;
        switch (K.tree_variant) {
            case GSM_TV_CONDITIONED: {
                double x0 = (K.flags & GSM_F_ORIGIN) ? K.origin : INFINITY;                 // STP:180
                double x1 = T;
                if (x0 < x1) { early = true; break; }                                       // STP:182-184
                const bool cond_root = (K.flags & GSM_F_COND_ROOT) != 0;
                if (cond_root && root_fake) { tree = gsm_neg_inf(); break; }                // STP:437-440
                double xq = cond_root ? x1 : x0;
                double e_neg = exp(-K.c1 * xq);
                tree = -log(gsm_q_from_e(K, e_neg, exp(K.c1 * xq)));                        // STP:435 / 442
                if (K.flags & GSM_F_COND_SAMPLING) {                                        // STP:446-453
                    double omp0 = 1 - gsm_p0_from_e(K, e_neg);
                    tree -= cond_root ? log(K.lambda * omp0 * omp0) : log(omp0);
                }
                if (K.flags & GSM_F_COND_RHO_SAMPLING) {                                    // STP:455-461, 492-494
                    double omp0h = K.rho * K.lmm / (K.lambda * K.rho + K.np_b * exp((K.mu - K.lambda) * xq));
                    tree -= cond_root ? log(K.lambda * omp0h * omp0h) : log(omp0h);
                }
                int32_t internal = A.n_leaf - A.n_da - 1;                                   // STP:463
                tree += internal * K.log_2;                                                 // STP:465
                tree += A.tree;
            } break;
            case GSM_TV_COMPLETE_RHO: {
                // Java's B(h) needs exp(lambda*T): when that overflows every branch integral is NaN (STP:420)
                if (!(exp(K.lambda * T) < INFINITY)) tree = NAN;
                else tree = A.tree;
            } break;
            case GSM_TV_NO_PSI: {
                if (A.n_extant != A.n_leaf) { tree = gsm_neg_inf(); break; }                // STP:360-362
                tree = log((double)A.n_extant) + K.log_lmm - K.lmm * T;                     // STP:367
                tree -= log(K.np_a + K.np_b * exp(-K.lmm * T));                             // STP:368
                tree += A.tree;
            } break;
            case GSM_TV_PSI: {
                if (A.n_extinct + A.n_extant + A.n_da != A.n_leaf) { tree = gsm_neg_inf(); break; }   // STP:313-316
                tree = log(4.0) + log((double)A.n_extant) + K.log_rho + K.log_lambda +
                       K.log_psi * (double)(A.n_da + A.n_extinct);                          // STP:325
                tree -= log(K.c1) + log(K.c2 + 1) + log(1 - K.c2 + (1 + K.c2) * exp(K.c1 * T));   // STP:326
                tree += A.tree;
            } break;
            default: break;
        }
    }
    if (early) {
        out[GSM_O_STP_LOGP] = gsm_neg_inf();
        out[GSM_O_TREE_LOGP] = gsm_neg_inf();
        out[GSM_O_STUB_LOGP] = 0.0;
        return;
    }
    if (K.stub_term) {
        if (A.bad & GSM_BAD_SA_STUBS) {                                                     // STP:232-240
            out[GSM_O_STP_LOGP] = gsm_neg_inf();
            out[GSM_O_TREE_LOGP] = tree;
            out[GSM_O_STUB_LOGP] = gsm_neg_inf();
            return;
        }
        stub = A.stub;
    }
    stp = tree + stub;                                                                      // STP:280
    out[GSM_O_STP_LOGP] = stp;
    out[GSM_O_TREE_LOGP] = tree;
    out[GSM_O_STUB_LOGP] = stub;
    (void)n;
}
