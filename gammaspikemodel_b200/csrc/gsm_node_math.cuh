// gsm_node_math.cuh — per-tree constants and per-node (per-branch) terms of the GammaSpikeModel hot path.
//
// Everything here is GSM_HD (__host__ __device__) so that the exact code the sm_100a kernels run can
// also be compiled by g++ for the CPU-side logic tests (tests/emu/gsm_emu.cpp).  The product only ever
// runs it inside the kernels of gsm_kernels.cu.
//
// Citations (relative to the reference tree):
//   PRCM  src/gammaspike/clockmodel/PunctuatedRelaxedClockModel.java
//   BSP   src/gammaspike/distribution/BranchSpikePrior.java
//   BRP   src/gammaspike/distribution/BranchRatePrior.java
//   STP   src/gammaspike/distribution/StumpedTreePrior.java
//   Stubs src/gammaspike/tree/Stubs.java
//   SPL   src/gammaspike/logger/SaltativeProportionLogger.java
//
// ---- the algebra (DESIGN.md §4 has the derivations and the conditioning argument) ----
// With c1, c2 of STP:496-502 define, for a node of height h,
//        G(h)   = (1+c2) + (1-c2) e^{-c1 h}              ( = -[(c2-1)e^{-c1 h} - c2 - 1], the bracket of STP:752 )
//        aux(h) = log G(h)                                one exp + one log per node, staged in shared memory.
// Then, exactly (in real arithmetic):
//   * q(h) of STP:489-491        = e^{c1 h} G(h)^2                  =>  log q  = c1 h + 2 aux
//   * p1(h) of STP:523-524       = 4 rho / q(h)                     =>  log p1 = log(4 rho) - c1 h - 2 aux
//   * p0(h) of STP:506-507       = [ (l+m+psi) G - c1 ((1+c2) - (1-c2) e^{-c1 h}) ] / (2 l G)
//   * E(branch) of STP:749-754   = (h0-h1)(l+m+psi-c1) + 2 (aux(h1) - aux(h0))          (no division, no extra log)
//   * psi = 0, rho = 1 (STP:758-769, 383-426): c1 = l-m, G = 2/(l-m) e^{-(l-m)h} (l e^{(l-m)h} - m), so
//       Lb(h) := log(l e^{(l-m)h} - m) = aux - log(2/(l-m)) + (l-m) h,
//       E = 2 (Lb(h1) - Lb(h0) + l (h0-h1))  is the same number as the general E above,
//       B(h1) - B(h0) of STP:409 = Lb(h0) - Lb(h1),  and  log(1 - Qt(h)) of STP:425 = log(l-m) + (l-m)h - Lb(h).
//   * psi = 0, rho < 1 (STP:351-379): rho l + (l(1-rho) - m) e^{-(l-m)h} = (l-m)/2 * G(h).
// G > 0 for every valid state (lambda > mu >= 0, 0 <= rho <= 1, h >= 0).  Each rewrite is at least as
// well conditioned as the Java expression it replaces, so results agree with the reference to the
// rounding noise of the reference itself (observed <= 1e-12 relative; tolerance 1e-10).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/gsm_b200.h"
#include "gsm_fastmath.cuh"

#define GSM_LUT_M 16           // per-tree table over m = 1..GSM_LUT_M spikes on a branch
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
    int32_t need_aux;         // some term needs aux(h)
    int32_t need_E;           // some term needs the expected stub count of a branch
    int32_t stp_early;        // 1: STP returns -inf before looking at the tree (STP:160-164, 176-178)
    int32_t spike_invalid;    // shape <= 0 (BSP:60-63)
    int32_t rate_invalid;     // sigma <= 0 (BRP:36-39)
    int32_t use_spike;
    int32_t spike_known_n;    // BSP:77 known-n branch (else mixture)
    int32_t stub_term;        // STP:219 stub density is evaluated
    int32_t stub_counts;      // stubs.estimateStubs()
    int32_t stub_nosampling;  // psi == 0 && rho == 1: the Java uses the closed forms of STP:758-769
    // scalars
    double base, spike_mean, shape, log_shape, sigma;
    double m_ln, inv_2s2, log_s_sqrt2pi;
    double lambda, mu, psi, rho, origin;
    double lmm;               // lambda - mu
    double c1, c2, kE, kS;    // kE = lambda + mu + psi - c1, kS = lambda + mu + psi
    double omc2, opc2;        // 1 - c2, 1 + c2
    double log_lambda, log_psi, log_rho, log_lmm, log_4rho, log_2lambda;
    double t_internal;        // per-internal-node constant of the active tree variant
    double t_leaf;            // per-leaf constant of the active tree variant
    double lb_shift;          // log(2/(l-m)): Lb(h) = aux - lb_shift + (l-m) h
    double T;                 // root height (gsm_finish_consts)
    // per-m tables (m = number of spikes on the branch): Gamma(shape*m, 1/shape) log-density is
    //   am1[m] * log(x*shape) - x*shape + lgc[m]
    double am1[GSM_LUT_M + 1];   // shape*m - 1
    double lgc[GSM_LUT_M + 1];   // log(shape) - logGamma(shape*m)
};

struct GsmAcc {
    double tree, stub, spike, rate, burst, dist;
    int32_t n_extant, n_leaf, n_da, n_extinct, bad;
};
// GsmAcc.bad bits
#define GSM_BAD_NEG_SPIKE 1   // BSP:72-75
#define GSM_BAD_NEG_RATE  2   // BRP:47-50
#define GSM_BAD_SA_STUBS  4   // STP:232-240
#define GSM_BAD_THROWS    8   // the Java throws inside calculateTreeLogLikelihood (Stubs:631-634) -> -inf (STP:258-265)

GSM_HD double gsm_neg_inf() { return -INFINITY; }

// ---- commons-math Gamma.logGamma (Lanczos, g = 607/128), as BEAST 2.7.7 ships it ----
#define GSM_LANCZOS_INIT {0.99999999999999709182,    57.156235665862923517,     -59.597960355475491248,  \
                          14.136097974741747174,     -0.49191381609762019978,   .33994649984811888699e-4, \
                          .46523628927048575665e-4,  -.98374475304879564677e-4, .15808870322491248884e-3, \
                          -.21026444172410488319e-3, .21743961811521264320e-3,  -.16431810653676389022e-3, \
                          .84418223983852743293e-4,  -.26190838401581408670e-4, .36899182659531622704e-5}
static const double gsm_lanczos_h[15] = GSM_LANCZOS_INIT;
#if defined(__CUDACC__)
static __constant__ double gsm_lanczos_d[15] = GSM_LANCZOS_INIT;
#endif
GSM_HD double gsm_log_gamma(double x) {
#if defined(__CUDA_ARCH__)
    const double* __restrict__ L = gsm_lanczos_d;
#else
    const double* L = gsm_lanczos_h;
#endif
    if (!(x > 0.0)) return NAN;
    double sum = 0.0;
#pragma unroll 1
    for (int i = 14; i > 0; --i) sum = sum + (L[i] / (x + i));
    sum = sum + L[0];
    double tmp = x + 607.0 / 128.0 + .5;
    return ((x + .5) * gsm_log(tmp)) - tmp + 0.91893853320467274178 /* 0.5*log(2*pi) */ + gsm_log(sum / x);
}

// sum_{i=2..n} log(i): table for small n (filled in the Java's summation order), loop beyond
GSM_HD double gsm_logfact(const double* __restrict__ logfact, int32_t n) {
    if (n < GSM_LOGFACT_N) return n < 2 ? 0.0 : logfact[n];
    double s = logfact[GSM_LOGFACT_N - 1];
    for (int32_t i = GSM_LOGFACT_N; i <= n; i++) s += gsm_log((double)i);
    return s;
}

// GammaDistributionImpl(alpha = shape*m, beta = 1/shape).logDensity(x), BSP:97-99:
//   log(x/beta)(alpha-1) - log(beta) - x/beta - logGamma(alpha)  with x/beta = x*shape, -log(beta) = log(shape).
// lxs = log(x*shape), xs = x*shape.  (x == 0: lxs = -inf, so alpha > 1 -> -inf, alpha < 1 -> +inf, alpha == 1 -> NaN,
// exactly as the Java expression; x < 0 -> -inf.)
GSM_HD double gsm_gamma_logdens(const GsmTreeConsts& K, int32_t m, double x, double xs, double lxs) {
    if (x < 0) return gsm_neg_inf();
    if (m <= GSM_LUT_M) return lxs * K.am1[m] - xs + K.lgc[m];
    const double alpha = K.shape * m;
    return lxs * (alpha - 1) - xs + (K.log_shape - gsm_log_gamma(alpha));
}

// ---- per-tree constants that do not depend on the tree's heights ----
// Split in three so that the lean kernel can evaluate the transcendental functions on the lanes of one warp:
//   gsm_setup_scalars  wiring + arithmetic-only constants
//   gsm_setup_arg(j)   argument of the j-th logarithm (j < GSM_SETUP_NLOG)
//   gsm_setup_finish   stores the logarithms and everything derived from them
#define GSM_SETUP_NLOG 9
GSM_HD void gsm_setup_scalars(GsmTreeConsts& K, const double* __restrict__ p, int32_t use_spike, uint32_t flags,
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
    const bool stubs_spike = (flags & GSM_F_STUBS_TO_SPIKE_PRIOR) != 0;
    K.stub_counts = (stub_mode != GSM_STUBS_NOSTUB) ? 1 : 0;                 // Stubs:648-650
    K.spike_known_n = (!stubs_spike || K.stub_counts) ? 1 : 0;               // BSP:77

    // BranchRatePrior BRP:33-41
    K.rate_invalid = (K.sigma <= 0) ? 1 : 0;
    K.m_ln = 0.0 - (0.5 * K.sigma * K.sigma);                                // log(1) - 0.5 s^2
    K.inv_2s2 = 1.0 / (2 * K.sigma * K.sigma);

    // StumpedTreePrior dispatcher STP:150-214
    K.lmm = K.lambda - K.mu;
    K.stp_early = (K.lambda <= K.mu) ? 1 : 0;                                // STP:160-164
    K.tree_variant = GSM_TV_NONE;
    if (!(flags & GSM_F_IGNORE_TREE_PRIOR)) {
        if (flags & (GSM_F_ORIGIN | GSM_F_COND_ROOT)) K.tree_variant = GSM_TV_CONDITIONED;
        else if (K.psi == 0 && K.rho == 1) K.tree_variant = GSM_TV_COMPLETE_RHO;
        else if (K.psi == 0) K.tree_variant = GSM_TV_NO_PSI;
        else K.tree_variant = GSM_TV_PSI;
    }
    K.stub_term = (!(flags & GSM_F_IGNORE_STUB_PRIOR) && (flags & GSM_F_STUBS_TO_TREE_PRIOR)) ? 1 : 0;  // STP:219
    K.stub_nosampling = (K.psi == 0 && K.rho == 1) ? 1 : 0;                  // STP:243

    // STP:496-502
    const double d = K.lambda - K.mu - K.psi;
    K.c1 = sqrt(fabs(d * d + 4 * K.lambda * K.psi));
    K.c2 = -(K.lambda - K.mu - 2 * K.lambda * K.rho - K.psi) / K.c1;
    K.kS = K.lambda + K.mu + K.psi;
    K.kE = K.kS - K.c1;
    K.omc2 = 1 - K.c2;
    K.opc2 = 1 + K.c2;
    K.need_E = ((K.stub_term && K.stub_counts) || (!K.spike_known_n)) ? 1 : 0;   // Poisson mean of the stub count
}

GSM_HD double gsm_setup_arg(const GsmTreeConsts& K, int32_t j) {
    switch (j) {
        case 0: return K.shape;
        case 1: return K.sigma * 2.50662827463100050242 /* sqrt(2*pi) */;
        case 2: return K.lambda;
        case 3: return K.psi;
        case 4: return K.rho;
        case 5: return K.lmm;
        case 6: return 4 * K.rho;
        case 7: return 2 * K.lambda;
        case 8: return 2.0 / K.lmm;
        default: return 1.0;
    }
}

// lg[j] = log(gsm_setup_arg(K, j)); exp_psi = exp(psi), exp_npsi = exp(-psi)
GSM_HD void gsm_setup_finish(GsmTreeConsts& K, const double* lg, double exp_psi, double exp_npsi) {
    K.log_shape = lg[0];
    K.log_s_sqrt2pi = lg[1];
    K.log_lambda = lg[2];
    K.log_psi = lg[3];
    K.log_rho = lg[4];
    K.log_lmm = lg[5];
    K.log_4rho = lg[6];
    K.log_2lambda = lg[7];
    K.lb_shift = lg[8];
    if (!(K.flags & GSM_F_IGNORE_TREE_PRIOR)) {
        if (exp_psi == INFINITY || exp_npsi == 0) K.stp_early = 1;           // STP:176-178
    }
    K.t_internal = 0.0;
    K.t_leaf = 0.0;
    switch (K.tree_variant) {
        case GSM_TV_COMPLETE_RHO:   // log l + log(1-Qt) = log l + log(l-m) + (l-m)h - Lb = [log l + log(l-m) + lb_shift] - aux
            K.t_internal = K.log_lambda + K.log_lmm + K.lb_shift; break;
        case GSM_TV_NO_PSI:         // log l + log rho + 2 log(l-m) - (l-m)h - 2 log((l-m)/2 G)   (STP:375-376)
            K.t_internal = K.log_lambda + K.log_rho + 2 * K.log_lmm - 2 * (K.log_lmm - 0.69314718055994530942); break;
        case GSM_TV_PSI:            // log l + log p1 = log l + log 4rho - c1 h - 2 aux               (STP:333)
            K.t_internal = K.log_lambda + K.log_4rho;
            K.t_leaf = -K.log_2lambda - K.log_4rho;                          // log p0 - log p1           (STP:342)
            break;
        case GSM_TV_CONDITIONED:    // log l - log q                                                   (STP:480)
            K.t_internal = K.log_lambda;
            K.t_leaf = K.log_psi - K.log_2lambda;                            // log psi + log q + log p0  (STP:471)
            break;
        default: break;
    }
    K.need_aux = (K.need_E || (K.tree_variant != GSM_TV_NONE && !K.stp_early)) ? 1 : 0;
}

GSM_HD void gsm_setup_consts(GsmTreeConsts& K, const double* __restrict__ p, int32_t use_spike, uint32_t flags,
                             int32_t stub_mode) {
    gsm_setup_scalars(K, p, use_spike, flags, stub_mode);
    double lg[GSM_SETUP_NLOG];
    for (int32_t j = 0; j < GSM_SETUP_NLOG; j++) lg[j] = gsm_log(gsm_setup_arg(K, j));
    gsm_setup_finish(K, lg, gsm_exp(K.psi), gsm_exp(-K.psi));
}

// one entry of the per-m tables (entries are independent: the kernel spreads m over the lanes of a warp)
GSM_HD void gsm_lut_entry(GsmTreeConsts& K, double shape, int32_t m) {
    if (m < 1) {
        K.am1[0] = -1.0;
        K.lgc[0] = 0.0;
        return;
    }
    K.am1[m] = shape * m - 1;
    K.lgc[m] = gsm_log(shape) - gsm_log_gamma(shape * m);
}

GSM_HD void gsm_finish_consts(GsmTreeConsts& K, double root_height) { K.T = root_height; }

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

// ---- pass A (needs height, parent height, node flags; no HBM traffic beyond the staged rows) --------------
// Leaf census, the node's own tree-density term, and aux(h) for the children / pass B.  Returns aux.
GSM_HD double gsm_node_pass_a(const GsmTreeConsts& K, GsmAcc& A, uint32_t nf, double h, double hp) {
    const bool root = (nf & GSM_NF_ROOT) != 0;
    const bool leaf = !(nf & GSM_NF_NONLEAF);
    const bool da = (nf & GSM_NF_DA) != 0;
    const double len = root ? 0.0 : hp - h;

    if (leaf) {                                                              // STP:190-195, 291-312, 352-359
        A.n_leaf++;
        if (da) A.n_da++;
        else if (h <= GSM_LEAF_EPS) A.n_extant++;
        else if (len > GSM_LEAF_EPS) A.n_extinct++;
    }
    if (!K.need_aux) return 0.0;

    const double e_neg = gsm_exp(-K.c1 * h);
    const double G = fma(K.omc2, e_neg, K.opc2);
    const double aux = gsm_log(G);

    // Where the Java's own expressions leave the finite range the aux algebra does not follow them by itself (it has no
    // exp(+c1 h) and no log(e^{(l-m)h} - 1)): those nodes are evaluated with the Java expressions, literally, so that the
    // state gets the reference's -inf / NaN (reproduce, do not fix).  Both conditions are false for every
    // state BEAST can reach (heights >= 0, c1 T far below 700); the lean kernel hands such trees to this path.
    const bool java_overflow = K.c1 * h > 700.0;                             // exp(c1 h) (1 + c2)^2 overflows near c1 h = 709
    if ((K.tree_variant == GSM_TV_PSI || K.tree_variant == GSM_TV_CONDITIONED) && java_overflow) {
        const double ep = exp(K.c1 * h), en = exp(-K.c1 * h);
        const double q = ep * K.opc2 * K.opc2 + en * K.omc2 * K.omc2 + 2 * (1 - K.c2 * K.c2);                 // STP:489-491
        const double p0 = (K.kS - K.c1 * (K.opc2 - en * K.omc2) / (K.opc2 + en * K.omc2)) / (2 * K.lambda);   // STP:506-507
        const double p1 = (4 * K.rho) / (2 * (1 - K.c2 * K.c2) + en * K.omc2 * K.omc2 + ep * K.opc2 * K.opc2);   // STP:523-524
        if (K.tree_variant == GSM_TV_PSI) {
            if (!leaf) {
                if (!(nf & GSM_NF_FAKE)) A.tree += log(K.lambda) + log(p1);                                   // STP:333
            } else if (!(h <= GSM_LEAF_EPS) && !da) {
                A.tree += log(p0) - log(p1);                                                                  // STP:342
            }
        } else {
            if (leaf) {
                if (!da) A.tree += log(K.psi) + log(q) + log(p0);                                             // STP:471 (h > 5e-12 here)
            } else if (nf & GSM_NF_FAKE) {
                A.tree += log(K.psi);                                                                         // STP:478
            } else {
                A.tree += log(K.lambda) - log(q);                                                             // STP:480
            }
        }
        return aux;
    }
    if (K.tree_variant == GSM_TV_COMPLETE_RHO && !leaf && h < 0.0) {         // STP:423-426, 826-831 with a negative height
        const double e = exp(K.lmm * h);
        const double log_qt = log(K.mu * (e - 1)) - log(K.lambda * e - K.mu);
        A.tree += log(K.lambda) + log(1 - exp(log_qt));
        return aux;
    }

    switch (K.tree_variant) {
        case GSM_TV_COMPLETE_RHO:                                            // STP:387-393, 423-426
            if (!leaf) A.tree += K.t_internal - aux;
            break;
        case GSM_TV_NO_PSI:                                                  // STP:371-377
            if (!leaf && !(nf & GSM_NF_FAKE)) A.tree += K.t_internal - K.lmm * h - 2 * aux;
            break;
        case GSM_TV_PSI:
            if (!leaf) {                                                     // STP:329-334
                if (!(nf & GSM_NF_FAKE)) A.tree += K.t_internal - K.c1 * h - 2 * aux;
            } else if (!(h <= GSM_LEAF_EPS) && !da) {                        // STP:337-343: log p0 - log p1
                const double numer = K.kS * G - K.c1 * (K.opc2 - e_neg * K.omc2);
                A.tree += gsm_log(numer) + K.t_leaf + K.c1 * h + aux;        // (log numer - log 2l - aux) - (log 4rho - c1 h - 2 aux)
            }
            break;
        case GSM_TV_CONDITIONED:                                             // STP:467-483
            if (leaf) {
                if (!da) {
                    if (h > 0.000000000005 || K.rho == 0.) {
                        const double numer = K.kS * G - K.c1 * (K.opc2 - e_neg * K.omc2);
                        A.tree += gsm_log(numer) + K.t_leaf + K.c1 * h + aux;   // log psi + (c1 h + 2 aux) + (log numer - log 2l - aux)
                    } else {
                        A.tree += K.log_4rho;
                    }
                }
            } else if (nf & GSM_NF_FAKE) {
                A.tree += K.log_psi;
            } else {
                A.tree += K.t_internal - K.c1 * h - 2 * aux;
            }
            break;
        default: break;
    }
    return aux;
}

// Mixture-mode spike density of one branch (BSP:102-172): log sum_k Poisson(k; mu) * Gamma(x; shape*m(k), 1/shape)
GSM_HD double gsm_spike_mixture(const GsmTreeConsts& K, const double* __restrict__ logfact, uint32_t nf, double x,
                                double xs, double lxs, double mu) {
    if (mu > 0) {                                                                        // BSP:112
        double branchP = 0;
        double cumsum = 0;
        const double log_mu = gsm_log(mu);
        int32_t k = 0;
        while (cumsum < GSM_MAX_CUM_SUM && k < GSM_K_CAP) {                              // BSP:117
            double p = -mu + k * log_mu;                                                 // BSP:120
            p -= gsm_logfact(logfact, k);                                                // BSP:121
            const double pReal = gsm_exp(p);
            cumsum += pReal;
            const int32_t m = gsm_n_spikes(nf, k);                                       // BSP:127
            if (m == 0) {
                if (!(x != 0)) branchP += pReal;                                         // BSP:128-135
            } else {
                const double g = gsm_gamma_logdens(K, m, x, xs, lxs);
                if (!(x == 0 || g == gsm_neg_inf() || g != g)) branchP += gsm_exp(p + g);   // BSP:140-144
            }
            k++;
        }
        return gsm_log(branchP);                                                         // BSP:151
    }
    const int32_t m = gsm_n_spikes(nf, 0);                                               // BSP:157
    if (m == 0) return (x != 0) ? gsm_neg_inf() : 0.0;                                   // BSP:158-164
    return gsm_gamma_logdens(K, 1, x, xs, lxs);                                          // BSP:166-167
}

// ---- pass B: everything that needs the branch's own parameters (rate, spike, stub count) -------------------
// i: node number, n: node count of this tree, nf: GSM_NF_* flags, h: height, hp: parent height (h for root),
// aux / auxp: aux(h) of node / parent.  WANT_RATES: also produce getRateForBranch (PRCM:225-242) and
// getBurstSize for all nodes (SpikeSize:48-56).
template <bool WANT_RATES>
GSM_HD void gsm_node_pass_b(const GsmTreeConsts& K, const double* __restrict__ logfact, GsmAcc& A, int32_t i, int32_t n,
                            uint32_t nf, double h, double hp, double aux, double auxp, double rate, double spike,
                            int32_t nst_raw, double* out_rate, double* out_wspike) {
    const bool root = (nf & GSM_NF_ROOT) != 0;
    const bool da = (nf & GSM_NF_DA) != 0;
    const double len = root ? 0.0 : hp - h;                                  // Node.getLength
    const int32_t nst = gsm_nstubs_on_branch(K, i, n, nst_raw);

    // ---- clock model: burst, SPL sums, optional per-branch outputs ----
    const double burst = gsm_burst(K, nf, h, spike, nst);
    const bool eligible = !(len <= 0 || da || root);                         // PRCM:229, SPL:41
    double r = 1.0;                                                          // PRCM:210-222
    if ((K.flags & GSM_F_HAS_RATES) && !(K.flags & GSM_F_RELAXED_FALSE)) r = rate;
    if (WANT_RATES) {
        double eff = K.base;
        if (eligible) {
            const double dist = len * K.base * r + burst;                    // PRCM:233
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
            const double lr = gsm_log(rate);
            const double x0 = lr - K.m_ln;
            A.rate += -(x0 * x0) * K.inv_2s2 - K.log_s_sqrt2pi - lr;         // BRP:171
        }
    }

    // expected number of stubs on this branch (STP:749-769), where some term needs it
    const bool have_len = hp > h && !root;                                   // STP:558 / 878 (h0 <= h1 -> no stubs)
    const bool rj = K.stub_mode == GSM_STUBS_EXACT;                          // Stubs.getReversibleJump
    double E = 0.0;
    // the reversible-jump branch of the Java has no h0 <= h1 rule: E is whatever STP:749-769 give (0 for the root)
    if (K.need_E && (have_len || (rj && !root))) E = fma(len, K.kE, 2 * (aux - auxp));

    // ---- BranchSpikePrior BSP:69-174 (all N entries incl. root) ----
    if (spike < 0) A.bad |= GSM_BAD_NEG_SPIKE;                               // BSP:72-75
    const double xs = spike * K.shape;
    if (K.spike_known_n) {
        const int32_t ns = (K.flags & GSM_F_STUBS_TO_SPIKE_PRIOR) ? nst : 0; // BSP:81
        const int32_t m = gsm_n_spikes(nf, ns);                              // BSP:85
        if (m == 0) {
            if (spike != 0) A.spike += gsm_neg_inf();                        // BSP:86-95
        } else {
            A.spike += gsm_gamma_logdens(K, m, spike, xs, gsm_log(xs));      // BSP:97-99
        }
    } else {
        A.spike += gsm_spike_mixture(K, logfact, nf, spike, xs, gsm_log(xs), have_len ? E : 0.0);   // BSP:106-108, STP:876-892
    }

    // ---- stub density, count mode: STP:555-571 / 621-637 ----
    if (K.stub_term) {
        if (da && nst > 0) A.bad |= GSM_BAD_SA_STUBS;                        // STP:232-240
        if (rj) {
            // per-branch part of STP:573-600 / 639-663: -E - sum_{i=2..count} log i (every node, the root with E = 0);
            // the per-stub birth-rate terms are added by gsm_stub_term
            A.stub += -E - gsm_logfact(logfact, nst);
        } else if (K.stub_counts && !root) {
            if (!have_len) {
                if (nst != 0) A.stub += gsm_neg_inf();                       // STP:558-561
            } else {
                double lp = nst * gsm_log(E) - E;                            // STP:566
                lp -= gsm_logfact(logfact, nst);                             // STP:567
                A.stub += lp;
            }
        }
    }

    // ---- psi = 0, rho = 1 tree density: the branch integral of STP:398-416 ----
    // B(h_i) - B(h_parent) = Lb(h_parent) - Lb(h_i), Lb = aux - lb_shift + (l-m)h  =>  -(integral) = (aux - auxp) - (l-m) len
    if (K.tree_variant == GSM_TV_COMPLETE_RHO && !root) A.tree += (aux - auxp) - K.lmm * len;
}

// ---- one listed stub of EXACT mode: STP:573-592 / 639-655 with Stubs.getAbsoluteTimeOfStub (Stubs:625-640) ----
// h / hp: heights of the stub's branch node and of its parent; leaf_da: the node isDirectAncestor.
GSM_HD void gsm_stub_term(const GsmTreeConsts& K, GsmAcc& A, bool root, bool leaf_da, double h, double hp, double rel_time) {
    if (root) {                                                              // Stubs:631-634 throws -> STP:258-265
        A.bad |= GSM_BAD_THROWS;
        return;
    }
    const double t = h + rel_time * (hp - h);                                // Stubs:636
    if (t < h || t > hp) { A.stub += gsm_neg_inf(); return; }                // STP:580-582 / 647-649
    double log_qt;
    if (K.stub_nosampling) {                                                 // STP:826-831
        const double e = gsm_exp(K.lmm * t);
        log_qt = gsm_log(K.mu * (e - 1)) - gsm_log(K.lambda * e - K.mu);
    } else {                                                                 // STP:813-820
        if (leaf_da) { A.stub += gsm_neg_inf(); return; }                    // STP:583-585
        const double e = gsm_exp(-K.c1 * t);
        const double top = e * K.omc2 - K.opc2, btm = e * K.omc2 + K.opc2;
        log_qt = gsm_log(K.kS + K.c1 * top / btm) - K.log_2lambda;
    }
    A.stub += 0.69314718055994530942 + K.log_lambda + log_qt;                // STP:796-806
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
                const double x0 = (K.flags & GSM_F_ORIGIN) ? K.origin : INFINITY;           // STP:180
                const double x1 = T;
                if (x0 < x1) { early = true; break; }                                       // STP:182-184
                const bool cond_root = (K.flags & GSM_F_COND_ROOT) != 0;
                if (cond_root && root_fake) { tree = gsm_neg_inf(); break; }                // STP:437-440
                const double xq = cond_root ? x1 : x0;
                const double e_neg = gsm_exp(-K.c1 * xq);
                const double e_pos = gsm_exp(K.c1 * xq);
                tree = -gsm_log(e_pos * K.opc2 * K.opc2 + e_neg * K.omc2 * K.omc2 + 2 * (1 - K.c2 * K.c2));   // STP:435 / 442, 489-491
                if (K.flags & GSM_F_COND_SAMPLING) {                                        // STP:446-453, 506-507
                    const double p0 = (K.kS - K.c1 * (K.opc2 - e_neg * K.omc2) / (K.opc2 + e_neg * K.omc2)) / (2 * K.lambda);
                    const double omp0 = 1 - p0;
                    tree -= cond_root ? gsm_log(K.lambda * omp0 * omp0) : gsm_log(omp0);
                }
                if (K.flags & GSM_F_COND_RHO_SAMPLING) {                                    // STP:455-461, 492-494
                    const double omp0h = K.rho * K.lmm / (K.lambda * K.rho + (K.lambda * (1 - K.rho) - K.mu) * gsm_exp((K.mu - K.lambda) * xq));
                    tree -= cond_root ? gsm_log(K.lambda * omp0h * omp0h) : gsm_log(omp0h);
                }
                const int32_t internal = A.n_leaf - A.n_da - 1;                             // STP:463
                tree += internal * 0.69314718055994530942;                                  // STP:465
                tree += A.tree;
            } break;
            case GSM_TV_COMPLETE_RHO: {
                // Java's B(h) needs gsm_exp(lambda*T): when that overflows every branch integral is NaN (STP:420)
                if (!(gsm_exp(K.lambda * T) < INFINITY)) tree = NAN;
                else tree = A.tree;
            } break;
            case GSM_TV_NO_PSI: {
                if (A.n_extant != A.n_leaf) { tree = gsm_neg_inf(); break; }                // STP:360-362
                tree = gsm_log((double)A.n_extant) + K.log_lmm - K.lmm * T;                     // STP:367
                tree -= gsm_log(K.rho * K.lambda + (K.lambda * (1 - K.rho) - K.mu) * gsm_exp(-K.lmm * T));   // STP:368
                tree += A.tree;
            } break;
            case GSM_TV_PSI: {
                if (A.n_extinct + A.n_extant + A.n_da != A.n_leaf) { tree = gsm_neg_inf(); break; }   // STP:313-316
                tree = gsm_log(4.0) + gsm_log((double)A.n_extant) + K.log_rho + K.log_lambda +
                       K.log_psi * (double)(A.n_da + A.n_extinct);                          // STP:325
                tree -= gsm_log(K.c1) + gsm_log(K.c2 + 1) + gsm_log(1 - K.c2 + (1 + K.c2) * gsm_exp(K.c1 * T));   // STP:326
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
        if (A.bad & (GSM_BAD_SA_STUBS | GSM_BAD_THROWS)) {                                  // STP:232-240, 258-265
            out[GSM_O_STP_LOGP] = gsm_neg_inf();
            out[GSM_O_TREE_LOGP] = tree;
            out[GSM_O_STUB_LOGP] = gsm_neg_inf();
            return;
        }
        stub = A.stub;
        // psi = 0, rho = 1: the Java's I(h) = 2(log(l e^{(l-m)h} - m) + l t) (STP:764-769) overflows to +inf at the
        // root's children when e^{(l-m)T} does, and inf - inf makes the whole stub sum NaN; reproduce, don't fix.
        if (K.stub_counts && K.stub_nosampling && n > 1 && !(gsm_exp(K.lmm * K.T) < INFINITY)) stub = NAN;
    }
    stp = tree + stub;                                                                      // STP:280
    out[GSM_O_STP_LOGP] = stp;
    out[GSM_O_TREE_LOGP] = tree;
    out[GSM_O_STUB_LOGP] = stub;
    (void)n;
}

