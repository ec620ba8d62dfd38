/*
 * gsm_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See gsm_oracle.h for the
 * scope and the parity-pinning statement.  Compile with -O2 -ffp-contract=off so that the
 * arithmetic is the plain IEEE-754 double sequence the JVM executes (no FMA contraction).
 *
 * Abbreviations for citations (paths relative to /root/reference):
 *   PRCM  src/gammaspike/clockmodel/PunctuatedRelaxedClockModel.java
 *   BSP   src/gammaspike/distribution/BranchSpikePrior.java
 *   BRP   src/gammaspike/distribution/BranchRatePrior.java
 *   STP   src/gammaspike/distribution/StumpedTreePrior.java
 *   Stubs src/gammaspike/tree/Stubs.java
 *   SPL   src/gammaspike/logger/SaltativeProportionLogger.java
 *   SpikeSize src/gammaspike/clockmodel/SpikeSize.java
 */
#include "gsm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEG_INF (-INFINITY)
#define POS_INF (INFINITY)

/* ------------------------------------------------------------------------------------------
 * Third-party arithmetic (BEAST.base >= 2.7.7 fork of Apache commons-math 2.x; not under
 * /root/reference; call sites BSP:98-99,138-139,166-167,349-350 and BRP:129,171).
 * Restated from the published algorithm.
 * ---------------------------------------------------------------------------------------- */

/* org.apache.commons.math.special.Gamma.logGamma: Lanczos approximation, g = 607/128. */
static const double LANCZOS[15] = {
    0.99999999999999709182,    57.156235665862923517,     -59.597960355475491248,
    14.136097974741747174,     -0.49191381609762019978,   .33994649984811888699e-4,
    .46523628927048575665e-4,  -.98374475304879564677e-4, .15808870322491248884e-3,
    -.21026444172410488319e-3, .21743961811521264320e-3,  -.16431810653676389022e-3,
    .84418223983852743293e-4,  -.26190838401581408670e-4, .36899182659531622704e-5};

double gsmo_log_gamma(double x) {
    if (isnan(x) || x <= 0.0) return NAN;
    const double g = 607.0 / 128.0;
    const double half_log_2pi = 0.5 * log(2.0 * M_PI);
    double sum = 0.0;
    for (int i = 14; i > 0; --i) sum = sum + (LANCZOS[i] / (x + i));
    sum = sum + LANCZOS[0];
    double tmp = x + g + .5;
    return ((x + .5) * log(tmp)) - tmp + half_log_2pi + log(sum / x);
}

/* GammaDistributionImpl(alpha, beta).logDensity(x): x<0 -> -inf; else
 * log(x/beta)*(alpha-1) - log(beta) - x/beta - logGamma(alpha).  (x==0: alpha>1 -> -inf,
 * alpha<1 -> +inf, alpha==1 -> NaN) */
double gsmo_gamma_log_density(double alpha, double beta, double x) {
    if (x < 0) return NEG_INF;
    return log(x / beta) * (alpha - 1) - log(beta) - x / beta - gsmo_log_gamma(alpha);
}

/* NormalDistributionImpl(mean, sd).logDensity(x) = -(x-m)^2/(2 sd^2) - log(sd*sqrt(2 pi)). */
double gsmo_normal_log_density(double mean, double sd, double x) {
    const double sqrt2pi = sqrt(2.0 * M_PI);
    double x0 = x - mean;
    return -(x0 * x0) / (2 * sd * sd) - log(sd * sqrt2pi);
}

/* BRP.LogNormalImpl.logDensity BRP:167-172. */
double gsmo_lognormal_log_density(double m, double s, double x) {
    if (x <= 0) return NEG_INF;
    return gsmo_normal_log_density(m, s, log(x)) - log(x);
}

/* ------------------------------------------------------------------------------------------
 * StumpedTreePrior helpers
 * ---------------------------------------------------------------------------------------- */
double gsmo_q(double t, double c1, double c2) { /* STP:489-491 */
    return exp(c1 * t) * (1 + c2) * (1 + c2) + exp(-c1 * t) * (1 - c2) * (1 - c2) + 2 * (1 - c2 * c2);
}
double gsmo_one_minus_p0hat(double t, double lambda, double mu, double rho) { /* STP:492-494 */
    return rho * (lambda - mu) / (lambda * rho + (lambda * (1 - rho) - mu) * exp((mu - lambda) * t));
}
double gsmo_c1(double lambda, double mu, double psi) { /* STP:496-498 */
    return sqrt(fabs(pow(lambda - mu - psi, 2) + 4 * lambda * psi));
}
double gsmo_c2(double lambda, double mu, double psi, double rho) { /* STP:500-502 */
    return -(lambda - mu - 2 * lambda * rho - psi) / gsmo_c1(lambda, mu, psi);
}
double gsmo_p0(double height, double lambda, double mu, double psi, double c1, double c2) { /* STP:506-507 */
    return (lambda + mu + psi -
            c1 * ((1 + c2) - exp(-c1 * height) * (1 - c2)) / ((1 + c2) + exp(-c1 * height) * (1 - c2))) /
           (2 * lambda);
}
double gsmo_p1(double height, double rho, double c1, double c2) { /* STP:523-524 */
    return (4 * rho) / (2 * (1 - c2 * c2) + exp(-c1 * height) * (1 - c2) * (1 - c2) +
                        exp(c1 * height) * (1 + c2) * (1 + c2));
}
double gsmo_log_qt_nosampling(double time, double lambda, double mu) { /* STP:826-831 */
    double e = exp((lambda - mu) * time);
    double top = mu * (e - 1);
    double btm = lambda * e - mu;
    return log(top) - log(btm);
}
double gsmo_log_qt_sampling(double time, double lambda, double mu, double psi, double rho) { /* STP:813-820 */
    double c1 = gsmo_c1(lambda, mu, psi);
    double c2 = gsmo_c2(lambda, mu, psi, rho);
    double e = exp(-c1 * time);
    double top = e * (1 - c2) - (1 + c2);
    double btm = e * (1 - c2) + (1 + c2);
    return log(lambda + mu + psi + c1 * top / btm) - log(2 * lambda);
}
/* getExpectedStubCountForBranch(h1,h0,lambda,mu,psi,rho) STP:749-754 */
double gsmo_expected_stubs_sampling(double h1, double h0, double lambda, double mu, double psi, double rho) {
    double c1 = gsmo_c1(lambda, mu, psi);
    double c2 = gsmo_c2(lambda, mu, psi, rho);
    double f = ((c2 - 1) * exp(-c1 * h1) - c2 - 1) / ((c2 - 1) * exp(-c1 * h0) - c2 - 1);
    return (h0 - h1) * (lambda + mu + psi - c1) + 2 * log(f);
}
/* getStubRateIntegral STP:764-769 */
static double stub_rate_integral(double height, double lambda, double mu, double root_height) {
    double t = root_height - height;
    double b = log(lambda * exp((lambda - mu) * height) - mu);
    return 2 * (b + lambda * t);
}
/* getExpectedStubCountForBranch(h1,h0,lambda,mu) STP:758-762 */
double gsmo_expected_stubs_nosampling(double h1, double h0, double lambda, double mu, double root_height) {
    return stub_rate_integral(h1, lambda, mu, root_height) - stub_rate_integral(h0, lambda, mu, root_height);
}
/* StumpedTreePrior.getMeanStubNumber STP:876-892 */
double gsmo_mean_stub_number(double node_h, double parent_h, double lambda, double mu, double psi, double rho,
                             double root_height) {
    if (parent_h <= node_h) return 0;
    if (psi == 0 && rho == 1) return gsmo_expected_stubs_nosampling(node_h, parent_h, lambda, mu, root_height);
    return gsmo_expected_stubs_sampling(node_h, parent_h, lambda, mu, psi, rho);
}

/* getLambda/getMu/getPsi/getRho STP:672-737.  NaN = input not wired. */
void gsmo_param_maps(double lambda_in, double net_div_in, double r0_in, double turnover_in, double psi_in,
                     double sampling_prop_in, double rho_in, double out[4]) {
    double lambda, mu, psi, rho;
    if (isnan(lambda_in)) lambda = net_div_in / (1 - turnover_in); /* STP:673-676 */
    else lambda = lambda_in;                                        /* STP:688 */
    if (isnan(r0_in)) mu = turnover_in * lambda;                    /* STP:693-696 */
    else mu = lambda / r0_in;                                       /* STP:697-700 */
    if (isnan(psi_in) && isnan(sampling_prop_in)) psi = 0.0;        /* STP:713-715 */
    else if (isnan(psi_in)) psi = sampling_prop_in * mu / (1 - sampling_prop_in); /* STP:717-720 */
    else psi = psi_in;                                              /* STP:722 */
    rho = isnan(rho_in) ? 1.0 : rho_in;                             /* STP:733-736 */
    out[0] = lambda; out[1] = mu; out[2] = psi; out[3] = rho;
}

/* ------------------------------------------------------------------------------------------
 * beast.base.evolution.tree.Node predicates on an array tree (BEAST.base 2.7.7)
 * ---------------------------------------------------------------------------------------- */
typedef struct node_info {
    int32_t n, root, leaf_count, da_count;
    int32_t* nchild;      /* isLeaf <=> nchild==0 */
    uint8_t* da;          /* isDirectAncestor: leaf, non-root, parent height == own height */
    uint8_t* fake;        /* isFake: non-leaf with a direct-ancestor child */
} node_info;

static void node_info_build(node_info* ni, int32_t n, const double* h, const int32_t* parent) {
    ni->n = n; ni->root = -1; ni->leaf_count = 0; ni->da_count = 0;
    ni->nchild = (int32_t*)calloc((size_t)n > 0 ? (size_t)n : 1, sizeof(int32_t));
    ni->da = (uint8_t*)calloc((size_t)n > 0 ? (size_t)n : 1, 1);
    ni->fake = (uint8_t*)calloc((size_t)n > 0 ? (size_t)n : 1, 1);
    for (int32_t i = 0; i < n; i++) {
        if (parent[i] < 0) { if (ni->root < 0) ni->root = i; }
        else ni->nchild[parent[i]]++;
    }
    for (int32_t i = 0; i < n; i++) {
        if (ni->nchild[i] == 0) {
            ni->leaf_count++;
            if (parent[i] >= 0 && h[parent[i]] == h[i]) { ni->da[i] = 1; ni->da_count++; }
        }
    }
    for (int32_t i = 0; i < n; i++)
        if (ni->da[i]) ni->fake[parent[i]] = 1;
}
static void node_info_free(node_info* ni) { free(ni->nchild); free(ni->da); free(ni->fake); }

static inline int is_root(const gsmo_tree* t, int32_t i) { return t->parent[i] < 0; }
static inline double node_length(const gsmo_tree* t, int32_t i) { /* Node.getLength */
    return is_root(t, i) ? 0.0 : t->height[t->parent[i]] - t->height[i];
}
static inline int parent_is_fake(const gsmo_tree* t, const node_info* ni, int32_t i) {
    return t->parent[i] >= 0 && ni->fake[t->parent[i]];
}

/* Stubs.getNStubsOnBranch Stubs:344-370 (nostub -> -1; branchCounts -> nr>=dim ? 0 : value[nr],
 * dim = nodeCount-1 per Stubs:200; exact -> number of listed stubs whose branch == nr, Stubs:357-364:
 * the O(S) scan of the Java, index 0 excluded by includeStub Stubs:285-287). */
static int32_t nstubs_on_branch(const gsmo_config* cfg, const gsmo_tree* t, int32_t nr) {
    switch (cfg->stub_mode) {
        case GSMO_STUBS_NOSTUB: return -1;
        case GSMO_STUBS_COUNTS:
            if (nr >= t->n - 1) return 0;
            return t->nstubs[nr];
        case GSMO_STUBS_EXACT: {
            int32_t nstubs = 0;
            for (int32_t i = 0; i < t->n_stub_entries; i++)
                if (t->stub_branch[i] == nr) nstubs++;
            return nstubs;
        }
    }
    return -1;
}
static inline int estimate_stubs(const gsmo_config* cfg) { return cfg->stub_mode != GSMO_STUBS_NOSTUB; } /* Stubs:648-650 */

/* ------------------------------------------------------------------------------------------
 * Clock model  (PRCM:167-242)
 * ---------------------------------------------------------------------------------------- */
static double burst_size(const gsmo_config* cfg, const gsmo_tree* t, const node_info* ni, int32_t i) { /* PRCM:178-208 */
    if ((cfg->flags & GSMO_F_HAS_INDICATOR) && !t->use_spike) return 0;                  /* PRCM:181-183 */
    if (cfg->flags & GSMO_F_NO_SPIKE_DATED_TIPS) {                                       /* PRCM:185-187 */
        if (ni->nchild[i] == 0 && t->height[i] > 0) return 0;
    }
    int stubs_wired = (cfg->flags & GSMO_F_STUBS_TO_CLOCK) != 0;
    if ((!stubs_wired && t->parent[i] >= 0 && parent_is_fake(t, ni, i)) ||               /* PRCM:192 */
        (stubs_wired && nstubs_on_branch(cfg, t, i) == 0)) {                             /* PRCM:195 */
        if (t->parent[i] >= 0 && parent_is_fake(t, ni, i)) return 0;                     /* PRCM:198-200 */
    }
    double spike_mean = t->params[GSMO_P_SPIKE_MEAN];
    double relative_spike = t->spike[i];
    return relative_spike * spike_mean;                                                  /* PRCM:204-206 */
}
static double branch_rate(const gsmo_config* cfg, const gsmo_tree* t, int32_t i) { /* PRCM:210-222 */
    if (!(cfg->flags & GSMO_F_HAS_RATES)) return 1;
    if (!(cfg->flags & GSMO_F_RELAXED_FALSE)) return t->rate[i];
    return 1;
}
static double rate_for_branch(const gsmo_config* cfg, const gsmo_tree* t, const node_info* ni, int32_t i) { /* PRCM:225-242 */
    double base = t->params[GSMO_P_CLOCK_RATE];
    double len = node_length(t, i);
    if (len <= 0 || ni->da[i] || is_root(t, i)) return base;                             /* PRCM:229 */
    double burst = burst_size(cfg, t, ni, i);
    double r = branch_rate(cfg, t, i);
    double dist = len * base * r + burst;                                                /* PRCM:233 */
    return dist / len;                                                                   /* PRCM:236 */
}

/* ------------------------------------------------------------------------------------------
 * BranchSpikePrior  (BSP:53-213)
 * ---------------------------------------------------------------------------------------- */
static int32_t n_spikes(const gsmo_tree* t, const node_info* ni, int32_t i, int32_t nstubs) { /* BSP:197-213 */
    if (is_root(t, i)) return nstubs + 1;
    if (ni->da[i]) return 0;
    if (ni->fake[t->parent[i]]) return nstubs;
    return nstubs + 1;
}

#define MAX_CUM_SUM 0.999  /* BSP:39 */
#define ORACLE_K_CAP 10000000 /* safety cap on the truncation loops (the Java has none) */

static double spike_prior_logp(const gsmo_config* cfg, const gsmo_tree* t, const node_info* ni) { /* BSP:53-187 */
    double logP = 0;
    double shape = t->params[GSMO_P_SPIKE_SHAPE];
    double mean = 1;                                                                     /* BSP:59 */
    if (shape <= 0 || mean <= 0) return NEG_INF;                                         /* BSP:60-63 */
    double scale = mean / shape;                                                         /* BSP:64 */
    int stubs_null = !(cfg->flags & GSMO_F_STUBS_TO_SPIKE_PRIOR);
    double root_h = ni->root >= 0 ? t->height[ni->root] : 0.0;
    for (int32_t nr = 0; nr < t->n; nr++) {                                              /* BSP:69 */
        double x = t->spike[nr];
        if (x < 0) return NEG_INF;                                                       /* BSP:72-75 */
        if (stubs_null || estimate_stubs(cfg)) {                                         /* BSP:77 */
            int32_t nst = stubs_null ? 0 : nstubs_on_branch(cfg, t, nr);                 /* BSP:81 */
            int32_t m = n_spikes(t, ni, nr, nst);                                        /* BSP:85 */
            if (m == 0) {
                logP += (x != 0) ? NEG_INF : 0;                                          /* BSP:86-95 */
            } else {
                double alpha = shape * m;                                                /* BSP:97 */
                logP += gsmo_gamma_log_density(alpha, scale, x);                         /* BSP:98-99 */
            }
        } else {                                                                         /* BSP:102-172 */
            double h0 = t->height[nr];
            double h1 = is_root(t, nr) ? h0 : t->height[t->parent[nr]];                  /* BSP:106-107 */
            double mu = gsmo_mean_stub_number(h0, h1, t->params[GSMO_P_LAMBDA], t->params[GSMO_P_MU],
                                              t->params[GSMO_P_PSI], t->params[GSMO_P_RHO], root_h); /* BSP:108 */
            if (mu > 0) {
                double branchP = 0;
                int k = 0;
                double cumsum = 0;
                while (cumsum < MAX_CUM_SUM && k < ORACLE_K_CAP) {                       /* BSP:117 */
                    double p = -mu + k * log(mu);                                        /* BSP:120 */
                    for (int i = 2; i <= k; i++) p += -log((double)i);                   /* BSP:121 */
                    double pReal = exp(p);
                    cumsum += pReal;
                    int32_t m = n_spikes(t, ni, nr, k);                                  /* BSP:127 */
                    if (m == 0) {
                        if (x != 0) branchP += 0; else branchP += exp(p);                /* BSP:128-135 */
                    } else {
                        double alpha = shape * m;
                        double g = gsmo_gamma_log_density(alpha, scale, x);
                        if (x == 0 || g == NEG_INF || isnan(g)) branchP += 0;            /* BSP:140-141 */
                        else branchP += exp(p + g);                                      /* BSP:143 */
                    }
                    k++;
                }
                logP += log(branchP);                                                    /* BSP:151 */
            } else {
                int32_t m = n_spikes(t, ni, nr, 0);                                      /* BSP:157 */
                if (m == 0) {
                    if (x != 0) logP += NEG_INF; else logP += 0;                         /* BSP:158-164 */
                } else {
                    logP += gsmo_gamma_log_density(shape, scale, x);                     /* BSP:166-167 */
                }
            }
        }
    }
    if (logP == POS_INF) logP = NEG_INF;                                                 /* BSP:177-179 */
    return logP;
}

/* BSP.getCumulativeProbs BSP:315-389.  `use_indicator_branch` = (indicator wired && indicator true), BSP:347. */
int32_t gsmo_cumulative_probs(double mu, double spike, double shape, int use_indicator_branch,
                              double* out, int32_t cap) {
    double scale = 1.0 / shape;                                                          /* BSP:321-325 */
    int32_t k = 0, len = 0;
    double poissonCumSum = 0, branchPSum = 0;
    while (poissonCumSum < MAX_CUM_SUM && k < ORACLE_K_CAP) {                            /* BSP:334 */
        double branchP = 0;
        double p = -mu + k * log(mu);                                                    /* BSP:339 */
        for (int i = 2; i <= k; i++) p += -log((double)i);                               /* BSP:340 */
        double pReal = exp(p);
        double alpha = shape * (k + 1);                                                  /* BSP:344 */
        if (use_indicator_branch) {
            double g = gsmo_gamma_log_density(alpha, scale, spike);                      /* BSP:349-350 */
            if (g == NEG_INF || isnan(g)) {
                if (poissonCumSum > 0) break;                                            /* BSP:352 */
                branchP += 0;
            } else {
                branchP += exp(p + g);                                                   /* BSP:355 */
            }
        } else {
            branchP += pReal;                                                            /* BSP:361 */
        }
        poissonCumSum += pReal;
        branchPSum += branchP;
        if (len < cap) out[len] = branchP;
        len++;
        k++;
    }
    int32_t w = len < cap ? len : cap;
    for (int32_t i = 0; i < w; i++) out[i] = out[i] / branchPSum;                        /* BSP:374-377 */
    double cumsum = 0;
    for (int32_t i = 0; i < w; i++) {                                                    /* BSP:381-387 */
        double p = out[i];
        out[i] = p + cumsum;
        cumsum += p;
    }
    return len;
}

/* Randomizer.randomChoice(double[] cf) of BEAST.base 2.7.7 (not under /root/reference; restated from memory of the BEAST 2
 * source): s = 0 if U <= cf[0]; else the first s >= 1 with cf[s-1] < U <= cf[s]; cf.length if none. */
int32_t gsmo_random_choice(const double* cf, int32_t len, double U) {
    int32_t s;
    if (len <= 0) return 0;
    if (U <= cf[0]) {
        s = 0;
    } else {
        for (s = 1; s < len; s++)
            if (U <= cf[s] && U > cf[s - 1]) break;
    }
    return s;
}

/* ------------------------------------------------------------------------------------------
 * BranchRatePrior (BRP:29-65) — also equals the template's core Prior + LogNormal(M=1,S=GSMclockSD,
 * meanInRealSpace) of fxtemplates/GammaSpikeClockModel.xml:60-62.
 * ---------------------------------------------------------------------------------------- */
static double rate_prior_logp(const gsmo_tree* t) {
    double logP = 0;
    double s = t->params[GSMO_P_CLOCK_SD];
    const double mean = 1;
    if (s <= 0) return NEG_INF;                                                          /* BRP:36-39 */
    double m = log(mean) - (0.5 * s * s);                                                /* BRP:40 */
    for (int32_t nr = 0; nr < t->n; nr++) {
        double rate = t->rate[nr];
        if (rate < 0) return NEG_INF;                                                    /* BRP:47-50 */
        logP += gsmo_lognormal_log_density(m, s, rate);                                  /* BRP:58 */
    }
    return logP;
}

/* ------------------------------------------------------------------------------------------
 * StumpedTreePrior (STP:150-665)
 * ---------------------------------------------------------------------------------------- */
#define HEIGHT_THRESHOLD_OF_LEAF 1e-10 /* STP:57 */

static double tree_logp_with_psi(const gsmo_tree* t, const node_info* ni, double lambda, double mu, double psi,
                                 double rho) { /* STP:290-347 */
    double m = 0, n = 0, k = 0;
    for (int32_t i = 0; i < t->n; i++) {
        if (ni->nchild[i] != 0) continue;
        if (ni->da[i]) k++;                                                              /* STP:297-299 */
        else if (t->height[i] <= HEIGHT_THRESHOLD_OF_LEAF) n++;                          /* STP:301-303 */
        else if (node_length(t, i) > HEIGHT_THRESHOLD_OF_LEAF) m++;                      /* STP:305-307 */
    }
    if (m + n + k != ni->leaf_count) return NEG_INF;                                     /* STP:313-316 */
    double treeLogP = 0;
    double c1 = gsmo_c1(lambda, mu, psi);
    double c2 = gsmo_c2(lambda, mu, psi, rho);
    double x1 = t->height[ni->root];
    treeLogP += log(4) + log(n) + log(rho) + log(lambda) + log(psi) * (k + m);           /* STP:325 */
    treeLogP -= log(c1) + log(c2 + 1) + log(1 - c2 + (1 + c2) * exp(c1 * x1));           /* STP:326 */
    for (int32_t i = 0; i < t->n; i++) {                                                 /* STP:329-334 */
        if (ni->nchild[i] == 0) continue;
        if (ni->fake[i]) continue;
        treeLogP += log(lambda) + log(gsmo_p1(t->height[i], rho, c1, c2));
    }
    for (int32_t i = 0; i < t->n; i++) {                                                 /* STP:337-343 */
        if (ni->nchild[i] != 0) continue;
        if (t->height[i] <= HEIGHT_THRESHOLD_OF_LEAF) continue;
        if (ni->da[i]) continue;
        double lh = t->height[i];
        treeLogP += log(gsmo_p0(lh, lambda, mu, psi, c1, c2)) - log(gsmo_p1(lh, rho, c1, c2));
    }
    return treeLogP;
}

static double tree_logp_without_psi(const gsmo_tree* t, const node_info* ni, double lambda, double mu, double rho) { /* STP:351-379 */
    double n = 0;
    for (int32_t i = 0; i < t->n; i++) {
        if (ni->nchild[i] != 0) continue;
        if (!ni->da[i] && t->height[i] <= HEIGHT_THRESHOLD_OF_LEAF) n++;
    }
    if (n != ni->leaf_count) return NEG_INF;                                             /* STP:360-362 */
    double treeLogP = 0;
    double x1 = t->height[ni->root];
    treeLogP += log(n) + log(lambda - mu) - (lambda - mu) * x1;                          /* STP:367 */
    treeLogP -= log(rho * lambda + (lambda * (1 - rho) - mu) * exp(-(lambda - mu) * x1)); /* STP:368 */
    for (int32_t i = 0; i < t->n; i++) {                                                 /* STP:371-377 */
        if (ni->nchild[i] == 0) continue;
        if (ni->fake[i]) continue;
        double nh = t->height[i];
        treeLogP += log(lambda) + log(rho) + 2 * log(lambda - mu) - (lambda - mu) * nh;
        treeLogP -= 2 * log(rho * lambda + (lambda * (1 - rho) - mu) * exp(-(lambda - mu) * nh));
    }
    return treeLogP;
}

static double blue_rate_integral(double height, double lambda, double mu, double root_height) { /* STP:417-422 */
    double t = root_height - height;
    double tlm = t * (lambda - mu);
    double b = log(lambda * exp(lambda * root_height) - mu * exp(tlm + mu * root_height));
    return tlm - b;
}
static double tree_logp_complete_rho(const gsmo_tree* t, const node_info* ni, double lambda, double mu) { /* STP:383-426 */
    double tree_h = t->height[ni->root];
    double blue = 0;
    for (int32_t i = 0; i < t->n; i++) {                                                 /* STP:403-412 */
        double h0 = is_root(t, i) ? t->height[i] : t->height[t->parent[i]];
        double h1 = t->height[i];
        double integral_branch = blue_rate_integral(h1, lambda, mu, tree_h) - blue_rate_integral(h0, lambda, mu, tree_h);
        blue += integral_branch;
    }
    double treeLogP = -blue;                                                             /* STP:385 */
    for (int32_t i = 0; i < t->n; i++) {                                                 /* STP:387-393 */
        if (ni->nchild[i] == 0) continue;
        double logQt = gsmo_log_qt_nosampling(t->height[i], lambda, mu);
        double logRate = log(lambda) + log(1 - exp(logQt));                              /* STP:423-426 */
        treeLogP += logRate;
    }
    return treeLogP;
}

static double tree_logp_conditioned(const gsmo_config* cfg, const gsmo_tree* t, const node_info* ni, double lambda,
                                    double mu, double psi, double rho, double x0, double x1) { /* STP:428-487 */
    double treeLogP;
    double c1 = gsmo_c1(lambda, mu, psi);
    double c2 = gsmo_c2(lambda, mu, psi, rho);
    int cond_root = (cfg->flags & GSMO_F_COND_ROOT) != 0;
    if (!cond_root) {
        treeLogP = -log(gsmo_q(x0, c1, c2));                                             /* STP:435 */
    } else {
        if (ni->fake[ni->root]) return NEG_INF;                                          /* STP:437-440 */
        treeLogP = -log(gsmo_q(x1, c1, c2));                                             /* STP:442 */
    }
    if (cfg->flags & GSMO_F_COND_SAMPLING) {                                             /* STP:446-453 */
        if (cond_root)
            treeLogP -= log(lambda * (1 - gsmo_p0(x1, lambda, mu, psi, c1, c2)) * (1 - gsmo_p0(x1, lambda, mu, psi, c1, c2)));
        else
            treeLogP -= log(1 - gsmo_p0(x0, lambda, mu, psi, c1, c2));
    }
    if (cfg->flags & GSMO_F_COND_RHO_SAMPLING) {                                         /* STP:455-461 */
        if (cond_root)
            treeLogP -= log(lambda * gsmo_one_minus_p0hat(x1, lambda, mu, rho) * gsmo_one_minus_p0hat(x1, lambda, mu, rho));
        else
            treeLogP -= log(gsmo_one_minus_p0hat(x0, lambda, mu, rho));
    }
    int internalNodeCount = ni->leaf_count - ni->da_count - 1;                           /* STP:463 */
    treeLogP += internalNodeCount * log(2);                                              /* STP:465 */
    for (int32_t i = 0; i < t->n; i++) {                                                 /* STP:467-483 */
        double h = t->height[i];
        if (ni->nchild[i] == 0) {
            if (!ni->da[i]) {
                if (h > 0.000000000005 || rho == 0.)
                    treeLogP += log(psi) + log(gsmo_q(h, c1, c2)) + log(gsmo_p0(h, lambda, mu, psi, c1, c2));
                else
                    treeLogP += log(4 * rho);
            }
        } else {
            if (ni->fake[i]) treeLogP += log(psi);
            else treeLogP += log(lambda) - log(gsmo_q(h, c1, c2));
        }
    }
    return treeLogP;
}

/* getStubLogBirthRate STP:796-806 */
double gsmo_stub_log_birth_rate_sampling(double t, double lambda, double mu, double psi, double rho) { /* STP:796-800 */
    double logQt = gsmo_log_qt_sampling(t, lambda, mu, psi, rho);
    return log(2) + log(lambda) + logQt;
}
double gsmo_stub_log_birth_rate_nosampling(double t, double lambda, double mu) { /* STP:802-806 */
    double logQt = gsmo_log_qt_nosampling(t, lambda, mu);
    return log(2) + log(lambda) + logQt;
}

/* getStubLogPForBranchWithSampling STP:539-601 / getStubLogPForBranchWithoutSampling STP:605-665.
 * *threw is set where the Java throws (Stubs.getAbsoluteTimeOfStub on the root, Stubs:631-634; the exception is caught
 * at STP:258-265 and calculateTreeLogLikelihood returns -inf). */
static double stub_logp_branch(const gsmo_config* cfg, const gsmo_tree* t, const node_info* ni, int32_t nr,
                               double lambda, double mu, double psi, double rho, int with_sampling, int* threw) {
    if (!estimate_stubs(cfg)) return 0;                                                  /* STP:547 / 613 */
    double tree_h = t->height[ni->root];
    double h0 = is_root(t, nr) ? tree_h : t->height[t->parent[nr]];                      /* STP:550 / 616 */
    double h1 = t->height[nr];
    if (cfg->stub_mode == GSMO_STUBS_EXACT) {                                            /* STP:573-600 / 639-663 */
        double stubLogP = 0;
        int stubCount = 0;
        for (int32_t s = 0; s < t->n_stub_entries; s++) {                                /* includeStub: index 0 never listed */
            if (t->stub_branch[s] != nr) continue;
            /* Stubs.getAbsoluteTimeOfStub Stubs:625-640 */
            if (is_root(t, nr)) { *threw = 1; return NEG_INF; }                          /* Stubs:631-634 throws */
            double parent_h = t->height[t->parent[nr]];
            double stubHeight = h1 + t->stub_time[s] * (parent_h - h1);                  /* Stubs:636 */
            if (stubHeight < h1 || stubHeight > parent_h) return NEG_INF;                /* STP:580-582 / 647-649 */
            if (with_sampling && ni->da[nr]) return NEG_INF;                             /* STP:583-585 (sampling variant only) */
            double logRate = with_sampling ? gsmo_stub_log_birth_rate_sampling(stubHeight, lambda, mu, psi, rho)
                                           : gsmo_stub_log_birth_rate_nosampling(stubHeight, lambda, mu);
            stubLogP += logRate;                                                         /* STP:587-589 / 651-653 */
            stubCount++;
        }
        double E = with_sampling ? gsmo_expected_stubs_sampling(h1, h0, lambda, mu, psi, rho)   /* STP:594 */
                                 : gsmo_expected_stubs_nosampling(h1, h0, lambda, mu, tree_h);  /* STP:658 */
        stubLogP += -E;                                                                  /* STP:596 / 659 */
        for (int i = 2; i <= stubCount; i++) stubLogP += -log((double)i);                /* STP:597 / 660 */
        return stubLogP;
    }
    if (is_root(t, nr)) return 0;                                                        /* STP:556 / 622 */
    int32_t cnt = nstubs_on_branch(cfg, t, nr);
    if (h0 <= h1) {                                                                      /* STP:558-561 */
        if (cnt == 0) return 0;
        return NEG_INF;
    }
    double E = with_sampling ? gsmo_expected_stubs_sampling(h1, h0, lambda, mu, psi, rho)
                             : gsmo_expected_stubs_nosampling(h1, h0, lambda, mu, tree_h);
    double lp = cnt * log(E) - E;                                                        /* STP:566 / 632 */
    for (int i = 2; i <= cnt; i++) lp += -log((double)i);                                /* STP:567 / 633 */
    return lp;
}

/* calculateTreeLogLikelihood STP:150-284.  The cross-call nExtantTaxa check (STP:196-200) is host
 * state and is NOT applied here; n is reported through *n_extant. */
static void stumped_tree_prior(const gsmo_config* cfg, const gsmo_tree* t, const node_info* ni, double* stp,
                               double* tree_lp, double* stub_lp, double* n_extant) {
    double lambda = t->params[GSMO_P_LAMBDA], mu = t->params[GSMO_P_MU];
    double psi = t->params[GSMO_P_PSI], rho = t->params[GSMO_P_RHO];
    double treeLogP = 0, stubLogP = 0;
    int n = 0;
    for (int32_t i = 0; i < t->n; i++)                                                   /* STP:190-195 */
        if (ni->nchild[i] == 0 && !ni->da[i] && t->height[i] <= HEIGHT_THRESHOLD_OF_LEAF) n++;
    *n_extant = (double)n;
    *tree_lp = 0; *stub_lp = 0;
    if (lambda <= mu) { *stp = NEG_INF; *tree_lp = NEG_INF; return; }                    /* STP:160-164 */
    if (!(cfg->flags & GSMO_F_IGNORE_TREE_PRIOR)) {
        if (exp(psi) == POS_INF || exp(-psi) == 0) { *stp = NEG_INF; *tree_lp = NEG_INF; return; } /* STP:176-178 */
        if ((cfg->flags & GSMO_F_ORIGIN) || (cfg->flags & GSMO_F_COND_ROOT)) {           /* STP:179-186 */
            double x0 = (cfg->flags & GSMO_F_ORIGIN) ? t->params[GSMO_P_ORIGIN] : POS_INF;
            double x1 = t->height[ni->root];
            if (x0 < x1) { *stp = NEG_INF; *tree_lp = NEG_INF; return; }
            treeLogP = tree_logp_conditioned(cfg, t, ni, lambda, mu, psi, rho, x0, x1);
        } else {
            if ((psi == 0) && (rho == 1)) treeLogP = tree_logp_complete_rho(t, ni, lambda, mu);      /* STP:203-205 */
            else if (psi == 0) treeLogP = tree_logp_without_psi(t, ni, lambda, mu, rho);              /* STP:207-209 */
            else treeLogP = tree_logp_with_psi(t, ni, lambda, mu, psi, rho);                          /* STP:211-213 */
        }
    }
    *tree_lp = treeLogP;
    if (!(cfg->flags & GSMO_F_IGNORE_STUB_PRIOR) && (cfg->flags & GSMO_F_STUBS_TO_TREE_PRIOR)) { /* STP:219 */
        for (int32_t i = 0; i < t->n; i++) {                                             /* STP:232-240 */
            if (ni->da[i]) {
                if (nstubs_on_branch(cfg, t, i) > 0) { *stp = NEG_INF; *stub_lp = NEG_INF; return; }
            }
        }
        int with_sampling = !((psi == 0) && (rho == 1));                                 /* STP:243-253 */
        int threw = 0;
        for (int32_t nr = 0; nr < t->n && !threw; nr++)
            stubLogP += stub_logp_branch(cfg, t, ni, nr, lambda, mu, psi, rho, with_sampling, &threw);
        if (threw) { *stp = NEG_INF; *stub_lp = NEG_INF; return; }                       /* STP:258-265 */
    }
    *stub_lp = stubLogP;
    *stp = treeLogP + stubLogP;                                                          /* STP:280 */
}

/* ------------------------------------------------------------------------------------------
 * One tree: all outputs
 * ---------------------------------------------------------------------------------------- */
void gsmo_eval_tree(const gsmo_config* cfg, const gsmo_tree* t, double* out, double* out_rates, double* out_wspikes) {
    node_info ni;
    node_info_build(&ni, t->n, t->height, t->parent);
    if (t->n <= 0 || ni.root < 0) {
        for (int i = 0; i < GSMO_NOUT; i++) out[i] = NAN;
        node_info_free(&ni);
        return;
    }
    stumped_tree_prior(cfg, t, &ni, &out[GSMO_O_STP_LOGP], &out[GSMO_O_TREE_LOGP], &out[GSMO_O_STUB_LOGP],
                       &out[GSMO_O_N_EXTANT]);
    out[GSMO_O_SPIKE_LOGP] = spike_prior_logp(cfg, t, &ni);
    out[GSMO_O_RATE_LOGP] = (cfg->flags & GSMO_F_HAS_RATES) ? rate_prior_logp(t) : 0.0;

    /* SaltativeProportionLogger.log SPL:33-61 */
    double total = 0, abrupt = 0;
    for (int32_t i = 0; i < t->n; i++) {
        double len = node_length(t, i);
        if (len <= 0 || ni.da[i] || is_root(t, i)) continue;                             /* SPL:41 */
        double branch_time = len;                                                        /* SPL:44 */
        double br = rate_for_branch(cfg, t, &ni, i);                                     /* SPL:45 */
        total += branch_time * br;                                                       /* SPL:46-49 */
        abrupt += burst_size(cfg, t, &ni, i);                                            /* SPL:53-54 */
    }
    out[GSMO_O_SUM_BURST] = abrupt;
    out[GSMO_O_SUM_DIST] = total;

    if (out_rates)                                                                       /* PRCM:267-276 */
        for (int32_t i = 0; i < t->n; i++) out_rates[i] = rate_for_branch(cfg, t, &ni, i);
    if (out_wspikes)                                                                     /* SpikeSize:48-56, PRCM:167-170 */
        for (int32_t i = 0; i < t->n; i++) out_wspikes[i] = burst_size(cfg, t, &ni, i);
    node_info_free(&ni);
}

int gsmo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int gsmo_eval_batch(const gsmo_config* cfg, int64_t n_trees, int32_t n_nodes, int64_t stride,
                    const double* height, const int32_t* parent, const double* rate, const double* spike,
                    const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                    const int32_t* node_count, double* out_tree, double* out_rates, double* out_wspikes,
                    int n_threads) {
    return gsmo_eval_batch_stubs(cfg, n_trees, n_nodes, stride, height, parent, rate, spike, nstubs, params, use_spike,
                                 node_count, NULL, NULL, NULL, out_tree, out_rates, out_wspikes, n_threads);
}

int gsmo_eval_batch_stubs(const gsmo_config* cfg, int64_t n_trees, int32_t n_nodes, int64_t stride,
                          const double* height, const int32_t* parent, const double* rate, const double* spike,
                          const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                          const int32_t* node_count, const int64_t* stub_ptr, const int32_t* stub_branch,
                          const double* stub_time, double* out_tree, double* out_rates, double* out_wspikes,
                          int n_threads) {
    if (n_trees < 0 || n_nodes < 0 || stride < n_nodes) return 1;
    if (cfg->stub_mode == GSMO_STUBS_EXACT && !stub_ptr) return 2;   /* exact mode is defined by its stub list */
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
#else
    (void)n_threads;
#endif
    for (int64_t ti = 0; ti < n_trees; ti++) {
        gsmo_tree t;
        t.n = node_count ? node_count[ti] : n_nodes;
        t.height = height + ti * stride;
        t.parent = parent + ti * stride;
        t.rate = rate ? rate + ti * stride : NULL;
        t.spike = spike + ti * stride;
        t.nstubs = nstubs ? nstubs + ti * stride : NULL;
        t.params = params + ti * GSMO_NPARAM;
        t.use_spike = use_spike ? use_spike[ti] : 1;
        t.n_stub_entries = stub_ptr ? (int32_t)(stub_ptr[ti + 1] - stub_ptr[ti]) : 0;
        t.stub_branch = stub_ptr ? stub_branch + stub_ptr[ti] : NULL;
        t.stub_time = stub_ptr ? stub_time + stub_ptr[ti] : NULL;
        gsmo_eval_tree(cfg, &t, out_tree + ti * GSMO_NOUT, out_rates ? out_rates + ti * stride : NULL,
                       out_wspikes ? out_wspikes + ti * stride : NULL);
    }
    return 0;
}
