/*
 * gsm_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C, fp64 restatement of the per-branch compute of jordandouglas/GammaSpikeModel
 * (a pure-Java BEAST 2 package).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (gammaspikemodel_b200/, include/gsm_b200.h) never calls into it.
 *
 * PARITY PINNING:
 *   - pinned by the reference's own tests: ONLY the origin/root-conditioned tree density
 *     (test/gammaspike/distribution/StumpedTreePriorTest.java:35,53,71,96,114 — five goldens,
 *     tol 1e-4) and the helpers under it (q, getP0, oneMinusP0Hat, getC1/getC2, turnover /
 *     samplingProportion parameter maps).
 *   - everything else on the path (clock-rate map, spike prior, rate prior, the three
 *     unconditioned tree densities, stub terms, ProportionOfSaltation, stub CDF) is
 *     "PARITY UNPINNED" by the reference: it has no test or fixture for them, and the reference
 *     (Java, needs BEAST 2.7.7 + a JVM) cannot be built or run in this environment, so there is
 *     no oracle/_ref.  Those functions are pinned instead by independent cross-checks in tests/
 *     (scipy.stats gamma/lognorm/poisson, mpmath 50-digit evaluation, algebraic identities).
 *   - arithmetic that lives in third-party code absent from /root/reference — BEAST.base >= 2.7.7's
 *     fork of Apache commons-math 2.x (GammaDistributionImpl.logDensity, NormalDistributionImpl.logDensity,
 *     special.Gamma.logGamma) and beast.base.evolution.tree.Node predicates — is restated from its
 *     published algorithm and anchored on the reference's call sites.
 */
#ifndef GSM_ORACLE_H
#define GSM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Wiring flags: which optional BEAST Inputs are connected.  Values mirror include/gsm_b200.h
 * (tests assert the two headers agree). */
#define GSMO_F_HAS_INDICATOR        (1u << 0)   /* PRCM.indicator wired (PunctuatedRelaxedClockModel.java:34) */
#define GSMO_F_NO_SPIKE_DATED_TIPS  (1u << 1)   /* PRCM.noSpikeOnDatedTips (PRCM:42) */
#define GSMO_F_HAS_RATES            (1u << 2)   /* PRCM.rates wired (PRCM:36) */
#define GSMO_F_RELAXED_FALSE        (1u << 3)   /* PRCM.relaxed wired AND false (PRCM:216) */
#define GSMO_F_STUBS_TO_CLOCK       (1u << 4)   /* PRCM.stubs wired (PRCM:30) */
#define GSMO_F_STUBS_TO_SPIKE_PRIOR (1u << 5)   /* BranchSpikePrior.stubs wired (BSP:28) */
#define GSMO_F_STUBS_TO_TREE_PRIOR  (1u << 6)   /* StumpedTreePrior.stubs wired (STP:46) */
#define GSMO_F_IGNORE_TREE_PRIOR    (1u << 7)   /* STP:49 */
#define GSMO_F_IGNORE_STUB_PRIOR    (1u << 8)   /* STP:50 */
#define GSMO_F_COND_SAMPLING        (1u << 9)   /* STP:26 */
#define GSMO_F_COND_RHO_SAMPLING    (1u << 10)  /* STP:28 */
#define GSMO_F_COND_ROOT            (1u << 11)  /* STP:31 */
#define GSMO_F_ORIGIN               (1u << 12)  /* STP.origin wired (STP:24); value = params[8] */
#define GSMO_F_BSP_HAS_INDICATOR    (1u << 13)  /* BranchSpikePrior.indicator wired (BSP:35) */

/* Stubs.StubMode (Stubs.java:36-40) */
#define GSMO_STUBS_NOSTUB  0
#define GSMO_STUBS_COUNTS  1
#define GSMO_STUBS_EXACT   2

/* per-tree parameter row */
#define GSMO_NPARAM 9
enum { GSMO_P_CLOCK_RATE = 0, GSMO_P_SPIKE_MEAN, GSMO_P_SPIKE_SHAPE, GSMO_P_CLOCK_SD,
       GSMO_P_LAMBDA, GSMO_P_MU, GSMO_P_PSI, GSMO_P_RHO, GSMO_P_ORIGIN };

/* per-tree result row */
#define GSMO_NOUT 8
enum { GSMO_O_STP_LOGP = 0,   /* value returned by StumpedTreePrior.calculateTreeLogLikelihood (STP:150-284),
                                 without the cross-call nExtantTaxa check (STP:196-200, host state) */
       GSMO_O_TREE_LOGP,      /* treeLogP component (STP:167) */
       GSMO_O_STUB_LOGP,      /* stubLogP component (STP:168) */
       GSMO_O_SPIKE_LOGP,     /* BranchSpikePrior.calculateLogP (BSP:53-187) */
       GSMO_O_RATE_LOGP,      /* BranchRatePrior.calculateLogP (BRP:29-65) */
       GSMO_O_SUM_BURST,      /* SaltativeProportionLogger abruptChange (SPL:36,54) */
       GSMO_O_SUM_DIST,       /* SaltativeProportionLogger totalChange (SPL:35,49) */
       GSMO_O_N_EXTANT };     /* extant-leaf count n (STP:190-195) */

typedef struct gsmo_config {
    uint32_t flags;
    int32_t  stub_mode;    /* GSMO_STUBS_*; meaningful only where a STUBS_TO_* flag is set */
} gsmo_config;

/* ---- scalar building blocks (exported for unit tests) ---- */
double gsmo_log_gamma(double x);                                       /* commons-math Gamma.logGamma */
double gsmo_gamma_log_density(double alpha, double beta, double x);    /* GammaDistributionImpl.logDensity */
double gsmo_normal_log_density(double mean, double sd, double x);      /* NormalDistributionImpl.logDensity */
double gsmo_lognormal_log_density(double m, double s, double x);       /* BRP:167-172 */
double gsmo_q(double t, double c1, double c2);                         /* STP:489-491 */
double gsmo_one_minus_p0hat(double t, double lambda, double mu, double rho); /* STP:492-494 */
double gsmo_c1(double lambda, double mu, double psi);                  /* STP:496-498 */
double gsmo_c2(double lambda, double mu, double psi, double rho);      /* STP:500-502 */
double gsmo_p0(double h, double lambda, double mu, double psi, double c1, double c2); /* STP:506-507 */
double gsmo_p1(double h, double rho, double c1, double c2);            /* STP:523-524 */
double gsmo_log_qt_nosampling(double t, double lambda, double mu);     /* STP:826-831 */
double gsmo_log_qt_sampling(double t, double lambda, double mu, double psi, double rho); /* STP:813-820 */
double gsmo_expected_stubs_sampling(double h1, double h0, double lambda, double mu, double psi, double rho); /* STP:749-754 */
double gsmo_expected_stubs_nosampling(double h1, double h0, double lambda, double mu, double root_height); /* STP:758-769 */
double gsmo_mean_stub_number(double node_h, double parent_h, double lambda, double mu, double psi, double rho,
                             double root_height);                      /* STP:876-892 */
/* parameter maps STP:672-737; pass NaN for an unwired input */
void gsmo_param_maps(double lambda_in, double net_div_in, double r0_in, double turnover_in, double psi_in,
                     double sampling_prop_in, double rho_in, double out_lmpr[4]);

/* ---- one tree ---- */
typedef struct gsmo_tree {
    int32_t n;                 /* node count */
    const double*  height;     /* [n] */
    const int32_t* parent;     /* [n], -1 for the root */
    const double*  rate;       /* [n] or NULL when rates are not wired */
    const double*  spike;      /* [n] */
    const int32_t* nstubs;     /* [n]; entry i >= n-1 is ignored (Stubs.java:200,354); NULL if no counts.  Not used in
                                  EXACT mode: there the counts come from the stub list (Stubs.java:357-364) */
    const double*  params;     /* [GSMO_NPARAM] */
    int32_t use_spike;         /* value of the indicator (GSMuseSpikeModel) */
    /* EXACT (reversible-jump) mode, Stubs.java:42-54: entries 1 .. dim-1 of Stubs.branchNr / Stubs.time (index 0 is the
     * dummy that Stubs.includeStub always rejects, Stubs.java:285-287) */
    int32_t n_stub_entries;
    const int32_t* stub_branch;  /* [n_stub_entries] branch (node number) the stub sits on */
    const double*  stub_time;    /* [n_stub_entries] relative position along the branch, Stubs.java:625-640 */
} gsmo_tree;

/* Evaluate everything for one tree.  out[GSMO_NOUT]; out_rates / out_wspikes are [n] or NULL
 * (getRatesArray PRCM:267-276; SpikeSize.getArrayValue SpikeSize:48-56 with clockModel wired). */
void gsmo_eval_tree(const gsmo_config* cfg, const gsmo_tree* t, double* out, double* out_rates, double* out_wspikes);

/* BranchSpikePrior.getCumulativeProbs(mu, nodeNr) BSP:315-389.  Writes up to cap entries,
 * returns the length K the Java array would have (may exceed cap; then only cap are written). */
int32_t gsmo_cumulative_probs(double mu, double spike, double shape, int use_indicator_branch,
                              double* out, int32_t cap);

/* Randomizer.randomChoice(cf) for a given uniform U (Stubs.sampleNStubsOnBranch Stubs:446-448) */
int32_t gsmo_random_choice(const double* cf, int32_t len, double U);

/* ---- batch (tree-major rows with a common stride) ---- */
int gsmo_eval_batch(const gsmo_config* cfg, int64_t n_trees, int32_t n_nodes, int64_t stride,
                    const double* height, const int32_t* parent, const double* rate, const double* spike,
                    const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                    const int32_t* node_count /* NULL => n_nodes for every tree */,
                    double* out_tree, double* out_rates, double* out_wspikes, int n_threads);
/* ... with the stub lists of EXACT mode, CSR by tree: entries [stub_ptr[t], stub_ptr[t+1]) belong to tree t */
int gsmo_eval_batch_stubs(const gsmo_config* cfg, int64_t n_trees, int32_t n_nodes, int64_t stride,
                          const double* height, const int32_t* parent, const double* rate, const double* spike,
                          const int32_t* nstubs, const double* params, const uint8_t* use_spike,
                          const int32_t* node_count, const int64_t* stub_ptr, const int32_t* stub_branch,
                          const double* stub_time, double* out_tree, double* out_rates, double* out_wspikes, int n_threads);
/* getStubLogBirthRate STP:796-806 (log 2 + log lambda + logQt) */
double gsmo_stub_log_birth_rate_sampling(double t, double lambda, double mu, double psi, double rho);
double gsmo_stub_log_birth_rate_nosampling(double t, double lambda, double mu);

int gsmo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
