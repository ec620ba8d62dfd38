// gsm_lean.cuh — the lean evaluation path: per-node bodies of gsm_lean_kernel (gsm_kernels.cu).
//
// The generic functions of gsm_node_math.cuh follow the Java statement by statement and handle every wiring and
// every special value; they cost ~450 instructions per branch.  This file is the same arithmetic for the default
// wiring family (rates wired, known-n spike prior, no noSpikeOnDatedTips: gsm_lean_wiring_ok) restructured so that
// a branch costs ~150 instructions:
//
//   * BEAST's node numbering is used instead of rediscovered: leaves are nodes 0..L-1, L = (n+1)/2
//     (beast.base.evolution.tree.Tree invariant; validated per tree: n odd, every parent index in [L, n)).
//   * per-branch terms are accumulated as sums of their factors (sum log r, sum log^2 r, sum len, sum dA, ...)
//     and combined with the per-tree constants once per thread, instead of one fused term per branch;
//   * log / exp use 512-entry tables staged in shared memory (|r| <= 2^-10: degree-4 polynomials, 8-9 fp64
//     instructions each, absolute error <= 3e-16);
//   * no per-branch special-value branches: every value that feeds a log / table lookup is folded into two
//     unsigned maxima per thread ("chk" words).  A tree whose maxima show anything but positive normal
//     arguments, stub counts in 0..31, non-negative branch lengths ... is NOT finished here: it is appended to
//     a redo list and evaluated by the generic kernel (gsm_eval_kernel) in a second launch.  So the reference's
//     behaviour for -inf / NaN / invalid states is inherited from the generic path, never re-implemented.
//   * two loop flavours: CLEAN (root is node n-1 and every other branch has positive length — an ordinary
//     BDWS tree) and GENERAL (sampled ancestors / fake nodes / zero-length branches / root anywhere).
//
// Everything is GSM_HD so that tests/emu/gsm_emu.cpp can walk the same code on the host.
#pragma once

#include "gsm_node_math.cuh"

#define GSM_LEAN_MAX_NST 63      // stub counts handled by the tables of the lean path (0..63); beyond: generic path
#define GSM_LEAN_MTAB_N 65       // m = 0..64 spikes on a branch
#define GSM_LEAN_MAX_NODES 32767 // parent index is 15 bits of the packed word

// ---- constant tables of log / exp: 64 entries of 8 bytes, read from 16-fold replicated copies ------------------
// A table lookup with a data-dependent index is the most expensive instruction of this path when it goes to shared
// memory as it is: 32 lanes asking for 32 random 16-byte entries cost ~10 shared-memory wavefronts
// (profiles/r01_v5_lean_ncu_summary.txt: 77 % of the L1/shared pipe, the top unit).  Here every lane of a half-warp
// owns a replica of the table that lives in its own pair of banks: entry j of replica q is at byte (j * 16 + q) * 8, so
// an LDS.64 is two wavefronts whatever the indices are.
//   log: x = 2^k z with z in [1, 2); entry j (64 linear sub-intervals of z) is invc_j ~ 2^(-i_j/64)
//        with i_j = round(64 log2(centre_j)) kept in the low 6 bits of the entry itself.  Then
//            log(x) = (64 k + i_j) ln2/64 + log1p(r),  r = z invc_j - 1,  |r| <= 0.01248,
//        i.e. no second table for log(c_j), and sums of logs collect the integer part in integer registers.
//   exp: 2^(j/64); |r| <= ln2/128.
#define GSM_T6_N 64
#define GSM_T6_REP 16
#define GSM_T6_WORDS (GSM_T6_N * GSM_T6_REP)   // uint64 words of the replicated log table (8 KB, on an 8 KB boundary)
// The exp table is looked up once per internal node (pass A) and per leaf above height zero, against ~3.5 log lookups per
// node: it gets 8 replicas only (entry j of replica q at byte (j * 8 + q) * 8; 4 KB on a 4 KB boundary; two lanes of a
// half-warp share a replica) and the 4 KB saved decide whether three CTAs fit an SM at 2,000 taxa.
#define GSM_T6E_REP 8
#define GSM_T6E_WORDS (GSM_T6_N * GSM_T6E_REP)

// device image of both replicated tables (one TMA bulk copy per CTA); filled by gsm_init_tables / the emulator
struct __attribute__((aligned(128))) GsmLeanBlob {
    uint64_t logtab[GSM_T6_WORDS];
    uint64_t exptab[GSM_T6E_WORDS];
};
static const uint64_t gsm_log6_tab_h[GSM_LOG6_TAB_N] = GSM_LOG6_TAB_INIT;
static const uint64_t gsm_exp6_tab_h[GSM_EXP6_TAB_N] = GSM_EXP6_TAB_INIT;
inline void gsm_lean_blob_fill(GsmLeanBlob& b) {
    for (int j = 0; j < GSM_T6_N; j++) {
        for (int q = 0; q < GSM_T6_REP; q++) b.logtab[j * GSM_T6_REP + q] = gsm_log6_tab_h[j];
        for (int q = 0; q < GSM_T6E_REP; q++) b.exptab[j * GSM_T6E_REP + q] = gsm_exp6_tab_h[j];
    }
}

GSM_HD bool gsm_lean_wiring_ok(uint32_t flags, int32_t stub_mode) {
    // known-n spike prior (BSP:77-100) and the Poisson mixture over the stub count (BSP:102-172, Stubs nostub mode: the
    // wiring of both BEAUti templates) are both covered; the mixture is a flavour of the pass B body (MIX)
    // reversible-jump placement terms (STP:573-600, 639-663) are evaluated by the generic path only
    const bool rj_terms = stub_mode == GSM_STUBS_EXACT && (flags & GSM_F_STUBS_TO_TREE_PRIOR) && !(flags & GSM_F_IGNORE_STUB_PRIOR);
    return !rj_terms && (flags & GSM_F_HAS_RATES) && !(flags & GSM_F_NO_SPIKE_DATED_TIPS);
}
GSM_HD bool gsm_lean_mixture(uint32_t flags, int32_t stub_mode) {   // BSP:77 false
    return (flags & GSM_F_STUBS_TO_SPIKE_PRIOR) != 0 && stub_mode == GSM_STUBS_NOSTUB;
}

// fp64 constants that do not fit an instruction immediate live in constant memory on the device, so that the hot
// loops use them as constant-bank operands instead of re-materialising them with moves
enum { GSM_LK_LN2_64 = 0, GSM_LK_KBIAS, GSM_LK_INVLN2N, GSM_LK_NLN2N, GSM_LK_SHIFT, GSM_LK_TWO52, GSM_LK_LEAF_EPS, GSM_LK_COND_EPS,
       GSM_LK_Q3, GSM_LK_Q4 = GSM_LK_Q3 + 4, GSM_LK_E2 = GSM_LK_Q4 + 5, GSM_LK_N = GSM_LK_E2 + 3 };
#define GSM_LK_INIT {GSM_LN2_64, 4503599628419072.0 /* 2^52 + 2^20 */, GSM_EXP6_INVLN2N, GSM_EXP6_NEGLN2N, 0x1.8p52, 0x1p52, \
                     GSM_LEAF_EPS, 0.000000000005 /* STP:470 */, \
                     GSM_LOG6_Q3, GSM_LOG6_Q4, GSM_EXP6_Q2}
static const double gsm_lean_k_h[GSM_LK_N] = GSM_LK_INIT;
#if defined(__CUDACC__)
static __constant__ double gsm_lean_k_d[GSM_LK_N] = GSM_LK_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define GSM_LK(i) gsm_lean_k_d[i]
#else
#define GSM_LK(i) gsm_lean_k_h[i]
#endif

// one thread's view of the two replicated tables
struct GsmT6 {
#if defined(__CUDACC__)
    // device: shared-window byte address of the table (log: 8 KB aligned, exp: 4 KB aligned: the entry offset is OR-ed in)
    // | this lane's replica offset (log: (lane & 15) * 8, exp: (lane & 7) * 8)
    uint32_t log_ar, exp_ar;
#else
    const uint64_t* log_t;   // host (CPU-side logic tests): replica 0
    const uint64_t* exp_t;
#endif
};
// d = (a & c) | b / (a & ~c) | (b & c) in one instruction
GSM_HD uint32_t gsm_and_or(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return (a & c) | b;
#endif
}
GSM_HD uint32_t gsm_bitsel(uint32_t a, uint32_t b, uint32_t c) {   // bits of b where c is set, bits of a elsewhere
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return (a & ~c) | (b & c);
#endif
}
// entry j of a table; jb = j << 7 plus arbitrary bits outside 0x1f80
GSM_HD uint64_t gsm_t6_log_ld(const GsmT6& t, uint32_t jb) {
#if defined(__CUDA_ARCH__)
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(gsm_and_or(jb, t.log_ar, 0x1f80u)));
    return v;
#elif defined(__CUDACC__)
    (void)t; (void)jb;
    return 0;
#else
    return t.log_t[(jb & 0x1f80u) >> 3];
#endif
}
// exp table: jb = j << 6 plus arbitrary bits outside 0xfc0
GSM_HD uint64_t gsm_t6_exp_ld(const GsmT6& t, uint32_t jb) {
#if defined(__CUDA_ARCH__)
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(gsm_and_or(jb, t.exp_ar, 0xfc0u)));
    return v;
#elif defined(__CUDACC__)
    (void)t; (void)jb;
    return 0;
#else
    return t.exp_t[(jb & 0xfc0u) >> 3];
#endif
}

// ---- table-driven log / exp (arguments must be "regular": see the chk words) ------------------------------------
// r = z invc - 1 and K = 64 k + i of a positive normal x: log(x) = K ln2/64 + log1p(r).  Integer work per call: two shifts,
// three three-input logic operations, one add.
GSM_HD double gsm_log6_r(const GsmT6& t, double x, int32_t& K) {
    const uint32_t hi = (uint32_t)gsm_hi32(x);
    const uint64_t e = gsm_t6_log_ld(t, hi >> 7);
    const uint32_t elo = (uint32_t)(e & 0xffffffffu);
    const double z = gsm_hilo((int32_t)gsm_bitsel(0x3ff00000u, hi, 0x000fffffu), gsm_lo32(x));   // 2^-k x in [1, 2)
    K = (int32_t)gsm_bitsel(hi >> 14, elo, 63u) - 64 * 1023;
    return fma(z, gsm_hilo((int32_t)(uint32_t)(e >> 32), (int32_t)elo), -1.0);
}
// log1p(r), |r| <= 0.01248: degree 5, absolute error <= 8e-14 (the per-branch sums)
GSM_HD double gsm_log6_y3(double r) {
    // Estrin's scheme: the same five operations as Horner's, a dependency chain of three instead of four
    const double r2 = r * r;
    const double a = fma(GSM_LK(GSM_LK_Q3 + 1), r, GSM_LK(GSM_LK_Q3 + 0));
    const double b = fma(GSM_LK(GSM_LK_Q3 + 3), r, GSM_LK(GSM_LK_Q3 + 2));
    return fma(fma(b, r2, a), r2, r);
}
// ... degree 6, absolute error <= 4e-16 (aux, which feeds differences)
GSM_HD double gsm_log6_y4(double r) {
    const double r2 = r * r;
    double q = fma(GSM_LK(GSM_LK_Q4 + 4), r, GSM_LK(GSM_LK_Q4 + 3));
    q = fma(q, r, GSM_LK(GSM_LK_Q4 + 2));
    q = fma(q, r, GSM_LK(GSM_LK_Q4 + 1));
    q = fma(q, r, -0.5);
    return fma(q, r2, r);
}
// int -> double: the conversion instruction runs on its own pipe, next to the fp64 and integer ones
GSM_HD double gsm_k2d(int32_t K) { return (double)K; }
GSM_HD double gsm_u2d(int32_t n) { return (double)n; }

GSM_HD double gsm_log6(const GsmT6& t, double x) {   // degree 5
    int32_t K;
    const double r = gsm_log6_r(t, x, K);
    return fma(gsm_k2d(K), GSM_LK(GSM_LK_LN2_64), gsm_log6_y3(r));
}
GSM_HD double gsm_log6_hi(const GsmT6& t, double x) {   // degree 6
    int32_t K;
    const double r = gsm_log6_r(t, x, K);
    return fma(gsm_k2d(K), GSM_LK(GSM_LK_LN2_64), gsm_log6_y4(r));
}

// exp(x), |x| < 700.  One-step argument reduction: its error, |x| 1e-16 relative, only matters where exp(x) does
// (x = -c1 h <= 0 here: G = (1 + c2) + (1 - c2) e^x).  Relative error <= 1e-14.
GSM_HD double gsm_exp6(const GsmT6& t, double x) {
    double kd = fma(x, GSM_LK(GSM_LK_INVLN2N), GSM_LK(GSM_LK_SHIFT));
    const int32_t ki = gsm_lo32(kd);
    kd -= GSM_LK(GSM_LK_SHIFT);
    const double r = fma(kd, GSM_LK(GSM_LK_NLN2N), x);
    const uint64_t e = gsm_t6_exp_ld(t, (uint32_t)ki << 6);
    const double scale = gsm_hilo((int32_t)((uint32_t)(e >> 32) + ((uint32_t)ki << 14)), (int32_t)(uint32_t)(e & 0xffffffffu));
    const double r2 = r * r;
    double q = fma(GSM_LK(GSM_LK_E2 + 2), r, GSM_LK(GSM_LK_E2 + 1));
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r2, r);   // expm1(r), |r| <= ln2/128
    return fma(scale, em1, scale);
}

// "specialness" of a log argument as an unsigned number: >= GSM_CHK_LIMIT unless x is a positive normal finite double
#define GSM_CHK_LIMIT 0x7fe00000u
GSM_HD uint32_t gsm_chk_pos(double x) { return (uint32_t)(gsm_hi32(x) - 0x00100000); }
GSM_HD uint32_t gsm_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

// ---- per-tree constants held in registers -----------------------------------------------------------------------
// per-tree tables in shared memory (built by the setup warp):
//   mtab[m]  = {shape m - 1, log(shape) - logGamma(shape m)}, m = 0..64; mtab[0] = {-1, 0}            (BSP:97-99)
//   ntab[n]  = {log(shape) - logGamma(shape (n + 1)), -sum_{i=2..n} log i, shape (n + 1) - 1, (double)n}, n = 0..63
//              (BSP:97-99, STP:567): everything a branch with n stubs needs from its count, two 16-byte loads
struct __attribute__((aligned(16))) GsmLeanN { double gcon, nlf, w, nd; };
struct GsmLeanC {
    // pass A
    double neg_c1, omc2, opc2, kS, c1, rho;
    // pass B
    double kEh;        // (lambda + mu + psi - c1) / 2: E/2 = len * kEh + (aux - aux_parent)   (STP:749-769 via aux)
    double shape, base, eff_mean;
    GsmT6 t6;
    const GsmLogEntry* mtab;
    const GsmLeanN* ntab;
    const GsmLogEntry* mixtab; // mixture mode (overlays ntab): [k] = {g[k+1] / (k+1), 1 / (k+1)}, [64 + k] = {g[k] / (k+1), 1 / (k+1)},
                               // g[j] = Gamma(shape j) / Gamma(shape (j+1)), k = 0..62
    int32_t L;                 // leaves are nodes 0..L-1
    int32_t leaf_term;         // 0: none, 1: psi > 0 (STP:337-343), 2: origin / root conditioned (STP:467-474)
    bool use_counts;           // stub counts exist (COUNTS / EXACT); else getNStubsOnBranch == -1 (Stubs:348-350)
    bool s2p, s2c;             // stubs wired to the spike prior (BSP:81) / the clock model (PRCM:192-195)
    bool stub_term;            // STP:219 && count mode
    bool use_rate;             // relaxed clock (PRCM:216)
    bool mix;                  // mixture-mode spike prior (BSP:102-172)
    bool sa_unscanned;         // the tree was not scanned for sampled ancestors (no fake flags): a direct-ancestor leaf met by
                               // a GENERAL body sends the tree to the generic path
};

// true when the per-tree constants allow the lean path at all (everything else goes to the generic kernel)
GSM_HD bool gsm_lean_consts_ok(const GsmTreeConsts& K) {
    if (K.stp_early || K.spike_invalid || K.rate_invalid) return false;
    const bool fin = (K.c1 >= 0.0) && (K.c1 < INFINITY) && (K.opc2 > 1e-300) && (K.opc2 < 1e300) && (K.omc2 > -1e300) &&
                     (K.omc2 < 1e300) && (K.opc2 + K.omc2 > 1e-300);   // G(h) stays a positive normal number < 2^1022
    const bool par = (K.shape > 1e-300) && (K.shape < 1e300) && (K.sigma < 1e150) && (K.base == K.base) &&
                     (K.spike_mean == K.spike_mean) && (fabs(K.kE) < INFINITY) && (fabs(K.kS) < INFINITY) &&
                     (K.psi < 700.0);   // exp(psi) overflow makes STP -inf (STP:176-178): generic path
    // mixture mode: g[m] = Gamma(shape m) / Gamma(shape (m + 1)) must stay inside the range of the table-driven exp
    const bool mixok = K.spike_known_n || (K.shape <= 80.0);
    return fin && par && mixok;
}

GSM_HD void gsm_lean_consts(const GsmTreeConsts& K, GsmLeanC& c, int32_t n) {
    c.neg_c1 = -K.c1; c.omc2 = K.omc2; c.opc2 = K.opc2; c.kS = K.kS; c.c1 = K.c1; c.rho = K.rho;
    c.kEh = 0.5 * K.kE; c.shape = K.shape; c.base = K.base;
    c.eff_mean = ((K.flags & GSM_F_HAS_INDICATOR) && !K.use_spike) ? 0.0 : K.spike_mean;   // PRCM:181-183
    c.L = (n + 1) >> 1;
    c.leaf_term = K.tree_variant == GSM_TV_PSI ? 1 : (K.tree_variant == GSM_TV_CONDITIONED ? 2 : 0);
    c.use_counts = K.stub_mode != GSM_STUBS_NOSTUB;
    c.s2p = (K.flags & GSM_F_STUBS_TO_SPIKE_PRIOR) != 0;
    c.s2c = (K.flags & GSM_F_STUBS_TO_CLOCK) != 0;
    c.stub_term = K.stub_term && K.stub_counts;
    c.use_rate = !(K.flags & GSM_F_RELAXED_FALSE);
    c.mix = !K.spike_known_n;
    c.sa_unscanned = false;
}

// ---- pass A: aux(h) = log G(h) for every node; extinct-leaf terms of the psi > 0 densities -----------------------
struct GsmLeanAccA {
    double leafterm;      // sum over leaves that need it of  log(numer) + c1 h + aux
    int32_t n_leafterm;   // ... their number (each adds t_leaf)
    int32_t n_plain;      // origin/root-conditioned: leaves that add log(4 rho) instead (STP:473)
    uint32_t chk;         // specialness of numer
    uint32_t hmax;        // max of the high words of the heights (>= 0x7ff00000: negative / inf / NaN)
};
GSM_HD void gsm_lean_acc_a_zero(GsmLeanAccA& S) { S.leafterm = 0.0; S.n_leafterm = S.n_plain = 0; S.chk = 0u; S.hmax = 0u; }

// aux of one node.  e_out / G_out feed the leaf term.
GSM_HD double gsm_lean_aux(const GsmLeanC& c, GsmLeanAccA& S, double h, double& e_out, double& G_out) {
    S.hmax = gsm_umax(S.hmax, (uint32_t)gsm_hi32(h));
    const double e = gsm_exp6(c.t6, c.neg_c1 * h);
    const double G = fma(c.omc2, e, c.opc2);
    e_out = e;
    G_out = G;
    return gsm_log6_hi(c.t6, G);
}

// aux of four nodes at once, stage by stage (the table loads of the four are issued back to back, the four polynomial
// chains interleave): the same operations per node as gsm_lean_aux, bit for bit.
GSM_HD void gsm_lean_aux4(const GsmLeanC& c, GsmLeanAccA& S, const double* h, double* aux) {
    double r[4];
    uint64_t e[4];
    int32_t ki[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        S.hmax = gsm_umax(S.hmax, (uint32_t)gsm_hi32(h[k]));
        const double x = c.neg_c1 * h[k];
        double kd = fma(x, GSM_LK(GSM_LK_INVLN2N), GSM_LK(GSM_LK_SHIFT));
        ki[k] = gsm_lo32(kd);
        kd -= GSM_LK(GSM_LK_SHIFT);
        r[k] = fma(kd, GSM_LK(GSM_LK_NLN2N), x);
        e[k] = gsm_t6_exp_ld(c.t6, (uint32_t)ki[k] << 6);
    }
    double G[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double scale = gsm_hilo((int32_t)((uint32_t)(e[k] >> 32) + ((uint32_t)ki[k] << 14)), (int32_t)(uint32_t)(e[k] & 0xffffffffu));
        const double r2 = r[k] * r[k];
        double q = fma(GSM_LK(GSM_LK_E2 + 2), r[k], GSM_LK(GSM_LK_E2 + 1));
        q = fma(q, r[k], 0.5);
        const double em1 = fma(q, r2, r[k]);
        G[k] = fma(c.omc2, fma(scale, em1, scale), c.opc2);
    }
    uint64_t le[4];
#pragma unroll
    for (int k = 0; k < 4; k++) le[k] = gsm_t6_log_ld(c.t6, (uint32_t)gsm_hi32(G[k]) >> 7);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t hi = (uint32_t)gsm_hi32(G[k]);
        const uint32_t elo = (uint32_t)(le[k] & 0xffffffffu);
        const double z = gsm_hilo((int32_t)gsm_bitsel(0x3ff00000u, hi, 0x000fffffu), gsm_lo32(G[k]));
        const int32_t K = (int32_t)gsm_bitsel(hi >> 14, elo, 63u) - 64 * 1023;
        const double rr = fma(z, gsm_hilo((int32_t)(uint32_t)(le[k] >> 32), (int32_t)elo), -1.0);
        aux[k] = fma(gsm_k2d(K), GSM_LK(GSM_LK_LN2_64), gsm_log6_y4(rr));
    }
}

// LT = c.leaf_term (compile-time in the kernel's loops); leaf and not a direct ancestor is decided by the caller
template <int LT>
GSM_HD void gsm_lean_leaf_term(const GsmLeanC& c, GsmLeanAccA& S, bool leaf_not_da, double h, double aux, double e, double G) {
    if (LT == 0) return;
    bool need;
    if (LT == 1) need = leaf_not_da && !(h <= GSM_LK(GSM_LK_LEAF_EPS));                       // STP:339
    else need = leaf_not_da && (h > GSM_LK(GSM_LK_COND_EPS) || c.rho == 0.);                  // STP:470
    const double numer = c.kS * G - c.c1 * (c.opc2 - e * c.omc2);                             // 2 lambda G p0(h), STP:506-507
    const double arg = need ? numer : 1.0;
    S.chk = gsm_umax(S.chk, gsm_chk_pos(arg));
    const double ln = gsm_log6_hi(c.t6, arg);
    if (need) {
        S.leafterm += ln + fma(c.c1, h, aux);
        S.n_leafterm++;
    } else if (LT == 2 && leaf_not_da) {
        S.n_plain++;
    }
}

// ---- pass B -----------------------------------------------------------------------------------------------------
// Logs enter the sums as log(x) = K ln2/64 + y: the y parts are summed in fp64, the K parts (times their integer
// weights) in integer registers.
struct GsmLeanAccB {
    double gam;        // sum y(x shape) (shape m - 1)                                           (BSP:97-99, the log term)
    double gamc;       // sum log(shape) - logGamma(shape m)
    double spk;        // sum of spikes                                                           (x shape term; burst in CLEAN)
    double spk_el;     // GENERAL: sum of spikes that count as bursts of eligible branches        (SPL:53-54)
    double lr, lr2;    // sum log r, sum log^2 r                                                  (BRP:171)
    double stub;       // sum n_i y(E_i / 2)                                                      (STP:566)
    double nlf;        // sum -log(n_i!)                                                          (STP:567)
    double len, dA;    // sum len, sum (aux - aux_parent)
    double lenr;       // sum len * r                                                             (SPL:44-49 / PRCM:233)
    double auxint, hint;     // sums over internal nodes of aux, h
    double auxfake, hfake;   // GENERAL: the same over fake internal nodes
    int32_t k_mk, k_k; // sum (m_i - 1) K_i, sum K_i of x_i shape: ln2/64 (shape (k_mk + k_k) - k_k) belongs to gam
    int32_t k_nk;      // sum n_i (K_i + 64) of E_i / 2: ln2/64 k_nk belongs to stub
    double mix;        // MIX: sum y(S_i), S_i = the mixture sum of a branch relative to its first term   (BSP:112-151)
    int32_t k_mix;     // ... and the K parts
    int32_t n_extant, n_extinct, n_da, n_fake;
    uint32_t chk;      // specialness of rate, spike*shape, E/2 (and, GENERAL, of impossible combinations)
    uint32_t chk_n;    // max stub count as unsigned (negative counts become huge)
    uint32_t chk_len;  // CLEAN: specialness of len (zero / negative lengths -> rerun as GENERAL)
};
GSM_HD void gsm_lean_acc_b_zero(GsmLeanAccB& S) {
    S.gam = S.gamc = S.spk = S.spk_el = S.lr = S.lr2 = S.stub = S.nlf = S.len = S.dA = S.lenr = 0.0;
    S.auxint = S.hint = S.auxfake = S.hfake = 0.0;
    S.mix = 0.0;
    S.k_mix = 0;
    S.k_mk = S.k_k = S.k_nk = S.n_extant = S.n_extinct = S.n_da = S.n_fake = 0;
    S.chk = S.chk_n = S.chk_len = 0u;
}

// One node i.  GENERAL only: root (the node has no parent: the caller passes hp = h, auxp = aux), pf (the parent is
// fake), fake_self.
//   GENERAL  false: ordinary tree (no zero-length branch in the main loop); true: any tree
//   LEAF     1: i < L, 0: i >= L (compile-time for chunks on one side of L), 2: decided here
//   STDW     the default wiring (stub counts wired to clock model, spike prior and tree prior; relaxed clock) is
//            compiled in; false: the wiring is read from c
// MIX (only with STDW == false): the spike prior is the Poisson mixture over the unknown stub count (BSP:102-172).  With
//   mu = E (expected stubs on the branch), x = spike, xs = x shape, lgc[m] = log shape - logGamma(shape m):
//   branchP = sum_{k < K} Pois(k; mu) Gamma(x; shape m(k), 1/shape),  K = first k with sum_{j<k} Pois(j; mu) >= 0.999,
//   term_k = exp(-mu - xs + (shape m - 1) log xs + lgc[m] + k log mu - log k!).  For m(k) = k + 1 (BSP:197-213)
//   log branchP = [-mu - xs + (shape - 1) log xs + lgc[1]] + log S,  S = sum_k w_k,  w_0 = 1,
//   w_{k+1} = w_k (mu xs^shape) g[k+1] / (k+1): one exp for xs^shape, one for Pois(0) = e^-mu, one log for S, and a
//   multiply-add per term -- instead of two exp and a table walk per term.  Children of a fake node have m(k) = k:
//   term_0 counts only for x == 0 (then log branchP = -mu), else log branchP = [...] + log(mu S'), S' = sum_{1<=k<K} w'_k,
//   w'_1 = 1, w'_{k+1} = w'_k (mu xs^shape) g[k] / (k+1).  The bracket is the known-n term with m = 1 and goes through the
//   same accumulators; -mu is recovered from the length / aux sums.
template <bool GENERAL, bool WANT_RATES, int LEAF, bool STDW, bool MIX = false>
GSM_HD void gsm_lean_b_node(const GsmLeanC& c, GsmLeanAccB& S, int32_t i, bool root, bool parent_fake, bool fake_self, double h,
                            double hp, double aux, double auxp, double rate, double spike, int32_t nst_raw, double* out_rate,
                            double* out_wspike) {
    const bool use_counts = STDW || c.use_counts, s2p = STDW || c.s2p, s2c = STDW || c.s2c, stub_term = STDW || c.stub_term,
               use_rate = STDW || c.use_rate;
    const double len = hp - h;                          // root: 0
    const double dA = aux - auxp;
    const double E2 = fma(len, c.kEh, dA);              // half the expected stub count (STP:749-769)
    const bool have_len = len > 0.0;                    // STP:558 / PRCM:229
    const int32_t nst = use_counts ? nst_raw : -1;      // Stubs:344-370
    const int32_t ns = s2p ? nst : 0;                   // BSP:81
    const bool leaf = LEAF == 2 ? (i < c.L) : (LEAF == 1);
    int32_t m = ns + 1;                                 // BSP:197-213
    bool da = false, pf = false;
    if (GENERAL) {
        const bool zero = len == 0.0;
        da = leaf && zero && !root;                                                     // Node.isDirectAncestor
        pf = parent_fake && !root;                                                      // parent isFake
        m = root ? ns + 1 : (da ? 0 : (pf ? ns : ns + 1));
        if (!(have_len || zero)) S.chk = 0xffffffffu;                                   // negative / NaN length
        if (da && c.sa_unscanned) S.chk = 0xffffffffu;                                  // its sibling's parent_fake is unknown here
        if (!MIX && m == 0 && spike != 0.0) S.chk = 0xffffffffu;                        // BSP:86-95 -> -inf
        if (stub_term && !have_len && nst != 0) S.chk = 0xffffffffu;                    // STP:232-240, 558-561
    } else {
        S.chk_len = gsm_umax(S.chk_len, gsm_chk_pos(len));
    }
    if (use_counts) S.chk_n = gsm_umax(S.chk_n, (uint32_t)nst);   // above GSM_LEAN_MAX_NST: tree goes to the generic path
    const uint32_t ni = (uint32_t)nst < (uint32_t)GSM_LEAN_MAX_NST ? (uint32_t)nst : (uint32_t)GSM_LEAN_MAX_NST;

    // ---- BranchRatePrior BRP:43-60 ----
    S.chk = gsm_umax(S.chk, gsm_chk_pos(rate));
    const double lr = gsm_log6(c.t6, rate);
    S.lr += lr;
    S.lr2 = fma(lr, lr, S.lr2);

    // ---- BranchSpikePrior: BSP:69-100 (known stub count) / BSP:102-172 (MIX) ----
    const double xs = spike * c.shape;
    bool mix_sum = false, mix_pf = false;
    if (MIX) {
        // which spike count the terms of this branch start from: m(0) = 1, or 0 for a sampled ancestor / a child of a fake
        // node whose only admissible spike is zero (BSP:128-135, 158-164)
        const bool zero_m = GENERAL && (da || (pf && (!have_len || spike == 0.0)));
        m = zero_m ? 0 : 1;
        if (zero_m && spike != 0.0) S.chk = 0xffffffffu;     // -inf (BSP:128-135, 158-164): generic path
        mix_sum = GENERAL ? (have_len && !zero_m) : true;   // mu > 0 and a Gamma term exists
        mix_pf = GENERAL && pf;
        const uint32_t chk_mu = gsm_chk_pos(E2);             // mu = 2 E2 must be a positive normal number where it is used
        S.chk = gsm_umax(S.chk, (GENERAL && !have_len) ? 0u : chk_mu);
    }
    const double arg_s = (GENERAL && m == 0) ? 1.0 : xs;
    S.chk = gsm_umax(S.chk, gsm_chk_pos(arg_s));
    int32_t ks;
    const double ys = gsm_log6_y3(gsm_log6_r(c.t6, arg_s, ks));
    S.k_k += ks;
    S.spk += spike;
    if (!GENERAL && STDW) {
        // m = n + 1: one 32-byte entry carries the Gamma constant and weight of the spike prior, -log(n!) of the stub prior and
        // n as a double (no integer -> double conversions, no shape m - 1 arithmetic in the loop)
        const GsmLeanN ne = c.ntab[ni];
        S.gam = fma(ys, ne.w, S.gam);
        S.gamc += ne.gcon;
        S.k_mk += (int32_t)ni * ks;                     // sum (m - 1) K: the K themselves are in k_k
        // ---- stub density, count mode: STP:555-571 / 621-637 ----
        S.chk = gsm_umax(S.chk, gsm_chk_pos(E2));
        int32_t ke;
        const double yE = gsm_log6_y3(gsm_log6_r(c.t6, E2, ke));
        S.stub = fma(ne.nd, yE, S.stub);
        S.nlf += ne.nlf;
        S.k_nk += (int32_t)ni * (ke + 64);
    } else {
        const uint32_t mi = (uint32_t)m < (uint32_t)(GSM_LEAN_MTAB_N - 1) ? (uint32_t)m : (uint32_t)(GSM_LEAN_MTAB_N - 1);
        // {shape m - 1, log shape - logGamma(shape m)}; m = 0: {-1, 0}.  Without an mtab (persistent kernel) the same numbers
        // come from ntab[m - 1].
        GsmLogEntry me;
        if (c.mtab) {
            me = c.mtab[mi];
        } else {
            const GsmLeanN q = c.ntab[mi == 0u ? 0u : mi - 1u];
            me.invc = mi == 0u ? -1.0 : q.w;
            me.logc = mi == 0u ? 0.0 : q.gcon;
        }
        S.gam = fma(ys, me.invc, S.gam);
        S.gamc += me.logc;
        S.k_mk += ((int32_t)mi - 1) * ks;               // (MIX: m is 0 or 1; log 1 is K = 1, y = -ln2/64 in this table)
        if (stub_term) {
            const uint32_t chkE = gsm_chk_pos(E2);
            S.chk = gsm_umax(S.chk, (GENERAL && !have_len) ? 0u : chkE);
            int32_t ke;
            const double yE = gsm_log6_y3(gsm_log6_r(c.t6, E2, ke));   // finite garbage when E2 == 0 (len == 0): times n = 0
            S.stub = fma(gsm_u2d((int32_t)ni), yE, S.stub);
            S.nlf += c.ntab[ni].nlf;
            S.k_nk += (int32_t)ni * (ke + 64);
        }
    }
    if (MIX) {
        const double mu = 2.0 * E2;
        const double lxs = fma(gsm_k2d(ks), GSM_LK(GSM_LK_LN2_64), ys);
        double tpow = c.shape * lxs;                       // log xs^shape
        if (!(fabs(tpow) < 690.0) || !(mu < 690.0)) {     // outside the range of the table-driven exp: generic path
            if (mix_sum) S.chk = 0xffffffffu;
            tpow = 0.0;
        }
        const double step = mu * gsm_exp6(c.t6, tpow);     // mu xs^shape
        double pk = gsm_exp6(c.t6, mix_sum ? -mu : 0.0);   // Pois(0; mu)
        double cum = 0.0, w = 1.0, Ssum = 0.0;
        int32_t k = 0;
        if (mix_sum) {
            if (!mix_pf) {
                do {                                       // BSP:117-149 with m = k + 1
                    cum += pk;
                    Ssum += w;
                    const GsmLogEntry r = c.mixtab[k];
                    pk *= mu * r.logc;                     // r.logc = 1 / (k + 1)
                    w *= step * r.invc;                    // r.invc = g[k+1] / (k + 1)
                    k++;
                } while (cum < GSM_MAX_CUM_SUM && k < GSM_LEAN_MAX_NST);
            } else {
                do {                                       // m = k: the k = 0 term is zero (x != 0)
                    cum += pk;
                    Ssum += (k >= 1) ? w : 0.0;
                    const GsmLogEntry r = c.mixtab[k + GSM_LEAN_MAX_NST + 1];   // {g[k] / (k + 1), 1 / (k + 1)}
                    pk *= mu * r.logc;
                    w = (k >= 1) ? w * step * r.invc : 1.0;
                    k++;
                } while (cum < GSM_MAX_CUM_SUM && k < GSM_LEAN_MAX_NST);
                Ssum *= mu;
            }
            if (!(cum >= GSM_MAX_CUM_SUM)) S.chk = 0xffffffffu;   // more than 63 terms: generic path
        } else {
            Ssum = 1.0;
        }
        S.chk = gsm_umax(S.chk, gsm_chk_pos(Ssum));        // 0 (-> log 0 = -inf, BSP:151), inf, NaN: generic path
        int32_t kq;
        const double yq = gsm_log6_y3(gsm_log6_r(c.t6, Ssum, kq));
        S.mix += yq;
        S.k_mix += kq;
    }
    S.len += len;
    S.dA += dA;

    // ---- clock model: SPL sums, optional per-branch outputs (PRCM:178-242, SPL:33-61) ----
    const double r = use_rate ? rate : 1.0;             // PRCM:210-222
    S.lenr = fma(len, r, S.lenr);                       // len == 0 for every ineligible branch
    double bsp = spike;
    if (GENERAL) {
        if (pf && (!s2c || nst == 0)) bsp = 0.0;        // PRCM:192-202
        S.spk_el += have_len ? bsp : 0.0;
    }
    if (WANT_RATES) {
        const double burst = bsp * c.eff_mean;          // PRCM:204-206
        double eff = c.base;
        if (have_len) eff = (len * c.base * r + burst) / len;   // PRCM:233-236
        *out_rate = eff;
        *out_wspike = burst;
    }

    // ---- census of the leaves (STP:190-195, 291-312, 352-359) / sums over internal nodes ----
    if (leaf) {
        if (GENERAL) {
            const bool low = h <= GSM_LK(GSM_LK_LEAF_EPS);
            const bool ext = !low && (len > GSM_LK(GSM_LK_LEAF_EPS));
            S.n_da += da ? 1 : 0;
            S.n_extant += (!da && low) ? 1 : 0;
            S.n_extinct += (!da && ext) ? 1 : 0;
        } else {
#if defined(__CUDA_ARCH__)
            // two compares, two predicated increments
            asm("{\n\t.reg .pred lo, ex;\n\t"
                "setp.le.f64 lo, %2, %4;\n\t"
                "setp.gt.and.f64 ex, %3, %4, !lo;\n\t"
                "@lo add.s32 %0, %0, 1;\n\t"
                "@ex add.s32 %1, %1, 1;\n\t}"
                : "+r"(S.n_extant), "+r"(S.n_extinct)
                : "d"(h), "d"(len), "d"(GSM_LK(GSM_LK_LEAF_EPS)));
#else
            const bool low = h <= GSM_LK(GSM_LK_LEAF_EPS);
            const bool ext = !low && (len > GSM_LK(GSM_LK_LEAF_EPS));
            S.n_extant += low ? 1 : 0;
            S.n_extinct += ext ? 1 : 0;
#endif
        }
    } else {
        S.auxint += aux;
        S.hint += h;
        if (GENERAL && fake_self) {
            S.auxfake += aux;
            S.hfake += h;
            S.n_fake++;
        }
    }
}

GSM_HD bool gsm_lean_std_wiring(const GsmLeanC& c) { return c.use_counts && c.s2p && c.s2c && c.stub_term && c.use_rate; }

// per-thread combination of the factor sums into the six result sums (constants per node are added once per tree
// by gsm_lean_finish).  spk_el: in CLEAN every main-loop node is eligible.
struct GsmLean6 { double tree, stub, spike, rate, burst, dist; };

template <bool GENERAL>
GSM_HD void gsm_lean_combine(const GsmTreeConsts& K, const GsmLeanC& c, const GsmLeanAccA& SA, const GsmLeanAccB& S, GsmLean6& o) {
    const double aux_nf = S.auxint - (GENERAL ? S.auxfake : 0.0), h_nf = S.hint - (GENERAL ? S.hfake : 0.0);
    double tree = 0.0;
    const int32_t variant = K.tree_variant;
    if (variant == GSM_TV_COMPLETE_RHO) tree = (S.dA - K.lmm * S.len) - S.auxint;            // STP:387-393, 398-426
    else if (variant == GSM_TV_NO_PSI) tree = -(K.lmm * h_nf) - 2 * aux_nf;                  // STP:371-377
    else if (variant == GSM_TV_PSI || variant == GSM_TV_CONDITIONED) tree = -(K.c1 * h_nf) - 2 * aux_nf + SA.leafterm;
    o.tree = tree;
    const double l64 = GSM_LN2_64;
    o.stub = fma(l64, (double)S.k_nk, S.stub + S.nlf) - 2 * fma(c.kEh, S.len, S.dA);         // sum n log E - log n! - E
    o.spike = fma(l64, fma(c.shape, (double)(S.k_mk + S.k_k), -(double)S.k_k), S.gam + S.gamc) - c.shape * S.spk;
    if (c.mix) o.spike += fma(l64, (double)S.k_mix, S.mix) - 2 * fma(c.kEh, S.len, S.dA);   // sum log S_i - sum mu_i
    o.rate = -K.inv_2s2 * (S.lr2 - 2 * K.m_ln * S.lr) - S.lr;                                // sum -(lr - m)^2/(2 s^2) - lr, without n m^2
    o.burst = c.eff_mean * (GENERAL ? S.spk_el : S.spk);
    o.dist = c.base * S.lenr;
}

// per-tree: add the per-node constants and the tail accumulator (node n-1, evaluated by the generic functions).
// nb = number of nodes the main loops covered (n-1), of which nb - L internal.
GSM_HD void gsm_lean_finish(const GsmTreeConsts& K, const GsmLeanC& c, const GsmLean6& r, int32_t n_nst, int32_t n_leafterm,
                            int32_t n_plain, int32_t n_fake, int32_t nb, GsmAcc& A /* in: tail + census; out: totals */) {
    const int32_t n_int = nb - c.L;
    double tree = r.tree;
    // count * constant, with 0 * (+-inf) = 0: a term that never occurs adds nothing (log psi, log 4 rho can be -inf)
    auto times = [](int32_t cnt, double k) { return cnt != 0 ? cnt * k : 0.0; };
    switch (K.tree_variant) {
        case GSM_TV_COMPLETE_RHO: tree += times(n_int, K.t_internal); break;
        case GSM_TV_NO_PSI: tree += times(n_int - n_fake, K.t_internal); break;
        case GSM_TV_PSI: tree += times(n_int - n_fake, K.t_internal) + times(n_leafterm, K.t_leaf); break;
        case GSM_TV_CONDITIONED:
            tree += times(n_int - n_fake, K.t_internal) + times(n_leafterm, K.t_leaf) + times(n_plain, K.log_4rho) +
                    times(n_fake, K.log_psi);
            break;
        default: tree = 0.0; break;
    }
    A.tree += tree;
    (void)n_nst;
    if (c.stub_term) A.stub += r.stub;
    A.spike += r.spike;
    A.rate += r.rate - nb * (K.inv_2s2 * K.m_ln * K.m_ln + K.log_s_sqrt2pi);
    A.burst += r.burst;
    A.dist += r.dist + r.burst;
}
