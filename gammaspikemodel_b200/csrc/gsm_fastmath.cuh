// gsm_fastmath.cuh — table-driven fp64 log / exp for the per-branch kernels.
//
// Why: libdevice's log()/exp() cost ~100 SASS instructions each on sm_100a, most of them integer /
// uniform-datapath moves that materialise polynomial coefficients and branches for special cases
// (profiles/r01_v1_*: 783 instructions per branch, only 208 of them fp64).  These versions spend
// ~13 fp64 instructions + a few integer ops + one table load, take their coefficients as
// constant-bank operands, and fall back to libdevice only for zero/negative/subnormal/inf/NaN
// arguments (log) or |x| >= 704 (exp), which keeps the reference's special-value behaviour.
// Accuracy (tests/test_fastmath.py): < 1.5 ulp on the fast path, so every use stays far inside the
// 1e-10 relative tolerance of the path.
//
// GSM_HD: the same code is compiled by g++ for the CPU-side logic tests.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "gsm_tables.inc"

#if defined(__CUDACC__)
#define GSM_HD __host__ __device__ __forceinline__
#else
#define GSM_HD inline
#endif

struct __attribute__((aligned(16))) GsmLogEntry {
    double invc, logc;
};

static const GsmLogEntry gsm_log_tab_h[GSM_LOG_TAB_N] = GSM_LOG_TAB_INIT;
static const uint64_t gsm_exp_tab_h[GSM_EXP_TAB_N] = GSM_EXP_TAB_INIT;
// log1p(r) = r + r^2 * (c2 + c3 r + ... + c8 r^6)
static const double gsm_log_poly_h[7] = {-0.5, 1.0 / 3, -0.25, 0.2, -1.0 / 6, 1.0 / 7, -0.125};
// exp(r) - 1 - r = r^2 (1/2 + r/6) + r^4 (1/24 + r/120)
static const double gsm_exp_poly_h[4] = {0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120};

#if defined(__CUDACC__)
static __device__ const double2 gsm_log_tab_d[GSM_LOG_TAB_N] = GSM_LOG_TAB_INIT;
static __device__ const unsigned long long gsm_exp_tab_d[GSM_EXP_TAB_N] = GSM_EXP_TAB_INIT;
static __constant__ double gsm_log_poly_d[7] = {-0.5, 1.0 / 3, -0.25, 0.2, -1.0 / 6, 1.0 / 7, -0.125};
static __constant__ double gsm_exp_poly_d[4] = {0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120};
static __constant__ double gsm_fm_consts_d[8] = {GSM_LN2_HI, GSM_LN2_LO, GSM_EXP_INVLN2N, GSM_EXP_NEGLN2HIN,
                                                 GSM_EXP_NEGLN2LON, 0x1.8p52, 4503599627371520.0 /* 2^52 + 1024 */, 0.0};
#endif

GSM_HD int32_t gsm_hi32(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int32_t)(b >> 32);
#endif
}
GSM_HD int32_t gsm_lo32(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int32_t)(b & 0xffffffffu);
#endif
}
GSM_HD double gsm_hilo(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

// Rare-argument fallbacks are kept out of line so that the hot loops stay small.
#if defined(__CUDACC__)
static __device__ __noinline__ double gsm_log_slow_d(double x) { return log(x); }
static __device__ __noinline__ double gsm_exp_slow_d(double x) { return exp(x); }
#endif
GSM_HD double gsm_log_slow(double x) {
#if defined(__CUDA_ARCH__)
    return gsm_log_slow_d(x);
#else
    return log(x);
#endif
}
GSM_HD double gsm_exp_slow(double x) {
#if defined(__CUDA_ARCH__)
    return gsm_exp_slow_d(x);
#else
    return exp(x);
#endif
}

// natural logarithm
GSM_HD double gsm_log(double x) {
    const int32_t hi = gsm_hi32(x);
    // anything but a positive normal finite number: libdevice (log(0) = -inf, log(<0) = NaN, ...)
    if ((uint32_t)(hi - 0x00100000) >= 0x7fe00000u) return gsm_log_slow(x);
    const int32_t tmp = hi - 0x3fe60000;
    const int32_t i = (tmp >> 13) & (GSM_LOG_TAB_N - 1);
    const int32_t k = tmp >> 20;  // arithmetic shift
    const double z = gsm_hilo(hi - (tmp & (int32_t)0xfff00000), gsm_lo32(x));
#if defined(__CUDA_ARCH__)
    const double2 e = __ldg(&gsm_log_tab_d[i]);
    const double invc = e.x, logc = e.y;
    const double* __restrict__ c = gsm_log_poly_d;
    const double ln2hi = gsm_fm_consts_d[0], ln2lo = gsm_fm_consts_d[1], kbias = gsm_fm_consts_d[6];
#else
    const double invc = gsm_log_tab_h[i].invc, logc = gsm_log_tab_h[i].logc;
    const double* c = gsm_log_poly_h;
    const double ln2hi = GSM_LN2_HI, ln2lo = GSM_LN2_LO, kbias = 4503599627371520.0;
#endif
    const double r = fma(z, invc, -1.0);
    const double kd = gsm_hilo(0x43300000, k + 1024) - kbias;  // (double)k without a conversion instruction
    const double hi_part = fma(kd, ln2hi, logc);
    const double r2 = r * r;
    double p = c[6];
    p = fma(p, r, c[5]);
    p = fma(p, r, c[4]);
    p = fma(p, r, c[3]);
    p = fma(p, r, c[2]);
    p = fma(p, r, c[1]);
    p = fma(p, r, c[0]);
    const double y = fma(p, r2, r);
    return hi_part + fma(kd, ln2lo, y);
}

// exponential
GSM_HD double gsm_exp(double x) {
    const uint32_t ahi = (uint32_t)gsm_hi32(x) & 0x7fffffffu;
    if (ahi >= 0x40860000u) return gsm_exp_slow(x);  // |x| >= 704, inf, NaN: libdevice handles overflow / underflow
#if defined(__CUDA_ARCH__)
    const double invln2n = gsm_fm_consts_d[2], nhi = gsm_fm_consts_d[3], nlo = gsm_fm_consts_d[4], shift = gsm_fm_consts_d[5];
    const double* __restrict__ c = gsm_exp_poly_d;
#else
    const double invln2n = GSM_EXP_INVLN2N, nhi = GSM_EXP_NEGLN2HIN, nlo = GSM_EXP_NEGLN2LON, shift = 0x1.8p52;
    const double* c = gsm_exp_poly_h;
#endif
    double kd = fma(x, invln2n, shift);
    const int32_t ki = gsm_lo32(kd);
    kd -= shift;
    double r = fma(kd, nhi, x);
    r = fma(kd, nlo, r);
#if defined(__CUDA_ARCH__)
    const unsigned long long t = __ldg(&gsm_exp_tab_d[ki & (GSM_EXP_TAB_N - 1)]);
#else
    const uint64_t t = gsm_exp_tab_h[ki & (GSM_EXP_TAB_N - 1)];
#endif
    const double scale = gsm_hilo((int32_t)(t >> 32) + (ki << 13), (int32_t)(t & 0xffffffffu));
    const double r2 = r * r;
    const double a = fma(r, c[1], c[0]);
    const double b = fma(r, c[3], c[2]);
    double tmp = fma(r2, a, r);
    tmp = fma(r2 * r2, b, tmp);
    return fma(scale, tmp, scale);
}
