/*
 * oracle_math.h -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * fp32 transcendental functions of the CPU oracle, in two groups.
 *
 * (1) ATen mirrors: what torch 2.11 (CPU, x86-64) executes for the reference's
 *     torch.sigmoid / F.binary_cross_entropy_with_logits calls
 *     (models/models_online_deep/fm_adam.py:65,80,86 and the same lines of the other four classes),
 *     restated operation by operation:
 *       - torch.sigmoid (UnaryOpsKernel.cpp sigmoid_kernel, cpu_kernel_vec): the vector body is
 *         1/(1 + Sleef_expf{8,16}_u10(0 - x)); the scalar tail is 1/(1 + expf(-x)) with glibc's expf.
 *         The kernel is built for AVX-512 as well (ALSO_REGISTER_AVX512_DISPATCH), so on the build
 *         container's CPU the body covers the first n - n % 32 elements (two 16-lane vectors per
 *         iteration) of a contiguous tensor below the 32768-element parallel grain.
 *       - at::log_sigmoid (Activation.cpp log_sigmoid_cpu_kernel): min(x,0) - Sleef_log1pf_u10(
 *         Sleef_expf_u10(-|x|)) on every element (the tail is a partial vector load).
 *     orc_expf_sleef / orc_log1pf_sleef restate Sleef 3.6's published algorithms (FMA build);
 *     orc_expf_glibc restates glibc 2.39's sysdeps/ieee754/flt-32/e_expf.c as its x86-64 FMA ifunc
 *     variant computes it.  Pinned: orc_expf_glibc equals this image's libm expf on ALL 2^32 inputs
 *     (exhaustive run, oracle/verify_math.c) and the sigmoid / log_sigmoid mirrors equal torch on
 *     millions of inputs (tests/test_oracle_math.py, which also runs on the GPU box: same image).
 * (2) Portable Cephes-style expf/logf (<= 2 ulp), used only where the reference's result feeds a
 *     continuous (not sign-step) update: nn.BCELoss values and torch.pow in the hedge step
 *     (deepfm_onn.py:117-120,147-150).  glibc's log1pf/logf/powf are not restated; the distance is
 *     ~1e-7 relative and never amplified (tests/test_oracle_vs_golden.py).
 *
 * Everything is written with explicit fmaf()/fma() and plain IEEE-754 + - * / so that a gcc build
 * with -ffp-contract=off produces the same bits on every host.  The CUDA product code
 * (fm_for_online_recommendation_b200/csrc/fmb_common.cuh) restates the same algorithms with
 * __fmaf_rn/__fmul_rn/__fadd_rn/__fma_rn; parity tests compare the two bit for bit.
 */
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float orc_bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t orc_f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* 2^n for n in [-126, 127] */
static inline float orc_pow2i(int n) { return orc_bits2f((uint32_t)(n + 127) << 23); }

/* ---- (2) portable functions ---- */
static inline float orc_expf(float x) {
    if (x != x) return x;
    if (x > 88.7228317f) return INFINITY;
    if (x < -103.972084f) return 0.0f;
    float fn = rintf(x * 1.44269504f);
    float r = fmaf(fn, -0.693359375f, x);
    r = fmaf(fn, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r);
    y = y + 1.0f;
    int n = (int)fn;
    int n1 = n / 2;
    int n2 = n - n1;
    y = y * orc_pow2i(n1);
    y = y * orc_pow2i(n2);
    return y;
}

static inline float orc_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return -INFINITY;
    if (x == INFINITY) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t b = orc_f2bits(x);
    e += (int)((b >> 23) & 0xffu) - 126;
    float m = orc_bits2f((b & 0x807fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < 0.707106781f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float)e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(0.693359375f, fe, r);
    return r;
}

/* log(1+u) for u >= 0 (used with u = exp(-|z|) in (0, 1]) */
static inline float orc_log1pf(float u) {
    float w = 1.0f + u;
    if (w == 1.0f) return u;
    float l = orc_logf(w);
    float c = ((w - 1.0f) - u) / w;
    return l - c;
}


/* b^e for b > 0 */
static inline float orc_powf(float b, float e) { return orc_expf(e * orc_logf(b)); }


/* ------------------------------------------------------------------------------------------ */
/* (1) ATen mirrors                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* Sleef_expf*_u10 (sleef 3.6 src/libm/sleefsimdsp.c xexpf, FMA helpers) */
static inline float orc_expf_sleef(float d) {
    const int q = (int)rintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
    float s, u;
    s = fmaf((float)q, -0.693145751953125f, d);
    s = fmaf((float)q, -1.428606765330187045e-06f, s);
    u = 0.000198527617612853646278381f;
    u = fmaf(u, s, 0.00139304355252534151077271f);
    u = fmaf(u, s, 0.00833336077630519866943359f);
    u = fmaf(u, s, 0.0416664853692054748535156f);
    u = fmaf(u, s, 0.166666671633720397949219f);
    u = fmaf(u, s, 0.5f);
    u = 1.0f + fmaf(s * s, u, s);
    u = (u * orc_pow2i(q >> 1)) * orc_pow2i(q - (q >> 1)); /* vldexp2 */
    if (d < -104.0f) u = 0.0f;
    if (100.0f < d) u = INFINITY;
    return u;
}

/* Sleef_log1pf*_u10 (xlog1pf, non-AVX512 helper path, double-float arithmetic with FMA) */
static inline float orc_log1pf_sleef(float d) {
    float dp1 = d + 1.0f;
    const int o = dp1 < 1.17549435e-38f;
    if (o) dp1 = dp1 * (4294967296.0f * 4294967296.0f);
    int e = (int)((orc_f2bits(dp1 * (1.0f / 0.75f)) >> 23) & 0xffu) - 0x7f;      /* vilogb2k */
    float t = orc_bits2f(orc_f2bits(1.0f) + ((uint32_t)(-e) << 23));            /* vldexp3(1, -e) */
    const float m = fmaf(d, t, t - 1.0f);
    if (o) e -= 64;
    float sx, sy, xx, xy;
    { /* s = (ln2_hi, ln2_lo) * e */
        const float fx = 0.69314718246459960938f, fy = -1.904654323148236017e-09f, y = (float)e;
        sx = fx * y;
        sy = fmaf(fy, y, fmaf(fx, y, -sx));
    }
    { /* x = (m, 0) / (2 + m) */
        const float dx = 2.0f + m, dy = (2.0f - dx) + m;
        const float r = 1.0f / dx, q = m * r, u = fmaf(r, m, -q);
        const float v = fmaf(-dy, r, fmaf(-dx, r, 1.0f));
        xx = q;
        xy = fmaf(q, v, fmaf(0.0f, r, u));
    }
    const float x2 = xx * xx;
    t = 0.3027294874e+0f;
    t = fmaf(t, x2, 0.3996108174e+0f);
    t = fmaf(t, x2, 0.6666694880e+0f);
    { /* s += 2x */
        const float bx = xx * 2.0f, by = xy * 2.0f, r = sx + bx;
        sy = (((sx - r) + bx) + sy) + by;
        sx = r;
    }
    { /* s += x^3 t */
        const float y = (x2 * xx) * t, r = sx + y;
        sy = ((sx - r) + y) + sy;
        sx = r;
    }
    float r = sx + sy;
    if (d > 1e+38f) r = INFINITY;
    if (-1.0f > d) r = NAN;
    if (d == -1.0f) r = -INFINITY;
    if (d == 0.0f && signbit(d)) r = -0.0f;
    return r;
}

/* glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c + e_exp2f_data.c, N = 32), x86-64 FMA variant:
 * gcc contracts r = InvLn2N*xd - kd into one fma there; the other products round separately or fused
 * without changing any result (all 16 combinations checked exhaustively, oracle/verify_math.c). */
static const uint64_t orc_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
static inline double orc_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t orc_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline float orc_expf_glibc(float x) {
    const uint32_t ax = orc_f2bits(x) & 0x7fffffffu;
    if (ax >= 0x42b00000u) { /* |x| >= 88 or nan */
        if (orc_f2bits(x) == 0xff800000u) return 0.0f;
        if (ax >= 0x7f800000u) return x + x;
        if (x > 0x1.62e42ep6f) return INFINITY;         /* overflow */
        if (x < -0x1.9fe368p6f) return 0.0f;            /* underflow */
        if (x < -0x1.9d1d9ep6f) return orc_bits2f(1u);  /* __math_may_uflowf: 0x1.4p-75f squared */
    }
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
                 C2 = 0x1.62e42ff0c52d6p-1 / 32;
    const double xd = (double)x;
    double z = InvLn2N * xd;
    double kd = z + Shift;
    const uint64_t ki = orc_d2u(kd);
    kd = kd - Shift;
    const double r = fma(InvLn2N, xd, -kd);
    const double s = orc_u2d(orc_exp2f_tab[ki % 32] + (ki << 47));
    z = C0 * r + C1;
    const double r2 = r * r;
    double y = C2 * r + 1.0;
    y = z * r2 + y;
    y = y * s;
    return (float)y;
}

/* The vector width of the ATen sigmoid kernel on the host that generated tests/golden (AVX-512: two
 * 16-lane vectors per loop iteration).  An AVX2-only host would use 16. */
#define ORC_SIGMOID_BLOCK 32
/* torch.sigmoid of element `idx` of a contiguous fp32 tensor of `n` (< 32768) elements */
static inline float orc_sigmoid_at(float x, int64_t idx, int64_t n) {
    if (idx < n - n % ORC_SIGMOID_BLOCK) {
        float a = 0.0f - x;
        a = orc_expf_sleef(a);
        a = 1.0f + a;
        return 1.0f / a;
    }
    return 1.0f / (1.0f + orc_expf_glibc(-x));
}
/* at::log_sigmoid of one element (position-independent) */
static inline float orc_log_sigmoid(float x) {
    const float mn = x < 0.0f ? x : 0.0f; /* vec::minimum(x, 0) (propagates nan; not reachable here) */
    return mn - orc_log1pf_sleef(orc_expf_sleef(-fabsf(x)));
}


/* MKL vsSqrt (VML_HA) as torch 2.11 calls it for Tensor.sqrt() on fp32 (see gen_rsqrt14_table.c): one
 * Newton step from VRSQRT14PS, scale-invariant in the exponent (verified for every normal exponent),
 * denormal inputs pre-scaled by an even power of two.  Equals torch.sqrt bit for bit on every
 * non-negative finite fp32 input (tests/test_oracle_math.py: all 2^24 mantissa/parity combinations,
 * all denormals).  0.59 % of inputs come out one ulp below the correctly rounded root. */
static const uint16_t orc_rsqrt14_tab[65536] = {
#include "rsqrt14_table.inc"
};
static inline float orc_sqrt_mkl(float x) {
    uint32_t b = orc_f2bits(x);
    if (x != x || b == 0x7f800000u || x == 0.0f) return x; /* nan, +inf, +-0 */
    if (b >> 31) return NAN;
    int q = 0;
    if (b < 0x00800000u) { x = x * 18446744073709551616.0f; b = orc_f2bits(x); q = -32; } /* 2^64 */
    const int E = (int)(b >> 23), p = (E + 1) & 1;  /* E - 127 = 2*qq + p */
    q += (E - 127 - p) / 2;
    const uint32_t man = b & 0x7fffffu;
    const float xn = orc_bits2f(((uint32_t)(127 + p) << 23) | man);            /* [1, 4) */
    const float y = (p == 0 && man == 0) ? 1.0f
                  : orc_bits2f((126u << 23) | ((uint32_t)orc_rsqrt14_tab[((uint32_t)p << 15) | (man >> 8)] << 7));
    const float S = xn * y, H = 0.5f * y;
    const float e = fmaf(-S, S, xn);
    const float r = fmaf(e, H, S);                                             /* [1, 2] */
    return orc_bits2f(orc_f2bits(r) + ((uint32_t)q << 23));                    /* r * 2^q, exact */
}

#endif
