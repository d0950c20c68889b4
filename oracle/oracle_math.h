/*
 * oracle_math.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Portable fp32 transcendental functions used by the CPU oracle.  They are
 * written with explicit fmaf() and plain IEEE-754 binary32 + - * / so that a
 * gcc build with -ffp-contract=off produces the same bits on every host.  The
 * CUDA product code (fm_for_online_recommendation_b200/csrc/fmb_math.cuh)
 * restates the same algorithms with __fmaf_rn/__fmul_rn/__fadd_rn; parity
 * tests compare the two bit for bit.
 *
 * Why not libm: the reference evaluates sigmoid/log through ATen
 * (models/models_online_deep/fm_adam.py:80,86 -> torch.sigmoid,
 * F.binary_cross_entropy_with_logits).  ATen's vectorised exp is Sleef/"u20"
 * on the vector body and glibc expf on the scalar tail, so its bits depend on
 * batch size and CPU ISA ("hard parts").  No single
 * function can match it bit for bit; the oracle therefore fixes ONE
 * well-defined expf/logf (Cephes-style, <= 2 ulp) and quantifies the distance
 * to torch in tests/test_oracle_vs_golden.py.
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

/* ATen sigmoid: 1 / (1 + exp(-x))  (UnaryOpsKernel.cpp sigmoid_kernel) */
static inline float orc_sigmoidf(float x) { return 1.0f / (1.0f + orc_expf(-x)); }

/* b^e for b > 0 */
static inline float orc_powf(float b, float e) { return orc_expf(e * orc_logf(b)); }

#endif
