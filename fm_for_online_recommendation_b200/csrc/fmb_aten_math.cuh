// fmb_aten_math.cuh -- ATen mirrors (torch 2.11 CPU): what the reference's torch.sigmoid /
// F.binary_cross_entropy_with_logits / torch.optim.Adam calls execute, restated operation by operation.
// Reference call sites: models/models_online_deep/fm_adam.py:60-68 (Adam), :65,80 (BCE-with-logits), :80,86
// (sigmoid) and the same lines of the other four classes.  oracle/oracle_math.h holds the same functions
// in C together with how each was identified and pinned (Sleef 3.6 xexpf/xlog1pf, glibc 2.39 expf, MKL
// vsSqrt); tests compare the two implementations bit for bit, and the oracle against torch itself.
//
// Why bits matter: the reference's update is Adam's first step, an eps-damped SIGN step; a last-ulp
// difference in delta or in the denominator flips coordinates whose gradient nearly cancels, and the
// trajectories separate (measured in round 1: 3 % of the weights beyond 1e-5 after 10 000 steps).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fmb {

__device__ __forceinline__ float aten_pow2i(int n) { return __int_as_float((n + 127) << 23); }

// Sleef_expf*_u10: the vector body of ATen's sigmoid kernel and of log_sigmoid
__device__ __forceinline__ float expf_sleef(float d) {
    const int q = __float2int_rn(__fmul_rn(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
    const float fq = (float)q;
    float s = __fmaf_rn(fq, -0.693145751953125f, d);
    s = __fmaf_rn(fq, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
    u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
    u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
    u = __fmaf_rn(u, s, 0.166666671633720397949219f);
    u = __fmaf_rn(u, s, 0.5f);
    u = __fadd_rn(1.0f, __fmaf_rn(__fmul_rn(s, s), u, s));
    u = __fmul_rn(__fmul_rn(u, aten_pow2i(q >> 1)), aten_pow2i(q - (q >> 1)));
    if (d < -104.0f) u = 0.0f;
    if (100.0f < d) u = __int_as_float(0x7f800000);
    return u;
}

// Sleef_log1pf*_u10 (double-float arithmetic)
__device__ __forceinline__ float log1pf_sleef(float d) {
    float dp1 = __fadd_rn(d, 1.0f);
    const bool o = dp1 < 1.17549435e-38f;
    if (o) dp1 = __fmul_rn(dp1, 18446744073709551616.0f);
    int e = (int)(((uint32_t)__float_as_int(__fmul_rn(dp1, 1.0f / 0.75f)) >> 23) & 0xffu) - 0x7f;
    float t = __int_as_float((int)(0x3f800000u + ((uint32_t)(-e) << 23)));
    const float m = __fmaf_rn(d, t, __fsub_rn(t, 1.0f));
    if (o) e -= 64;
    const float fx = 0.69314718246459960938f, fy = -1.904654323148236017e-09f, fe = (float)e;
    float sx = __fmul_rn(fx, fe);
    float sy = __fmaf_rn(fy, fe, __fmaf_rn(fx, fe, -sx));
    const float dx = __fadd_rn(2.0f, m), dy = __fadd_rn(__fsub_rn(2.0f, dx), m);
    const float r = __fdiv_rn(1.0f, dx), qx = __fmul_rn(m, r), u = __fmaf_rn(r, m, -qx);
    const float v = __fmaf_rn(-dy, r, __fmaf_rn(-dx, r, 1.0f));
    const float xx = qx, xy = __fmaf_rn(qx, v, __fmaf_rn(0.0f, r, u));
    const float x2 = __fmul_rn(xx, xx);
    t = 0.3027294874e+0f;
    t = __fmaf_rn(t, x2, 0.3996108174e+0f);
    t = __fmaf_rn(t, x2, 0.6666694880e+0f);
    {
        const float bx = __fmul_rn(xx, 2.0f), by = __fmul_rn(xy, 2.0f), rr = __fadd_rn(sx, bx);
        sy = __fadd_rn(__fadd_rn(__fadd_rn(__fsub_rn(sx, rr), bx), sy), by);
        sx = rr;
    }
    {
        const float y = __fmul_rn(__fmul_rn(x2, xx), t), rr = __fadd_rn(sx, y);
        sy = __fadd_rn(__fadd_rn(__fsub_rn(sx, rr), y), sy);
        sx = rr;
    }
    float res = __fadd_rn(sx, sy);
    if (d > 1e+38f) res = __int_as_float(0x7f800000);
    if (-1.0f > d) res = __int_as_float(0x7fc00000);
    if (d == -1.0f) res = __int_as_float(0xff800000);
    if (d == 0.0f && (__float_as_int(d) < 0)) res = -0.0f;
    return res;
}

// glibc 2.39 expf (x86-64 FMA variant): the scalar tail of ATen's sigmoid kernel.  fp64 arithmetic.
static __device__ const unsigned long long fmb_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
static __device__ __noinline__ float expf_glibc(float x) {
    const uint32_t bx = (uint32_t)__float_as_int(x), ax = bx & 0x7fffffffu;
    if (ax >= 0x42b00000u) {  // |x| >= 88 or nan
        if (bx == 0xff800000u) return 0.0f;
        if (ax >= 0x7f800000u) return __fadd_rn(x, x);
        if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);
        if (x < -0x1.9fe368p6f) return 0.0f;
        if (x < -0x1.9d1d9ep6f) return __int_as_float(1);
    }
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
    const double xd = (double)x;
    double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, Shift);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, Shift);
    const double r = __fma_rn(InvLn2N, xd, -kd);
    const double s = __longlong_as_double((long long)(fmb_exp2f_tab[ki & 31] + (ki << 47)));
    z = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(z, r2), y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
}

// torch.sigmoid of element idx of a contiguous fp32 tensor of n (< 32768) elements: AVX-512 kernel, two
// 16-lane vectors per iteration (Sleef), the last n % 32 elements through the scalar lambda (glibc).
#define FMB_SIGMOID_BLOCK 32
__device__ __forceinline__ float sigmoid_at(float x, int idx, int n) {
    if (idx < n - n % FMB_SIGMOID_BLOCK) return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_sleef(__fsub_rn(0.0f, x))));
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_glibc(-x)));
}

// at::log_sigmoid (every element goes through the Sleef vector path, partial vectors included)
__device__ __forceinline__ float log_sigmoid(float x) {
    return __fsub_rn(x < 0.0f ? x : 0.0f, log1pf_sleef(expf_sleef(-fabsf(x))));
}

// MKL vsSqrt as Tensor.sqrt() runs it (oracle/gen_rsqrt14_table.c): one Newton step from VRSQRT14PS, whose
// 65536 values are tabulated (128 KB, L1/L2 resident).  One ulp below the rounded root on 0.59 % of inputs.
static __device__ const unsigned short fmb_rsqrt14_tab[65536] = {
#include "rsqrt14_table.inc"
};
// normal inputs only (0x00800000 <= bits < 0x7f800000).  An exact power of four needs no special case: the table's
// first entry gives S = 0x1.fffap-1 and the Newton step rounds back to exactly 1.
__device__ __forceinline__ float sqrt_mkl_normal(float x) {
    const uint32_t b = (uint32_t)__float_as_int(x);
    const uint32_t E = b >> 23, man = b & 0x7fffffu;
    const uint32_t par = ~E & 1u;                                  // E - 127 = 2q + par
    const float xn = __int_as_float((int)(((127u + par) << 23) | man));   // [1, 4)
    const float y = __int_as_float((int)(0x3f000000u | ((uint32_t)__ldg(&fmb_rsqrt14_tab[(par << 15) | (man >> 8)]) << 7)));
    const float S = __fmul_rn(xn, y), H = __fmul_rn(0.5f, y);
    const float e = __fmaf_rn(-S, S, xn);
    const float r = __fmaf_rn(e, H, S);
    const int q = ((int)E - 127 - (int)par) >> 1;                  // even numerator: the shift is exact
    return __int_as_float(__float_as_int(r) + (q << 23));
}
__device__ __forceinline__ float sqrt_mkl(float x) {
    uint32_t b = (uint32_t)__float_as_int(x);
    if (b - 0x00800000u < 0x7f000000u) return sqrt_mkl_normal(x);
    if (x != x || b == 0x7f800000u || x == 0.0f) return x;
    if (b >> 31) return __int_as_float(0x7fc00000);
    // denormal: MKL pre-scales by an even power of two
    return __fmul_rn(sqrt_mkl_normal(__fmul_rn(x, 18446744073709551616.0f)), 2.3283064365386963e-10f);   // 2^64, 2^-32
}

}  // namespace fmb
