/*
 * verify_math.c -- exhaustive / large-sample checks behind the arithmetic claims of oracle_math.h and
 * csrc/fmb_common.cuh (TEST INFRASTRUCTURE ONLY; run by hand, minutes of CPU time):
 *
 *     gcc -O2 -ffp-contract=off -mfma oracle/verify_math.c -o /tmp/verify_math -lm
 *     /tmp/verify_math expf      # orc_expf_glibc == this host's libm expf on ALL 2^32 inputs
 *     /tmp/verify_math divc      # s / c == fma(fma(-q, c, s), rc, q), q = s*rc, for all mantissas, binades 2^-104 .. 2^122
 *     /tmp/verify_math window    # the Adam window test never disagrees with the full pipeline (4e8 random triples,
 *                                # reciprocal perturbed by +-1 ulp)
 *     /tmp/verify_math qbound    # the invariant behind the window test: the pipeline's step q lies within 2^-19 (relative)
 *                                # of the cheap estimate U = (A*|g|) * rcp(|g| + c0) for EVERY binade of |g| in
 *                                # [1e-25, 1e15) (2^18 random mantissas per binade x 6 learning rates x rcp +-1 ulp)
 * Results recorded in DESIGN.md section 4 (round 2, build container: glibc 2.39, x86-64 with FMA).
 */
#include "oracle_math.h"
#include <stdio.h>
#include <stdlib.h>

static int check_expf(void) {
    long bad = 0, tot = 0;
    for (uint64_t u = 0; u < 0x100000000ull; ++u) {
        const float x = orc_bits2f((uint32_t)u);
        if (x != x) continue;
        const float w = expf(x), m = orc_expf_glibc(x);
        ++tot;
        if (orc_f2bits(w) != orc_f2bits(m)) { if (++bad < 10) printf("x=%a libm=%a mine=%a\n", x, w, m); }
    }
    printf("expf: tested %ld inputs, %ld mismatches\n", tot, bad);
    return bad != 0;
}

static int check_divc(void) {
    const float c = 0.03162277660168381f, rc = 0x1.f9f6e6p+4f;
    long bad = 0;
    for (int e = 23; e <= 249; ++e)   /* 2^-104 .. 2^122 */
        for (uint32_t m = 0; m < (1u << 23); ++m) {
            const float s = orc_bits2f(((uint32_t)e << 23) | m);
            const float q = s * rc;
            if (orc_f2bits(s / c) != orc_f2bits(fmaf(fmaf(-q, c, s), rc, q))) ++bad;
        }
    printf("divc: %ld mismatches over 227 binades x 2^23 mantissas\n", bad);
    return bad != 0;
}

static float adam_exact(float p, float g, float lr) {
    const float bc2s = 0.03162277660168381f;
    const float m = 0.1f * g, v = (0.001f * g) * g, d = (orc_sqrt_mkl(v) / bc2s) + 1e-8f, a = -(lr / 0.1f);
    return p + ((a * m) / d);
}
static int adam_window(float p, float g, float lr, int pert, float* out) {
    const float a = -(lr / 0.1f), ag = fabsf(g);
    if (!(ag >= 7.62939453125e-06f && ag < 1e15f)) return 0;
    float rc = 1.0f / ag;
    rc = orc_bits2f(orc_f2bits(rc) + pert);               /* rcp.approx: up to 1 ulp off */
    const float tau = 9.99999905e-09f * rc, r1 = fmaf(tau, tau, -tau), A = a * 0.099999994f;
    float U = fmaf(A, r1, A);
    if (g < 0.f) U = -U;
    const float ra = p + fmaf(U, 1.9073486328125e-06f, U), rb = p + fmaf(-U, 1.9073486328125e-06f, U);
    if (ra == rb) { *out = ra; return 1; }
    return 0;
}
/* general window (round 2, second form): one reciprocal of (|g| + c0), valid for every tau */
static float adam_U(float g, float lr, int pert) {
    const float a = -(lr / 0.1f), ag = fabsf(g);
    const float den = ag + 9.99999905e-09f;
    float rc = 1.0f / den;
    rc = orc_bits2f(orc_f2bits(rc) + pert);
    const float A = a * 0.099999994f;
    float U = (A * ag) * rc;
    return g < 0.f ? -U : U;
}
static int adam_window2(float p, float g, float lr, int pert, float* out) {
    const float ag = fabsf(g);
    if (!(ag >= 1e-25f && ag < 1e15f)) return 0;
    const float U = adam_U(g, lr, pert);
    const float ra = p + fmaf(U, 1.9073486328125e-06f, U), rb = p + fmaf(-U, 1.9073486328125e-06f, U);
    if (ra == rb) { *out = ra; return 1; }
    return 0;
}
static int check_qbound(void) {
    const float lrs[] = {1e-5f, 1e-4f, 1e-3f, 1e-2f, 0.05f, 3e-3f};
    const float bc2s = 0.03162277660168381f;
    double worst = 0; long n = 0, bad = 0;
    srand48(7);
    for (int e = -84; e < 50; ++e) {          /* 2^-84 = 5e-26 .. 2^50 = 1.1e15 */
        double wb = 0;
        for (int it = 0; it < (1 << 18); ++it) {
            float g = (float)ldexp(1.0 + drand48(), e);
            if (!(g >= 1e-25f && g < 1e15f)) continue;
            if (it & 1) g = -g;
            for (int l = 0; l < 6; ++l) {
                const float lr = lrs[l], a = -(lr / 0.1f);
                const float m = 0.1f * g, v = (0.001f * g) * g, d = (orc_sqrt_mkl(v) / bc2s) + 1e-8f;
                const float q = (a * m) / d;
                for (int pert = -1; pert <= 1; ++pert) {
                    const float U = adam_U(g, lr, pert);
                    const double rel = fabs((double)q - (double)U) / fabs((double)U) * 16777216.0;   /* units of 2^-24 */
                    if (rel > wb) wb = rel;
                    ++n;
                    if (!(rel < 32.0)) ++bad;
                }
            }
        }
        if (wb > worst) worst = wb;
        if (wb > 14.0) printf("binade 2^%d: worst |q-U|/|U| = %.2f * 2^-24\n", e, wb);
    }
    printf("qbound: %ld cases, worst %.2f * 2^-24 (window half-width 32 * 2^-24), %ld outside\n", n, worst, bad);
    return bad != 0;
}

static int check_window(void) {
    const float lrs[] = {1e-4f, 1e-3f, 1e-2f, 0.05f, 3e-3f};
    long n = 0, fast = 0, bad = 0;
    srand48(1);
    for (long it = 0; it < 400000000L; ++it) {
        const float lr = lrs[it % 5];
        float g = (float)(exp((drand48() * (it & 1 ? 60 : 24) - (it & 1 ? 53 : 17)) * 0.6931471805599453) * (0.5 + drand48()));
        if (lrand48() & 1) g = -g;
        const float p = (float)((drand48() * 2 - 1) * exp((drand48() * 12 - 10) * 0.6931471805599453) * 4);
        const float ex = adam_exact(p, g, lr);
        ++n;
        for (int pert = -1; pert <= 1; ++pert) {
            float o;
            if ((it & 1 ? adam_window2 : adam_window)(p, g, lr, pert, &o)) {
                if (pert == 0) ++fast;
                if (orc_f2bits(o) != orc_f2bits(ex)) { if (++bad < 10) printf("BAD p=%a g=%a lr=%g\n", p, g, lr); }
            }
        }
    }
    printf("window: %ld triples, fast path taken on %.1f %%, %ld disagreements\n", n, 100.0 * fast / n, bad);
    return bad != 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: verify_math expf|divc|window|qbound\n"); return 2; }
    if (argv[1][0] == 'e') return check_expf();
    if (argv[1][0] == 'd') return check_divc();
    if (argv[1][0] == 'q') return check_qbound();
    return check_window();
}
