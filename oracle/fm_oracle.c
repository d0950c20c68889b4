/*
 * fm_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the deep FM family of haan6/fm-for-online-recommendation
 * (models/models_online_deep/{fm_adam,deepfm_adam,nfm_adam,deepfm_onn,nfm_onn}.py)
 * with every fp32 operation written in the order ATen executes it (SURVEY.md
 * section 7 "hard parts", section 8 rows A1-A8, A12).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file; the product (fm_for_online_recommendation_b200) never does.
 *
 * Parity status: PINNED against outputs of the reference itself, generated in the
 * build container by tests/golden/make_golden.py (imports /root/reference) and
 * committed under tests/golden (npz files).  The reference ships no tests or golden
 * vectors of its own.
 *
 * What is mirrored bit for bit (torch 2.11 CPU as it runs in the build container, one thread):
 * torch.sum's accumulator cascade (orc_sum_aten), torch.sigmoid (Sleef vector body + glibc scalar
 * tail, split by position in the batch), log_sigmoid, the fresh-state Adam step.  Not mirrored:
 * MKL's sgemm order (MLP tower) and glibc logf/log1pf/powf in nn.BCELoss / torch.pow (hedge step).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (oracle/Makefile).
 */
#include "oracle_math.h"
#include <stdlib.h>

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* canonical reduction orders                                          */
/* ------------------------------------------------------------------ */

/* ATen's CPU sum of a contiguous fp32 row on x86 (aten/src/ATen/native/cpu/SumKernel.cpp, the
 * AVX2 build that torch dispatches to even on AVX-512 hosts: Vectorized<float>::size() == 8).
 *   n >= 8: vectorized_inner_sum -- 4 (ilp) x 8 (lanes) accumulators walk the row 32 floats at a
 *           time through multi_row_sum's cascade, left-over vectors go to accumulator 0, the four
 *           accumulators are folded 0+1+2+3, then the scalar tail (n % 8) is summed from 0 and the
 *           eight lanes are added to it in lane order.
 *   n <  8: scalar_inner_sum -- row_sum with 4 scalar accumulators (ilp), same folding.
 * Verified bit-exact against torch.sum on every golden fixture (tests/test_oracle_vs_golden.py).
 * This is the single reduction order used for Sum_f first[b,f], Sum_j bi[b,j], Sum_j x_l[b,j],
 * loss.mean(), the bias gradient and Sum(alpha). */
/* ------------------------------------------------------------------ */
/* host threads for the bench's CPU baseline                            */
/* ------------------------------------------------------------------ */
/* ORC_THREADS (default 1) host threads share the batch step: samples are independent in the forward pass and in
 * the loss, fields are independent in the backward pass (see orc_fm_backward_update).  The arithmetic and the
 * order of every sample / row are untouched, so results do not depend on the thread count (tests run both). */
#include <pthread.h>
typedef void (*orc_body)(int tid, int nthreads, void* ctx);
typedef struct { orc_body fn; void* ctx; int tid, nthreads; } orc_task;
static void* orc_tramp(void* a) { orc_task* t = (orc_task*)a; t->fn(t->tid, t->nthreads, t->ctx); return NULL; }
static int orc_threads(void) {
    const char* e = getenv("ORC_THREADS");
    int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : (n > 256 ? 256 : n);
}
static void orc_parallel(int work_items, orc_body fn, void* ctx) {
    int n = orc_threads();
    if (work_items < 256 && work_items < 8 * n) n = 1;   /* tiny problems: not worth the thread start-up */
    if (n == 1) { fn(0, 1, ctx); return; }
    pthread_t th[256];
    orc_task tk[256];
    for (int t = 0; t < n; ++t) { tk[t].fn = fn; tk[t].ctx = ctx; tk[t].tid = t; tk[t].nthreads = n; }
    for (int t = 1; t < n; ++t) pthread_create(&th[t], NULL, orc_tramp, &tk[t]);
    fn(0, n, ctx);
    for (int t = 1; t < n; ++t) pthread_join(th[t], NULL);
}

static int ceil_log2_i(int64_t x) {
    int r = 0;
    while (((int64_t)1 << r) < x) ++r;
    return r;
}
/* multi_row_sum: `size` rows of A floats (row i at x + i*A), accumulated with the 4-level cascade */
static void multi_row_sum(const float* x, int64_t size, int A, float* out /*[A]*/) {
    float acc[4][32];
    for (int j = 0; j < 4; ++j) for (int a = 0; a < A; ++a) acc[j][a] = 0.f;
    int lp = ceil_log2_i(size) / 4;
    const int level_power = lp > 4 ? lp : 4;
    const int64_t level_step = (int64_t)1 << level_power, level_mask = level_step - 1;
    int64_t i = 0;
    while (i + level_step <= size) {
        for (int64_t j = 0; j < level_step; ++j, ++i)
            for (int a = 0; a < A; ++a) acc[0][a] = acc[0][a] + x[i * A + a];
        for (int j = 1; j < 4; ++j) {
            for (int a = 0; a < A; ++a) { acc[j][a] = acc[j][a] + acc[j - 1][a]; acc[j - 1][a] = 0.f; }
            const int64_t mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size; ++i)
        for (int a = 0; a < A; ++a) acc[0][a] = acc[0][a] + x[i * A + a];
    for (int j = 1; j < 4; ++j)
        for (int a = 0; a < A; ++a) acc[0][a] = acc[0][a] + acc[j][a];
    for (int a = 0; a < A; ++a) out[a] = acc[0][a];
}
API float orc_sum_aten(const float* x, int64_t n) {
    float p[32];
    if (n < 8) {
        const int64_t size_ilp = n / 4;
        multi_row_sum(x, size_ilp, 4, p);
        for (int64_t i = size_ilp * 4; i < n; ++i) p[0] = p[0] + x[i];
        for (int t = 1; t < 4; ++t) p[0] = p[0] + p[t];
        return p[0];
    }
    const int64_t vec_size = n / 8, size_ilp = vec_size / 4;
    multi_row_sum(x, size_ilp, 32, p);
    for (int64_t v = size_ilp * 4; v < vec_size; ++v)
        for (int l = 0; l < 8; ++l) p[l] = p[l] + x[v * 8 + l];
    for (int t = 1; t < 4; ++t)
        for (int l = 0; l < 8; ++l) p[l] = p[l] + p[t * 8 + l];
    float fin = 0.f;
    for (int64_t i = vec_size * 8; i < n; ++i) fin = fin + x[i];
    for (int l = 0; l < 8; ++l) fin = fin + p[l];
    return fin;
}

/* ------------------------------------------------------------------ */
/* model                                                               */
/* ------------------------------------------------------------------ */

enum { K_FM = 0, K_DEEPFM = 1, K_NFM = 2, K_DEEPFM_ONN = 3, K_NFM_ONN = 4 };

typedef struct {
    int32_t kind, F, k, L, H, R;
    int32_t batch_size; /* ONN ctor arg (deepfm_onn.py:14) */
    int32_t update_mode; /* 0 = fresh-Adam sign step (reference), 1 = plain SGD */
    float* w1;    /* [R]      first_order_embeddings, fields concatenated */
    float* V;     /* [R*k]    second_order_embeddings */
    float* mlp;   /* W0[H,k] c0[H] W1[H,H] c1[H] ... */
    float* bias;  /* [1] */
    float* alpha; /* [L] (ONN) */
    float lr, hb, hs;
    /* zero-initialised scratch owned by the caller */
    float* gA;        /* [R*(k+1)] */
    float* gB;        /* [R*(k+1)] */
    uint8_t* touched; /* [R] */
    int32_t shard_G;  /* > 1: forward sums in the row-sharded multi-GPU order (owner-major), see
                         fm_for_online_recommendation_b200/csrc/sharded.cu; 0/1: reference order */
    int32_t rank_B;   /* > 0: the batch is the concatenation of per-rank batches of rank_B samples and duplicate rows
                         are summed in RANK-PARTIAL order (csrc/shard2.cu): every rank sums its own entries of a
                         row in sample order, the partial sums are added in rank order.  0: reference order
                         (one chain over all samples).  Identical to the reference when a row is hit by one rank. */
    /* update_mode 2 (per-coordinate FTRL-Proximal): state, zero-initialised by the caller */
    float *fz_V, *fn_V;   /* [R*k] */
    float *fz_w1, *fn_w1; /* [R]   */
    float* f_bias;        /* [2] z, n of the bias */
    float f_beta, f_l1, f_l2;
} orc_model;

static size_t mlp_w_off(const orc_model* m, int l) {
    if (l == 0) return 0;
    return (size_t)m->H * m->k + m->H + (size_t)(l - 1) * ((size_t)m->H * m->H + m->H);
}
static int mlp_in(const orc_model* m, int l) { return l == 0 ? m->k : m->H; }
static size_t mlp_c_off(const orc_model* m, int l) { return mlp_w_off(m, l) + (size_t)m->H * mlp_in(m, l); }
API int64_t orc_mlp_numel(const orc_model* m) { return m->L > 0 ? (int64_t)mlp_w_off(m, m->L) : 0; }

/* A1-A3: fm_adam.py:34-54, deepfm_adam.py:46-77 */
typedef struct {
    const orc_model* m; const int32_t* ids; const float* xv; int B;
    float *first, *S, *bi, *sum_first, *sum_bi, *z_fm;
} fwd_ctx;
static void fm_forward_slice(int tid, int nthreads, void* vctx) {
    const fwd_ctx* c = (const fwd_ctx*)vctx;
    const orc_model* m = c->m; const int32_t* ids = c->ids; const float* xv = c->xv; const int B = c->B;
    float *first = c->first, *S = c->S, *bi = c->bi, *sum_first = c->sum_first, *sum_bi = c->sum_bi, *z_fm = c->z_fm;
    const int b_lo = (int)((int64_t)B * tid / nthreads), b_hi = (int)((int64_t)B * (tid + 1) / nthreads);
    const int F = m->F, k = m->k;
    float* Sj = (float*)malloc(sizeof(float) * k);
    float* Qj = (float*)malloc(sizeof(float) * k);
    float* fo = (float*)malloc(sizeof(float) * F);
    float* bj = (float*)malloc(sizeof(float) * k);
    for (int b = b_lo; b < b_hi; ++b) {
        for (int j = 0; j < k; ++j) { Sj[j] = 0.f; Qj[j] = 0.f; }
        const int G = m->shard_G > 1 ? m->shard_G : 1;
        float sf_sharded = 0.f;
        for (int o = 0; o < G; ++o) {
            /* G == 1: the reference's left-to-right sum over fields.  G > 1: owner o first sums the rows
             * it holds (row r lives on rank r % G) in field order, then the owners' partials are added
             * in owner order -- the order the multi-GPU path uses. */
            float pS[256], pQ[256], pf = 0.f;
            for (int j = 0; j < k; ++j) { pS[j] = 0.f; pQ[j] = 0.f; }
            for (int f = 0; f < F; ++f) {
                const int32_t r = ids[(size_t)b * F + f];
                if (G > 1 && r % G != o) continue;
                const float x = xv[(size_t)b * F + f];
                fo[f] = m->w1[r] * x; /* deepfm_adam.py:50 */
                pf = pf + fo[f];
                const float* v = m->V + (size_t)r * k;
                for (int j = 0; j < k; ++j) {
                    float e = v[j] * x;     /* deepfm_adam.py:60 */
                    pS[j] = pS[j] + e;      /* python sum(), left to right, :62 */
                    float sq = e * e;       /* :66 */
                    pQ[j] = pQ[j] + sq;     /* :67 */
                }
            }
            if (G > 1) {
                for (int j = 0; j < k; ++j) { Sj[j] = Sj[j] + pS[j]; Qj[j] = Qj[j] + pQ[j]; }
                sf_sharded = sf_sharded + pf;
            } else {
                for (int j = 0; j < k; ++j) { Sj[j] = pS[j]; Qj[j] = pQ[j]; }
            }
        }
        for (int j = 0; j < k; ++j) bj[j] = ((Sj[j] * Sj[j]) - Qj[j]) * 0.5f; /* :68 */
        float sf = G > 1 ? sf_sharded : orc_sum_aten(fo, F);
        float sb = orc_sum_aten(bj, k);
        if (first) for (int f = 0; f < F; ++f) first[(size_t)b * F + f] = fo[f];
        if (S) for (int j = 0; j < k; ++j) S[(size_t)b * k + j] = Sj[j];
        if (bi) for (int j = 0; j < k; ++j) bi[(size_t)b * k + j] = bj[j];
        if (sum_first) sum_first[b] = sf;
        if (sum_bi) sum_bi[b] = sb;
        if (z_fm) z_fm[b] = (sf + sb) + m->bias[0]; /* :75 */
    }
    free(Sj); free(Qj); free(fo); free(bj);
}

API void orc_fm_forward(const orc_model* m, const int32_t* ids, const float* xv, int B,
                        float* first, float* S, float* bi, float* sum_first, float* sum_bi, float* z_fm) {
    fwd_ctx c = {m, ids, xv, B, first, S, bi, sum_first, sum_bi, z_fm};
    orc_parallel(B, fm_forward_slice, &c);
}

/* A4: deepfm_adam.py:82-86.  act is [L][B][H]; head[l*B+b] = Sum_j act[l][b][j] (ATen row order). */
API void orc_mlp_forward(const orc_model* m, const float* bi, int B, float* act, float* head) {
    const int H = m->H;
    for (int l = 0; l < m->L; ++l) {
        const int nin = mlp_in(m, l);
        const float* W = m->mlp + mlp_w_off(m, l);
        const float* c = m->mlp + mlp_c_off(m, l);
        const float* xin = l == 0 ? bi : act + (size_t)(l - 1) * B * H;
        float* out = act + (size_t)l * B * H;
        for (int b = 0; b < B; ++b) {
            for (int o = 0; o < H; ++o) {
                float a = 0.f;
                for (int i = 0; i < nin; ++i) a = fmaf(xin[(size_t)b * nin + i], W[(size_t)o * nin + i], a);
                a = a + c[o];
                out[(size_t)b * H + o] = a > 0.f ? a : 0.f;
            }
            head[(size_t)l * B + b] = orc_sum_aten(out + (size_t)b * H, H);
        }
    }
}

/* backward of head `top` (grad gtop[b] on Sum_j act[top][b][j]) through layers top..0.
 * gmlp has the layout of m->mlp and is fully overwritten (zeros above `top`). gbi may be NULL. */
API void orc_mlp_backward(const orc_model* m, const float* bi, const float* act, const float* gtop, int top,
                          int B, float* gmlp, float* gbi) {
    const int H = m->H;
    const int64_t n = orc_mlp_numel(m);
    for (int64_t i = 0; i < n; ++i) gmlp[i] = 0.f;
    float* g = (float*)malloc(sizeof(float) * (size_t)B * H);  /* grad wrt act[l] */
    float* gp = (float*)malloc(sizeof(float) * (size_t)B * H); /* grad wrt pre-activation */
    for (int b = 0; b < B; ++b)
        for (int o = 0; o < H; ++o) g[(size_t)b * H + o] = gtop[b];
    for (int l = top; l >= 0; --l) {
        const int nin = mlp_in(m, l);
        const float* W = m->mlp + mlp_w_off(m, l);
        const float* xin = l == 0 ? bi : act + (size_t)(l - 1) * B * H;
        const float* out = act + (size_t)l * B * H;
        float* gW = gmlp + mlp_w_off(m, l);
        float* gc = gmlp + mlp_c_off(m, l);
        for (size_t i = 0; i < (size_t)B * H; ++i) gp[i] = out[i] > 0.f ? g[i] : 0.f; /* threshold_backward */
        for (int o = 0; o < H; ++o) {
            for (int i = 0; i < nin; ++i) {
                float a = 0.f;
                for (int b = 0; b < B; ++b) a = fmaf(gp[(size_t)b * H + o], xin[(size_t)b * nin + i], a);
                gW[(size_t)o * nin + i] = a;
            }
            float a = 0.f;
            for (int b = 0; b < B; ++b) a = a + gp[(size_t)b * H + o];
            gc[o] = a;
        }
        if (l > 0 || gbi) {
            float* gx = l > 0 ? g : gbi;
            /* note: g (layer l-1 grad) has row length nin == H for l > 0 */
            float* tmp = (float*)malloc(sizeof(float) * (size_t)B * nin);
            for (int b = 0; b < B; ++b)
                for (int i = 0; i < nin; ++i) {
                    float a = 0.f;
                    for (int o = 0; o < H; ++o) a = fmaf(gp[(size_t)b * H + o], W[(size_t)o * nin + i], a);
                    tmp[(size_t)b * nin + i] = a;
                }
            for (size_t i = 0; i < (size_t)B * nin; ++i) gx[i] = tmp[i];
            free(tmp);
        }
    }
    free(g); free(gp);
}

/* F.binary_cross_entropy_with_logits value + gradient (A6).
 * kind 0: loss(z)           delta = ((sigmoid(z) - y) / B)
 * kind 1: loss(sigmoid(z))  delta = (((sigmoid(p) - y) / B) * (1 - p)) * p,  p = sigmoid(z)
 * returns mean loss (ATen row order). */
typedef struct { int kind; const float* z; const float* y; int B; float* delta; float* lv; } loss_ctx;
static void loss_slice(int tid, int nthreads, void* vctx) {
    const loss_ctx* c = (const loss_ctx*)vctx;
    const int kind = c->kind, B = c->B;
    const float* z = c->z; const float* y = c->y; float* delta = c->delta; float* lv = c->lv;
    const float fB = (float)B;
    const int b_lo = (int)((int64_t)B * tid / nthreads), b_hi = (int)((int64_t)B * (tid + 1) / nthreads);
    for (int b = b_lo; b < b_hi; ++b) {
        float in = z[b], p = 0.f;
        if (kind == 1) { p = orc_sigmoid_at(z[b], b, B); in = p; }   /* torch.sigmoid(output), fm_adam.py:80 */
        /* (1 - y) * x - log_sigmoid(x)  (ATen Loss.cpp binary_cross_entropy_with_logits) */
        float ls = orc_log_sigmoid(in);
        lv[b] = ((1.0f - y[b]) * in) - ls;
        float d = (orc_sigmoid_at(in, b, B) - y[b]) / fB;         /* backward: (input.sigmoid() - target) * grad */
        if (kind == 1) d = (d * (1.0f - p)) * p; /* sigmoid_backward: grad * (1 - out) * out */
        delta[b] = d;
    }
}
API float orc_loss_delta(int kind, const float* z, const float* y, int B, float* delta) {
    float* lv = (float*)malloc(sizeof(float) * B);
    const float fB = (float)B;
    loss_ctx c = {kind, z, y, B, delta, lv};
    orc_parallel(B, loss_slice, &c);
    float loss = orc_sum_aten(lv, B) / fB;
    free(lv);
    return loss;
}

/* torch.optim.Adam first step with fresh state (A12): lr is an fp32 0-dim tensor. */
static inline float adam1(float p, float g, float lr) {
    const float bc2s = 0.03162277660168381f; /* float((1 - 0.999) ** 0.5) */
    float m = 0.1f * g;                      /* lerp(0, g, 1 - beta1) */
    float v = (0.001f * g) * g;              /* addcmul(value = 1 - beta2) */
    float d = (orc_sqrt_mkl(v) / bc2s) + 1e-8f;  /* exp_avg_sq.sqrt(): MKL vsSqrt, not sqrtf */
    float a = -(lr / 0.1f);                  /* -(lr / bias_correction1) */
    return p + ((a * m) / d);
}
/* mode 1: torch.optim.SGD.  param.add_(grad, alpha=-lr) is ONE fused multiply-add in ATen's CPU add kernel
 * (BinaryOpsKernel.cpp: vec::fmadd(b, alpha, a)); pinned in tests/test_oracle_math.py. */
static inline float upd(float p, float g, float lr, int mode) {
    return mode == 0 ? adam1(p, g, lr) : fmaf(g, -lr, p);
}
/* element-wise exports for tests/test_oracle_math.py */
API void orc_vec_sqrt_mkl(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_sqrt_mkl(x[i]); }
API void orc_vec_sigmoid(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_sigmoid_at(x[i], i, n); }
API void orc_vec_log_sigmoid(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_log_sigmoid(x[i]); }
API void orc_vec_expf_glibc(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_expf_glibc(x[i]); }
/* per-coordinate FTRL-Proximal (McMahan et al. 2013), every operation rounded once in this order; the CUDA side is
 * fmb::ftrl_update (csrc/fmb_common.cuh).  alpha = the learning rate. */
static inline float orc_ftrl_update(float w, float g, float* z, float* n, float alpha, float beta, float l1, float l2) {
    const float nn = *n + (g * g);
    const float sn = sqrtf(*n), snn = sqrtf(nn);
    const float sigma = (snn - sn) / alpha;
    const float zz = *z + (g - (sigma * w));
    *z = zz; *n = nn;
    if (fabsf(zz) <= l1) return 0.f;
    const float num = zz - copysignf(l1, zz);
    const float den = ((beta + snn) / alpha) + l2;
    return -(num / den);
}
API void orc_vec_ftrl(float* w, const float* g, float* z, float* nacc, int64_t n, float alpha, float beta, float l1, float l2) {
    for (int64_t i = 0; i < n; ++i) w[i] = orc_ftrl_update(w[i], g[i], z + i, nacc + i, alpha, beta, l1, l2);
}
API void orc_update_dense(float* p, const float* g, int64_t n, float lr, int mode) {
    for (int64_t i = 0; i < n; ++i) p[i] = upd(p[i], g[i], lr, mode);
}

/* A6 sparse part: embedding_dense_backward sums duplicate rows in sample order, then one step per row.
 * gs[b]: gradient on z wrt the FM logit (first-order path, and the Sum_j bi path when use_fm2).
 * gvec[b*k+j]: gradient on bi coming from the MLP (second evaluation of second_order, deepfm_adam.py:81). */
typedef struct {
    orc_model* m; const int32_t* ids; const float* xv; int B; const float* S; const float* gs; int use_fm2;
    const float* gvec; int32_t* tl; size_t* ntf; int by_field;
    float* tot; int32_t* cur_rank;   /* rank-partial order (m->rank_B > 0): folded partials of finished ranks, rank of ga */
} bwd_ctx;
/* contribution of entry (b, f) to the gradient sums of its row */
static inline void bwd_entry(const bwd_ctx* c, int b, int f, int32_t* tl, size_t* nt) {
    orc_model* m = c->m;
    const int F = m->F, k = m->k, kp = k + 1;
    const int32_t r = c->ids[(size_t)b * F + f];
    const float x = c->xv[(size_t)b * F + f];
    if (!m->touched[r]) { m->touched[r] = 1; tl[(*nt)++] = r; if (c->cur_rank) c->cur_rank[r] = -1; }
    const float* v = m->V + (size_t)r * k;
    float* ga = m->gA + (size_t)r * kp;
    float* gb = m->gB + (size_t)r * kp;
    if (c->cur_rank) {   /* a new rank starts on this row: fold the previous rank's partial into the total */
        const int rk = b / m->rank_B;
        if (c->cur_rank[r] != rk) {
            float* tt = c->tot + (size_t)r * kp;
            if (c->cur_rank[r] >= 0)
                for (int j = 0; j < kp; ++j) { tt[j] = (tt[j] != tt[j]) ? ga[j] : tt[j] + ga[j]; ga[j] = 0.f; }
            else
                for (int j = 0; j < kp; ++j) tt[j] = NAN;   /* "no partial yet" marker */
            c->cur_rank[r] = rk;
        }
    }
    ga[k] = ga[k] + (c->gs[b] * x);
    for (int j = 0; j < k; ++j) {
        float e = v[j] * x;
        float s = c->S[(size_t)b * k + j];
        if (c->use_fm2) {
            float ge = (c->gs[b] * s) - (c->gs[b] * e);
            ga[j] = ga[j] + (ge * x);
        }
        if (c->gvec) {
            float gv = c->gvec[(size_t)b * k + j];
            float ge = (gv * s) - (gv * e);
            gb[j] = gb[j] + (ge * x);
        }
    }
}
static void bwd_rows_update(const bwd_ctx* c, const int32_t* tl, size_t nt) {
    orc_model* m = c->m;
    const int k = m->k, kp = k + 1;
    for (size_t t = 0; t < nt; ++t) {
        const int32_t r = tl[t];
        float* v = m->V + (size_t)r * k;
        float* ga = m->gA + (size_t)r * kp;
        float* gb = m->gB + (size_t)r * kp;
        if (c->cur_rank) {   /* total = partial of the first rank, then + the later ranks' partials in rank order */
            const float* tt = c->tot + (size_t)r * kp;
            for (int j = 0; j < kp; ++j) if (!(tt[j] != tt[j])) ga[j] = tt[j] + ga[j];
        }
        for (int j = 0; j < k; ++j) {
            float g = (c->use_fm2 && c->gvec) ? ga[j] + gb[j] : (c->gvec ? gb[j] : ga[j]);
            if (m->update_mode == 2)
                v[j] = orc_ftrl_update(v[j], g, m->fz_V + (size_t)r * k + j, m->fn_V + (size_t)r * k + j, m->lr, m->f_beta,
                                       m->f_l1, m->f_l2);
            else
                v[j] = upd(v[j], g, m->lr, m->update_mode);
            ga[j] = 0.f; gb[j] = 0.f;
        }
        if (m->update_mode == 2)
            m->w1[r] = orc_ftrl_update(m->w1[r], ga[k], m->fz_w1 + r, m->fn_w1 + r, m->lr, m->f_beta, m->f_l1, m->f_l2);
        else
            m->w1[r] = upd(m->w1[r], ga[k], m->lr, m->update_mode);
        ga[k] = 0.f;
        m->touched[r] = 0;
    }
}
/* thread t takes fields t, t + T, ...: accumulate (samples ascending), then update that field's touched rows */
static void bwd_fields(int tid, int nthreads, void* vctx) {
    const bwd_ctx* c = (const bwd_ctx*)vctx;
    const int F = c->m->F, B = c->B;
    for (int f = tid; f < F; f += nthreads) {
        int32_t* tlf = c->tl + (size_t)f * B;
        size_t nt = 0;
        for (int b = 0; b < B; ++b) bwd_entry(c, b, f, tlf, &nt);
        bwd_rows_update(c, tlf, nt);
    }
}
API void orc_fm_backward_update(orc_model* m, const int32_t* ids, const float* xv, int B, const float* S,
                                const float* gs, int use_fm2, const float* gvec) {
    const int F = m->F;
    int32_t* tl = (int32_t*)malloc(sizeof(int32_t) * (size_t)B * F);
    bwd_ctx c = {m, ids, xv, B, S, gs, use_fm2, gvec, tl, NULL, 0, NULL, NULL};
    if (m->rank_B > 0 && !gvec) {   /* FM-only step of the multi-GPU path */
        c.tot = (float*)malloc(sizeof(float) * (size_t)m->R * (m->k + 1));
        c.cur_rank = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->R);
    }
    /* A row's contributions must be added in sample order.  The reference walk is b-major, f-minor; when the rows
     * of different fields are disjoint (global id = field offset + local id), walking field by field with the samples
     * ascending inside gives every row the same order, and the fields become independent (one thread each).  That
     * is only used when the batch's id ranges are disjoint and ordered field by field (checked here). */
    int by_field = orc_threads() > 1 && B >= 256;
    if (by_field) {
        int32_t* lo = (int32_t*)malloc(sizeof(int32_t) * F);
        int32_t* hi = (int32_t*)malloc(sizeof(int32_t) * F);
        for (int f = 0; f < F; ++f) { lo[f] = ids[f]; hi[f] = ids[f]; }
        for (int b = 1; b < B; ++b)
            for (int f = 0; f < F; ++f) {
                const int32_t r = ids[(size_t)b * F + f];
                if (r < lo[f]) lo[f] = r;
                if (r > hi[f]) hi[f] = r;
            }
        for (int f = 0; f + 1 < F; ++f)
            if (hi[f] >= lo[f + 1]) by_field = 0;
        free(lo); free(hi);
    }
    if (by_field) {
        orc_parallel(B, bwd_fields, &c);
    } else {
        size_t nt = 0;
        for (int b = 0; b < B; ++b)
            for (int f = 0; f < F; ++f) bwd_entry(&c, b, f, tl, &nt);
        bwd_rows_update(&c, tl, nt);
    }
    free(tl); free(c.tot); free(c.cur_rank);
}

/* ------------------------------------------------------------------ */
/* the reference method surface                                        */
/* ------------------------------------------------------------------ */

static int is_onn(const orc_model* m) { return m->kind >= K_DEEPFM_ONN; }
static int is_nfm(const orc_model* m) { return m->kind == K_NFM || m->kind == K_NFM_ONN; }

typedef struct {
    float *first, *S, *bi, *sf, *sb, *zfm, *act, *head, *base;
} fwd_buf;

static void fb_alloc(const orc_model* m, int B, fwd_buf* w) {
    w->S = (float*)malloc(sizeof(float) * (size_t)B * m->k);
    w->bi = (float*)malloc(sizeof(float) * (size_t)B * m->k);
    w->sf = (float*)malloc(sizeof(float) * B);
    w->sb = (float*)malloc(sizeof(float) * B);
    w->zfm = (float*)malloc(sizeof(float) * B);
    w->base = (float*)malloc(sizeof(float) * B);
    w->first = NULL;
    w->act = m->L > 0 ? (float*)malloc(sizeof(float) * (size_t)m->L * B * m->H) : NULL;
    w->head = m->L > 0 ? (float*)malloc(sizeof(float) * (size_t)m->L * B) : NULL;
}
static void fb_free(fwd_buf* w) {
    free(w->S); free(w->bi); free(w->sf); free(w->sb); free(w->zfm); free(w->base); free(w->act); free(w->head);
}

/* full forward. z[b]: logit (Adam family, fm_adam.py:53 / deepfm_adam.py:88 / nfm_adam.py:87) or the
 * last head's probability (ONN, deepfm_onn.py:102). players: [L*B] head probabilities (ONN) or NULL. */
static void full_forward(const orc_model* m, const int32_t* ids, const float* xv, int B, fwd_buf* w, float* z,
                         float* players) {
    orc_fm_forward(m, ids, xv, B, NULL, w->S, w->bi, w->sf, w->sb, w->zfm);
    for (int b = 0; b < B; ++b) w->base[b] = is_nfm(m) ? w->sf[b] + m->bias[0] : w->zfm[b];
    if (m->kind == K_FM) { for (int b = 0; b < B; ++b) z[b] = w->zfm[b]; return; }
    orc_mlp_forward(m, w->bi, B, w->act, w->head);
    if (!is_onn(m)) {
        for (int b = 0; b < B; ++b) z[b] = w->base[b] + w->head[(size_t)(m->L - 1) * B + b];
    } else {
        for (int l = 0; l < m->L; ++l)
            for (int b = 0; b < B; ++b) {
                float p = orc_sigmoid_at(w->base[b] + w->head[(size_t)l * B + b], b, B);
                if (players) players[(size_t)l * B + b] = p;
                if (l == m->L - 1) z[b] = p;
            }
    }
}

API void orc_forward(const orc_model* m, const int32_t* ids, const float* xv, int B, float* z, float* players) {
    fwd_buf w; fb_alloc(m, B, &w);
    full_forward(m, ids, xv, B, &w, z, players);
    fb_free(&w);
}

/* predict: fm_adam.py:84-88 (sigmoid(z) > 0.5); deepfm_onn.py:171-175 (sigmoid(p_last) > 0.5) */
API void orc_predict(const orc_model* m, const int32_t* ids, const float* xv, int B, uint8_t* pred) {
    float* z = (float*)malloc(sizeof(float) * B);
    orc_forward(m, ids, xv, B, z, NULL);
    for (int b = 0; b < B; ++b) pred[b] = orc_sigmoid_at(z[b], b, B) > 0.5f;
    free(z);
}

/* update_embedding: fm_adam.py:56-69, deepfm_adam.py:91-104, nfm_adam.py:90-103,
 * deepfm_onn.py:156-169, nfm_onn.py:158-171.  Loss on forward_fm only; MLP untouched. */
API float orc_update_embedding(orc_model* m, const int32_t* ids, const float* xv, const float* y, int B) {
    fwd_buf w; fb_alloc(m, B, &w);
    orc_fm_forward(m, ids, xv, B, NULL, w.S, w.bi, w.sf, w.sb, w.zfm);
    float* delta = (float*)malloc(sizeof(float) * B);
    const int kind = is_nfm(m) ? 1 : 0; /* NFM*: BCEWL(sigmoid(z_fm)); others BCEWL(z_fm) */
    float loss = orc_loss_delta(kind, w.zfm, y, B, delta);
    float gbias = orc_sum_aten(delta, B);
    orc_fm_backward_update(m, ids, xv, B, w.S, delta, 1, NULL);
    if (m->update_mode == 2)
        m->bias[0] = orc_ftrl_update(m->bias[0], gbias, m->f_bias, m->f_bias + 1, m->lr, m->f_beta, m->f_l1, m->f_l2);
    else
        m->bias[0] = upd(m->bias[0], gbias, m->lr, m->update_mode);
    free(delta); fb_free(&w);
    return loss;
}

static float bce_prob(float p, float y) { /* nn.BCELoss per element (ATen Loss.cpp binary_cross_entropy) */
    float w = 1.0f + (-p);
    float l1 = (w == 1.0f) ? -p : (w == 0.f ? -INFINITY : orc_logf(w) - (((w - 1.0f) - (-p)) / w));
    float l0 = orc_logf(p);
    l1 = fmaxf(l1, -100.f); l0 = fmaxf(l0, -100.f);
    return ((y - 1.0f) * l1) - (y * l0);
}

/* Adam-family fit (fm_adam.py:71-82, deepfm_adam.py:106-117, nfm_adam.py:105-116) and
 * hedge-backprop fit (deepfm_onn.py:109-154, nfm_onn.py:111-156). */
API void orc_fit(orc_model* m, const int32_t* ids, const float* xv, const float* y, int B) {
    fwd_buf w; fb_alloc(m, B, &w);
    float* z = (float*)malloc(sizeof(float) * B);
    float* delta = (float*)malloc(sizeof(float) * B);
    if (!is_onn(m)) {
        full_forward(m, ids, xv, B, &w, z, NULL);
        const int kind = (m->kind == K_NFM) ? 0 : 1; /* FM, DeepFM: BCEWL(sigmoid(z)); NFM: BCEWL(z) */
        (void)orc_loss_delta(kind, z, y, B, delta);
        float gbias = orc_sum_aten(delta, B);
        if (m->kind == K_FM) {
            orc_fm_backward_update(m, ids, xv, B, w.S, delta, 1, NULL);
        } else {
            const int64_t n = orc_mlp_numel(m);
            float* gmlp = (float*)malloc(sizeof(float) * n);
            float* gbi = (float*)malloc(sizeof(float) * (size_t)B * m->k);
            orc_mlp_backward(m, w.bi, w.act, delta, m->L - 1, B, gmlp, gbi);
            orc_fm_backward_update(m, ids, xv, B, w.S, delta, m->kind == K_DEEPFM, gbi);
            orc_update_dense(m->mlp, gmlp, n, m->lr, m->update_mode);
            free(gmlp); free(gbi);
        }
        m->bias[0] = upd(m->bias[0], gbias, m->lr, m->update_mode);
    } else {
        const int L = m->L;
        const int64_t n = orc_mlp_numel(m);
        float* pl = (float*)malloc(sizeof(float) * (size_t)L * B);
        full_forward(m, ids, xv, B, &w, z, pl);
        float* loss = (float*)malloc(sizeof(float) * L);
        float* lv = (float*)malloc(sizeof(float) * B);
        float* gmlp = (float*)malloc(sizeof(float) * n);
        float* acc = (float*)malloc(sizeof(float) * n);
        const float invB = 1.0f / (float)B; /* mean backward */
        for (int i = 0; i < L; ++i) {
            for (int b = 0; b < B; ++b) lv[b] = bce_prob(pl[(size_t)i * B + b], y[b]);
            loss[i] = orc_sum_aten(lv, B) / (float)B;
            for (int b = 0; b < B; ++b) {
                float p = pl[(size_t)i * B + b];
                float den = fmaxf((1.0f - p) * p, 1e-12f);
                float dp = (invB * (p - y[b])) / den;      /* binary_cross_entropy_backward */
                delta[b] = (dp * (1.0f - p)) * p;          /* sigmoid_backward */
            }
            orc_mlp_backward(m, w.bi, w.act, delta, i, B, gmlp, NULL);
            /* deepfm_onn.py:132-139: w[j] (+)= alpha[i] * grad for j <= i */
            for (int j = 0; j <= i; ++j) {
                /* head j is the first one whose backward reaches layer j (w[j] is None before) */
                const size_t lo = mlp_w_off(m, j), hi = mlp_w_off(m, j + 1);
                for (size_t t = lo; t < hi; ++t) {
                    float term = m->alpha[i] * gmlp[t];
                    acc[t] = (j == i) ? term : acc[t] + term;
                }
            }
        }
        for (int64_t t = 0; t < n; ++t) m->mlp[t] = m->mlp[t] - (m->lr * acc[t]); /* :143-145 */
        const float floorv = m->hs / (float)L;
        for (int i = 0; i < L; ++i) {
            float a = m->alpha[i] * orc_powf(m->hb, loss[i]); /* :148 */
            m->alpha[i] = fmaxf(a, floorv);                   /* :149-150 */
        }
        float zt = orc_sum_aten(m->alpha, L);
        for (int i = 0; i < L; ++i) m->alpha[i] = m->alpha[i] / zt; /* :152-154 */
        free(pl); free(loss); free(lv); free(gmlp); free(acc);
    }
    free(z); free(delta); fb_free(&w);
}

/* run_experiment: fm_adam.py:90-119 (identical in all five classes).  conf = {tp, fp, tn, fn}. */
API void orc_run_experiment(orc_model* m, const int32_t* ids, const float* xv, const float* y, int N, int64_t* conf,
                            uint8_t* preds) {
    conf[0] = conf[1] = conf[2] = conf[3] = 0;
    for (int i = 0; i < N; ++i) {
        uint8_t p;
        orc_predict(m, ids + (size_t)i * m->F, xv + (size_t)i * m->F, 1, &p);
        orc_fit(m, ids + (size_t)i * m->F, xv + (size_t)i * m->F, y + i, 1);
        const int yi = y[i] == 1.0f;
        if ((int)p == yi) { if (yi) conf[0]++; else conf[2]++; }
        else { if (yi) conf[3]++; else conf[1]++; }
        if (preds) preds[i] = p;
    }
}
