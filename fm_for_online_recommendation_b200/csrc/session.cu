// session.cu -- a training session: owns every temporary of the hot path (device workspaces and
// pinned host staging) so that one C call performs a whole `update_embedding` / `fit` step.
//
// This is the drop-in boundary for the FM training step: the entry points take plain pointers and
// sizes; `*_host` variants take HOST buffers and do the H2D/D2H copies themselves (what a maintainer
// of the reference would bind in place of fm_adam.py:56-82 / deepfm_adam.py:91-117).
#include "fmb_common.cuh"
#include <cstdlib>
#include <cstring>
#include <new>

extern "C" {
int fmb_fm_forward(const int32_t*, const float*, const float*, const float*, int, int, int, float*, float*, float*,
                   float*, float*, const float*, int, float*, float*, cudaStream_t);
size_t fmb_sort_workspace_bytes(int64_t);
int fmb_sort_segment(const int32_t*, int64_t, int, void*, size_t, int32_t*, int32_t*, int32_t*, int32_t*,
                     cudaStream_t);
size_t fmb_bwd_workspace_bytes(int64_t, int);
int fmb_sort_fields_max_batch(void);
int fmb_sort_fields(const int32_t*, int, int, const int32_t*, int32_t*, int32_t*, cudaStream_t);
int fmb_fm_backward_update(const int32_t*, const int32_t*, int64_t, const float*, float*, int, int, const float*,
                           const float*, int, const float*, float, int, void*, size_t, cudaStream_t);
int fmb_finish_step(const float*, const float*, int, float*, float, int, float*, cudaStream_t);
}

#define FMB_GRAPH_CACHE 128
struct StepKey {
    const void *ids, *xv, *y, *table, *bias, *loss;
    int B, key_bits, loss_kind, mode, sort_buf, skip_sort;
    float lr;
};

struct fmb_session {
    int F, k, rowp, kp4;
    int64_t maxB;
    // device
    int32_t* d_ids;
    float* d_xv;
    float* d_y;
    float* d_S;
    float* d_z;
    float* d_delta;
    float* d_lossv;
    float* d_loss;
    int32_t* d_skeys;   // sorted keys / permutation of the step being run (points into the two buffers below)
    int32_t* d_perm;
    int32_t* d_skeys_buf[2];
    int32_t* d_perm_buf[2];
    // pre-sort: the sort depends on the ids only, so the sort of batch t+1 may run (on st1) while step t is
    // still in its backward kernels.  presort_ids names the batch whose sorted form sits in presort_buf.
    const int32_t* presort_ids;
    int presort_B, presort_buf, last_buf;
    cudaEvent_t ev_presort, ev_buf_free[2];
    int buf_used[2];
    void* d_sort_ws;
    size_t sort_ws_bytes;
    void* d_bwd_ws;
    size_t bwd_ws_bytes;
    int32_t* d_field_off;  // [F+1] global row offset of every field (enables the per-field sort)
    // pinned host staging
    int32_t* h_ids;
    float* h_xv;
    float* h_y;
    float* h_loss;
    // second input slot + copy stream for the pipelined host entry point (slot 0 = the buffers above)
    int32_t* d_ids2; float* d_xv2; float* d_y2; float* d_loss2;
    int32_t* h_ids2; float* h_xv2; float* h_y2;
    cudaStream_t st_copy;
    cudaEvent_t ev_h2d[2], ev_done[2];
    int slot_used[2];
    int64_t launches;  // kernels launched through this session (bench.py's gpu_launches)
    // CUDA-graph cache of whole steps, keyed by every argument that is baked into the kernels
    cudaStream_t st0, st1, st2;   // st0/st1: graph capture (main / side branch); st2: pre-sorts
    int64_t steps_done;
    cudaEvent_t ev_fork, ev_fwd, ev_sort, ev_join;
    int use_graph;
    int ngraphs, next_evict;
    StepKey gkey[FMB_GRAPH_CACHE];
    cudaGraphExec_t gexec[FMB_GRAPH_CACHE];
    int glaunches[FMB_GRAPH_CACHE];
};

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) { fmb_set_error("%s: %s", #call, cudaGetErrorString(e__)); return FMB_ERR_CUDA; } \
    } while (0)

FMB_API void fmb_session_destroy(fmb_session* s) {
    if (!s) return;
    cudaFree(s->d_ids); cudaFree(s->d_xv); cudaFree(s->d_y); cudaFree(s->d_S); cudaFree(s->d_z);
    cudaFree(s->d_delta); cudaFree(s->d_lossv); cudaFree(s->d_loss);
    for (int i = 0; i < 2; ++i) { cudaFree(s->d_skeys_buf[i]); cudaFree(s->d_perm_buf[i]); if (s->ev_buf_free[i]) cudaEventDestroy(s->ev_buf_free[i]); }
    if (s->ev_presort) cudaEventDestroy(s->ev_presort);
    cudaFree(s->d_sort_ws); cudaFree(s->d_bwd_ws); cudaFree(s->d_field_off);
    cudaFreeHost(s->h_ids); cudaFreeHost(s->h_xv); cudaFreeHost(s->h_y); cudaFreeHost(s->h_loss);
    cudaFree(s->d_ids2); cudaFree(s->d_xv2); cudaFree(s->d_y2); cudaFree(s->d_loss2);
    cudaFreeHost(s->h_ids2); cudaFreeHost(s->h_xv2); cudaFreeHost(s->h_y2);
    if (s->st_copy) cudaStreamDestroy(s->st_copy);
    for (int i = 0; i < 2; ++i) { if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]); if (s->ev_done[i]) cudaEventDestroy(s->ev_done[i]); }
    for (int i = 0; i < s->ngraphs; ++i) cudaGraphExecDestroy(s->gexec[i]);
    if (s->st0) cudaStreamDestroy(s->st0);
    if (s->st1) cudaStreamDestroy(s->st1);
    if (s->st2) cudaStreamDestroy(s->st2);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_fwd) cudaEventDestroy(s->ev_fwd);
    if (s->ev_sort) cudaEventDestroy(s->ev_sort);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    delete s;
}

// F fields, embedding size k, up to max_batch samples per step.  field_off_host [F+1] (nullable):
// global row offset of each field; when given and B <= fmb_sort_fields_max_batch() the step uses
// the one-kernel per-field sort instead of the generic radix sort (same result).
FMB_API int fmb_session_create(fmb_session** out, int F, int k, int64_t max_batch, const int32_t* field_off_host) {
    FMB_CHECK_ARG(out && F > 0 && k > 0 && max_batch > 0, "fmb_session_create: bad arguments");
    FMB_CHECK_ARG(max_batch * F < ((int64_t)1 << 31), "fmb_session_create: max_batch*F must fit int32");
    fmb_session* s = new (std::nothrow) fmb_session();
    FMB_CHECK_ARG(s, "fmb_session_create: out of host memory");
    memset(s, 0, sizeof(*s));
    s->F = F; s->k = k; s->rowp = fmb_round_up(k + 1, 16); s->kp4 = fmb_round_up(k, 4); s->maxB = max_batch;
    const int64_t N = max_batch * F;
    s->sort_ws_bytes = fmb_sort_workspace_bytes(N);
    s->bwd_ws_bytes = fmb_bwd_workspace_bytes(N, k);
    cudaError_t e = cudaSuccess;
    auto dm = [&](void** p, size_t n) { if (e == cudaSuccess) e = cudaMalloc(p, n); };
    auto hm = [&](void** p, size_t n) { if (e == cudaSuccess) e = cudaMallocHost(p, n); };
    dm((void**)&s->d_ids, N * 4); dm((void**)&s->d_xv, N * 4); dm((void**)&s->d_y, max_batch * 4);
    dm((void**)&s->d_S, max_batch * s->kp4 * 4); dm((void**)&s->d_z, max_batch * 4);
    dm((void**)&s->d_delta, max_batch * 4); dm((void**)&s->d_lossv, max_batch * 4); dm((void**)&s->d_loss, 256);
    for (int i = 0; i < 2; ++i) { dm((void**)&s->d_skeys_buf[i], N * 4); dm((void**)&s->d_perm_buf[i], N * 4); }
    s->d_skeys = s->d_skeys_buf[0]; s->d_perm = s->d_perm_buf[0];
    dm(&s->d_sort_ws, s->sort_ws_bytes); dm(&s->d_bwd_ws, s->bwd_ws_bytes);
    hm((void**)&s->h_ids, N * 4); hm((void**)&s->h_xv, N * 4); hm((void**)&s->h_y, max_batch * 4);
    hm((void**)&s->h_loss, 256);
    dm((void**)&s->d_ids2, N * 4); dm((void**)&s->d_xv2, N * 4); dm((void**)&s->d_y2, max_batch * 4);
    dm((void**)&s->d_loss2, 256);
    hm((void**)&s->h_ids2, N * 4); hm((void**)&s->h_xv2, N * 4); hm((void**)&s->h_y2, max_batch * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st_copy, cudaStreamNonBlocking);
    for (int i = 0; i < 2; ++i) {
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming);
    }
    if (field_off_host) {
        dm((void**)&s->d_field_off, (size_t)(F + 1) * 4);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_field_off, field_off_host, (size_t)(F + 1) * 4, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st0, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st1, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fwd, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_sort, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_presort, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_buf_free[i], cudaEventDisableTiming);
    {
        const char* ng = getenv("FMB_NO_GRAPH");
        s->use_graph = !(ng && ng[0] == '1');
    }
    if (e != cudaSuccess) {
        fmb_set_error("fmb_session_create: %s", cudaGetErrorString(e));
        fmb_session_destroy(s);
        return FMB_ERR_CUDA;
    }
    *out = s;
    return FMB_OK;
}

FMB_API int64_t fmb_session_launches(const fmb_session* s) { return s ? s->launches : 0; }
FMB_API int fmb_session_graph_count(const fmb_session* s) { return s ? s->ngraphs : 0; }

// the kernels of one FM-only step.  `side` (nullable) is a second stream: the sort does not depend on
// the forward pass and the bias/loss epilogue does not depend on the row updates, so they fork.
static int fm_step_launch(fmb_session* s, const int32_t* ids, const float* xv, const float* y, int B, float* table,
                          float* bias, int key_bits, int loss_kind, float lr, int mode, float* loss_dev,
                          cudaStream_t main, cudaStream_t side, int sort_buf, bool skip_sort, int* nlaunch) {
    const int64_t N = (int64_t)B * s->F;
    int32_t* skeys = s->d_skeys_buf[sort_buf];
    int32_t* perm = s->d_perm_buf[sort_buf];
    const bool by_field = s->d_field_off && B <= fmb_sort_fields_max_batch();
    cudaStream_t sort_st = side ? side : main;
    if (side) { cudaEventRecord(s->ev_fork, main); cudaStreamWaitEvent(side, s->ev_fork, 0); }
    int rc = fmb_fm_forward(ids, xv, table, bias, B, s->F, s->k, nullptr, s->d_S, nullptr, nullptr, s->d_z, y,
                            loss_kind, s->d_delta, s->d_lossv, main);
    if (rc) return rc;
    if (side) cudaEventRecord(s->ev_fwd, main);
    if (!skip_sort) {
        if (by_field)
            rc = fmb_sort_fields(ids, B, s->F, s->d_field_off, skeys, perm, sort_st);
        else
            rc = fmb_sort_segment(ids, N, key_bits, s->d_sort_ws, s->sort_ws_bytes, skeys, perm, nullptr, nullptr,
                                  sort_st);
        if (rc) return rc;
    }
    if (side) { cudaEventRecord(s->ev_sort, side); cudaStreamWaitEvent(main, s->ev_sort, 0); }
    rc = fmb_fm_backward_update(skeys, perm, N, xv, table, s->F, s->k, s->d_S, s->d_delta, 1, nullptr, lr,
                                mode, s->d_bwd_ws, s->bwd_ws_bytes, main);
    if (rc) return rc;
    if (side) cudaStreamWaitEvent(side, s->ev_fwd, 0);
    rc = fmb_finish_step(s->d_delta, s->d_lossv, B, bias, lr, mode, loss_dev ? loss_dev : s->d_loss,
                         side ? side : main);
    if (rc) return rc;
    if (side) { cudaEventRecord(s->ev_join, side); cudaStreamWaitEvent(main, s->ev_join, 0); }
    *nlaunch = 1 + (skip_sort ? 0 : (by_field ? 1 : 3 * ((key_bits + 7) / 8))) + 2 + 1;
    return FMB_OK;
}

// One FM-only training step with DEVICE inputs (FMAdam.update_embedding/fit, and the
// update_embedding of DeepFM/NFM/ONN classes whose loss is on forward_fm only).
//   loss_kind 0: BCEWithLogits(z_fm)   (fm_adam.py:66, deepfm_adam.py:101, deepfm_onn.py:166)
//   loss_kind 1: BCEWithLogits(sigmoid(z_fm))   (fm_adam.py:80, nfm_adam.py:100, nfm_onn.py:168)
//   loss_dev (nullable): receives the mean loss (device scalar).
// The step is captured once per distinct argument set into a CUDA graph (forward || sort ->
// backward/update || bias+loss) and replayed afterwards; FMB_NO_GRAPH=1 launches the kernels directly.
FMB_API int fmb_session_fm_step(fmb_session* s, const int32_t* ids, const float* xv, const float* y, int B,
                                float* table, float* bias, int key_bits, int loss_kind, float lr, int mode,
                                float* loss_dev, cudaStream_t stream) {
    FMB_CHECK_ARG(s && ids && y && table && bias, "fmb_session_fm_step: null pointer");
    FMB_CHECK_ARG(B > 0 && B <= s->maxB, "fmb_session_fm_step: B=%d exceeds session max_batch", B);
    int nl = 0;
    // sorted form of this batch: already produced by fmb_session_presort, or sorted inside the step
    const bool pre = s->presort_ids == ids && s->presort_B == B;
    const int buf = pre ? s->presort_buf : 1 - s->last_buf;
    if (pre) {
        CU(cudaStreamWaitEvent(stream, s->ev_presort, 0));
        s->presort_ids = nullptr;
    } else {
        if (s->buf_used[buf]) CU(cudaStreamWaitEvent(stream, s->ev_buf_free[buf], 0));
        if (s->presort_ids && s->presort_buf == buf) {
            // a pre-sort of some other batch targets the buffer this step is about to overwrite: order the
            // step behind it and forget its result
            CU(cudaStreamWaitEvent(stream, s->ev_presort, 0));
            s->presort_ids = nullptr;
        }
    }
    s->last_buf = buf;
    int rc = FMB_OK;
    // the very first step runs eagerly: it sets the kernels' function attributes outside any capture
    if (!s->use_graph || s->steps_done == 0) {
        // direct launches; the sort (when not pre-sorted) and the bias/loss epilogue still fork to the side stream
        rc = fm_step_launch(s, ids, xv, y, B, table, bias, key_bits, loss_kind, lr, mode, loss_dev, stream, s->st1,
                            buf, pre, &nl);
        s->launches += nl;
    } else {
        StepKey key;
        memset(&key, 0, sizeof(key));
        key.ids = ids; key.xv = xv; key.y = y; key.table = table; key.bias = bias; key.loss = nullptr;
        key.B = B; key.key_bits = key_bits; key.loss_kind = loss_kind; key.mode = mode; key.lr = lr;
        key.sort_buf = buf; key.skip_sort = pre;
        int slot = -1;
        for (int i = 0; i < s->ngraphs; ++i)
            if (memcmp(&s->gkey[i], &key, sizeof(key)) == 0) { slot = i; break; }
        if (slot < 0) {
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(s->st0, cudaStreamCaptureModeThreadLocal));
            // the graph always writes the loss to the session's own scalar: the caller's pointer (often a fresh
            // allocation every step) must not be part of the cache key
            rc = fm_step_launch(s, ids, xv, y, B, table, bias, key_bits, loss_kind, lr, mode, s->d_loss, s->st0,
                                s->st1, buf, pre, &nl);
            cudaError_t e = cudaStreamEndCapture(s->st0, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) { fmb_set_error("graph capture: %s", cudaGetErrorString(e)); return FMB_ERR_CUDA; }
            cudaGraphExec_t exec = nullptr;
            e = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { fmb_set_error("graph instantiate: %s", cudaGetErrorString(e)); return FMB_ERR_CUDA; }
            if (s->ngraphs < FMB_GRAPH_CACHE) slot = s->ngraphs++;
            else { slot = s->next_evict; s->next_evict = (s->next_evict + 1) % FMB_GRAPH_CACHE; cudaGraphExecDestroy(s->gexec[slot]); }
            s->gkey[slot] = key; s->gexec[slot] = exec; s->glaunches[slot] = nl;
        }
        CU(cudaGraphLaunch(s->gexec[slot], stream));
        if (loss_dev && loss_dev != s->d_loss) CU(cudaMemcpyAsync(loss_dev, s->d_loss, 4, cudaMemcpyDeviceToDevice, stream));
        s->launches += s->glaunches[slot];
    }
    if (rc) return rc;
    CU(cudaEventRecord(s->ev_buf_free[buf], stream));
    s->buf_used[buf] = 1;
    s->steps_done += 1;
    return FMB_OK;
}

// Sort batch `ids` now, on the session's side stream, into the sorted-buffer the running step is NOT using;
// the next fmb_session_fm_step called with the same `ids` pointer and B skips its own sort.  The sort
// depends on the ids only, so this overlaps the backward kernels of the step in flight.  `ready` (nullable)
// is an event the ids depend on (e.g. their H2D copy).  The ids must not change until that step has run.
static int presort_impl(fmb_session* s, const int32_t* ids, int B, int key_bits, cudaEvent_t ready) {
    const int buf = 1 - s->last_buf;
    if (s->buf_used[buf]) CU(cudaStreamWaitEvent(s->st2, s->ev_buf_free[buf], 0));
    if (ready) CU(cudaStreamWaitEvent(s->st2, ready, 0));
    const int64_t N = (int64_t)B * s->F;
    int rc;
    if (s->d_field_off && B <= fmb_sort_fields_max_batch())
        rc = fmb_sort_fields(ids, B, s->F, s->d_field_off, s->d_skeys_buf[buf], s->d_perm_buf[buf], s->st2);
    else
        rc = fmb_sort_segment(ids, N, key_bits, s->d_sort_ws, s->sort_ws_bytes, s->d_skeys_buf[buf], s->d_perm_buf[buf],
                              nullptr, nullptr, s->st2);
    if (rc) return rc;
    CU(cudaEventRecord(s->ev_presort, s->st2));
    s->launches += (s->d_field_off && B <= fmb_sort_fields_max_batch()) ? 1 : 3 * ((key_bits + 7) / 8);
    s->presort_ids = ids; s->presort_B = B; s->presort_buf = buf;
    return FMB_OK;
}

FMB_API int fmb_session_presort(fmb_session* s, const int32_t* ids, int B, int key_bits) {
    FMB_CHECK_ARG(s && ids && B > 0 && B <= s->maxB, "fmb_session_presort: bad arguments");
    return presort_impl(s, ids, B, key_bits, nullptr);
}

// true when `p` is page-locked host memory the copy engine can read directly (no staging copy needed)
static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Pipelined step with HOST inputs.  `slot` (0 or 1) selects one of two device input buffers: the copies of
// step t+1 (on the session's copy stream) overlap the kernels of step t (on `stream`).  Pinned / registered
// host buffers are read in place, pageable ones go through the session's pinned staging area.  The loss
// arrives in pinned memory; fmb_session_wait_loss(slot) waits for it.  Every step still moves its own
// 4*B*F (+4*B*F) + 4*B bytes in and 4 bytes out.
FMB_API int fmb_session_fm_step_host_async(fmb_session* s, int slot, const int32_t* ids_host, const float* xv_host,
                                           const float* y_host, int B, float* table, float* bias, int key_bits,
                                           int loss_kind, float lr, int mode, cudaStream_t stream) {
    FMB_CHECK_ARG(s && ids_host && y_host && table && bias, "fmb_session_fm_step_host_async: null pointer");
    FMB_CHECK_ARG(slot == 0 || slot == 1, "fmb_session_fm_step_host_async: slot must be 0 or 1");
    FMB_CHECK_ARG(B > 0 && B <= s->maxB, "fmb_session_fm_step_host_async: B=%d exceeds session max_batch", B);
    const size_t N = (size_t)B * s->F;
    int32_t* d_ids = slot ? s->d_ids2 : s->d_ids;
    float* d_xv = slot ? s->d_xv2 : s->d_xv;
    float* d_y = slot ? s->d_y2 : s->d_y;
    float* d_loss = slot ? s->d_loss2 : s->d_loss;
    int32_t* h_ids = slot ? s->h_ids2 : s->h_ids;
    float* h_xv = slot ? s->h_xv2 : s->h_xv;
    float* h_y = slot ? s->h_y2 : s->h_y;
    // the previous user of this slot (two steps ago) must be done with the device buffers and the staging area
    if (s->slot_used[slot]) CU(cudaEventSynchronize(s->ev_done[slot]));
    const void* src_ids = ids_host;
    const void* src_y = y_host;
    const void* src_xv = xv_host;
    if (!host_ptr_is_pinned(ids_host)) { memcpy(h_ids, ids_host, N * 4); src_ids = h_ids; }
    if (!host_ptr_is_pinned(y_host)) { memcpy(h_y, y_host, (size_t)B * 4); src_y = h_y; }
    if (xv_host && !host_ptr_is_pinned(xv_host)) { memcpy(h_xv, xv_host, N * 4); src_xv = h_xv; }
    CU(cudaMemcpyAsync(d_ids, src_ids, N * 4, cudaMemcpyHostToDevice, s->st_copy));
    CU(cudaMemcpyAsync(d_y, src_y, (size_t)B * 4, cudaMemcpyHostToDevice, s->st_copy));
    if (xv_host) CU(cudaMemcpyAsync(d_xv, src_xv, N * 4, cudaMemcpyHostToDevice, s->st_copy));
    CU(cudaEventRecord(s->ev_h2d[slot], s->st_copy));
    CU(cudaStreamWaitEvent(stream, s->ev_h2d[slot], 0));
    // the sort of this batch starts as soon as its ids have landed: it overlaps the step still in flight
    int rc = presort_impl(s, d_ids, B, key_bits, s->ev_h2d[slot]);
    if (rc) return rc;
    rc = fmb_session_fm_step(s, d_ids, xv_host ? d_xv : nullptr, d_y, B, table, bias, key_bits, loss_kind, lr,
                                 mode, d_loss, stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s->h_loss + 8 * slot, d_loss, 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaEventRecord(s->ev_done[slot], stream));
    s->slot_used[slot] = 1;
    return FMB_OK;
}

// waits for the step last submitted on `slot` and returns its mean loss
FMB_API int fmb_session_wait_loss(fmb_session* s, int slot, float* loss_host) {
    FMB_CHECK_ARG(s && (slot == 0 || slot == 1) && s->slot_used[slot], "fmb_session_wait_loss: nothing submitted on slot %d", slot);
    CU(cudaEventSynchronize(s->ev_done[slot]));
    if (loss_host) *loss_host = s->h_loss[8 * slot];
    return FMB_OK;
}

// Same step with HOST inputs: copies ids/xv/y in (xv_host NULL = all ones, nothing copied), runs
// the step, copies the loss back and waits for it.  Bytes moved per step: H2D 4*B*F (+4*B*F) + 4*B,
// D2H 4.
FMB_API int fmb_session_fm_step_host(fmb_session* s, const int32_t* ids_host, const float* xv_host,
                                     const float* y_host, int B, float* table, float* bias, int key_bits,
                                     int loss_kind, float lr, int mode, float* loss_host, cudaStream_t stream) {
    FMB_CHECK_ARG(s && ids_host && y_host && table && bias, "fmb_session_fm_step_host: null pointer");
    FMB_CHECK_ARG(B > 0 && B <= s->maxB, "fmb_session_fm_step_host: B=%d exceeds session max_batch", B);
    const size_t N = (size_t)B * s->F;
    memcpy(s->h_ids, ids_host, N * 4);
    memcpy(s->h_y, y_host, (size_t)B * 4);
    CU(cudaMemcpyAsync(s->d_ids, s->h_ids, N * 4, cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(s->d_y, s->h_y, (size_t)B * 4, cudaMemcpyHostToDevice, stream));
    if (xv_host) {
        memcpy(s->h_xv, xv_host, N * 4);
        CU(cudaMemcpyAsync(s->d_xv, s->h_xv, N * 4, cudaMemcpyHostToDevice, stream));
    }
    int rc = fmb_session_fm_step(s, s->d_ids, xv_host ? s->d_xv : nullptr, s->d_y, B, table, bias, key_bits,
                                 loss_kind, lr, mode, s->d_loss, stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s->h_loss, s->d_loss, 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    if (loss_host) *loss_host = s->h_loss[0];
    return FMB_OK;
}
