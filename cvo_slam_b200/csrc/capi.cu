// capi.cu — the extern "C" boundary declared in include/cvo_b200.h.
//
// A handle mirrors one cvo::cvo object (thirdparty/cvo/include/cvo.hpp:82-282): three cloud
// slots (fixed / moving / previous), the persistent R, T, ell, and one CUDA stream.  A batch
// holds many frames and aligns many independent pairs in one launch
// (the loop-closure verification pattern of src/keyframe_graph.cpp:622-731).

#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <utility>
#include <thread>
#include <vector>

namespace cvo_b200 {

static thread_local char g_err[512] = "";

void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int arena_alloc(CloudArena &a, int frames, int cap) {
    a = CloudArena();
    a.cap = cap;
    a.frames = frames;
    size_t n = (size_t)frames * cap;
    CVO_CUDA_TRY(cudaMalloc(&a.pos, n * sizeof(float4)));
    CVO_CUDA_TRY(cudaMalloc(&a.f03, n * sizeof(float4)));
    CVO_CUDA_TRY(cudaMalloc(&a.f4, n * sizeof(float)));
    CVO_CUDA_TRY(cudaMalloc(&a.pix, n * sizeof(float2)));
    CVO_CUDA_TRY(cudaMalloc(&a.n, 2 * frames * sizeof(int)));   // counts, then overflow flags
    CVO_CUDA_TRY(cudaMemset(a.n, 0, 2 * frames * sizeof(int)));
    a.ovf = a.n + frames;
    return CVO_OK;
}

void arena_free(CloudArena &a) {
    cudaFree(a.pos); cudaFree(a.f03); cudaFree(a.f4); cudaFree(a.pix); cudaFree(a.n);
    a = CloudArena();
}

static void identity_RT(float R[9], float T[3]) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.f : 0.f;
    T[0] = T[1] = T[2] = 0.f;
}

}  // namespace cvo_b200

using namespace cvo_b200;

struct cvo_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cvo_calib cal;
    cvo_params prm;
    CloudArena arena;             // 3 clouds
    int slot_idx[3] = {-1, -1, -1};
    SelWorkspace *sel = nullptr;
    int sel_w = 0, sel_h = 0, sel_slot = -1;
    // pinned frame staging of cvo_set_frame_png: two buffers alternate, an event per buffer marks "H2D done"
    uint8_t *png_bgr[2] = {nullptr, nullptr};
    uint16_t *png_depth[2] = {nullptr, nullptr};
    cudaEvent_t png_done[2] = {nullptr, nullptr};
    size_t png_px = 0;
    int png_next = 0;
    AlignWorkspace *aws = nullptr;
    float R[9], T[3], ell;
    // device + pinned staging
    AlignTask *d_task = nullptr;
    cvo_align_result *d_res = nullptr;
    cvo_iter_record *d_trace = nullptr;
    QueryTask *d_q = nullptr;
    QueryOut *d_qo = nullptr;
    char *pinned = nullptr;       // host staging
    int trace_cap = 0;
    int64_t launches = 0;
    int last_nnz = 0;
};

struct cvo_batch {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, sel_stream = nullptr;
    cvo_calib cal;
    cvo_params prm;
    int w = 0, h = 0, max_frames = 0, max_pairs = 0, chunk = 0;
    // Uploads of host images run beside the alignment of other frames: copies on copy_stream, selection on
    // sel_stream.  A consumer on `stream` waits for the uploads whose frame range it touches; an upload waits for
    // the consumers of the frames it overwrites (compute_done covers everything enqueued on `stream` so far).
    struct Upload { int lo = 0, hi = 0; cudaEvent_t ev = nullptr; bool valid = false; } up[2];
    int up_next = 0;
    int busy_lo = 0, busy_hi = 0;         // frames referenced by work on `stream` since the last upload that waited
    cudaEvent_t compute_done = nullptr;
    CloudArena arena;
    SelWorkspace *sel[2] = {nullptr, nullptr};
    cudaEvent_t sel_done[2] = {nullptr, nullptr}, ev0 = nullptr, ev1 = nullptr, ev_user[2] = {nullptr, nullptr};
    AlignWorkspace *aws = nullptr;
    AlignTask *d_tasks = nullptr;
    cvo_align_result *d_results = nullptr;
    QueryTask *d_q = nullptr;
    QueryOut *d_qo = nullptr;
    AlignTask *h_tasks = nullptr;         // pinned
    cvo_align_result *h_results = nullptr; // pinned
    QueryTask *h_q = nullptr;
    QueryOut *h_qo = nullptr;
    LcTask *d_lc = nullptr, *h_lc = nullptr;           // loop-closure verification staging (first use)
    LcOut *d_lco = nullptr, *h_lco = nullptr;
    QueryTask *d_lcq = nullptr, *h_lcq = nullptr;      // its self inner products
    QueryOut *d_lcqo = nullptr, *h_lcqo = nullptr;
    int64_t launches = 0;
    float last_align_ms = 0.f;
};

static const int kPinnedBytes = 1 << 16;

// before enqueuing work on b->stream that touches frames [lo, hi)
static int batch_consume(cvo_batch *b, int lo, int hi) {
    for (int u = 0; u < 2; u++)
        if (b->up[u].valid && b->up[u].lo < hi && lo < b->up[u].hi) CVO_CUDA_TRY(cudaStreamWaitEvent(b->stream, b->up[u].ev, 0));
    if (b->busy_hi <= b->busy_lo) { b->busy_lo = lo; b->busy_hi = hi; }
    else { if (lo < b->busy_lo) b->busy_lo = lo; if (hi > b->busy_hi) b->busy_hi = hi; }
    return CVO_OK;
}
// after enqueuing it
static int batch_consumed(cvo_batch *b) {
    CVO_CUDA_TRY(cudaEventRecord(b->compute_done, b->stream));
    return CVO_OK;
}
static void pair_frame_range(const cvo_pair_desc *pairs, int n, int &lo, int &hi) {
    lo = 1 << 30; hi = -1;
    for (int i = 0; i < n; i++) {
        const int a = pairs[i].fixed_frame < pairs[i].moving_frame ? pairs[i].fixed_frame : pairs[i].moving_frame;
        const int c = pairs[i].fixed_frame > pairs[i].moving_frame ? pairs[i].fixed_frame : pairs[i].moving_frame;
        if (a < lo) lo = a;
        if (c + 1 > hi) hi = c + 1;
    }
}

// k_query reports "this cloud was truncated (selection overflow, or larger than the align scratch)" by
// returning -1 - count: restore the counts, tell the caller
static bool decode_query_counts(QueryOut *q, int n) {
    bool trunc = false;
    for (int i = 0; i < n; i++)
        if (q[i].count < 0) { q[i].count = -1 - q[i].count; trunc = true; }
    return trunc;
}
static int truncated_rc(const char *who) {
    set_last_error("%s: a cloud was truncated (more selected points than the cloud arena / align scratch holds); "
                   "the result covers the stored subset", who);
    return CVO_ERR_CAPACITY;
}

static int handle_ensure_arena(cvo_handle *h, int need_cap) {
    if (h->arena.pos && h->arena.cap >= need_cap) return CVO_OK;
    // grow: allocate a new arena and copy the live clouds
    CloudArena na;
    int cap = need_cap;
    int rc = arena_alloc(na, 3, cap);
    if (rc != CVO_OK) return rc;
    if (h->arena.pos) {
        CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int k = 0; k < 3; k++) {
            size_t so = (size_t)k * h->arena.cap, d_o = (size_t)k * cap;
            size_t n = h->arena.cap;
            CVO_CUDA_TRY(cudaMemcpy(na.pos + d_o, h->arena.pos + so, n * sizeof(float4), cudaMemcpyDeviceToDevice));
            CVO_CUDA_TRY(cudaMemcpy(na.f03 + d_o, h->arena.f03 + so, n * sizeof(float4), cudaMemcpyDeviceToDevice));
            CVO_CUDA_TRY(cudaMemcpy(na.f4 + d_o, h->arena.f4 + so, n * sizeof(float), cudaMemcpyDeviceToDevice));
            CVO_CUDA_TRY(cudaMemcpy(na.pix + d_o, h->arena.pix + so, n * sizeof(float2), cudaMemcpyDeviceToDevice));
        }
        CVO_CUDA_TRY(cudaMemcpy(na.n, h->arena.n, 3 * sizeof(int), cudaMemcpyDeviceToDevice));
        CVO_CUDA_TRY(cudaMemcpy(na.ovf, h->arena.ovf, 3 * sizeof(int), cudaMemcpyDeviceToDevice));
        arena_free(h->arena);
    }
    h->arena = na;
    if (h->aws) { align_ws_destroy(h->aws); h->aws = nullptr; }
    return CVO_OK;
}

// The align scratch is sized by the largest cloud it has to hold.  For ordinary clouds that is the
// arena capacity (no synchronisation); for large arenas (dense selection) the actual point
// counts are read back so that the scratch follows the data, not the worst case.
static int handle_ensure_aws(cvo_handle *h) {
    int need = h->arena.cap;
    if (h->arena.cap > 8192) {
        int n[3] = {0, 0, 0};
        CVO_CUDA_TRY(cudaMemcpyAsync(n, h->arena.n, sizeof(n), cudaMemcpyDeviceToHost, h->stream));
        CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
        int m = 0;
        for (int s = 0; s < 3; s++)
            if (h->slot_idx[s] >= 0 && n[h->slot_idx[s]] > m) m = n[h->slot_idx[s]];
        need = m + m / 4 + 256;
        if (need > h->arena.cap) need = h->arena.cap;
        if (need < 4096) need = 4096;
    }
    if (h->aws && align_ws_max_points(h->aws) >= need) return CVO_OK;
    if (h->aws) { CVO_CUDA_TRY(cudaStreamSynchronize(h->stream)); align_ws_destroy(h->aws); h->aws = nullptr; }
    // a handle aligns one pair at a time: on one thread-block cluster of up to 16 CTAs, or — for clouds
    // large enough to feed them (dense selection, C3) — on a cooperative grid of 128 CTAs
    if (need > 8192) return align_ws_create(&h->aws, need, h->device, 0, 128);
    return align_ws_create(&h->aws, need, h->device, 16);
}

static int handle_slot_arena_index(cvo_handle *h, int slot) {
    if (h->slot_idx[slot] >= 0) return h->slot_idx[slot];
    bool used[3] = {false, false, false};
    for (int s = 0; s < 3; s++)
        if (h->slot_idx[s] >= 0) used[h->slot_idx[s]] = true;
    for (int k = 0; k < 3; k++)
        if (!used[k]) { h->slot_idx[slot] = k; return k; }
    return -1;
}

static bool slot_ok(int s) { return s >= 0 && s < CVO_NUM_SLOTS; }

extern "C" {

void cvo_default_params(cvo_params *p) {
    if (!p) return;
    p->ell_init = 0.15f;      // cvo.cpp:35
    p->sigma = 0.1f;          // :36
    p->sp_thres = 8e-3f;      // :37
    p->c = 7.0f;              // :38
    p->d = 7.0f;              // :39
    p->c_ell = 200.f;         // :41
    p->c_sigma = 1.f;         // :42
    p->max_iter = 2000;       // :48
    p->min_step = 2 * 1.0e-1f;  // :49
    p->max_step = 0.8f;       // :333
    p->eps = 5 * 1.0e-5f;     // :50
    p->eps_2 = 1.0e-5f;       // :51
    p->ell_after_k2 = 0.10f;  // :810
    p->ell_after_k9 = 0.06f;  // :811
    p->ell_after_k19 = 0.03f; // :812
    p->num_want = 3000;       // pcd_generator.cpp:22
    p->feature_type = 1;      // cvo.cpp:355,366
    p->gray_mode = 0;
    p->exp_mode = 0;
}

const char *cvo_last_error(void) { return g_err; }

int cvo_random_pattern(uint8_t *out, int n) {
    if (!out || n < 0) return CVO_ERR_INVALID;
    host_random_pattern(out, n);
    return CVO_OK;
}

int cvo_create(const cvo_calib *calib, const cvo_params *params, int device, cvo_handle **out) {
    if (!calib || !out) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(device));
    cvo_handle *h = new cvo_handle();
    h->device = device;
    h->cal = *calib;
    if (params) h->prm = *params;
    else cvo_default_params(&h->prm);
    identity_RT(h->R, h->T);      // cvo.cpp:66-67
    h->ell = h->prm.ell_init;     // cvo.cpp:35
    h->trace_cap = h->prm.max_iter > 4096 ? 4096 : h->prm.max_iter;
    if (h->trace_cap < 1) h->trace_cap = 1;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_task, sizeof(AlignTask));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_res, sizeof(cvo_align_result));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_trace, sizeof(cvo_iter_record) * h->trace_cap);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_q, sizeof(QueryTask) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_qo, sizeof(QueryOut) * 8);
    if (e == cudaSuccess) e = cudaMallocHost(&h->pinned, kPinnedBytes);
    if (e != cudaSuccess) {
        set_last_error("cvo_create: %s", cudaGetErrorString(e));
        cvo_destroy(h);
        return CVO_ERR_CUDA;
    }
    *out = h;
    return CVO_OK;
}

int cvo_destroy(cvo_handle *h) {
    if (!h) return CVO_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    sel_destroy(h->sel);
    align_ws_destroy(h->aws);
    if (h->arena.pos) arena_free(h->arena);
    cudaFree(h->d_task); cudaFree(h->d_res); cudaFree(h->d_trace); cudaFree(h->d_q); cudaFree(h->d_qo);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (int i = 0; i < 2; i++) {
        if (h->png_bgr[i]) cudaFreeHost(h->png_bgr[i]);
        if (h->png_depth[i]) cudaFreeHost(h->png_depth[i]);
        if (h->png_done[i]) cudaEventDestroy(h->png_done[i]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return CVO_OK;
}

static int handle_prepare_frame(cvo_handle *h, int slot, int w, int hgt, int *arena_index) {
    if (!h || !slot_ok(slot) || w < 64 || hgt < 64) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    if (!h->sel || h->sel_w != w || h->sel_h != hgt) {
        if (h->sel) { cudaStreamSynchronize(h->stream); sel_destroy(h->sel); h->sel = nullptr; }
        int rc = sel_create(&h->sel, w, hgt, 1);
        if (rc != CVO_OK) return rc;
        h->sel_w = w; h->sel_h = hgt;
    }
    int rc = handle_ensure_arena(h, cloud_capacity_for(h->prm, w, hgt));
    if (rc != CVO_OK) return rc;
    int k = handle_slot_arena_index(h, slot);
    if (k < 0) return CVO_ERR_INVALID;
    *arena_index = k;
    return CVO_OK;
}

int cvo_set_frame(cvo_handle *h, int slot, const uint8_t *bgr, size_t bgr_stride, const uint16_t *depth,
                  size_t depth_stride, int width, int height) {
    if (!bgr || !depth) return CVO_ERR_INVALID;
    int k;
    int rc = handle_prepare_frame(h, slot, width, height, &k);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpy2DAsync(sel_bgr_ptr(h->sel, 0), (size_t)width * 3, bgr, bgr_stride, (size_t)width * 3,
                                   height, cudaMemcpyHostToDevice, h->stream));
    CVO_CUDA_TRY(cudaMemcpy2DAsync(sel_depth_ptr(h->sel, 0), (size_t)width * 2, depth, depth_stride,
                                   (size_t)width * 2, height, cudaMemcpyHostToDevice, h->stream));
    h->launches += 0;
    rc = sel_run(h->sel, 1, nullptr, nullptr, h->cal, h->prm, h->arena, k, h->stream, &h->launches);
    h->sel_slot = slot;
    return rc;
}

// run_SLAM.cpp:134-143 + cvo::set_pcd in one call: the colour PNG (-> BGR8, as cv::imread) and the depth PNG
// (-> u16, as cv::imread with ANYDEPTH) are decoded on two host threads straight into pinned staging memory, then
// uploaded and selected on the handle's stream.  Two staging buffers alternate, so the decode of the next frame may
// start while the copy of this one is still in flight.
int cvo_set_frame_png(cvo_handle *h, int slot, const uint8_t *rgb_png, size_t rgb_bytes, const uint8_t *depth_png,
                      size_t depth_bytes) {
    if (!h || !slot_ok(slot) || !rgb_png || !depth_png) return CVO_ERR_INVALID;
    int w = 0, hgt = 0, wd = 0, hd = 0, ch = 0, bits = 0;
    int rc = cvo_png_info(rgb_png, rgb_bytes, &w, &hgt, &ch, &bits);
    if (rc == CVO_OK) rc = cvo_png_info(depth_png, depth_bytes, &wd, &hd, &ch, &bits);
    if (rc != CVO_OK) return rc;
    if (w != wd || hgt != hd) { set_last_error("cvo_set_frame_png: colour is %d x %d, depth %d x %d", w, hgt, wd, hd); return CVO_ERR_INVALID; }
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    const size_t px = (size_t)w * hgt;
    if (h->png_px != px) {
        CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 2; i++) {
            if (h->png_bgr[i]) cudaFreeHost(h->png_bgr[i]);
            if (h->png_depth[i]) cudaFreeHost(h->png_depth[i]);
            h->png_bgr[i] = nullptr; h->png_depth[i] = nullptr;
            CVO_CUDA_TRY(cudaMallocHost(&h->png_bgr[i], px * 3));
            CVO_CUDA_TRY(cudaMallocHost(&h->png_depth[i], px * 2));
            if (!h->png_done[i]) CVO_CUDA_TRY(cudaEventCreateWithFlags(&h->png_done[i], cudaEventDisableTiming));
        }
        h->png_px = px;
    }
    const int b = h->png_next;
    h->png_next ^= 1;
    CVO_CUDA_TRY(cudaEventSynchronize(h->png_done[b]));   // the copy that last read this buffer has finished
    int rc_d = CVO_OK;
    char err_d[256] = "";
    std::thread td([&] {
        rc_d = cvo_png_decode_depth16(depth_png, depth_bytes, h->png_depth[b], px, nullptr, nullptr);
        if (rc_d != CVO_OK) { strncpy(err_d, cvo_last_error(), sizeof(err_d) - 1); }
    });
    rc = cvo_png_decode_bgr8(rgb_png, rgb_bytes, h->png_bgr[b], px * 3, nullptr, nullptr);
    td.join();
    if (rc == CVO_OK && rc_d != CVO_OK) { set_last_error("%s", err_d); rc = rc_d; }
    if (rc != CVO_OK) return rc;
    rc = cvo_set_frame(h, slot, h->png_bgr[b], (size_t)w * 3, h->png_depth[b], (size_t)w * 2, w, hgt);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaEventRecord(h->png_done[b], h->stream));
    return CVO_OK;
}

int cvo_set_frame_device(cvo_handle *h, int slot, const uint8_t *bgr_dev, const uint16_t *depth_dev, int width,
                         int height) {
    if (!bgr_dev || !depth_dev) return CVO_ERR_INVALID;
    int k;
    int rc = handle_prepare_frame(h, slot, width, height, &k);
    if (rc != CVO_OK) return rc;
    rc = sel_run(h->sel, 1, bgr_dev, depth_dev, h->cal, h->prm, h->arena, k, h->stream, &h->launches);
    h->sel_slot = slot;
    return rc;
}

int cvo_set_cloud(cvo_handle *h, int slot, int n, const float *positions, const float *features) {
    if (!h || !slot_ok(slot) || n < 0 || (n > 0 && (!positions || !features))) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    int need = (n + 31) / 32 * 32;
    int dflt = cloud_capacity_for(h->prm, 640, 480);
    if (need < dflt) need = dflt;
    int rc = handle_ensure_arena(h, need);
    if (rc != CVO_OK) return rc;
    int k = handle_slot_arena_index(h, slot);
    if (k < 0) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::vector<float4> pos(n), f03(n);
    std::vector<float> f4(n);
    std::vector<float2> pix(n, make_float2(0.f, 0.f));
    for (int i = 0; i < n; i++) {
        pos[i] = make_float4(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2], 0.f);
        f03[i] = make_float4(features[5 * i], features[5 * i + 1], features[5 * i + 2], features[5 * i + 3]);
        f4[i] = features[5 * i + 4];
    }
    size_t o = (size_t)k * h->arena.cap;
    // on the handle's own (non-blocking) stream, so that the uploads are ordered before the kernels the next
    // calls launch there; the host vectors die at return, so wait for the copies
    const int zero = 0;
    if (n > 0) {
        CVO_CUDA_TRY(cudaMemcpyAsync(h->arena.pos + o, pos.data(), n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
        CVO_CUDA_TRY(cudaMemcpyAsync(h->arena.f03 + o, f03.data(), n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
        CVO_CUDA_TRY(cudaMemcpyAsync(h->arena.f4 + o, f4.data(), n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        CVO_CUDA_TRY(cudaMemcpyAsync(h->arena.pix + o, pix.data(), n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    }
    CVO_CUDA_TRY(cudaMemcpyAsync(h->arena.n + k, &n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CVO_CUDA_TRY(cudaMemcpyAsync(h->arena.ovf + k, &zero, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CVO_OK;
}

// LocalTracker feeds the same image to two cvo objects (local_tracker.cpp:356,415),
// so the reference selects its points twice.  A caller that knows this can select once and copy the
// device cloud (140 KB) into the other handle.
int cvo_copy_cloud(cvo_handle *dst, int dst_slot, cvo_handle *src, int src_slot) {
    if (!dst || !src || !slot_ok(dst_slot) || !slot_ok(src_slot)) return CVO_ERR_INVALID;
    if (src->slot_idx[src_slot] < 0) return CVO_ERR_NOT_INIT;
    if (dst->device != src->device) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(src->device));
    int rc = handle_ensure_arena(dst, src->arena.cap);
    if (rc != CVO_OK) return rc;
    const int kd = handle_slot_arena_index(dst, dst_slot);
    if (kd < 0) return CVO_ERR_INVALID;
    const int ks = src->slot_idx[src_slot];
    CVO_CUDA_TRY(cudaStreamSynchronize(src->stream));   // the cloud may still be in flight on the source stream
    const size_t so = (size_t)ks * src->arena.cap, d_o = (size_t)kd * dst->arena.cap, n = src->arena.cap;
    cudaStream_t st = dst->stream;
    CVO_CUDA_TRY(cudaMemcpyAsync(dst->arena.pos + d_o, src->arena.pos + so, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CVO_CUDA_TRY(cudaMemcpyAsync(dst->arena.f03 + d_o, src->arena.f03 + so, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CVO_CUDA_TRY(cudaMemcpyAsync(dst->arena.f4 + d_o, src->arena.f4 + so, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CVO_CUDA_TRY(cudaMemcpyAsync(dst->arena.pix + d_o, src->arena.pix + so, n * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    CVO_CUDA_TRY(cudaMemcpyAsync(dst->arena.n + kd, src->arena.n + ks, sizeof(int), cudaMemcpyDeviceToDevice, st));
    CVO_CUDA_TRY(cudaMemcpyAsync(dst->arena.ovf + kd, src->arena.ovf + ks, sizeof(int), cudaMemcpyDeviceToDevice, st));
    // the source handle's owner may overwrite the slot (set_frame on ITS stream) as soon as this returns
    CVO_CUDA_TRY(cudaStreamSynchronize(st));
    return CVO_OK;
}

int cvo_slot_move(cvo_handle *h, int dst, int src) {
    if (!h || !slot_ok(dst) || !slot_ok(src)) return CVO_ERR_INVALID;
    if (dst == src) return CVO_OK;
    h->slot_idx[dst] = h->slot_idx[src];   // the previous content of dst is dropped (unique_ptr move)
    h->slot_idx[src] = -1;
    return CVO_OK;
}

int cvo_slot_size(cvo_handle *h, int slot, int *n) {
    if (!h || !slot_ok(slot) || !n) return CVO_ERR_INVALID;
    *n = 0;
    if (h->slot_idx[slot] < 0) return CVO_ERR_NOT_INIT;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    int ovf = 0;
    CVO_CUDA_TRY(cudaMemcpyAsync(n, h->arena.n + h->slot_idx[slot], sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA_TRY(cudaMemcpyAsync(&ovf, h->arena.ovf + h->slot_idx[slot], sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ovf ? truncated_rc("cvo_slot_size") : CVO_OK;   // *n = the points kept
}

int cvo_set_RT(cvo_handle *h, const float R[9], const float T[3]) {
    if (!h || !R || !T) return CVO_ERR_INVALID;
    memcpy(h->R, R, sizeof(h->R));
    memcpy(h->T, T, sizeof(h->T));
    return CVO_OK;
}
int cvo_get_RT(cvo_handle *h, float R[9], float T[3]) {
    if (!h || !R || !T) return CVO_ERR_INVALID;
    memcpy(R, h->R, sizeof(h->R));
    memcpy(T, h->T, sizeof(h->T));
    return CVO_OK;
}
int cvo_set_ell(cvo_handle *h, float ell) { if (!h) return CVO_ERR_INVALID; h->ell = ell; return CVO_OK; }
int cvo_get_ell(cvo_handle *h, float *ell) { if (!h || !ell) return CVO_ERR_INVALID; *ell = h->ell; return CVO_OK; }

static int handle_run_align(cvo_handle *h, const float R[9], const float T[3], float ell, bool single,
                            cvo_align_result *out, cvo_iter_record *trace, int trace_cap) {
    if (!h) return CVO_ERR_INVALID;
    if (h->slot_idx[CVO_SLOT_FIXED] < 0 || h->slot_idx[CVO_SLOT_MOVING] < 0) return CVO_ERR_NOT_INIT;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    int rc = handle_ensure_aws(h);
    if (rc != CVO_OK) return rc;
    AlignTask *ht = reinterpret_cast<AlignTask *>(h->pinned);
    cvo_align_result *hr = reinterpret_cast<cvo_align_result *>(h->pinned + 1024);
    ht->fixed = h->arena.view(h->slot_idx[CVO_SLOT_FIXED]);
    ht->moving = h->arena.view(h->slot_idx[CVO_SLOT_MOVING]);
    memcpy(ht->R, R, sizeof(ht->R));
    memcpy(ht->T, T, sizeof(ht->T));
    ht->ell = ell;
    CVO_CUDA_TRY(cudaMemcpyAsync(h->d_task, ht, sizeof(AlignTask), cudaMemcpyHostToDevice, h->stream));
    int tc = trace ? (trace_cap < h->trace_cap ? trace_cap : h->trace_cap) : 0;
    rc = align_run(h->aws, h->prm, 1, h->d_task, h->d_res, tc ? h->d_trace : nullptr, tc, single, h->stream,
                   &h->launches);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpyAsync(hr, h->d_res, sizeof(cvo_align_result), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (out) *out = *hr;
    h->last_nnz = hr->A_nonzero;
    if (tc) {
        int n = hr->iterations < tc ? hr->iterations : tc;
        if (n > 0) {
            CVO_CUDA_TRY(cudaMemcpyAsync(trace, h->d_trace, sizeof(cvo_iter_record) * n, cudaMemcpyDeviceToHost, h->stream));
            CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
    }
    if (!single) {   // R, T and ell persist between align() calls
        memcpy(h->R, hr->R, sizeof(h->R));
        memcpy(h->T, hr->T, sizeof(h->T));
        h->ell = hr->ell;
    }
    return hr->status;
}

int cvo_align(cvo_handle *h, cvo_align_result *out, cvo_iter_record *trace, int trace_cap) {
    if (!h) return CVO_ERR_INVALID;
    return handle_run_align(h, h->R, h->T, h->ell, false, out, trace, trace_cap);
}

int cvo_iteration_at(cvo_handle *h, const float R[9], const float T[3], float ell, cvo_iter_record *out) {
    if (!h || !R || !T) return CVO_ERR_INVALID;
    cvo_iter_record rec;
    memset(&rec, 0, sizeof(rec));
    int rc = handle_run_align(h, R, T, ell, true, nullptr, &rec, 1);
    if (out) *out = rec;
    return rc;
}

int cvo_last_pattern(cvo_handle *h, int32_t *ij, float *a, int cap, int *n) {
    if (!h || !h->aws || !n) return CVO_ERR_INVALID;
    return align_last_pattern(h->aws, h->last_nnz, ij, a, cap, n, h->stream);
}

static int handle_query(cvo_handle *h, int slot_a, const float *Ta, int slot_b, int kind, QueryOut *res) {
    if (!h || !slot_ok(slot_a) || !slot_ok(slot_b)) return CVO_ERR_INVALID;
    if (h->slot_idx[slot_a] < 0 || h->slot_idx[slot_b] < 0) return CVO_ERR_NOT_INIT;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    int rc = handle_ensure_aws(h);
    if (rc != CVO_OK) return rc;
    QueryTask *hq = reinterpret_cast<QueryTask *>(h->pinned + 2048);
    QueryOut *ho = reinterpret_cast<QueryOut *>(h->pinned + 4096);
    hq->a = h->arena.view(h->slot_idx[slot_a]);
    hq->b = h->arena.view(h->slot_idx[slot_b]);
    static const float I34[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(hq->Ta, Ta ? Ta : I34, sizeof(hq->Ta));
    hq->ell = h->ell;
    hq->kind = kind;
    CVO_CUDA_TRY(cudaMemcpyAsync(h->d_q, hq, sizeof(QueryTask), cudaMemcpyHostToDevice, h->stream));
    rc = query_run(h->aws, h->prm, 1, h->d_q, h->d_qo, h->stream, &h->launches);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpyAsync(ho, h->d_qo, sizeof(QueryOut), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const bool trunc = decode_query_counts(ho, 1);
    *res = *ho;
    return trunc ? CVO_ERR_CAPACITY : CVO_OK;
}

int cvo_inner_product(cvo_handle *h, int slot_a, const float *Ta, int slot_b, float *value, int *num) {
    if (!value || !num) return CVO_ERR_INVALID;
    QueryOut q;
    int rc = handle_query(h, slot_a, Ta, slot_b, 0, &q);
    if (rc != CVO_OK && rc != CVO_ERR_CAPACITY) return rc;
    *value = (float)q.sum;                 // inn_p(float v, int n, int n_e)   cvo.hpp:71
    *num = q.count == 0 ? 1 : q.count;     // `if (sum == 0) sum = 1;`          cvo.cpp:455
    return rc == CVO_ERR_CAPACITY ? truncated_rc("cvo_inner_product") : CVO_OK;
}

int cvo_hessian(cvo_handle *h, int slot_a, const float *Ta, int slot_b, double H[36], int *inliers) {
    if (!H || !inliers) return CVO_ERR_INVALID;
    QueryOut q;
    int rc = handle_query(h, slot_a, Ta, slot_b, 1, &q);
    if (rc != CVO_OK && rc != CVO_ERR_CAPACITY) return rc;
    *inliers = q.count;
    finish_hessian_host(q, H);
    return rc == CVO_ERR_CAPACITY ? truncated_rc("cvo_hessian") : CVO_OK;
}

// cvo::compute_innerproduct (cvo.cpp:475-503) in one launch: the four inner products and the
// Hessian are five independent queries, one CTA each.
int cvo_compute_innerproduct(cvo_handle *h, const float tran[16], float values[4], int nums[4], double H[36],
                             int *inliers) {
    if (!h || !tran || !values || !nums || !H || !inliers) return CVO_ERR_INVALID;
    if (h->slot_idx[CVO_SLOT_FIXED] < 0 || h->slot_idx[CVO_SLOT_MOVING] < 0) return CVO_ERR_NOT_INIT;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    int rc = handle_ensure_aws(h);
    if (rc != CVO_OK) return rc;
    QueryTask *hq = reinterpret_cast<QueryTask *>(h->pinned + 8192);
    QueryOut *ho = reinterpret_cast<QueryOut *>(h->pinned + 16384);
    static const float I34[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const CloudView fx = h->arena.view(h->slot_idx[CVO_SLOT_FIXED]), mv = h->arena.view(h->slot_idx[CVO_SLOT_MOVING]);
    // {inn_pre, inn_post, inn_fixed_pcd, inn_moving_pcd, post_hessian}
    const CloudView qa[5] = {mv, mv, fx, mv, mv}, qb[5] = {fx, fx, fx, mv, fx};
    const float *qt[5] = {I34, tran, I34, I34, tran};
    for (int k = 0; k < 5; k++) {
        hq[k].a = qa[k];
        hq[k].b = qb[k];
        memcpy(hq[k].Ta, qt[k], sizeof(hq[k].Ta));
        hq[k].ell = h->ell;
        hq[k].kind = k == 4 ? 1 : 0;
    }
    CVO_CUDA_TRY(cudaMemcpyAsync(h->d_q, hq, sizeof(QueryTask) * 5, cudaMemcpyHostToDevice, h->stream));
    rc = query_run(h->aws, h->prm, 5, h->d_q, h->d_qo, h->stream, &h->launches);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpyAsync(ho, h->d_qo, sizeof(QueryOut) * 5, cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const bool trunc = decode_query_counts(ho, 5);
    for (int k = 0; k < 4; k++) {
        values[k] = (float)ho[k].sum;
        nums[k] = ho[k].count == 0 ? 1 : ho[k].count;
    }
    *inliers = ho[4].count;
    finish_hessian_host(ho[4], H);
    return trunc ? truncated_rc("cvo_compute_innerproduct") : CVO_OK;
}

// the accept rule of the reference's only caller (src/keyframe_graph.cpp:711-712)
static void finish_lc_record(cvo_lc_result *o) {
    o->cos_angle = o->value[3] / (sqrtf(o->value[4]) * sqrtf(o->value[5]));
    const bool reject = (o->value[3] <= o->value[2]) || (o->value[3] <= o->value[1]) ||
                        (o->value[3] <= o->value[0]) || o->cos_angle < 0.1f;
    o->accept = reject ? 0 : 1;
}

// cvo::compute_innerproduct_lc (cvo.cpp:505-561) in one launch: six inner products and two
// Hessians are eight independent queries, one CTA each.
int cvo_compute_innerproduct_lc(cvo_handle *h, const float prior_tran[16], const float lc_prior_tran[16],
                                const float lc_prior_tran_2[16], const float lc_tran[16], cvo_lc_result *out) {
    if (!h || !prior_tran || !lc_prior_tran || !lc_prior_tran_2 || !lc_tran || !out) return CVO_ERR_INVALID;
    if (h->slot_idx[CVO_SLOT_FIXED] < 0 || h->slot_idx[CVO_SLOT_MOVING] < 0) return CVO_ERR_NOT_INIT;
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    int rc = handle_ensure_aws(h);
    if (rc != CVO_OK) return rc;
    QueryTask *hq = reinterpret_cast<QueryTask *>(h->pinned + 8192);
    QueryOut *ho = reinterpret_cast<QueryOut *>(h->pinned + 16384);
    static_assert(sizeof(QueryTask) * 8 <= 8192 && sizeof(QueryOut) * 8 <= kPinnedBytes - 16384, "pinned staging");
    static const float I34[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const CloudView fx = h->arena.view(h->slot_idx[CVO_SLOT_FIXED]), mv = h->arena.view(h->slot_idx[CVO_SLOT_MOVING]);
    // {inn_prior, inn_lc_prior, inn_lc_pre, inn_lc_post, inn_fixed_pcd, inn_moving_pcd, post_hessian, pnp-ransac Hessian}
    const CloudView qa[8] = {mv, mv, mv, mv, fx, mv, mv, mv}, qb[8] = {fx, fx, fx, fx, fx, mv, fx, fx};
    const float *qt[8] = {prior_tran, lc_prior_tran, I34, lc_tran, I34, I34, lc_tran, lc_prior_tran_2};
    for (int k = 0; k < 8; k++) {
        hq[k].a = qa[k];
        hq[k].b = qb[k];
        memcpy(hq[k].Ta, qt[k], sizeof(hq[k].Ta));
        hq[k].ell = h->ell;
        hq[k].kind = k >= 6 ? 1 : 0;
    }
    CVO_CUDA_TRY(cudaMemcpyAsync(h->d_q, hq, sizeof(QueryTask) * 8, cudaMemcpyHostToDevice, h->stream));
    rc = query_run(h->aws, h->prm, 8, h->d_q, h->d_qo, h->stream, &h->launches);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpyAsync(ho, h->d_qo, sizeof(QueryOut) * 8, cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const bool trunc = decode_query_counts(ho, 8);
    for (int k = 0; k < 6; k++) {
        out->value[k] = (float)ho[k].sum;
        out->num[k] = ho[k].count == 0 ? 1 : ho[k].count;
    }
    out->inliers_svd = ho[6].count;
    finish_hessian_host(ho[6], out->post_hessian);
    out->inliers_pnpransac = ho[7].count;
    finish_lc_record(out);
    return trunc ? truncated_rc("cvo_compute_innerproduct_lc") : CVO_OK;
}

int cvo_get_selected_points(cvo_handle *h, int slot, float *xy, int cap, int *n) {
    if (!h || !slot_ok(slot) || !n) return CVO_ERR_INVALID;
    int m = 0;
    int rc = cvo_slot_size(h, slot, &m);
    if (rc != CVO_OK) return rc;
    *n = m;
    int c = m < cap ? m : cap;
    if (c > 0 && xy) {
        size_t o = (size_t)h->slot_idx[slot] * h->arena.cap;
        CVO_CUDA_TRY(cudaMemcpyAsync(xy, h->arena.pix + o, sizeof(float2) * c, cudaMemcpyDeviceToHost, h->stream));
        CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return CVO_OK;
}

int cvo_get_selected_points_device(cvo_handle *h, int slot, const float **xy_dev, int *n) {
    if (!h || !slot_ok(slot) || !xy_dev || !n) return CVO_ERR_INVALID;
    int m = 0;
    int rc = cvo_slot_size(h, slot, &m);   // synchronises the handle's stream
    if (rc != CVO_OK) return rc;
    *n = m;
    *xy_dev = reinterpret_cast<const float *>(h->arena.pix + (size_t)h->slot_idx[slot] * h->arena.cap);
    return CVO_OK;
}

int cvo_get_cloud(cvo_handle *h, int slot, float *positions, float *features, int cap, int *n) {
    if (!h || !slot_ok(slot) || !n) return CVO_ERR_INVALID;
    int m = 0;
    int rc = cvo_slot_size(h, slot, &m);
    if (rc != CVO_OK) return rc;
    *n = m;
    int c = m < cap ? m : cap;
    if (c <= 0) return CVO_OK;
    size_t o = (size_t)h->slot_idx[slot] * h->arena.cap;
    std::vector<float4> pos(c), f03(c);
    std::vector<float> f4(c);
    CVO_CUDA_TRY(cudaMemcpy(pos.data(), h->arena.pos + o, sizeof(float4) * c, cudaMemcpyDeviceToHost));
    CVO_CUDA_TRY(cudaMemcpy(f03.data(), h->arena.f03 + o, sizeof(float4) * c, cudaMemcpyDeviceToHost));
    CVO_CUDA_TRY(cudaMemcpy(f4.data(), h->arena.f4 + o, sizeof(float) * c, cudaMemcpyDeviceToHost));
    for (int i = 0; i < c; i++) {
        if (positions) { positions[3 * i] = pos[i].x; positions[3 * i + 1] = pos[i].y; positions[3 * i + 2] = pos[i].z; }
        if (features) {
            features[5 * i] = f03[i].x; features[5 * i + 1] = f03[i].y; features[5 * i + 2] = f03[i].z;
            features[5 * i + 3] = f03[i].w; features[5 * i + 4] = f4[i];
        }
    }
    return CVO_OK;
}

// Keyframe re-does the gray conversion of the image the CVO front end has just converted
// (include/keyframe.h:34-55).  The selection keeps the gray image of the LAST frame set on this handle on the device.
int cvo_get_gray_device(cvo_handle *h, int slot, const uint8_t **gray_dev, int *width, int *height) {
    if (!h || !slot_ok(slot) || !gray_dev) return CVO_ERR_INVALID;
    if (!h->sel || h->sel_slot != slot) return CVO_ERR_NOT_INIT;   // only the last selected frame is kept
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    CVO_CUDA_TRY(cudaStreamSynchronize(h->stream));   // work queued on the handle's stream has completed
    *gray_dev = sel_gray_ptr(h->sel, 0);
    if (width) *width = h->sel_w;
    if (height) *height = h->sel_h;
    return CVO_OK;
}

int cvo_get_selection_debug(cvo_handle *h, int slot, uint8_t *map, int32_t info[5]) {
    if (!h || !slot_ok(slot) || !info) return CVO_ERR_INVALID;
    if (!h->sel || h->sel_slot != slot) return CVO_ERR_NOT_INIT;   // only the last selected frame is kept
    CVO_CUDA_TRY(cudaSetDevice(h->device));
    return sel_debug(h->sel, 0, map, info, h->stream);
}

int cvo_handle_stats(cvo_handle *h, int64_t stats[4]) {
    if (!h || !stats) return CVO_ERR_INVALID;
    stats[0] = h->launches;
    stats[1] = stats[2] = stats[3] = 0;
    if (h->aws) align_ws_stats(h->aws, h->stream, stats + 1);
    return CVO_OK;
}

int cvo_handle_phase_cycles(cvo_handle *h, int64_t cycles[8]) {
    if (!h || !cycles) return CVO_ERR_INVALID;
    for (int i = 0; i < 8; i++) cycles[i] = 0;
    if (h->aws) align_ws_phase_cycles(h->aws, h->stream, cycles);
    return CVO_OK;
}

int cvo_batch_phase_cycles(cvo_batch *b, int64_t cycles[8]) {
    if (!b || !cycles) return CVO_ERR_INVALID;
    align_ws_phase_cycles(b->aws, b->stream, cycles);
    return CVO_OK;
}

// ---- batches -----------------------------------------------------------------------------------

int cvo_batch_create(const cvo_calib *calib, const cvo_params *params, int device, int max_frames, int max_pairs,
                     int width, int height, cvo_batch **out) {
    if (!calib || !out || max_frames < 1 || max_pairs < 1 || width < 64 || height < 64) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(device));
    cvo_batch *b = new cvo_batch();
    b->device = device;
    b->cal = *calib;
    if (params) b->prm = *params;
    else cvo_default_params(&b->prm);
    b->w = width; b->h = height; b->max_frames = max_frames; b->max_pairs = max_pairs;
    b->chunk = max_frames < 148 ? max_frames : 148;   // one k_compact CTA per frame: a chunk fills the SMs
    int rc = arena_alloc(b->arena, max_frames, cloud_capacity_for(b->prm, width, height));
    cudaError_t e = cudaSuccess;
    if (rc == CVO_OK) rc = sel_create(&b->sel[0], width, height, b->chunk);
    if (rc == CVO_OK) rc = sel_create(&b->sel[1], width, height, b->chunk);
    if (rc == CVO_OK) rc = align_ws_create(&b->aws, b->arena.cap < 65536 ? b->arena.cap : 65536, device, 0);
    if (rc == CVO_OK) {
        e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->sel_stream, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&b->up[i].ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->compute_done, cudaEventDisableTiming);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&b->sel_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
        if (e == cudaSuccess) e = cudaEventCreate(&b->ev_user[0]);
        if (e == cudaSuccess) e = cudaEventCreate(&b->ev_user[1]);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_tasks, sizeof(AlignTask) * max_pairs);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_results, sizeof(cvo_align_result) * max_pairs);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_q, sizeof(QueryTask) * max_pairs);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_qo, sizeof(QueryOut) * max_pairs);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_tasks, sizeof(AlignTask) * max_pairs);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_results, sizeof(cvo_align_result) * max_pairs);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_q, sizeof(QueryTask) * max_pairs);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_qo, sizeof(QueryOut) * max_pairs);
        if (e != cudaSuccess) { set_last_error("cvo_batch_create: %s", cudaGetErrorString(e)); rc = CVO_ERR_CUDA; }
    }
    if (rc != CVO_OK) { cvo_batch_destroy(b); return rc; }
    *out = b;
    return CVO_OK;
}

int cvo_batch_destroy(cvo_batch *b) {
    if (!b) return CVO_OK;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    sel_destroy(b->sel[0]); sel_destroy(b->sel[1]);
    align_ws_destroy(b->aws);
    if (b->arena.pos) arena_free(b->arena);
    cudaFree(b->d_tasks); cudaFree(b->d_results); cudaFree(b->d_q); cudaFree(b->d_qo);
    cudaFree(b->d_lc); cudaFree(b->d_lco); cudaFree(b->d_lcq); cudaFree(b->d_lcqo);
    if (b->h_lc) cudaFreeHost(b->h_lc);
    if (b->h_lco) cudaFreeHost(b->h_lco);
    if (b->h_lcq) cudaFreeHost(b->h_lcq);
    if (b->h_lcqo) cudaFreeHost(b->h_lcqo);
    if (b->h_tasks) cudaFreeHost(b->h_tasks);
    if (b->h_results) cudaFreeHost(b->h_results);
    if (b->h_q) cudaFreeHost(b->h_q);
    if (b->h_qo) cudaFreeHost(b->h_qo);
    for (int i = 0; i < 2; i++) if (b->sel_done[i]) cudaEventDestroy(b->sel_done[i]);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    for (int i = 0; i < 2; i++) if (b->ev_user[i]) cudaEventDestroy(b->ev_user[i]);
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    if (b->sel_stream) cudaStreamDestroy(b->sel_stream);
    for (int i = 0; i < 2; i++) if (b->up[i].ev) cudaEventDestroy(b->up[i].ev);
    if (b->compute_done) cudaEventDestroy(b->compute_done);
    delete b;
    return CVO_OK;
}

// Host images: chunks alternate between two staging workspaces so that the H2D copy of chunk
// c+1 (copy stream) overlaps the selection kernels of chunk c (selection stream).  The call only enqueues:
// it returns while the copies are in flight (the host images must stay valid until a call that consumes
// these frames has returned), and neither stream is ordered behind the alignment of OTHER frames — a caller
// that uploads the frames of step k+1 into a second range of the arena before it aligns step k has the
// upload run beside that alignment.
int cvo_batch_set_frames(cvo_batch *b, int first, int n, const uint8_t *bgr, const uint16_t *depth) {
    if (!b || !bgr || !depth || first < 0 || n < 0 || first + n > b->max_frames) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    if (n == 0) return CVO_OK;
    if (b->busy_lo < first + n && first < b->busy_hi) {   // these frames are being read: behind everything enqueued so far
        CVO_CUDA_TRY(cudaStreamWaitEvent(b->sel_stream, b->compute_done, 0));
        b->busy_lo = b->busy_hi = 0;
    }
    cvo_batch::Upload &up = b->up[b->up_next & 1];
    b->up_next++;
    if (up.valid) CVO_CUDA_TRY(cudaStreamWaitEvent(b->sel_stream, up.ev, 0));   // (same stream: keeps the slot's meaning simple)
    const size_t fb = (size_t)b->w * b->h * 3, fd = (size_t)b->w * b->h;
    int ci = 0;
    cudaEvent_t copied;
    CVO_CUDA_TRY(cudaEventCreateWithFlags(&copied, cudaEventDisableTiming));
    for (int off = 0; off < n; off += b->chunk, ci++) {
        const int m = (n - off) < b->chunk ? (n - off) : b->chunk;
        SelWorkspace *ws = b->sel[ci & 1];
        // staging buffer `ci & 1` is free once the selection that last used it has finished
        CVO_CUDA_TRY(cudaStreamWaitEvent(b->copy_stream, b->sel_done[ci & 1], 0));
        CVO_CUDA_TRY(cudaMemcpy2DAsync(sel_bgr_ptr(ws, 0), sel_frame_bytes_bgr(ws), bgr + (size_t)off * fb, fb, fb, m,
                                       cudaMemcpyHostToDevice, b->copy_stream));
        const size_t dstride = (size_t)((char *)sel_depth_ptr(ws, 1) - (char *)sel_depth_ptr(ws, 0));
        CVO_CUDA_TRY(cudaMemcpy2DAsync(sel_depth_ptr(ws, 0), dstride, depth + (size_t)off * fd, fd * 2, fd * 2, m,
                                       cudaMemcpyHostToDevice, b->copy_stream));
        CVO_CUDA_TRY(cudaEventRecord(copied, b->copy_stream));
        CVO_CUDA_TRY(cudaStreamWaitEvent(b->sel_stream, copied, 0));
        int rc = sel_run(ws, m, nullptr, nullptr, b->cal, b->prm, b->arena, first + off, b->sel_stream, &b->launches);
        if (rc != CVO_OK) { cudaEventDestroy(copied); return rc; }
        CVO_CUDA_TRY(cudaEventRecord(b->sel_done[ci & 1], b->sel_stream));
    }
    cudaEventDestroy(copied);
    up.lo = first; up.hi = first + n; up.valid = true;
    CVO_CUDA_TRY(cudaEventRecord(up.ev, b->sel_stream));
    return CVO_OK;
}

int cvo_batch_set_frames_device(cvo_batch *b, int first, int n, const uint8_t *bgr_dev, const uint16_t *depth_dev) {
    if (!b || !bgr_dev || !depth_dev || first < 0 || n < 0 || first + n > b->max_frames) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    const size_t fb = (size_t)b->w * b->h * 3, fd = (size_t)b->w * b->h;
    if (n == 0) return CVO_OK;
    int rc0 = batch_consume(b, first, first + n);   // (a writer on the compute stream: ordered like a consumer)
    if (rc0 != CVO_OK) return rc0;
    // the selection workspace is shared with the uploads of host images
    for (int u = 0; u < 2; u++) if (b->up[u].valid) CVO_CUDA_TRY(cudaStreamWaitEvent(b->stream, b->up[u].ev, 0));
    for (int off = 0; off < n; off += b->chunk) {
        const int m = (n - off) < b->chunk ? (n - off) : b->chunk;
        int rc = sel_run(b->sel[0], m, bgr_dev + (size_t)off * fb, depth_dev + (size_t)off * fd, b->cal, b->prm,
                         b->arena, first + off, b->stream, &b->launches);
        if (rc != CVO_OK) return rc;
    }
    CVO_CUDA_TRY(cudaEventRecord(b->sel_done[0], b->stream));   // (sel[0]'s staging / scratch is busy until here)
    return batch_consumed(b);
}

int cvo_batch_frame_size(cvo_batch *b, int frame, int *n) {
    if (!b || !n || frame < 0 || frame >= b->max_frames) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    int ovf = 0;
    { int rc = batch_consume(b, frame, frame + 1); if (rc != CVO_OK) return rc; }
    CVO_CUDA_TRY(cudaMemcpyAsync(n, b->arena.n + frame, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CVO_CUDA_TRY(cudaMemcpyAsync(&ovf, b->arena.ovf + frame, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return ovf ? truncated_rc("cvo_batch_frame_size") : CVO_OK;
}

int cvo_batch_align(cvo_batch *b, int n_pairs, const cvo_pair_desc *pairs, cvo_align_result *results) {
    if (!b || !pairs || !results || n_pairs < 0 || n_pairs > b->max_pairs) return CVO_ERR_INVALID;
    if (n_pairs == 0) return CVO_OK;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    for (int i = 0; i < n_pairs; i++) {
        const cvo_pair_desc &p = pairs[i];
        if (p.fixed_frame < 0 || p.fixed_frame >= b->max_frames || p.moving_frame < 0 || p.moving_frame >= b->max_frames)
            return CVO_ERR_INVALID;
        AlignTask &t = b->h_tasks[i];
        t.fixed = b->arena.view(p.fixed_frame);
        t.moving = b->arena.view(p.moving_frame);
        memcpy(t.R, p.R, sizeof(t.R));
        memcpy(t.T, p.T, sizeof(t.T));
        t.ell = p.ell;
    }
    {
        int lo, hi;
        pair_frame_range(pairs, n_pairs, lo, hi);
        int rc = batch_consume(b, lo, hi);
        if (rc != CVO_OK) return rc;
    }
    CVO_CUDA_TRY(cudaMemcpyAsync(b->d_tasks, b->h_tasks, sizeof(AlignTask) * n_pairs, cudaMemcpyHostToDevice, b->stream));
    CVO_CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    int rc = align_run(b->aws, b->prm, n_pairs, b->d_tasks, b->d_results, nullptr, 0, false, b->stream, &b->launches);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    { int rc2 = batch_consumed(b); if (rc2 != CVO_OK) return rc2; }
    CVO_CUDA_TRY(cudaMemcpyAsync(b->h_results, b->d_results, sizeof(cvo_align_result) * n_pairs, cudaMemcpyDeviceToHost,
                                 b->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(b->stream));
    memcpy(results, b->h_results, sizeof(cvo_align_result) * n_pairs);
    cudaEventElapsedTime(&b->last_align_ms, b->ev0, b->ev1);
    return CVO_OK;
}

int cvo_batch_inner_product(cvo_batch *b, int n_pairs, const cvo_pair_desc *pairs, const cvo_align_result *results,
                            float *values, int *nums) {
    if (!b || !pairs || !results || !values || !nums || n_pairs < 0 || n_pairs > b->max_pairs) return CVO_ERR_INVALID;
    if (n_pairs == 0) return CVO_OK;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    for (int i = 0; i < n_pairs; i++) {
        if (pairs[i].fixed_frame < 0 || pairs[i].fixed_frame >= b->max_frames || pairs[i].moving_frame < 0 ||
            pairs[i].moving_frame >= b->max_frames)
            return CVO_ERR_INVALID;
        QueryTask &q = b->h_q[i];
        q.a = b->arena.view(pairs[i].moving_frame);
        q.b = b->arena.view(pairs[i].fixed_frame);
        memcpy(q.Ta, results[i].transform, sizeof(q.Ta));   // first 3 rows of the 4x4
        q.ell = results[i].ell;
        q.kind = 0;
    }
    {
        int lo, hi;
        pair_frame_range(pairs, n_pairs, lo, hi);
        int rc0 = batch_consume(b, lo, hi);
        if (rc0 != CVO_OK) return rc0;
    }
    CVO_CUDA_TRY(cudaMemcpyAsync(b->d_q, b->h_q, sizeof(QueryTask) * n_pairs, cudaMemcpyHostToDevice, b->stream));
    int rc = query_run(b->aws, b->prm, n_pairs, b->d_q, b->d_qo, b->stream, &b->launches);
    if (rc != CVO_OK) return rc;
    rc = batch_consumed(b);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpyAsync(b->h_qo, b->d_qo, sizeof(QueryOut) * n_pairs, cudaMemcpyDeviceToHost, b->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(b->stream));
    const bool trunc = decode_query_counts(b->h_qo, n_pairs);
    for (int i = 0; i < n_pairs; i++) {
        values[i] = (float)b->h_qo[i].sum;
        nums[i] = b->h_qo[i].count == 0 ? 1 : b->h_qo[i].count;
    }
    return trunc ? truncated_rc("cvo_batch_inner_product") : CVO_OK;
}

// compute_innerproduct_lc for every pair of a batch.  One k_verify_lc CTA per pair evaluates the
// four inner products and two Hessians against the fixed frame on one grid and finishes the Hessian
// on the device; the self inner products <fixed,fixed>, <moving,moving> run once per distinct
// (frame, ell) through k_query on the same stream.
int cvo_batch_verify_lc(cvo_batch *b, int n_pairs, const cvo_pair_desc *pairs, const cvo_align_result *results,
                        const float *prior_tran, const float *lc_prior_tran, const float *lc_prior_tran_2,
                        cvo_lc_result *out) {
    if (!b || !pairs || !results || !prior_tran || !lc_prior_tran || !lc_prior_tran_2 || !out || n_pairs < 0 ||
        n_pairs > b->max_pairs)
        return CVO_ERR_INVALID;
    if (n_pairs == 0) return CVO_OK;
    for (int i = 0; i < n_pairs; i++)
        if (pairs[i].fixed_frame < 0 || pairs[i].fixed_frame >= b->max_frames || pairs[i].moving_frame < 0 ||
            pairs[i].moving_frame >= b->max_frames)
            return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    static const float I34[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (!b->d_lc) {   // staging sized for max_pairs, allocated at the first verification (all or nothing)
        const size_t np = (size_t)b->max_pairs;
        cudaError_t e = cudaMalloc(&b->d_lc, sizeof(LcTask) * np);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_lco, sizeof(LcOut) * np);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_lc, sizeof(LcTask) * np);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_lco, sizeof(LcOut) * np);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_lcq, sizeof(QueryTask) * 2 * np);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_lcqo, sizeof(QueryOut) * 2 * np);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_lcq, sizeof(QueryTask) * 2 * np);
        if (e == cudaSuccess) e = cudaMallocHost(&b->h_lcqo, sizeof(QueryOut) * 2 * np);
        if (e != cudaSuccess) {
            cudaFree(b->d_lc); cudaFree(b->d_lco); cudaFree(b->d_lcq); cudaFree(b->d_lcqo);
            if (b->h_lc) cudaFreeHost(b->h_lc);
            if (b->h_lco) cudaFreeHost(b->h_lco);
            if (b->h_lcq) cudaFreeHost(b->h_lcq);
            if (b->h_lcqo) cudaFreeHost(b->h_lcqo);
            b->d_lc = nullptr; b->d_lco = nullptr; b->d_lcq = nullptr; b->d_lcqo = nullptr;
            b->h_lc = nullptr; b->h_lco = nullptr; b->h_lcq = nullptr; b->h_lcqo = nullptr;
            set_last_error("cvo_batch_verify_lc: staging allocation failed: %s", cudaGetErrorString(e));
            return CVO_ERR_CUDA;
        }
    }
    std::map<std::pair<int, uint32_t>, int> self_of;
    std::vector<int> self_fx(n_pairs), self_mv(n_pairs);
    int nq = 0;
    auto self_task = [&](int frame, float ell) {
        uint32_t bits;
        memcpy(&bits, &ell, 4);
        auto key = std::make_pair(frame, bits);
        auto it = self_of.find(key);
        if (it != self_of.end()) return it->second;
        QueryTask &t = b->h_lcq[nq];
        t.a = b->arena.view(frame);
        t.b = t.a;
        memcpy(t.Ta, I34, sizeof(t.Ta));
        t.ell = ell;
        t.kind = 0;
        self_of[key] = nq;
        return nq++;
    };
    for (int i = 0; i < n_pairs; i++) {
        LcTask &t = b->h_lc[i];
        t.a = b->arena.view(pairs[i].moving_frame);
        t.b = b->arena.view(pairs[i].fixed_frame);
        const float *qt[6] = {prior_tran + 16 * (size_t)i, lc_prior_tran + 16 * (size_t)i, I34, results[i].transform,
                              results[i].transform, lc_prior_tran_2 + 16 * (size_t)i};
        for (int k = 0; k < 6; k++) memcpy(t.T[k], qt[k], sizeof(t.T[k]));
        t.ell = results[i].ell;
        self_fx[i] = self_task(pairs[i].fixed_frame, results[i].ell);
        self_mv[i] = self_task(pairs[i].moving_frame, results[i].ell);
    }
    {
        int lo, hi;
        pair_frame_range(pairs, n_pairs, lo, hi);
        int rc0 = batch_consume(b, lo, hi);
        if (rc0 != CVO_OK) return rc0;
    }
    CVO_CUDA_TRY(cudaMemcpyAsync(b->d_lc, b->h_lc, sizeof(LcTask) * n_pairs, cudaMemcpyHostToDevice, b->stream));
    CVO_CUDA_TRY(cudaMemcpyAsync(b->d_lcq, b->h_lcq, sizeof(QueryTask) * nq, cudaMemcpyHostToDevice, b->stream));
    int rc = lc_run(b->aws, b->prm, n_pairs, b->d_lc, b->d_lco, b->stream, &b->launches);
    if (rc != CVO_OK) return rc;
    rc = query_run(b->aws, b->prm, nq, b->d_lcq, b->d_lcqo, b->stream, &b->launches);
    if (rc == CVO_OK) rc = batch_consumed(b);
    if (rc != CVO_OK) return rc;
    CVO_CUDA_TRY(cudaMemcpyAsync(b->h_lco, b->d_lco, sizeof(LcOut) * n_pairs, cudaMemcpyDeviceToHost, b->stream));
    CVO_CUDA_TRY(cudaMemcpyAsync(b->h_lcqo, b->d_lcqo, sizeof(QueryOut) * nq, cudaMemcpyDeviceToHost, b->stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(b->stream));
    bool lc_trunc = decode_query_counts(b->h_lcqo, nq);
    for (int i = 0; i < n_pairs; i++) lc_trunc = lc_trunc || b->h_lco[i].truncated != 0;
    for (int i = 0; i < n_pairs; i++) {
        const LcOut &o = b->h_lco[i];
        cvo_lc_result &r = out[i];
        for (int k = 0; k < 4; k++) {
            r.value[k] = (float)o.sum[k];
            r.num[k] = o.count[k] == 0 ? 1 : o.count[k];
        }
        const QueryOut *self[2] = {b->h_lcqo + self_fx[i], b->h_lcqo + self_mv[i]};
        for (int k = 0; k < 2; k++) {
            r.value[4 + k] = (float)self[k]->sum;
            r.num[4 + k] = self[k]->count == 0 ? 1 : self[k]->count;
        }
        r.inliers_svd = o.inliers[0];
        memcpy(r.post_hessian, o.H, sizeof(r.post_hessian));
        r.inliers_pnpransac = o.inliers[1];
        finish_lc_record(&r);
    }
    return lc_trunc ? truncated_rc("cvo_batch_verify_lc") : CVO_OK;
}

int cvo_batch_stats(cvo_batch *b, int64_t stats[4]) {
    if (!b || !stats) return CVO_ERR_INVALID;
    stats[0] = b->launches;
    align_ws_stats(b->aws, b->stream, stats + 1);
    return CVO_OK;
}

int cvo_batch_mark(cvo_batch *b, int which) {
    if (!b || which < 0 || which > 1) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaSetDevice(b->device));
    CVO_CUDA_TRY(cudaEventRecord(b->ev_user[which], b->stream));
    return CVO_OK;
}

int cvo_batch_elapsed_ms(cvo_batch *b, float *ms) {
    if (!b || !ms) return CVO_ERR_INVALID;
    CVO_CUDA_TRY(cudaEventSynchronize(b->ev_user[1]));
    CVO_CUDA_TRY(cudaEventElapsedTime(ms, b->ev_user[0], b->ev_user[1]));
    return CVO_OK;
}

int cvo_batch_last_align_ms(cvo_batch *b, float *ms) {
    if (!b || !ms) return CVO_ERR_INVALID;
    *ms = b->last_align_ms;
    return CVO_OK;
}

}  // extern "C"
